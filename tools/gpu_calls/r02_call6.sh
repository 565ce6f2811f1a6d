#!/bin/bash
# Round-2 GPU call 6 (1 GPU): resident-A forward kernel, pairs for dX only; A/B timing; head + pfc_sgd + model tests.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py tests/test_gpu_pfc_sgd.py tests/test_gpu_model.py tests/test_gpu_margins.py -q -x > $O/r02f_pytest.log 2>&1
echo "tests rc=$? : $(tail -1 $O/r02f_pytest.log)"
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload head --batch 1024 --steps 30 --warmup 5 --no-head-check $ARGS > $O/r02f_head_$tag.json 2> $O/r02f_head_$tag.err
  echo "head $tag rc=$? : $(python - <<PY
import json
try:
    d=json.load(open("$O/r02f_head_$tag.json"))
    print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"]])
except Exception as e: print("ERR", e)
PY
)"
}
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
run 125k_default X=1
run 125k_noares MSML_HEAD_ARES=0
run 125k_default2 X=1
ARGS="--classes 11679 --sample-rate 1.0 --fused-sgd"
run 11679_default X=1
