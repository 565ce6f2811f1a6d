#!/bin/bash
set -u
O=gpurun_out
MSML_HEAD_EW16=1 timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py tests/test_gpu_pfc_sgd.py -q -x -k "not gemm" > $O/r02r_pytest_ew16.log 2>&1
echo "tests ew16 rc=$? : $(tail -1 $O/r02r_pytest_ew16.log)"
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload head --batch 1024 --steps 30 --warmup 5 --no-head-check $ARGS > $O/r02r_head_$tag.json 2> $O/r02r_head_$tag.err
  echo "head $tag rc=$? : $(python - <<PY
import json
try:
    d=json.load(open("$O/r02r_head_$tag.json"))
    print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"]])
except Exception as e: print("ERR", e)
PY
)"
}
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
run 125k_raw_ew16 MSML_HEAD_EW16=1
run 125k_raw_ew8 X=1
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd --exact-head-grad"
run 125k_exact_ew16 MSML_HEAD_EW16=1
