#!/bin/bash
# Round-2 GPU call 7 (1 GPU): warp-converged elected-lane MMA issue; A/B over pair modes.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py tests/test_gpu_margins.py -q -x > $O/r02g_pytest.log 2>&1
echo "tests rc=$? : $(tail -1 $O/r02g_pytest.log)"
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload head --batch 1024 --steps 30 --warmup 5 --no-head-check $ARGS > $O/r02g_head_$tag.json 2> $O/r02g_head_$tag.err
  echo "head $tag rc=$? : $(python - <<PY
import json
try:
    d=json.load(open("$O/r02g_head_$tag.json"))
    print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"]])
except Exception as e: print("ERR", e)
PY
)"
}
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
run 125k_default X=1
run 125k_pairall MSML_HEAD_PAIR=all
run 125k_nopair MSML_HEAD_PAIR=0
ARGS="--classes 11679 --sample-rate 1.0 --fused-sgd"
run 11679_default X=1
run 11679_pairall MSML_HEAD_PAIR=all
ARGS="--classes 93431 --batch 128 --sample-rate 1.0 --fused-sgd"
run 93431_w1 X=1
