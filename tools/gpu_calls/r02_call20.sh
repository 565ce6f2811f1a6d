#!/bin/bash
set -u
O=gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload head --batch 1024 --steps 40 --warmup 5 --no-head-check $ARGS > $O/r02t_head_$tag.json 2> $O/r02t_head_$tag.err
  echo "head $tag rc=$? : $(python - <<PY
import json
try:
    d=json.load(open("$O/r02t_head_$tag.json"))
    print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"] if "gemm" in r["kernel"]])
except Exception as e: print("ERR", e)
PY
)"
}
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
run raw_ew16_a X=1
run raw_ew8_a MSML_HEAD_EW16=0
run raw_ew16_b X=1
run raw_ew8_b MSML_HEAD_EW16=0
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
