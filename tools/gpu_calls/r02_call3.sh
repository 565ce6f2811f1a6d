#!/bin/bash
# Round-2 GPU call 3 (1 GPU): CTA-pair GEMM bring-up, blocked dcos layout, sampled path after the select fix.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -k "gemm" > $O/r02c_pytest_gemm.log 2>&1
echo "gemm tests rc=$? : $(tail -1 $O/r02c_pytest_gemm.log)"
timeout 600 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py tests/test_gpu_pfc_sgd.py tests/test_gpu_margins.py -q -x -k "not gemm" > $O/r02c_pytest_head.log 2>&1
echo "head tests rc=$? : $(tail -1 $O/r02c_pytest_head.log)"
MSML_HEAD_PAIR=0 timeout 600 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py -q -x -k "not gemm" > $O/r02c_pytest_head_nopair.log 2>&1
echo "head tests (no pair) rc=$? : $(tail -1 $O/r02c_pytest_head_nopair.log)"
run() {  # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload head --batch 1024 --steps 20 --warmup 5 --no-head-check $ARGS > $O/r02c_head_$tag.json 2> $O/r02c_head_$tag.err
  echo "head $tag rc=$? : $(python - <<PY
import json
try:
    d=json.load(open("$O/r02c_head_$tag.json"))
    print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"]])
except Exception as e: print("ERR", e)
PY
)"
}
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
run 125k_default X=1
run 125k_nopair MSML_HEAD_PAIR=0
run 125k_r1 MSML_HEAD_PAIR=0 MSML_HEAD_DCOS_ROWMAJOR=1
run 125k_pair_rowmajor MSML_HEAD_DCOS_ROWMAJOR=1
ARGS="--classes 11679 --sample-rate 1.0 --fused-sgd"
run 11679_default X=1
run 11679_nopair MSML_HEAD_PAIR=0
run 11679_r1 MSML_HEAD_PAIR=0 MSML_HEAD_DCOS_ROWMAJOR=1
ARGS="--classes 125000 --sample-rate 0.1 --fused-sgd"
run 125k_sr01 X=1
ARGS="--classes 1000000 --sample-rate 0.1"
run 1m_sr01 X=1
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x > $O/r02c_pytest_engine.log 2>&1
echo "engine tests rc=$? : $(tail -1 $O/r02c_pytest_engine.log)"
