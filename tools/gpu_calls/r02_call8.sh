#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_head_edges.py -q -x -k "not gemm" > $O/r02h_pytest.log 2>&1
echo "tests rc=$? : $(tail -1 $O/r02h_pytest.log)"
timeout 300 python tools/gemm_major_bench.py > $O/r02h_gemm_major.json 2> $O/r02h_gemm_major.err
echo "major bench rc=$?"; cat $O/r02h_gemm_major.json | tr -d '\n ' | head -c 3000; echo
ARGS="--classes 125000 --sample-rate 1.0 --fused-sgd"
timeout 300 python bench.py --workload head --batch 1024 --steps 30 --warmup 5 --no-head-check $ARGS > $O/r02h_head_125k.json 2> $O/r02h_head_125k.err
python - <<PY
import json
d=json.load(open("$O/r02h_head_125k.json"))
print(d["ms_per_step_median"], d["head_algorithmic_tflops_over_gemm_time"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"]) for r in d["rooflines"]])
PY
