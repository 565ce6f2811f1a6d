#!/bin/bash
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02o_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02o_pytest_gpu.log)"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r02o_bench_n1.json 2> $O/r02o_bench_n1.err
echo "bench rc=$? : $(head -c 300 $O/r02o_bench_n1.json)"
python - <<PY
import json
d=json.load(open("$O/r02o_bench_n1.json"))
for k in ("head_microbench","head_microbench_fused_projection"):
    hm=d[k]; print(k, hm["gemm_us"], hm["tflops_over_gemm_time"], hm["frac_of_bf16_burst_peak"], [(r["kernel"].replace("head_","").replace("_gemm",""), r["avg_us"], r["frac"]) for r in hm["rooflines"]])
PY
