#!/bin/bash
# Round-2 GPU call 9 (1 GPU): whole GPU suite + headline bench at N=1 with the round-2 head.
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > $O/r02i_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02i_pytest_gpu.log)"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r02i_bench_n1.json 2> $O/r02i_bench_n1.err
echo "bench rc=$? : $(head -c 400 $O/r02i_bench_n1.json)"
timeout 300 python bench.py --steps 30 --warmup 5 --stock-head-sgd --no-cpu-baseline > $O/r02i_bench_n1_stocksgd.json 2> $O/r02i_bench_n1_stocksgd.err
echo "bench stock sgd rc=$? : $(head -c 300 $O/r02i_bench_n1_stocksgd.json)"
