#!/bin/bash
# Round-2 GPU call 32 (1 GPU): the committed final state (BN L2 hints on by default): whole GPU suite + headline bench.
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/r02ag_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02ag_pytest_gpu.log)"
timeout 300 python bench.py --steps 30 --warmup 5 > $O/r02ag_bench_n1.json 2> $O/r02ag_bench_n1.err
echo "bench rc=$? : $(head -c 220 $O/r02ag_bench_n1.json)"
