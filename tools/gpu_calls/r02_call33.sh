#!/bin/bash
# Round-2 GPU call 33 (1 GPU, last of the round): BN slab loads on ONE code path with a block-uniform L2 policy (the two-path
# version of call 31 spilled registers): BN parity, then the headline step with the hints off / on.
set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x -k "bn" > $O/r02ah_pytest_bn.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02ah_pytest_bn.log)"
MSML_BN_L2_KEEP=0 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02ah_bench_keep0.json 2> $O/r02ah_bench_keep0.err
echo "keep=0 rc=$? : $(head -c 180 $O/r02ah_bench_keep0.json | tail -c 60)"
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02ah_bench_keep1.json 2> $O/r02ah_bench_keep1.err
echo "default rc=$? : $(head -c 180 $O/r02ah_bench_keep1.json | tail -c 60)"
