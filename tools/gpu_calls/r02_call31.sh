#!/bin/bash
# Round-2 GPU call 31 (1 GPU): L2 residency hints for the two-pass BN kernels (statistics pass evict_last, apply pass evict_first):
# BN parity with the hints on, then the headline step without / with (64 MB and 110 MB slab limits), two alternating repetitions.
set -u
O=gpurun_out
MSML_BN_L2_KEEP=1 timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x -k "bn" > $O/r02af_pytest_l2.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02af_pytest_l2.log)"
for rep in 1 2; do
  MSML_BN_L2_KEEP=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02af_bench_keep0_$rep.json 2> $O/r02af_bench_keep0_$rep.err
  echo "keep=0 $rep rc=$? : $(head -c 180 $O/r02af_bench_keep0_$rep.json | tail -c 60)"
  MSML_BN_L2_KEEP=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02af_bench_keep64_$rep.json 2> $O/r02af_bench_keep64_$rep.err
  echo "keep=1/64MB $rep rc=$? : $(head -c 180 $O/r02af_bench_keep64_$rep.json | tail -c 60)"
  MSML_BN_L2_KEEP=1 MSML_BN_L2_KEEP_MB=110 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02af_bench_keep110_$rep.json 2> $O/r02af_bench_keep110_$rep.err
  echo "keep=1/110MB $rep rc=$? : $(head -c 180 $O/r02af_bench_keep110_$rep.json | tail -c 60)"
done
