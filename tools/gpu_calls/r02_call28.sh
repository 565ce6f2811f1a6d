#!/bin/bash
# Round-2 GPU call 28 (1 GPU): the first BN launch of every op as a programmatic dependent of its predecessor (MSML_BN_PDL_FIRST=1):
# BN parity under it, then the headline step without / with, two alternating repetitions.
set -u
O=gpurun_out
MSML_BN_PDL_FIRST=1 timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_engine.py -m gpu -q -x -k "bn or graph or flat" > $O/r02ac_pytest_pdl.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02ac_pytest_pdl.log)"
for rep in 1 2; do
  MSML_BN_PDL_FIRST=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02ac_bench_pdl0_$rep.json 2> $O/r02ac_bench_pdl0_$rep.err
  echo "pdl_first=0 $rep rc=$? : $(head -c 200 $O/r02ac_bench_pdl0_$rep.json)"
  MSML_BN_PDL_FIRST=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02ac_bench_pdl1_$rep.json 2> $O/r02ac_bench_pdl1_$rep.err
  echo "pdl_first=1 $rep rc=$? : $(head -c 200 $O/r02ac_bench_pdl1_$rep.json)"
done
