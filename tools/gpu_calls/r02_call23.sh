#!/bin/bash
# Round-2 GPU call 23 (1 GPU): chained BN statistics (msml_bn_fwd_ex) - parity, then the headline step with / without.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_model.py tests/test_gpu_engine.py -m gpu -q -x --durations=5 > $O/r02w_pytest_bn.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02w_pytest_bn.log)"
for rep in 1 2; do
  MSML_BN_CHAIN=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02w_bench_nochain_$rep.json 2> $O/r02w_bench_nochain_$rep.err
  echo "nochain $rep rc=$? : $(head -c 200 $O/r02w_bench_nochain_$rep.json)"
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02w_bench_chain_$rep.json 2> $O/r02w_bench_chain_$rep.err
  echo "chain $rep rc=$? : $(head -c 200 $O/r02w_bench_chain_$rep.json)"
done
