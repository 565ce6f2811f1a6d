#!/bin/bash
# Round-2 GPU call 24 (1 GPU): engine.FlatSGD (msml_sgd_flat) - parity, then the headline step with torch's fused SGD / FlatSGD.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x --durations=5 > $O/r02y_pytest_engine.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02y_pytest_engine.log)"
for rep in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --stock-backbone-sgd > $O/r02y_bench_stock_$rep.json 2> $O/r02y_bench_stock_$rep.err
  echo "stock $rep rc=$? : $(head -c 200 $O/r02y_bench_stock_$rep.json)"
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02y_bench_flat_$rep.json 2> $O/r02y_bench_flat_$rep.err
  echo "flat $rep rc=$? : $(head -c 200 $O/r02y_bench_flat_$rep.json)"
done
