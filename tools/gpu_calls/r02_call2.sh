#!/bin/bash
# Round-2 GPU call 2 (1 GPU): the whole GPU suite with the new tests, head workload (sampled / full, stock / fused SGD).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > $O/r02_pytest_gpu_a.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02_pytest_gpu_a.log)"
for sr in 0.1 1.0; do
  for f in "" "--fused-sgd"; do
    tag="sr${sr}${f:+_fused}"
    timeout 300 python bench.py --workload head --classes 125000 --batch 1024 --sample-rate $sr --steps 20 --warmup 5 $f \
      > $O/r02_head_125k_$tag.json 2> $O/r02_head_125k_$tag.err
    echo "head $tag rc=$? : $(head -c 400 $O/r02_head_125k_$tag.json)"
  done
done
timeout 300 python bench.py --workload head --classes 1000000 --batch 1024 --sample-rate 0.1 --steps 20 --warmup 5 \
  > $O/r02_head_1m_sr0.1.json 2> $O/r02_head_1m_sr0.1.err
echo "head 1M sr0.1 rc=$? : $(head -c 300 $O/r02_head_1m_sr0.1.json)"
