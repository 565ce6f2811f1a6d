#!/bin/bash
# Round-2 GPU call 21 (8 GPUs): final build — training step at N = 8, BASELINE config 4 (raw mode), per-collective device times.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29741 bench.py --gpus 8 --steps 30 --warmup 5 > $O/r02u_bench_n8.json 2> $O/r02u_bench_n8.err
echo "train n8 rc=$? : $(head -c 300 $O/r02u_bench_n8.json)"
for sr in 1.0 0.1; do
  timeout 300 $TR --master-port 29742 bench.py --gpus 8 --workload head --classes 1000000 --sample-rate $sr --batch 128 --steps 50 --warmup 10 --fused-sgd --no-head-check \
    > $O/r02u_head_1m_sr${sr}_n8.json 2> $O/r02u_head_1m_sr${sr}_n8.err
  echo "config4 sr=$sr rc=$? : $(head -c 260 $O/r02u_head_1m_sr${sr}_n8.json)"
done
