#!/bin/bash
# Round-2 GPU call 10 (8 GPUs): NCCL parity at W = 4 / 8, BASELINE config 4 as written (1M classes, sample_rate 0.1 / 1.0,
# class-sharded over 8 B200), training step at N = 8 (default and with NCCL's CTA count capped).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02j_gpus.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py -q -x -k "(8-1.0 or 8-0.1 or 4-1.0 or captured) and not torchdist" --durations=10 > $O/r02j_pytest_nccl8.log 2>&1
echo "nccl tests rc=$? : $(tail -1 $O/r02j_pytest_nccl8.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for sr in 1.0 0.1; do
  timeout 300 $TR --master-port 29721 bench.py --gpus 8 --workload head --classes 1000000 --sample-rate $sr --batch 128 --steps 50 --warmup 10 --fused-sgd --no-head-check \
    > $O/r02j_head_1m_sr${sr}_n8.json 2> $O/r02j_head_1m_sr${sr}_n8.err
  echo "config4 sr=$sr rc=$? : $(head -c 600 $O/r02j_head_1m_sr${sr}_n8.json)"
done
timeout 300 $TR --master-port 29722 bench.py --gpus 8 --workload head --classes 1000000 --sample-rate 0.1 --batch 128 --steps 50 --warmup 10 --no-head-check \
    > $O/r02j_head_1m_sr0.1_n8_stocksgd.json 2> $O/r02j_head_1m_sr0.1_n8_stocksgd.err
echo "config4 sr=0.1 stock sgd rc=$? : $(head -c 300 $O/r02j_head_1m_sr0.1_n8_stocksgd.json)"
timeout 400 $TR --master-port 29723 bench.py --gpus 8 --steps 30 --warmup 5 > $O/r02j_bench_n8.json 2> $O/r02j_bench_n8.err
echo "train n8 rc=$? : $(head -c 300 $O/r02j_bench_n8.json)"
NCCL_MAX_CTAS=8 timeout 400 $TR --master-port 29724 bench.py --gpus 8 --steps 30 --warmup 5 > $O/r02j_bench_n8_maxctas8.json 2> $O/r02j_bench_n8_maxctas8.err
echo "train n8 maxctas8 rc=$? : $(head -c 300 $O/r02j_bench_n8_maxctas8.json)"
