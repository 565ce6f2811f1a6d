#!/bin/bash
# Round-2 GPU call 4 (2 GPUs): the class-sharded head over REAL NCCL ranks (both transports), captured 2-rank train step, N=2 bench.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02d_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_nccl.py -q -x --durations=20 > $O/r02d_pytest_nccl.log 2>&1
echo "nccl tests rc=$? : $(tail -1 $O/r02d_pytest_nccl.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 \
  > $O/r02d_bench_n2.json 2> $O/r02d_bench_n2.err
echo "bench n2 rc=$? : $(head -c 700 $O/r02d_bench_n2.json)"
