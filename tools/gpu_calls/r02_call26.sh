#!/bin/bash
# Round-2 GPU call 26 (2 GPUs): FlatSGD under NCCL - the two-rank captured step (both backbone optimizers) and the N=2 headline line.
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -k "captured" > $O/r02aa_pytest_nccl_captured.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02aa_pytest_nccl_captured.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > $O/r02aa_bench_n2.json 2> $O/r02aa_bench_n2.err
echo "bench n2 rc=$? : $(head -c 300 $O/r02aa_bench_n2.json)"
