#!/bin/bash
# Round-2 GPU call 17 (2 GPUs): NCCL tests with the final build (raw mode over NCCL in the captured step), N=2 benches.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py -q -x --durations=5 > $O/r02q_pytest_nccl2.log 2>&1
echo "nccl tests rc=$? : $(tail -1 $O/r02q_pytest_nccl2.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29731 bench.py --gpus 2 --steps 30 --warmup 5 > $O/r02q_bench_n2.json 2> $O/r02q_bench_n2.err
echo "train n2 rc=$? : $(head -c 300 $O/r02q_bench_n2.json)"
timeout 300 $TR --master-port 29732 bench.py --gpus 2 --workload head --classes 1000000 --sample-rate 0.1 --batch 128 --steps 50 --warmup 10 --fused-sgd --no-head-check > $O/r02q_head_1m_sr0.1_n2.json 2> $O/r02q_head_1m_sr0.1_n2.err
echo "head n2 rc=$? : $(head -c 300 $O/r02q_head_1m_sr0.1_n2.json)"
