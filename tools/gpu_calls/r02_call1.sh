#!/bin/bash
# Round-2 GPU call 1: first real run of the kernels written blind in round 1 (output kept), the sampled-path diagnosis.
set -u
O=gpurun_out
mkdir -p $O
for f in check_pfc_sgd check_consensus check_dataloaderx check_head_edge_cases; do
  timeout 300 python -m pytest -x -q -p no:cacheprovider tests/unverified/$f.py > $O/r02_$f.log 2>&1
  echo "$f rc=$? : $(tail -1 $O/r02_$f.log)"
done
timeout 200 python tools/diag_sampled.py > $O/r02_diag_sampled.json 2> $O/r02_diag_sampled.err
echo "diag rc=$?"
timeout 200 python tools/diag_sampled.py --fused > $O/r02_diag_sampled_fused.json 2> $O/r02_diag_sampled_fused.err
echo "diag fused rc=$?"
timeout 200 python tools/diag_sampled.py --classes 125000 --steps 6 > $O/r02_diag_sampled_125k.json 2> $O/r02_diag_sampled_125k.err
echo "diag 125k rc=$?"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r02_smi.txt
