#!/bin/bash
# Round-2 GPU call 5 (1 GPU): ncu --set full of the head GEMMs at config-4 per-rank shapes, CTA pairs vs single CTA.
set -u
O=gpurun_out
mkdir -p $O
CMD="python bench.py --workload head --classes 125000 --batch 1024 --sample-rate 1.0 --fused-sgd --steps 2 --warmup 5 --no-head-check"
MSML_HEAD_PAIR=1 $CMD > $O/r02e_plain_pair.json 2> $O/r02e_plain_pair.err &&
MSML_HEAD_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm -s 20 -c 4 -f -o $O/r02e_head_pair $CMD > $O/r02e_ncu_pair.log 2>&1
echo "ncu pair rc=$?"
MSML_HEAD_PAIR=0 $CMD > $O/r02e_plain_nopair.json 2> $O/r02e_plain_nopair.err &&
MSML_HEAD_PAIR=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm -s 20 -c 4 -f -o $O/r02e_head_nopair $CMD > $O/r02e_ncu_nopair.log 2>&1
echo "ncu nopair rc=$?"
ls -la $O/*.ncu-rep
