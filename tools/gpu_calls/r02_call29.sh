#!/bin/bash
# Round-2 GPU call 29 (1 GPU): DRAM traffic of EVERY BN launch of one training step (the dominant kernel families), so that
# roofline.traffic is averaged over the same launches as roofline.achieved.  Counters only (no --set full: 700 launches).
set -u
O=gpurun_out
TRAIN="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
export MSML_PROFILER_RANGE=1
$TRAIN > $O/r02ad_plain_train.json 2> $O/r02ad_plain_train.err &&
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"bn_|sgd_flat|pfc_sgd|accum_bf16|fm_gate|fm_cat" -c 900 --csv --log-file $O/r02ad_bn_traffic.csv $TRAIN > $O/r02ad_ncu_bn_traffic.log 2>&1
echo "bn traffic rc=$? : $(wc -l < $O/r02ad_bn_traffic.csv) lines"
