#!/bin/bash
# Round-2 GPU call 30 (1 GPU): call 29 again with --cache-control none: ncu's default flushes every cache before each profiled
# kernel, so the BN apply pass's second read of its slab always missed under the profiler.  One pass per kernel (DRAM counters only).
set -u
O=gpurun_out
TRAIN="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
export MSML_PROFILER_RANGE=1
$TRAIN > $O/r02ae_plain_train.json 2> $O/r02ae_plain_train.err &&
ncu --profile-from-start off --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"bn_|sgd_flat|pfc_sgd|accum_bf16|fm_gate|fm_cat" -c 900 --csv --log-file $O/r02ae_traffic_warm.csv $TRAIN > $O/r02ae_ncu_traffic_warm.log 2>&1
echo "traffic rc=$? : $(wc -l < $O/r02ae_traffic_warm.csv) lines; $(grep -c 'pass' $O/r02ae_ncu_traffic_warm.log) pass lines"
grep -m3 "pass" $O/r02ae_ncu_traffic_warm.log
