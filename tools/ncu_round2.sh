#!/bin/bash
# Round-2 ncu passes (1 GPU).  Every profiled command first runs plain and must exit 0 (`&&`).  Outputs: gpurun_out/r02p_*.
#   1. launch list of the timed region of the training step (gpu__time_duration per kernel: compare SHARES with bench.py's rooflines)
#   2. --set full of this library's in-step kernels (BN, fusion, concat, optimizer, accumulate)
#   2b. DRAM byte counters of every launch of this library in one step, caches not flushed (roofline.traffic)
#   3. --set full of the head GEMMs at BASELINE config-4 per-rank shapes, exact and raw (fused-projection) mode
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02p_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $O/r02p_smoke.log)"
TRAIN="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
HEAD="python bench.py --workload head --classes 125000 --batch 1024 --sample-rate 1.0 --fused-sgd --steps 2 --warmup 5 --no-head-check"
export MSML_PROFILER_RANGE=1
$TRAIN > $O/r02p_plain_train.json 2> $O/r02p_plain_train.err &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/r02p_train_launches.csv $TRAIN > $O/r02p_ncu_train_list.log 2>&1
echo "launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"bn_|fm_gate|fm_cat|pfc_sgd|sgd_flat|accum_bf16|dap_" -c 60 -f -o $O/r02p_train_own $TRAIN > $O/r02p_ncu_train_own.log 2>&1
echo "train own kernels rc=$?"
ncu -i $O/r02p_train_own.ncu-rep --page raw --csv > $O/r02p_train_own_raw.csv 2>/dev/null
rm -f $O/r02p_train_own.ncu-rep
# 2b. DRAM counters of EVERY launch of this library in one step, L2 left as the step leaves it (--cache-control none, one pass per
#     kernel): the per-family totals become roofline.traffic (profiles/r02_traffic.json); ncu's default flushes the caches before
#     each profiled kernel, which makes every second read of a BN slab miss
ncu --profile-from-start off --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"bn_|sgd_flat|pfc_sgd|accum_bf16|fm_gate|fm_cat" -c 900 --csv --log-file $O/r02p_step_dram_traffic_warm.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02p_ncu_traffic_warm.log 2>&1
echo "step traffic rc=$?"
unset MSML_PROFILER_RANGE
$HEAD > $O/r02p_plain_head_raw.json 2> $O/r02p_plain_head_raw.err &&
ncu --set full --clock-control none --import-source on -k regex:"gemm|pfc_sgd" -s 25 -c 5 -f -o $O/r02p_head_raw $HEAD > $O/r02p_ncu_head_raw.log 2>&1
echo "head raw rc=$?"
ncu -i $O/r02p_head_raw.ncu-rep --page raw --csv > $O/r02p_head_raw_raw.csv 2>/dev/null
$HEAD --exact-head-grad > $O/r02p_plain_head_exact.json 2> $O/r02p_plain_head_exact.err &&
ncu --set full --clock-control none --import-source on -k regex:"gemm|pfc_sgd" -s 25 -c 5 -f -o $O/r02p_head_exact $HEAD --exact-head-grad > $O/r02p_ncu_head_exact.log 2>&1
echo "head exact rc=$?"
ncu -i $O/r02p_head_exact.ncu-rep --page raw --csv > $O/r02p_head_exact_raw.csv 2>/dev/null
rm -f $O/r02p_head_exact.ncu-rep
du -sh $O
