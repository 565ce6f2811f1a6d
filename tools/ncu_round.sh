#!/bin/bash
# ncu passes of commands that already exited 0 on a B200 in an earlier call.  Outputs under gpurun_out/ (kept < 64 MiB:
# reports are exported to CSV on the box; only the two small microbench reports are kept as .ncu-rep).
set -u
O=gpurun_out
TRAIN="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
FUSION="python bench.py --workload fusion --steps 5"
HEAD="python bench.py --workload head --classes 125000 --batch 1024 --steps 3 --warmup 3"
export MSML_PROFILER_RANGE=1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r01c_train_launches.csv $TRAIN > $O/ncu_train_list.log 2>&1
echo "launch list rc=$?"
for k in bn_fwd_fused bn_bwd_fused fm_gate; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 18 -o $O/r01c_train_$k -f $TRAIN > $O/ncu_train_$k.log 2>&1
  echo "train $k rc=$?"
  ncu -i $O/r01c_train_$k.ncu-rep --page raw --csv > $O/r01c_train_${k}_raw.csv 2>/dev/null
  rm -f $O/r01c_train_$k.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:"fm_gate" -s 8 -c 4 -o $O/r01c_fusion_cfg2 -f $FUSION > $O/ncu_fusion_full.log 2>&1
echo "fusion full rc=$?"
ncu -i $O/r01c_fusion_cfg2.ncu-rep --page raw --csv > $O/r01c_fusion_cfg2_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 12 -c 4 -o $O/r01c_head_cfg4 -f $HEAD > $O/ncu_head_full.log 2>&1
echo "head full rc=$?"
ncu -i $O/r01c_head_cfg4.ncu-rep --page raw --csv > $O/r01c_head_cfg4_raw.csv 2>/dev/null
rm -f $O/r01b_*.ncu-rep
du -sh $O
