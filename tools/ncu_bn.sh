set -u
O=gpurun_out
TRAIN="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
export MSML_PROFILER_RANGE=1
for k in bn_fwd_fused bn_bwd_fused; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 18 -o $O/r01d_train_$k -f $TRAIN > $O/ncu_train_$k.log 2>&1
  echo "train $k rc=$?"
  ncu -i $O/r01d_train_$k.ncu-rep --page raw --csv > $O/r01d_train_${k}_raw.csv 2>/dev/null
  rm -f $O/r01d_train_$k.ncu-rep
done
rm -f $O/*.ncu-rep
du -sh $O
