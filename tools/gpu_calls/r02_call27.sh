#!/bin/bash
# Round-2 GPU call 27 (1 GPU): final-build verification - smoke, whole GPU suite, headline bench, then the ncu launch list of the
# timed region and a --set full capture of the kernels added late in the round (sgd_flat, BN apply with chained statistics).
set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02ab_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $O/r02ab_smoke.log)"
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > $O/r02ab_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02ab_pytest_gpu.log)"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r02ab_bench_n1.json 2> $O/r02ab_bench_n1.err
echo "bench rc=$? : $(head -c 400 $O/r02ab_bench_n1.json)"
TRAIN="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
export MSML_PROFILER_RANGE=1
$TRAIN > $O/r02ab_plain_train.json 2> $O/r02ab_plain_train.err &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/r02ab_train_launches.csv $TRAIN > $O/r02ab_ncu_train_list.log 2>&1
echo "launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"sgd_flat|bn_fwd_fused_kernel.*Lb1ELb0ELi3ELb1|bn_fwd_fused_kernel.*Li2E" -c 12 -f -o $O/r02ab_new_kernels $TRAIN > $O/r02ab_ncu_new_kernels.log 2>&1
echo "new kernels rc=$?"
ncu -i $O/r02ab_new_kernels.ncu-rep --page raw --csv > $O/r02ab_new_kernels_raw.csv 2>/dev/null
rm -f $O/r02ab_new_kernels.ncu-rep
du -sh $O
