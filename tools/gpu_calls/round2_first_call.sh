#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget ran out gets its first real run, and
# the round-1 numbers are re-measured on the same box.  Usage (from the repo root, ~6 GPU-minutes):
#     gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Outputs under gpurun_out/r02_*.  A step that fails does not stop the later ones.
set -u
O=gpurun_out
mkdir -p $O
# 1. the pending checks, verbosely and OUTSIDE the xfail wrapper (each file in its own process)
for f in check_consensus check_pfc_sgd check_dataloaderx check_head_edge_cases; do
  timeout 300 python -m pytest -x -q -p no:cacheprovider tests/unverified/$f.py > $O/r02_$f.log 2>&1
  echo "$f rc=$? : $(tail -1 $O/r02_$f.log)"
done
# 2. the regular GPU suite
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? : $(tail -1 $O/r02_pytest_gpu.log)"
# 3. headline bench + the microbenches of the kernels around the hot ops (ATen twins timed beside them)
timeout 400 python bench.py --steps 30 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
echo "bench rc=$? : $(head -c 300 $O/r02_bench_n1.json)"
timeout 200 python bench.py --workload aux --steps 10 > $O/r02_bench_aux.json 2> $O/r02_bench_aux.err
echo "aux rc=$? : $(head -c 600 $O/r02_bench_aux.json)"
# 4. ncu: launch list of the step, full-set captures of the new kernels (each command exited 0 without ncu above)
export MSML_PROFILER_RANGE=1
TRAIN="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv \
  --log-file $O/r02_train_launches.csv $TRAIN > $O/r02_ncu_train_list.log 2>&1
echo "launch list rc=$?"
unset MSML_PROFILER_RANGE
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fm_cat|seg_|pfc_sgd" -c 24 -o $O/r02_aux -f \
  python bench.py --workload aux --steps 5 > $O/r02_ncu_aux.log 2>&1
echo "aux full rc=$?"
ncu -i $O/r02_aux.ncu-rep --page raw --csv > $O/r02_aux_raw.csv 2>/dev/null
rm -f $O/r02_aux.ncu-rep
du -sh $O
