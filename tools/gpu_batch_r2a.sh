#!/bin/bash
# Round-2 GPU batch A (run under gpurun on one B200): the -m gpu suite, A/B timings of kernel variants, ncu captures.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
( timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
for v in ctas5 ctas6 knn8 knn40; do
  SCAB_LIB_PATH=$PWD/_ab/lib_$v.so $B > $O/bench_$v.json 2> $O/bench_$v.err
done
( cd _r1 && python bench.py --no-cpu --steps 20 --warmup 5 > ../$O/bench_r1.json 2> ../$O/bench_r1.err )
L="python bench.py --quick --policy orca_lp --steps 20 --warmup 5"
$L > $O/lp_default.json 2> $O/lp_default.err
SCAB_LP_COMPACT=1 $L > $O/lp_compact.json 2> $O/lp_compact.err
SCAB_LP_PASS1_CTAS8=1 $L > $O/lp_ctas8.json 2> $O/lp_ctas8.err
SCAB_LP_COMPACT=1 SCAB_LP_PASS1_CTAS8=1 $L > $O/lp_both.json 2> $O/lp_both.err
# ncu: launch lists (shares), then full captures of the hot kernels
S="python bench.py --quick --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_sca.csv $S > $O/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_lp.csv $S --policy orca_lp > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_select_fast|k_neighbors_tile|k_status|k_integrate" -s 13 -c 4 -o $O/prof_sca $S > $O/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_policy_lp" -s 6 -c 2 -o $O/prof_lp $S --policy orca_lp > $O/ncu_p2.log 2>&1
ls -la $O
