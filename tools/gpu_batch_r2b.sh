#!/bin/bash
# Round-2 GPU batch B: second round of kernel variants (select v2, KNN tile v2 / direct), tier_delta sweep, e2e through the API
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2b
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
for v in direct tile6 tile20 sel5 sel6 direct_sel5; do
  SCAB_LIB_PATH=$PWD/_ab/lib_$v.so $B > $O/bench_$v.json 2> $O/bench_$v.err
done
for td in 0.2 0.25 0.35 0.45; do SCAB_TIER_DELTA=$td $B > $O/bench_td$td.json 2> $O/bench_td$td.err; done
python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err
SCAB_BENCH_E2E_OVERLAP=0 python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_full_serial.json 2> $O/bench_full_serial.err
python bench.py --quick --policy orca_lp --steps 20 --warmup 5 > $O/lp_default.json 2> $O/lp_default.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:"k_select_fast|k_neighbors_tile|k_neighbors_direct" -s 7 -c 2 -o $O/prof_sca $S > $O/ncu_p1.log 2>&1
SCAB_LIB_PATH=$PWD/_ab/lib_direct.so ncu --set full --clock-control none --import-source on -k regex:"k_neighbors_direct" -s 4 -c 1 -o $O/prof_direct $S > $O/ncu_p2.log 2>&1
ls -la $O
