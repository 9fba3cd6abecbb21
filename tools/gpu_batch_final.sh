#!/bin/bash
# Round-2 final single-GPU batch: the -m gpu suite, the bench lines of every BASELINE config, launch lists and ncu captures
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2final}
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --policy orca_lp > $O/bench_n1_orca_lp.json 2> $O/bench_n1_orca_lp.err
python bench.py --quick --skip-steps 500 > $O/bench_n1_late.json 2> $O/bench_n1_late.err
for c in 1 2 3 4; do python bench.py --config $c > $O/bench_config$c.json 2> $O/bench_config$c.err; done
python bench.py --config 4 --policy srvo3d --no-cpu > $O/bench_config4_srvo3d.json 2> $O/bench_config4_srvo3d.err
python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_sca.csv $S > $O/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_lp.csv $S --policy orca_lp > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_select_fast|k_neighbors_direct|k_status|k_integrate" -s 20 -c 6 -o $O/prof_sca $S > $O/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_policy_lp" -s 6 -c 2 -o $O/prof_lp $S --policy orca_lp > $O/ncu_p2.log 2>&1
ls -la $O
