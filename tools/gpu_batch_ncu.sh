#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2k}
mkdir -p $O
python bench.py --policy orca_lp --no-cpu > $O/bench_n1_orca_lp.json 2> $O/bench_n1_orca_lp.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_lp.csv $S --policy orca_lp > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_select_fast|k_neighbors_direct|k_status|k_integrate" -s 20 -c 6 -o $O/prof_sca $S > $O/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_policy_lp" -s 6 -c 2 -o $O/prof_lp $S --policy orca_lp > $O/ncu_p2.log 2>&1
ls -la $O
