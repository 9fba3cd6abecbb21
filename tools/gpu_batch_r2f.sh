#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2f
mkdir -p $O
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
$B --skip-steps 500 > $O/bench_late.json 2> $O/bench_late.err
python bench.py --config 4 --steps 20 --warmup 5 --quick > $O/config4.json 2> $O/config4.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_sca.csv $S > $O/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_late.csv $S --skip-steps 300 > $O/ncu_l2.log 2>&1
ls $O
