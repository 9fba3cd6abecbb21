#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2h
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
$B --skip-steps 500 > $O/bench_late.json 2> $O/bench_late.err
python bench.py --config 4 --steps 20 --warmup 5 --quick > $O/config4.json 2> $O/config4.err
python bench.py --config 3 --steps 20 --warmup 5 --quick > $O/config3.json 2> $O/config3.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_sca.csv $S > $O/ncu_l1.log 2>&1
ls $O
