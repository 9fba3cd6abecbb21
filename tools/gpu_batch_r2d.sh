#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2d
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err
python bench.py --quick --steps 20 --warmup 5 --skip-steps 500 > $O/bench_late.json 2> $O/bench_late.err
python bench.py --quick --policy orca_lp --steps 20 --warmup 5 > $O/lp.json 2> $O/lp.err
for c in 1 2 3 4; do python bench.py --config $c --steps 20 --warmup 5 > $O/config$c.json 2> $O/config$c.err; done
python bench.py --config 4 --policy srvo3d --steps 20 --warmup 5 --no-cpu > $O/config4_srvo3d.json 2> $O/config4_srvo3d.err
ls -la $O
