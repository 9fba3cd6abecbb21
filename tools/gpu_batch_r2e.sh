#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2e
mkdir -p $O
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_fork.json 2> $O/bench_fork.err
SCAB_NO_FORK=1 $B > $O/bench_nofork.json 2> $O/bench_nofork.err
SCAB_NO_FORK=2 $B > $O/bench_fork2.json 2> $O/bench_fork2.err
$B --skip-steps 500 > $O/bench_late.json 2> $O/bench_late.err
python bench.py --config 4 --steps 20 --warmup 5 --quick > $O/config4.json 2> $O/config4.err
( timeout 600 python -m pytest tests/test_gpu_knn_oracle.py tests/test_gpu_full_size.py -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
ls $O
