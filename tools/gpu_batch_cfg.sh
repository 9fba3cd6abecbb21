#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2cfg}
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_step_from_host.py -m gpu -q ) > $O/pytest_gpu.log 2>&1
for c in 1 2 3 4; do python bench.py --config $c > $O/bench_config$c.json 2> $O/bench_config$c.err; done
python bench.py --config 4 --policy srvo3d --no-cpu > $O/bench_config4_srvo3d.json 2> $O/bench_config4_srvo3d.err
ls -la $O
