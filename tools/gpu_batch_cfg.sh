#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2cfg}
mkdir -p $O
for c in 1 2; do python bench.py --config $c > $O/bench_config$c.json 2> $O/bench_config$c.err; done
ls -la $O
