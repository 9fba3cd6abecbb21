#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2e2e}
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_step_from_host.py tests/test_gpu_parity.py tests/test_gpu_knn_oracle.py -m gpu -q ) > $O/pytest_gpu.log 2>&1
python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_e2e_chunks4.json 2> $O/bench_e2e_chunks4.err
SCAB_E2E_CHUNKS=1 python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_e2e_chunks1.json 2> $O/bench_e2e_chunks1.err
SCAB_E2E_CHUNKS=8 python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_e2e_chunks8.json 2> $O/bench_e2e_chunks8.err
SCAB_E2E_CHUNKS=2 python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_e2e_chunks2.json 2> $O/bench_e2e_chunks2.err
python bench.py --no-cpu --policy orca_lp --steps 20 --warmup 5 > $O/bench_lp_e2e.json 2> $O/bench_lp_e2e.err
ls $O
