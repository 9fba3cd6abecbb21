#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2g
mkdir -p $O
echo "CUDA_DEVICE_MAX_CONNECTIONS=$CUDA_DEVICE_MAX_CONNECTIONS" > $O/env.txt
env | grep -i "cuda\|nccl" >> $O/env.txt
B="python bench.py --quick --steps 20 --warmup 5"
SCAB_BENCH_SIDE_STREAM=1 $B > $O/side_fork.json 2> $O/side_fork.err
SCAB_BENCH_SIDE_STREAM=1 SCAB_NO_FORK=1 $B > $O/side_nofork.json 2> $O/side_nofork.err
CUDA_DEVICE_MAX_CONNECTIONS=8 $B > $O/conn8_fork.json 2> $O/conn8_fork.err
CUDA_DEVICE_MAX_CONNECTIONS=8 SCAB_BENCH_SIDE_STREAM=1 $B > $O/conn8_side_fork.json 2> $O/conn8_side_fork.err
ls $O
