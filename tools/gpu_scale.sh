#!/bin/bash
# usage: gpu_scale.sh N tag  -- sharded bench at N GPUs (+ per-phase eager profile) under torchrun
cd "${GRAFT_REPO_ROOT:-.}"
N=$1; T=${2:-r2}
O=gpurun_out/$T
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 40 --warmup 5 --quick > $O/scale_n$N.json 2> $O/scale_n$N.err
SCAB_BENCH_PHASES=1 run --steps 20 --warmup 5 --quick > $O/phases_n$N.json 2> $O/phases_n$N.err
grep PHASES $O/phases_n$N.err | head -8
tail -1 $O/scale_n$N.json | cut -c1-200
