#!/bin/bash
# 2+ GPUs: sharded-run parity (peer-to-peer and NCCL transports) + scaling bench with both transports
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}; T=${2:-r2p}
O=gpurun_out/$T; mkdir -p $O
chk() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/sharding_check.py; }
( SHARD_TRANSPORT=p2p SHARD_POLICY=sca chk; SHARD_TRANSPORT=p2p SHARD_POLICY=orca_lp chk;
  SHARD_TRANSPORT=p2p SHARD_POLICY=sca SHARD_DUBINS=1 SHARD_N=8000 SHARD_MIG_CAP=1024 chk;
  SHARD_TRANSPORT=p2p SHARD_POLICY=rvo3d SHARD_WAYPOINTS=1 SHARD_N=8000 SHARD_MIG_CAP=1024 chk;
  SHARD_TRANSPORT=nccl SHARD_POLICY=sca chk ) 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" > $O/check_n$N.log
grep "sharding_check\|Error\|error" $O/check_n$N.log | cut -c1-300
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 40 --warmup 5 --quick > $O/scale_p2p_n$N.json 2> $O/scale_p2p_n$N.err
SCAB_TRANSPORT=nccl run --steps 40 --warmup 5 --quick > $O/scale_nccl_n$N.json 2> $O/scale_nccl_n$N.err
SCAB_BENCH_PHASES=1 run --steps 20 --warmup 5 --quick > $O/phases_n$N.json 2> $O/phases_n$N.err
grep PHASES $O/phases_n$N.err | head -2
python tools/bench_table.py $O/scale_p2p_n$N.json $O/scale_nccl_n$N.json
tail -3 $O/scale_p2p_n$N.err
