#!/bin/bash
# A/B: library variants (_ab/*.so) and run-time switches (NAME=VALUE) in one call: gpu_batch_ab2.sh tag item...
# (AB_LATE=1: also the late-rollout sample, steps 505-525; AB_CFG4=1: also BASELINE config 4)
cd "${GRAFT_REPO_ROOT:-.}"
T=$1; shift
O=gpurun_out/$T
mkdir -p $O
B="python bench.py --quick --steps 30 --warmup 5"
one() {  # one <prefix> <bench args...>
  local pre=$1; shift
  "$@" > $O/${pre}base.json 2> $O/${pre}base.err
  for it in $ITEMS; do
    case "$it" in
      *.so) n=$(basename $it .so); SCAB_LIB_PATH=$PWD/$it "$@" > $O/${pre}${n}.json 2> $O/${pre}${n}.err ;;
      *=*)  n=$(echo $it | tr '=' '_'); env $it "$@" > $O/${pre}${n}.json 2> $O/${pre}${n}.err ;;
    esac
  done
}
ITEMS="$*"
one r1_ $B
one r2_ $B
[ -n "$AB_LATE" ] && one late_ $B --skip-steps 500
[ -n "$AB_CFG4" ] && one cfg4_ python bench.py --config 4 --quick --steps 30 --warmup 5
python tools/bench_table.py $O/*.json
