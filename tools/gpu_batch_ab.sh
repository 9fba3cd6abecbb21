#!/bin/bash
# A/B of library variants built into _ab/ (python -m sca_b200.build -D... --out=_ab/lib_X.so): usage gpu_batch_ab.sh tag lib1 lib2 ...
# (AB_LATE=1: also the late-rollout sample, steps 505-525)
cd "${GRAFT_REPO_ROOT:-.}"
T=$1; shift
O=gpurun_out/$T
mkdir -p $O
B="python bench.py --quick --steps 30 --warmup 5"
for rep in 1 2; do
  $B > $O/base_$rep.json 2> $O/base_$rep.err
  for l in "$@"; do
    n=$(basename $l .so)
    SCAB_LIB_PATH=$PWD/$l $B > $O/${n}_$rep.json 2> $O/${n}_$rep.err
  done
done
if [ -n "$AB_LATE" ]; then
  $B --skip-steps 500 > $O/late_base.json 2> $O/late_base.err
  for l in "$@"; do
    n=$(basename $l .so)
    SCAB_LIB_PATH=$PWD/$l $B --skip-steps 500 > $O/late_${n}.json 2> $O/late_${n}.err
  done
fi
python tools/bench_table.py $O/*.json
