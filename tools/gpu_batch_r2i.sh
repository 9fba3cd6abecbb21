#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2i
mkdir -p $O
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
SCAB_LIB_PATH=$PWD/_ab/lib_knn6b.so $B > $O/bench_knn6b.json 2> $O/bench_knn6b.err
S="python bench.py --quick --steps 2 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:"k_neighbors_direct" -s 8 -c 1 -o $O/prof_knn $S > $O/ncu_p1.log 2>&1
ls $O
