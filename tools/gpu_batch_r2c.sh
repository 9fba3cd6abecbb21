#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --quick --steps 20 --warmup 5"
$B > $O/bench_new.json 2> $O/bench_new.err
for v in sel5 knn_u8 knn_c6 knn_c4u8; do
  SCAB_LIB_PATH=$PWD/_ab/lib_$v.so $B > $O/bench_$v.json 2> $O/bench_$v.err
done
S="python bench.py --quick --steps 2 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:"k_select_fast|k_neighbors_direct" -s 7 -c 2 -o $O/prof_sca $S > $O/ncu_p1.log 2>&1
ls -la $O
