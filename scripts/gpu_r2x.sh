#!/bin/bash
# rows in flight per thread in restriction / prolongation (CSB_RESTRICT_U, CSB_PROLONG_U): same box, same cube
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "" _r4 _r4p8 _r6p8; do
  echo "== variant libcsb200$v.so"
  lib=code_saturne_b200/csrc/libcsb200$v.so
  if [ -z "$v" ]; then unset CSB_LIB; else export CSB_LIB=$lib; fi
  timeout -k 5 300 python scripts/fcg_mg_phases.py --solves 3 --out $OUT/phases$v.md 2>&1 | tail -3 | cut -c1-200
  sed -n 5,8p $OUT/phases$v.md | cut -c1-260
done
