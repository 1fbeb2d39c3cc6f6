#!/bin/bash
OUT=gpurun_out/exp2; mkdir -p $OUT
export CSB_LIB=code_saturne_b200/csrc/libcsb200_u2.so
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mg_persistent -s 1 -c 1 \
  -o $OUT/mg_u2 python scripts/ncu_mg.py --cube 256 --precision 1e-2 > $OUT/ncu.log 2>&1
tail -5 $OUT/ncu.log
ls -la $OUT
