#!/bin/bash
OUT=gpurun_out/exp3; mkdir -p $OUT
for v in p256x2 p256x2u2 p256x2e6 p128x4; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout 300 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -17 | cut -c1-330 | head -9
done
