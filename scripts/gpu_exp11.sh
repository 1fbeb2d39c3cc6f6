#!/bin/bash
OUT=gpurun_out/exp11; mkdir -p $OUT
for v in g12t4 g12t2 g10t4 g8t4; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -32 | cut -c1-250 | grep "^# mult\|^| [0-3] \|whole-cycle"
done
