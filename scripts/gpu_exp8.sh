#!/bin/bash
OUT=gpurun_out/exp8; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1
echo "== parity (default)"; timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee $OUT/parity.log
echo "== default"; timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -32 | cut -c1-250
for v in g10 g12 ns2; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -32 | cut -c1-250 | grep -v "^| [4-9] \|^| 1[01] \|^$"
done
