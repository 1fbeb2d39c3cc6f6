#!/bin/bash
OUT=gpurun_out/exp4; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1
echo "== parity (default)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8 | tee $OUT/parity.log
echo "== default"; timeout 300 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -17 | cut -c1-330 | head -9
for v in ns2x3 p128x4 p128ns2x6 p128ns4x3; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout 300 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -17 | cut -c1-330 | head -9
done
