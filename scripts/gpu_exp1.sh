#!/bin/bash
# experiment: parity tests + per-phase breakdown for CTA-shape variants of the persistent kernels
OUT=gpurun_out/exp1; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8 | tee $OUT/parity.log
echo "== default"; timeout 300 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -18
for v in p512x1 p256x3 u2 p512x2 e6; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout 300 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -17 | cut -c1-400
done
