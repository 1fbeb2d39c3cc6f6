#!/bin/bash
OUT=gpurun_out/exp6; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1
echo "== smoke-size check"; timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== parity (default)"; timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8 | tee $OUT/parity.log
echo "== default"; timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -17 | cut -c1-330 | head -9
for v in ns2 ns4 g6 p512ns2; do
  echo "== $v"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$v.so timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_$v.md 2>&1 | tail -17 | cut -c1-330 | head -9
done
