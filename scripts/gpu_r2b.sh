#!/bin/bash
# Round-2 GPU visit B: fixed tests (+ 256^3 parity, shim), gather-depth / stage-count variants.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest r2 + shim + pcg_npc"; timeout -k 5 1500 python -m pytest tests/test_gpu_r2.py tests/test_gpu_shim.py "tests/test_gpu_parity.py::test_pcg_without_preconditioner_matches_oracle" -m gpu -q -s 2>&1 | tail -40 | tee $OUT/pytest_r2.log
for v in "ns4 3" "ns4 2"; do
  set -- $v
  echo "== variant lib=$1 gather_depth=$2"
  CSB_LIB=code_saturne_b200/csrc/libcsb200_$1.so CSB_GATHER_DEPTH=$2 timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_256_$1_gd$2.md 2>&1 | tail -22 | cut -c1-200
done
ls -la $OUT
