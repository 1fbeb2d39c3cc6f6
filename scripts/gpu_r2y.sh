#!/bin/bash
# direct-load path of the small levels in k_fcg_mg_persistent (CSB_DIRECT_ROWS): parity first, then timing
TAG=${1:-r02y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity with CSB_DIRECT_ROWS=70000"
CSB_DIRECT_ROWS=70000 timeout -k 5 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_native.py -m gpu -q -x -k "not 256" 2>&1 | tail -5 | tee $OUT/pytest_direct.log
for dr in 0 8000 25000 70000 250000; do
  echo "== CSB_DIRECT_ROWS=$dr"
  CSB_DIRECT_ROWS=$dr timeout -k 5 300 python scripts/fcg_mg_phases.py --solves 3 --out $OUT/phases_dr$dr.md 2>&1 | tail -2 | cut -c1-200
  sed -n 9,16p $OUT/phases_dr$dr.md | cut -c1-200
done
