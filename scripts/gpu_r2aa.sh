#!/bin/bash
# restriction through the slab pipeline (CSB_RESTRICT_RING=1, default) against the register-batched form (=0)
TAG=${1:-r02aa}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity (ring restriction)"
timeout -k 5 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_native.py tests/test_gpu_dist.py -m gpu -q -x -k "not 256" 2>&1 | tail -5 | tee $OUT/pytest_ring.log
for rr in 1 0; do
  echo "== CSB_RESTRICT_RING=$rr"
  CSB_RESTRICT_RING=$rr timeout -k 5 300 python scripts/fcg_mg_phases.py --solves 3 --out $OUT/phases_rr$rr.md 2>&1 | tail -2 | cut -c1-200
  sed -n 5,10p $OUT/phases_rr$rr.md | cut -c1-230
done
