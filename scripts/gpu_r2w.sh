#!/bin/bash
# one-CTA threshold of the coarse sub-cycle in k_fcg_mg_persistent (CSB_SB_ROWS): same box, same cube
TAG=${1:-r02w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for sb in 512 1000 2500 7000; do
  echo "== CSB_SB_ROWS=$sb"
  CSB_SB_ROWS=$sb timeout -k 5 300 python scripts/fcg_mg_phases.py --solves 3 --out $OUT/phases_sb$sb.md 2>&1 | tail -3 | cut -c1-200
done
