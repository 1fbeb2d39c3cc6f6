#!/bin/bash
# ncu --set full of the level-0 restriction and prolongation (kernel-per-phase forms k_restrict / k_prolong_set|inc)
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export CSB_FCG_MG_MODE_MULTI=1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_restrict -c 1 \
  -o $OUT/restrict_l0 python scripts/fcg_mg_phases.py --solves 1 --max-iter 2 > $OUT/ncu_restrict.log 2>&1
tail -2 $OUT/ncu_restrict.log | cut -c1-200
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_prolong -s 10 -c 1 \
  -o $OUT/prolong_l0 python scripts/fcg_mg_phases.py --solves 1 --max-iter 2 > $OUT/ncu_prolong.log 2>&1
tail -2 $OUT/ncu_prolong.log | cut -c1-200
ls -la $OUT
