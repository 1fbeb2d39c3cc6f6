#!/bin/bash
OUT=gpurun_out/exp5; mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mg_persistent -s 1 -c 1 \
  -o $OUT/mg_tma python scripts/ncu_mg.py --cube 256 --precision 1e-2 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
