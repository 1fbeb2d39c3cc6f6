#!/bin/bash
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt
echo "== setup trace"; OMP_NUM_THREADS=$(nproc) timeout -k 5 600 python scripts/setup_trace.py 256 2>&1 | grep -v "^$" | tee $OUT/setup_trace_256.log | tail -120
