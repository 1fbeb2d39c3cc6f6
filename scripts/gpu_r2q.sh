#!/bin/bash
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== setup trace"; OMP_NUM_THREADS=$(nproc) timeout -k 5 600 python scripts/setup_trace.py 256 2>&1 | grep -v "^$" > $OUT/setup_trace_256.log; grep "== setup" $OUT/setup_trace_256.log
echo "== pytest all"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_all.log
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json | cut -c1-1800
ls $OUT
