#!/bin/bash
TAG=${1:-r02rhs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest rhs"; timeout -k 5 600 python -m pytest tests/test_gpu_rhs.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_rhs.log
echo "== times"; timeout -k 5 300 python scripts/rhs_times.py 128 256 2>&1 | tail -3 | tee $OUT/rhs_times.md
