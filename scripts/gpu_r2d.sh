#!/bin/bash
# Round-2 GPU visit D: distributed FCG + V-cycle (in-process ranks on one GPU), merge.
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dist fcgmg"; timeout -k 5 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "fcg_multigrid_pc" 2>&1 | tail -40 | tee $OUT/pytest_dist.log
echo "== pytest 256 parity"; timeout -k 5 1200 python -m pytest tests/test_gpu_r2.py -m gpu -q -s -k 256 2>&1 | tail -8 | tee $OUT/pytest_256.log
ls -la $OUT
