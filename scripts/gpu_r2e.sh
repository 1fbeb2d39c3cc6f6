#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dist fcgmg"; timeout -k 5 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "fcg_multigrid_pc" 2>&1 | tail -30 | tee $OUT/pytest_dist.log
echo "== pytest all (except 256)"; timeout -k 5 1800 python -m pytest tests -m gpu -q -k "not 256 and not fcg_multigrid_pc" 2>&1 | tail -15 | tee $OUT/pytest_all.log
ls -la $OUT
