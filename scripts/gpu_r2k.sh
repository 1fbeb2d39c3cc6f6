#!/bin/bash
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dist (overlap on)"; timeout -k 5 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_dist.log
echo "== pytest dist (overlap off)"; CSB_HALO_OVERLAP=0 timeout -k 5 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "fcg_multigrid_pc or distributed_pcg" 2>&1 | tail -5 | tee $OUT/pytest_dist_noovl.log
echo "== pytest r2 quick"; timeout -k 5 900 python -m pytest tests/test_gpu_r2.py -m gpu -q -k "not 256 and not 128" 2>&1 | tail -4
ls $OUT
