#!/bin/bash
# Full GPU suite + bench after the Gauss-Seidel family went in.
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gs"; timeout -k 5 900 python -m pytest tests/test_gpu_gs.py -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_gs.log
echo "== pytest all"; timeout -k 5 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gs.py 2>&1 | tail -12 | tee $OUT/pytest_all.log
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
echo "== gs timing"; timeout -k 5 300 python scripts/gs_times.py 2>&1 | tail -20 | tee $OUT/gs_times.md
ls $OUT
