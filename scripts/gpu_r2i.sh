#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest all"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_all.log
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
echo "== phases"; timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_256.md 2>&1 | tail -24 | cut -c1-220
ls $OUT
