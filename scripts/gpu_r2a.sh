#!/bin/bash
# Round-2 GPU visit A: new persistent FCG + V-cycle kernel (tests, bench, phase table).
# Run as: gpurun --timeout 2400 -- bash scripts/gpu_r2a.sh <tag>
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; free -g >> $OUT/gpu.txt
echo "== pytest r2 (no 256)"; timeout -k 5 1200 python -m pytest tests/test_gpu_r2.py -m gpu -q -k "not 256" 2>&1 | tail -40 | tee $OUT/pytest_r2.log
echo "== pytest rest"; timeout -k 5 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_r2.py 2>&1 | tail -40 | tee $OUT/pytest_rest.log
echo "== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== phases"; timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_256.md 2>&1 | tail -30 | cut -c1-220
ls -la $OUT
