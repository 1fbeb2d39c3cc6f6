#!/bin/bash
OUT=gpurun_out/exp9; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
echo "== pytest -m gpu (all)"; ( time timeout -k 5 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) 2>&1 | tee $OUT/pytest_gpu.log
echo "== phases"; timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -32 | cut -c1-250 | head -8
