#!/bin/bash
OUT=gpurun_out/exp13; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
echo "== pytest -m gpu"; timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== phases"; timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -32 | cut -c1-250 | grep "^# mult\|^| [0-9]* \|whole-cycle"
echo "== trace L7"; CSB_LIB=code_saturne_b200/csrc/libcsb200_trace.so timeout -k 5 200 python scripts/phase_trace.py 2>&1 | tail -7
