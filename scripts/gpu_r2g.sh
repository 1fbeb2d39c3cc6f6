#!/bin/bash
# Round-2 GPU visit G: split-ring SpMV pipeline: correctness, then A/B against the previous build.
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (persistent-kernel tests)"; timeout -k 5 1500 python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_widen.py -m gpu -q -x -k "not 256" 2>&1 | tail -12 | tee $OUT/pytest.log
echo "== phases: split ring, gather depth 3 (default)"
timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_split_gd3.md 2>&1 | tail -22 | cut -c1-200
echo "== phases: split ring, gather depth 2"
CSB_GATHER_DEPTH=2 timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_split_gd2.md 2>&1 | tail -22 | cut -c1-200
echo "== phases: previous build (single ring)"
CSB_LIB=code_saturne_b200/csrc/libcsb200_prev.so timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_prev.md 2>&1 | tail -8 | cut -c1-200
ls -la $OUT
