#!/bin/bash
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest quick (r2 tests)"; timeout -k 5 900 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "not 256 and not 128" 2>&1 | tail -5 | tee $OUT/pytest.log
run() { echo "== $1 ${2:-}"; env $2 timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/phases_$1.md 2>&1 | tail -22 | cut -c1-200 | grep -v "^$" | sed -n '1p;4,8p;17,30p'; }
run split_gd3
run split_gd2 CSB_GATHER_DEPTH=2
run prevnf "CSB_LIB=code_saturne_b200/csrc/libcsb200_prevnf.so"
run prevnf4 "CSB_LIB=code_saturne_b200/csrc/libcsb200_prevnf4.so"
ls $OUT
