#!/bin/bash
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest r2 + dist quick"; timeout -k 5 900 python -m pytest tests/test_gpu_r2.py tests/test_gpu_dist.py -m gpu -q -k "not 256 and not 128" 2>&1 | tail -6 | tee $OUT/pytest.log
echo "== spmv variants"; timeout -k 5 600 python scripts/spmv_variants.py 2>&1 | tee $OUT/spmv_rows_variants.md
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
ls $OUT
