#!/bin/bash
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest fallback + blockn + gs"; timeout -k 5 900 python -m pytest tests/test_gpu_fallback.py tests/test_gpu_blockn.py tests/test_gpu_gs.py -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest.log
echo "== spmv variants"; timeout -k 5 600 python scripts/spmv_variants.py 2>&1 | tee $OUT/spmv_rows_variants.md
ls $OUT
