#!/bin/bash
# 1-GPU check of the rows added late in round 2 (native-array matrices, matrix assembler) and a
# timing of the native-path setup at bench size.
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest new"; timeout -k 5 900 python -m pytest tests/test_gpu_native.py tests/test_gpu_assembler.py -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_new.log
echo "== native / assembler times"; timeout -k 5 900 python scripts/native_times.py 2>&1 | tee $OUT/native_times.md | tail -30
ls -la $OUT
