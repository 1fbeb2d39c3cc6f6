#!/bin/bash
OUT=gpurun_out/exp10; mkdir -p $OUT
echo "== parity"; timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== bench (pipeline spmv)"; timeout -k 5 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['spmv'])"
echo "== bench (stream spmv)"; CSB_SPMV_STREAM=1 timeout -k 5 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['spmv'])"
