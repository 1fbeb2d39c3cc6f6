#!/bin/bash
# BASELINE config 5: ~200 M cells, weak scaling -- 584^3 = 199.2 M rows in 8 bricks of 292^3, one per GPU
N=${1:-8}; CUBE=${2:-584}
OUT=gpurun_out/r02c5; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
timeout -k 5 185 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --cube $CUBE --steps 2 --warmup 1 --no-cpu-baseline --no-vcycle-solver --spmv-launches 2 2>&1 \
  | grep -v "^\*\|OMP_NUM\|^$" | tail -3 | tee $OUT/bench_c5_${N}gpu.json | cut -c1-700
