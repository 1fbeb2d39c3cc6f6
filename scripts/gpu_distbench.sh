#!/bin/bash
# bench only, N GPUs, one process per GPU
N=${1:-2}
OUT=gpurun_out/distb$N; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -2 | tee $OUT/bench$N.log
