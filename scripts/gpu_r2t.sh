#!/bin/bash
# 2-GPU box: NCCL halo worker (full output), overlap A/B of the bench.
TAG=${1:-r02t}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CSB_SPIN_TIMEOUT_S=20
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$2 --master-addr 127.0.0.1 --master-port $3 "${@:4}" 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== proc worker nccl N=24 R=2"; tr 300 2 29522 tests/dist_proc_worker.py 24 neumann 1e-8 nccl > $OUT/worker_nccl_r2.log 2>&1; grep -v "^  \|^===\|^---" $OUT/worker_nccl_r2.log | tail -25
echo "== bench 2 GPUs, exchange waited for first (default)"
tr 600 2 29531 bench.py --gpus 2 --steps 10 --warmup 3 | tail -1 | tee $OUT/bench2_wait_first.json | cut -c1-200
echo "== bench 2 GPUs, overlapped exchange"
CSB_HALO_OVERLAP=1 tr 600 2 29532 bench.py --gpus 2 --steps 10 --warmup 3 | tail -1 | tee $OUT/bench2_overlap.json | cut -c1-200
ls $OUT
