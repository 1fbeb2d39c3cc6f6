#!/bin/bash
# 2-GPU box: NCCL halo worker, in-process dist tests (both exchange forms), A/B of the bench.
TAG=${1:-r02u}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CSB_SPIN_TIMEOUT_S=20
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$2 --master-addr 127.0.0.1 --master-port $3 "${@:4}" 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== proc worker nccl N=24 R=2"; tr 300 2 29522 tests/dist_proc_worker.py 24 neumann 1e-8 nccl > $OUT/worker_nccl_r2.log 2>&1; grep -v "^  \|^===\|^---" $OUT/worker_nccl_r2.log | tail -8
echo "== pytest dist"; timeout -k 5 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "not one_process" 2>&1 | tail -4 | tee $OUT/pytest_dist.log
for k in 1 2; do
echo "== bench 2 GPUs, in-line ghost reads (default), run $k"
tr 600 2 2953$k bench.py --gpus 2 --steps 10 --warmup 3 | tail -1 | tee $OUT/bench2_inline_$k.json | cut -c1-160
echo "== bench 2 GPUs, overlapped exchange, run $k"
CSB_HALO_OVERLAP=1 tr 600 2 2954$k bench.py --gpus 2 --steps 10 --warmup 3 | tail -1 | tee $OUT/bench2_overlap_$k.json | cut -c1-160
done
ls $OUT
