#!/bin/bash
# 8-GPU box: process-per-GPU parity (merge + overlap + NCCL halo path), then the bench at 8 / 4 / 2.
TAG=${1:-r02r}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CSB_SPIN_TIMEOUT_S=20
nvidia-smi topo -m 2>&1 | head -12 > $OUT/topo.txt
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$2 --master-addr 127.0.0.1 --master-port $3 "${@:4}" 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== proc worker fcgmg N=64 R=8"; tr 300 8 29521 tests/dist_proc_worker.py 64 dirichlet 1e-8 fcgmg 3000 | tail -3 | tee $OUT/worker64_r8.log
echo "== proc worker nccl N=32 R=8"; tr 300 8 29522 tests/dist_proc_worker.py 32 dirichlet 1e-8 nccl | tail -3 | tee $OUT/worker_nccl_r8.log
echo "== proc worker mg N=48 R=4"; tr 300 4 29527 tests/dist_proc_worker.py 48 neumann 1e-8 mg | tail -3 | tee $OUT/worker48_r4.log
for n in 8 4 2; do
  echo "== bench $n GPUs"
  tr 600 $n $((29530+n)) bench.py --gpus $n --steps 10 --warmup 3 | tail -1 | tee $OUT/bench$n.json | cut -c1-600
done
ls $OUT
