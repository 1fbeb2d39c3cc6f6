#!/bin/bash
# N-GPU box: one process per GPU over cudaIpc peer memory.  Usage: gpu_distN.sh N
N=${1:-4}
OUT=gpurun_out/dist$N; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
nvidia-smi topo -m 2>&1 | head -12 > $OUT/topo.txt
echo "== proc worker N=48 R=$N"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29521 \
  tests/dist_proc_worker.py 48 neumann 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6 | tee $OUT/worker48.log
echo "== proc worker N=64 R=$N dirichlet 1e-12"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29522 \
  tests/dist_proc_worker.py 64 dirichlet 1e-12 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -4 | tee $OUT/worker64.log
echo "== bench $N GPUs"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29523 \
  bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3 | tee $OUT/bench$N.log
