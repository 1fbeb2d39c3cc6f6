#!/bin/bash
# 2-GPU box: one process per GPU over cudaIpc peer memory
OUT=gpurun_out/dist2; mkdir -p $OUT
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=20
nvidia-smi topo -m 2>&1 | head -8 > $OUT/topo.txt
echo "== proc worker N=32 R=2"
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_proc_worker.py 32 dirichlet 2>&1 | tail -12 | tee $OUT/worker32.log
echo "== proc worker N=64 R=2 neumann"
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
  tests/dist_proc_worker.py 64 neumann 2>&1 | tail -6 | tee $OUT/worker64.log
echo "== bench 2 GPUs"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -8 | tee $OUT/bench2.log
echo "== bench 1 GPU"
CUDA_VISIBLE_DEVICES=0 timeout -k 5 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench1.log
