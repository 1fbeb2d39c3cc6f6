#!/bin/bash
# N-GPU box: one process per GPU.  Usage: gpu_r2f.sh N [tag]
N=${1:-2}
TAG=${2:-r02f}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CSB_SPIN_TIMEOUT_S=20
nvidia-smi topo -m 2>&1 | head -12 > $OUT/topo.txt
echo "== proc worker fcgmg N=64 R=$N"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29521 \
  tests/dist_proc_worker.py 64 dirichlet 1e-8 fcgmg 3000 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6 | tee $OUT/worker64.log
for mg in 250000 0; do
echo "== bench $N GPUs merge_glob=$mg"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29523 \
  bench.py --gpus $N --steps 5 --warmup 3 --merge-glob $mg 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3 | tee $OUT/bench${N}_merge$mg.log
done
