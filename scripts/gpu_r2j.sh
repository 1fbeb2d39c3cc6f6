#!/bin/bash
# 8-GPU box: FCG + V-cycle with coarse-level merge.  Usage: gpu_r2j.sh [tag]
TAG=${1:-r02j}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CSB_SPIN_TIMEOUT_S=20
nvidia-smi topo -m 2>&1 | head -12 > $OUT/topo.txt
tr() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$2 --master-addr 127.0.0.1 --master-port $3 "${@:4}" 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== proc worker fcgmg N=64 R=8"
tr 300 8 29521 tests/dist_proc_worker.py 64 dirichlet 1e-8 fcgmg 3000 | tail -4 | tee $OUT/worker64_r8.log
echo "== bench 8 GPUs merge_glob=250000"
tr 600 8 29523 bench.py --gpus 8 --steps 10 --warmup 3 --merge-glob 250000 | tail -2 | tee $OUT/bench8_merge250000.log
echo "== bench 8 GPUs merge_glob=70000"
tr 600 8 29524 bench.py --gpus 8 --steps 10 --warmup 3 --merge-glob 70000 | tail -2 | tee $OUT/bench8_merge70000.log
echo "== bench 4 GPUs merge_glob=250000"
tr 600 4 29525 bench.py --gpus 4 --steps 10 --warmup 3 --merge-glob 250000 | tail -2 | tee $OUT/bench4_merge250000.log
echo "== bench 2 GPUs merge_glob=250000"
tr 600 2 29526 bench.py --gpus 2 --steps 10 --warmup 3 --merge-glob 250000 | tail -2 | tee $OUT/bench2_merge250000.log
