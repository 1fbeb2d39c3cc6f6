#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the
# SpMV kernel.  Run as: gpurun --timeout 1500 -- bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -2 | tee $OUT/bench_ref.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --spmv-launches 5 > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full spmv"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 2 -c 2 \
  -o $OUT/spmv_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spmv-launches 3 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
