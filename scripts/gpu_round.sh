#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list of the bench
# command, full ncu captures of the persistent V-cycle kernel and of the SpMV kernel.
# Run as: gpurun --timeout 1800 -- bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
if [ "$2" = "ref" ]; then echo "== bench reference"; timeout -k 5 600 python bench.py --impl reference 2>&1 | tail -2 | tee $OUT/bench_ref.json; fi
echo "== phases"; timeout -k 5 200 python scripts/mg_phases.py --out $OUT/mg_phases_256.md > /dev/null 2>&1; tail -16 $OUT/mg_phases_256.md | cut -c1-200
echo "== ncu launches"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --spmv-launches 5 > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full: persistent V-cycle kernel (one whole solve)"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:k_mg_persistent -s 1 -c 1 \
  -o $OUT/mg_full python scripts/ncu_mg.py --cube 256 --precision 1e-8 --solves 2 > $OUT/ncu_full_mg.log 2>&1
tail -2 $OUT/ncu_full_mg.log
if [ "$2" = "ref" ]; then echo "== ncu full: spmv"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 2 -c 2 \
  -o $OUT/spmv_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spmv-launches 3 > $OUT/ncu_full_bench.log 2>&1; fi
ls -la $OUT
