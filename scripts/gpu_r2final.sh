#!/bin/bash
# Final 1-GPU visit of round 2: full parity suite, smoke, launch list of the bench command, phase table, bench (both
# arms), and the 1-GPU point of the weak-scaling series of BASELINE config 5 (292^3 = 24.9 M rows per GPU).
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest all"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_all.log
echo "== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== ncu launches of the bench command"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vcycle-solver --spmv-launches 5 > $OUT/ncu_launch_bench.log 2>&1
tail -2 $OUT/launches.csv | cut -c1-300
echo "== phases"; timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_256.md 2>&1 | tail -4 | cut -c1-220
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-600
echo "== bench reference arm"; timeout -k 5 900 python bench.py --impl reference --steps 2 --warmup 0 2>&1 | tail -1 | tee $OUT/bench_ref.json | cut -c1-300
echo "== config 5, 1 GPU: 292^3"; timeout -k 5 600 python bench.py --cube 292 --steps 3 --warmup 3 --no-cpu-baseline --no-vcycle-solver --spmv-launches 5 2>&1 | tail -1 | tee $OUT/bench_c5_1gpu.json | cut -c1-400
ls -la $OUT
