#!/bin/bash
# Final 1-GPU visit of round 2: full parity suite, ncu captures of the two headline kernels,
# launch list of the bench command, phase table, timings of the other rows, config 4, bench.
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest all"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_all.log
echo "== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== ncu full: k_spmv_rows"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_rows -s 3 -c 1 \
  -o $OUT/spmv_rows_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vcycle-solver --spmv-launches 3 > $OUT/ncu_spmv_rows.log 2>&1
tail -1 $OUT/ncu_spmv_rows.log | cut -c1-200
echo "== ncu full: k_fcg_mg_persistent (one whole solve)"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:k_fcg_mg_persistent -s 1 -c 1 \
  -o $OUT/fcg_mg_full python scripts/fcg_mg_phases.py --solves 2 > $OUT/ncu_fcg_mg.log 2>&1
tail -2 $OUT/ncu_fcg_mg.log | cut -c1-200
echo "== ncu launches of the bench command"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spmv-launches 5 > $OUT/ncu_launch_bench.log 2>&1
tail -2 $OUT/launches.csv | cut -c1-300
echo "== phases"; timeout -k 5 300 python scripts/fcg_mg_phases.py --out $OUT/fcg_mg_phases_256.md 2>&1 | tail -4 | cut -c1-220
echo "== widen times (128^3)"; timeout -k 5 900 python scripts/widen_times.py --cube 128 --reps 3 2>&1 | tee $OUT/widen_times_128.md | tail -30
echo "== gs times"; timeout -k 5 300 python scripts/gs_times.py 2>&1 | tee $OUT/gs_times.md | tail -8
echo "== config 4 (thread per block row, with the reference)"; timeout -k 5 1500 python scripts/config4_times.py --out $OUT/config4_times_b3.md 2>&1 | tail -8
echo "== config 4 (thread per component)"; CSB_SPMV_B3_BN=1 timeout -k 5 900 python scripts/config4_times.py --no-ref --out $OUT/config4_times_bn.md 2>&1 | tail -6
echo "== bench"; timeout -k 5 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-400
echo "== bench reference arm"; timeout -k 5 900 python bench.py --impl reference --steps 2 --warmup 0 2>&1 | tail -1 | tee $OUT/bench_ref.json | cut -c1-400
ls -la $OUT
