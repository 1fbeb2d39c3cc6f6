#!/bin/bash
# Round-2 GPU visit C: 256^3 parity, ncu captures (SpMV pipeline with source, FCG+MG kernel),
# launch list of the bench command.
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest 256 parity"; timeout -k 5 1200 python -m pytest tests/test_gpu_r2.py -m gpu -q -s -k 256 2>&1 | tail -15 | tee $OUT/pytest_256.log
echo "== ncu full: k_spmv_pipeline (level-0 256^3 SpMV through the solver's phase code)"
CSB_SPMV_PIPELINE=1 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_pipeline -s 3 -c 1 \
  -o $OUT/spmv_pipeline_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vcycle-solver --spmv-launches 3 > $OUT/ncu_spmv_pipeline.log 2>&1
tail -2 $OUT/ncu_spmv_pipeline.log
echo "== ncu full: k_fcg_mg_persistent (one whole solve)"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:k_fcg_mg_persistent -s 1 -c 1 \
  -o $OUT/fcg_mg_full python scripts/fcg_mg_phases.py --solves 2 > $OUT/ncu_fcg_mg.log 2>&1
tail -2 $OUT/ncu_fcg_mg.log
echo "== ncu launches of the bench command"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spmv-launches 5 > $OUT/ncu_launch_bench.log 2>&1
tail -3 $OUT/launches.csv | cut -c1-300
ls -la $OUT
