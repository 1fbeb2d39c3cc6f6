#!/bin/bash
OUT=gpurun_out/exp7; mkdir -p $OUT
echo "== default"; timeout -k 5 150 python scripts/mg_phases.py --out $OUT/phases_default.md 2>&1 | tail -32 | cut -c1-330
