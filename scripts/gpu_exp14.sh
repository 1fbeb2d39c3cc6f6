#!/bin/bash
for v in 2048 512 128 0; do
  echo "== CSB_SB_ROWS=$v"
  CSB_SB_ROWS=$v timeout -k 5 150 python scripts/mg_phases.py 2>&1 | tail -32 | cut -c1-250 | grep "^# mult\|^| [89] \|^| 1[01] " | head -5
done
