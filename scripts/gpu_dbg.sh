#!/bin/bash
export CSB_VERBOSE_ERRORS=1 CSB_SPIN_TIMEOUT_S=10
echo "== mg 16 2"; timeout -k 5 120 python tests/dist_worker.py mg 16 2 dirichlet 1e-8 2>&1 | tail -25
echo "== mg 24 1"; timeout -k 5 120 python tests/dist_worker.py mg 24 1 dirichlet 1e-8 2>&1 | tail -5
echo "== mg 32 2"; timeout -k 5 120 python tests/dist_worker.py mg 32 2 dirichlet 1e-8 2>&1 | tail -8
