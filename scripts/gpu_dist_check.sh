#!/bin/bash
# Distributed scenarios with R ranks as host threads on the visible GPU(s).
export CSB_VERBOSE_ERRORS=1
export CSB_SPIN_TIMEOUT_S=${CSB_SPIN_TIMEOUT_S:-8}
timeout 1500 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
