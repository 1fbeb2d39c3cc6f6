#!/bin/bash
# last visit: full GPU suite on the final code, smoke, bench line; then memcheck of the kernels added late
TAG=${1:-r02final2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest all"; timeout -k 5 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_all.log
echo "== smoke"; timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench"; timeout -k 5 300 python bench.py --no-vcycle-solver 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-300
echo "== memcheck (new kernels)"
timeout -k 5 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rhs.py tests/test_gpu_assembler.py -m gpu -q -x -k "golden" 2>&1 | tail -6 | tee $OUT/memcheck.log
