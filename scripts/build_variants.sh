#!/bin/bash
# Experiment helper: link libcsb200_<tag>.so variants that differ only in the CTA shape of
# the persistent kernels (persistent.cu).  Select one at run time with CSB_LIB=<path>.
#   scripts/build_variants.sh "p512x1:-DCSB_PT=512 -DCSB_MIN_CTAS=1" "u2:-DCSB_ROW_U=2"
set -e
HERE=code_saturne_b200/csrc
make -f $HERE/Makefile -j8 > /dev/null
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
    -Xcompiler -fPIC -Iinclude -I$HERE -ccbin /usr/bin/g++ $defs -c $HERE/persistent.cu -o /tmp/persistent_$tag.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ \
    -o $HERE/libcsb200_$tag.so /tmp/persistent_$tag.o $HERE/kernels.o $HERE/runtime.o $HERE/sles.o \
    $HERE/halo.o $HERE/comm.o $HERE/assembler.o $HERE/grid_host.o -Xcompiler -fopenmp -lgomp -ldl
  echo "built $HERE/libcsb200_$tag.so ($defs)"
done
