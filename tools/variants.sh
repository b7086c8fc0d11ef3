#!/bin/bash
# Developer tool: build timing variants of the library (debug macros in nbr_engine.cuh) into tools/bin/.
#   tools/variants.sh NBR_DBG_NO_APPEND NBR_DBG_NO_DRAIN ...   ->  tools/bin/libv_<macro>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for m in "$@"; do
  n=$(echo $m | tr "=," "__"); d=$(echo $m | sed "s/,/ -D/g")
  ( nvcc $FLAGS -D$d -shared -o tools/bin/libv_$n.so mocopci_b200/csrc/common.cu mocopci_b200/csrc/knn.cu \
      mocopci_b200/csrc/fps.cu mocopci_b200/csrc/gather.cu mocopci_b200/csrc/emd.cu mocopci_b200/csrc/probe.cu -lcudart 2>&1 | grep -E "error" || true ) &
done
wait
ls -la tools/bin/libv_*.so
