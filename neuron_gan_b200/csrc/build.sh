#!/bin/bash
# Builds libngan_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --use_fast_math -Xptxas -v"
mkdir -p build
pids=()
for f in api conv3x3_umma conv3x3_fold wgrad elementwise linear adam; do
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libngan_b200.so build/*.o -lcudart
echo "built $(cd .. && pwd)/libngan_b200.so"
