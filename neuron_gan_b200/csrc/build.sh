#!/bin/bash
# Builds libngan_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
#   NGAN_DEBUG_BUILD=1 bash build.sh   -> libngan_b200_dbg.so with the conv timing/trace hooks compiled in
#                                         (load it with NGAN_LIB=<path>; scripts/trace_conv.py)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --use_fast_math -Xptxas -v"
OUT=../libngan_b200.so
BUILD=build
if [ -n "$NGAN_DEBUG_BUILD" ]; then
  FLAGS="$FLAGS -DNGAN_CONV_DEBUG_BUILD"
  OUT=../libngan_b200_dbg.so
  BUILD=build_dbg
fi
if [ -n "$NGAN_VARIANT" ]; then      # tuning variants: NGAN_VARIANT=name NGAN_EXTRA_FLAGS="-D..." bash build.sh
  FLAGS="$FLAGS $NGAN_EXTRA_FLAGS"
  OUT=../libngan_b200_$NGAN_VARIANT.so
  BUILD=build_$NGAN_VARIANT
fi
mkdir -p $BUILD
pids=()
for f in api conv3x3_umma conv3x3_fold wgrad elementwise linear adam augment; do
  ( $NVCC $FLAGS -c $f.cu -o $BUILD/$f.o > $BUILD/$f.log 2>&1 || { cat $BUILD/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $BUILD/*.o -lcudart
echo "built $(cd .. && pwd)/$(basename $OUT)"
