#!/bin/bash
# Build libies_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${IES_EXTRA_FLAGS} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in engine spectral_f32 spectral_f64 spectral_c64 spectral_c128 spectral_fused; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer build/$f.o)" ] || [ ../../include/ies_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libies_b200.so build/engine.o build/spectral_f32.o build/spectral_f64.o build/spectral_c64.o build/spectral_c128.o build/spectral_fused.o -gencode arch=compute_100a,code=sm_100a
echo "built $(cd .. && pwd)/libies_b200.so"
