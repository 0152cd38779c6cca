#!/bin/bash
# Build libb2dsp.so for sm_100a in-tree (the .so is git-ignored but travels with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${B2_NVCC_EXTRA:-}"
OBJS=""
for src in api.cu welch_fft.cu big_fft.cu delay_fir.cu csm_herk.cu heig_jacobi.cu cleansc.cu das_map.cu das_map_tc.cu; do
  obj="build/${src%.cu}.o"
  mkdir -p build
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ common.cuh -nt "$obj" ] || [ fft_small.cuh -nt "$obj" ] || [ das_common.cuh -nt "$obj" ] || [ ../../include/b2dsp.h -nt "$obj" ]; then
    echo "nvcc $src"
    $NVCC $FLAGS -c "$src" -o "$obj" &
  fi
  OBJS="$OBJS $obj"
done
wait
$NVCC -shared -o libb2dsp.so $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $(pwd)/libb2dsp.so"
