#!/bin/bash
# dev: build kernel variants (resident blocks x threads per block of the arrival-table kernels) into build/variants/
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -shared"
build() {  # name, block, min blocks
  nvcc $FLAGS -DPRT_BLOCK_ARR=$2 -DPRT_MIN_BLOCKS_ARR=$3 -o build/variants/libprt_$1.so pygenray_b200/csrc/prt_kernels.cu &
}
build b128x5 128 5
build b128x4 128 4
build b256x2 256 2
build b320x2 320 2
wait
build b64x10 64 10
build b64x8 64 8
build b640x1 640 1
build b512x1 512 1
wait
ls -la build/variants/
