#!/bin/bash
# dev: variants of the chain-shortening options (all 128 threads x 4 blocks)
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -shared"
rm -f build/variants/*.so
nvcc $FLAGS -DPRT_FOLD_GUESS=0 -o build/variants/libprt_fold0.so pygenray_b200/csrc/prt_kernels.cu &
nvcc $FLAGS -DPRT_FOLD_GUESS=1 -o build/variants/libprt_fold1.so pygenray_b200/csrc/prt_kernels.cu &
nvcc $FLAGS -DPRT_FOLD_GUESS=1 -DPRT_BLOCK_ARR=256 -DPRT_MIN_BLOCKS_ARR=2 -o build/variants/libprt_fold1_b256x2.so pygenray_b200/csrc/prt_kernels.cu &
nvcc $FLAGS -DPRT_FOLD_GUESS=1 -DPRT_MIN_BLOCKS_ARR=5 -o build/variants/libprt_fold1_b128x5.so pygenray_b200/csrc/prt_kernels.cu &
wait
ls -la build/variants/
