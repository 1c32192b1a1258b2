#!/bin/bash
# GPU session 12: cold paths out of line, second step (code 92 -> 80 KB)
mkdir -p gpurun_out
for v in coldrhs cold2; do
  for n in 1048576 131072; do PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"; done
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_refine.py 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_refine.py --max-iter 20 2>&1 | tail -1 | sed "s/^/$v full /"
done > gpurun_out/r02_variants_i.log 2>&1
cat gpurun_out/r02_variants_i.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
