#!/bin/bash
# GPU session 11: exact trial step with an out-of-line field lookup (code 132 -> 92 KB) vs fully inlined vs fully out-of-line
mkdir -p gpurun_out
for v in inl coldrhs nif; do
  for n in 1048576 131072; do PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"; done
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_refine.py 2>&1 | tail -1 | sed "s/^/$v /"
done > gpurun_out/r02_variants_h.log 2>&1
cat gpurun_out/r02_variants_h.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_pytest_gpu.log
