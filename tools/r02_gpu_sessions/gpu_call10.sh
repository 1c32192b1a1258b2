#!/bin/bash
# GPU session 10: un-inlined exact field lookup (code size 132 -> 78 KB) and 5 resident blocks on the final kernel
mkdir -p gpurun_out
for v in inl nif inl_x5 nif_x5; do
  for n in 1048576 131072; do PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"; done
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_refine.py 2>&1 | tail -1 | sed "s/^/$v /"
done > gpurun_out/r02_variants_g.log 2>&1
cat gpurun_out/r02_variants_g.log
PRT_LIB_PATH=$PWD/build/variants/libprt_nif.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
