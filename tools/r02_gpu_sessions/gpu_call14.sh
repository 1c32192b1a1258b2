#!/bin/bash
# GPU session 14: the guess fold again, now that the L1 pipe is no longer saturated
mkdir -p gpurun_out
for v in default fold1; do
  lib=$PWD/pygenray_b200/libprt.so; [ $v = fold1 ] && lib=$PWD/build/variants/libprt_fold1.so
  for n in 1048576 131072; do PRT_LIB_PATH=$lib timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"; done
  PRT_LIB_PATH=$lib timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$lib timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/$v /"
done > gpurun_out/r02_variants_j.log 2>&1
cat gpurun_out/r02_variants_j.log
