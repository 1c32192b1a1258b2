#!/bin/bash
# GPU session 4 of round 2: parity suite, bounds-check build over the parity cases, texture-path variant
mkdir -p gpurun_out; rm -f gpurun_out/r02_bounds_check.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
PRT_LIB_PATH=$PWD/build/variants/libprt_boundscheck.so timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_boundscheck.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_boundscheck.log; cat gpurun_out/r02_bounds_check.txt
for n in 1048576 131072; do PRT_LIB_PATH=$PWD/build/variants/libprt_tex.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/tex /"; done > gpurun_out/r02_variants_d.log 2>&1
PRT_LIB_PATH=$PWD/build/variants/libprt_tex.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/tex /" >> gpurun_out/r02_variants_d.log
PRT_LIB_PATH=$PWD/build/variants/libprt_tex.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "config3 or config2_range or steep or ensemble_launch" 2>&1 | tail -2 | sed "s/^/tex /" >> gpurun_out/r02_variants_d.log
for n in 1048576 131072; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/default /"; done >> gpurun_out/r02_variants_d.log 2>&1
cat gpurun_out/r02_variants_d.log
