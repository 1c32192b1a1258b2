#!/bin/bash
# GPU session 16: bounds-check build of the FINAL code over the whole GPU suite; 1000 more fuzz seeds
mkdir -p gpurun_out; rm -f gpurun_out/r02_bounds_check.txt
PRT_LIB_PATH=$PWD/build/variants/libprt_boundscheck.so timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_boundscheck.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_boundscheck.log; cat gpurun_out/r02_bounds_check.txt
timeout 1500 python tests/tools/fuzz_gpu.py 1000 2000 > gpurun_out/r02_fuzz_gpu_1000.log 2>&1; tail -2 gpurun_out/r02_fuzz_gpu_1000.log
