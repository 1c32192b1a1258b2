#!/bin/bash
# GPU session 1 of round 2: parity suite, kernel variants, default bench, launch list
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt
nproc >> gpurun_out/r02_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
for v in b128x5 b128x4 b256x2 b320x2 b64x10 b64x8 b640x1 b512x1; do
  for n in 1048576 262144 131072; do
    PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  done
done > gpurun_out/r02_variants_cfg3.log 2>&1
cat gpurun_out/r02_variants_cfg3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/r02_bench_cfg3_n1.err
tail -c 3000 gpurun_out/r02_bench_cfg3_n1.json; tail -5 gpurun_out/r02_bench_cfg3_n1.err
