#!/bin/bash
# GPU session 24: records on the FINAL tree -- bounds-check pass over the whole GPU suite, ncu launch list of the default
# bench, ncu --set full of the fan kernel (full 1M-ray config-3 launch) and of the refinement kernel (config 4, round 3)
mkdir -p gpurun_out
PRT_LIB_PATH=$PWD/build/variants/libprt_boundscheck.so timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_boundscheck_final.log 2>&1
grep -a "PRT_BOUNDS_CHECK\|passed\|failed" gpurun_out/r02_pytest_gpu_boundscheck_final.log | tail -3
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench_cfg3_final2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/plain_fan.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_fan_kernel -s 1 -c 1 -f -o gpurun_out/r02_fan_cfg3_1M_final2 \
    python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/ncu_fan.log 2>&1
timeout 300 python tools/prof_refine.py > gpurun_out/plain_refine.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_refine_round -s 3 -c 1 -f -o gpurun_out/r02_refine_cfg4_round3_final2 \
    python tools/prof_refine.py > gpurun_out/ncu_refine.log 2>&1
tail -2 gpurun_out/plain_fan.log gpurun_out/plain_refine.log
ls -la gpurun_out/*final2*
