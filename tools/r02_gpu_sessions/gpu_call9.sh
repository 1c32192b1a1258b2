#!/bin/bash
# GPU session 9: final kernel (grid mode 3) -- parity suite, extra fuzz seeds, ncu launch list + full captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python tests/tools/fuzz_gpu.py 600 760 > gpurun_out/r02_fuzz_gpu.log 2>&1; tail -2 gpurun_out/r02_fuzz_gpu.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench_cfg3_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/plain_fan.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_fan_kernel -s 1 -c 1 -o gpurun_out/r02_fan_cfg3_1M_final \
    python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/ncu_fan.log 2>&1
timeout 300 python tools/prof_refine.py > gpurun_out/plain_refine.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_refine_round -s 3 -c 1 -o gpurun_out/r02_refine_cfg4_round3_final \
    python tools/prof_refine.py > gpurun_out/ncu_refine.log 2>&1
ls -la gpurun_out/ | tail -12
