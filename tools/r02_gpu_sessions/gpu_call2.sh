#!/bin/bash
# GPU session 2 of round 2: parity suite, chain-shortening variants, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
for v in fold1 fold0 fold1_b256x2 fold1_b128x5; do
  for n in 1048576 131072; do
    PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  done
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/$v /"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/$v /"
done > gpurun_out/r02_variants_b.log 2>&1
cat gpurun_out/r02_variants_b.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/r02_bench_cfg3_n1.err
tail -c 1500 gpurun_out/r02_bench_cfg3_n1.json; tail -3 gpurun_out/r02_bench_cfg3_n1.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg2_n1.json 2> gpurun_out/r02_bench_cfg2_n1.err
tail -c 600 gpurun_out/r02_bench_cfg2_n1.json; tail -3 gpurun_out/r02_bench_cfg2_n1.err
timeout 300 python tools/prof_refine.py > gpurun_out/r02_prof_refine.log 2>&1; cat gpurun_out/r02_prof_refine.log
# ncu: launch list of the bench command, then full captures of the two hot kernels
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/plain_fan.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_fan_kernel -s 1 -c 1 -o gpurun_out/r02_fan_cfg3_1M \
    python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 > gpurun_out/ncu_fan.log 2>&1
timeout 300 python tools/prof_refine.py > gpurun_out/plain_refine.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prt_refine_round -s 3 -c 1 -o gpurun_out/r02_refine_cfg4_round3 \
    python tools/prof_refine.py > gpurun_out/ncu_refine.log 2>&1
ls -la gpurun_out/
