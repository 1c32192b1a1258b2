#!/bin/bash
# GPU session 8: uniform range+depth mode (ZM = 3) -- parity suite, timings, final N = 1 bench lines, smoke
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
for n in 1048576 262144 131072; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/zm3 /"; done > gpurun_out/r02_variants_f.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/zm0 /" >> gpurun_out/r02_variants_f.log
timeout 300 python tools/prof_refine.py >> gpurun_out/r02_variants_f.log 2>&1
cat gpurun_out/r02_variants_f.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/r02_bench_cfg3_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg3_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['eigenrays']['ms'], d['eigenrays']['ztol_1m']['ms'])"; tail -3 gpurun_out/r02_bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg3_n1_reference_arm.json 2> gpurun_out/r02_bench_cfg3_n1_reference_arm.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_cfg2_n1.json 2> gpurun_out/r02_bench_cfg2_n1.err
head -c 250 gpurun_out/r02_bench_cfg2_n1.json; echo
timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_cfg1_n1.json 2> gpurun_out/r02_bench_cfg1_n1.err
head -c 250 gpurun_out/r02_bench_cfg1_n1.json; echo; tail -2 gpurun_out/r02_bench_cfg1_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
