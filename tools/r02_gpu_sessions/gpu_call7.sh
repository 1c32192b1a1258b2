#!/bin/bash
# GPU session 7: persistent-warp work queue -- parity suite + timings
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
for n in 1048576 262144 131072 65536; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/queue /"; done > gpurun_out/r02_variants_e.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/queue /" >> gpurun_out/r02_variants_e.log
timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/queue /" >> gpurun_out/r02_variants_e.log
timeout 300 python tools/prof_fan.py --workload cfg2 --n 100000 --save 0 --reps 4 2>&1 | tail -1 | sed "s/^/queue /" >> gpurun_out/r02_variants_e.log
PRT_REORDER=0 timeout 300 python tools/prof_fan.py --workload cfg3 --n 131072 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/queue-noreorder /" >> gpurun_out/r02_variants_e.log
timeout 300 python tools/prof_refine.py >> gpurun_out/r02_variants_e.log 2>&1
cat gpurun_out/r02_variants_e.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_cfg3_n1_queue.json 2> gpurun_out/r02_bench_cfg3_n1_queue.err
tail -c 1200 gpurun_out/r02_bench_cfg3_n1_queue.json | head -c 700; tail -3 gpurun_out/r02_bench_cfg3_n1_queue.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_cfg2_n1_queue.json 2> gpurun_out/r02_bench_cfg2_n1_queue.err
head -c 300 gpurun_out/r02_bench_cfg2_n1_queue.json
