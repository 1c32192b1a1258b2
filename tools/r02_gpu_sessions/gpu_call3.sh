#!/bin/bash
# GPU session 3 of round 2: parity suite, uniform-grid kernel + angle-ordered refinement timings, bench, memcheck
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
for n in 1048576 262144 131072; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1; done > gpurun_out/r02_variants_c.log 2>&1
timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 >> gpurun_out/r02_variants_c.log
timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 >> gpurun_out/r02_variants_c.log
timeout 300 python tools/prof_refine.py >> gpurun_out/r02_variants_c.log 2>&1
cat gpurun_out/r02_variants_c.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/r02_bench_cfg3_n1.err
tail -c 1800 gpurun_out/r02_bench_cfg3_n1.json; tail -3 gpurun_out/r02_bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg3_reference_arm.json 2> gpurun_out/r02_bench_cfg3_reference_arm.err
tail -c 600 gpurun_out/r02_bench_cfg3_reference_arm.json
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5_n1.json 2> gpurun_out/r02_bench_cfg5_n1.err
tail -c 700 gpurun_out/r02_bench_cfg5_n1.json; tail -3 gpurun_out/r02_bench_cfg5_n1.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 66 python -m pytest tests/test_gpu_parity.py -q -x \
   -k "status_codes or irregular_depths or node_aligned or bracket_fill_edge or eigenray_table or source_on_surface" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r02_memcheck.log
