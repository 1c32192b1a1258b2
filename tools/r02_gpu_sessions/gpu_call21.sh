#!/bin/bash
# GPU session 21: the reference's own physics tests against the product + flat_earth_c golden; shared-memory carve-out sweep;
# default bench (both arms) on the final tree
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu_call21.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_call21.log
for cv in default 0 4 7 14 28; do
  if [ $cv = default ]; then unset PRT_SMEM_CARVEOUT; else export PRT_SMEM_CARVEOUT=$cv; fi
  for n in 1048576 131072; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/carve=$cv /"; done
  timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/carve=$cv /"
done > gpurun_out/r02_variants_n.log 2>&1
cat gpurun_out/r02_variants_n.log
unset PRT_SMEM_CARVEOUT
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default_n1_call21.json 2> gpurun_out/r02_bench_default_n1_call21.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_default_n1_call21.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['eigenrays']['ms'], d['eigenrays']['ztol_1m']['ms'])"; tail -3 gpurun_out/r02_bench_default_n1_call21.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_default_n1_reference_arm_call21.json 2> gpurun_out/r02_bench_default_n1_reference_arm_call21.err
head -c 300 gpurun_out/r02_bench_default_n1_reference_arm_call21.json; echo
