#!/bin/bash
# GPU session 15: final tree -- what the driver runs at round end: pytest -m gpu, smoke(), bench.py (both arms)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r02_bench_default_reference_arm.json 2> gpurun_out/r02_bench_default_reference_arm.err; tail -4 gpurun_out/r02_bench_default_reference_arm.err
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -4 gpurun_out/r02_bench_default.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_default.json')); r=json.load(open('gpurun_out/r02_bench_default_reference_arm.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])
print('same config', d['config']==r['config'], 'e2e ratio', d['e2e']['value']/r['e2e']['value'], 'value ratio', d['value']/r['value'])
print(d['cpu_baseline']['python_reference'])"
