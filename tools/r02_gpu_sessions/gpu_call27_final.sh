#!/bin/bash
# GPU session 27: the round's final tree exactly as the driver runs it -- pytest -m gpu, smoke(), default bench both arms
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final2.log 2>&1; tail -1 gpurun_out/r02_smoke_final2.log
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_default_n1_reference_arm_final2.json 2> gpurun_out/ref_arm.err
wc -l gpurun_out/r02_bench_default_n1_reference_arm_final2.json
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_n1_final2.json 2> gpurun_out/bench.err ) 2>&1 | grep real
wc -l gpurun_out/r02_bench_default_n1_final2.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_default_n1_final2.json')); r=json.load(open('gpurun_out/r02_bench_default_n1_reference_arm_final2.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','warmup')}, d['roofline']['frac'], d['e2e']['value'], d['eigenrays']['ms'], d['eigenrays']['ztol_1m']['ms'], d['eigenrays']['parity_vs_cpu'], 'ref', r['value'], 'same_config', d['config']==r['config'])"
