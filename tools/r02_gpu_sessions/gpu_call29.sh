#!/bin/bash
# GPU session 29: the other workloads on the final tree (N = 1)
mkdir -p gpurun_out
for w in cfg1 cfg2 cfg5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_${w}_n1_final.json 2> gpurun_out/r02_bench_${w}_n1_final.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${w}_n1_final.json')); e=d.get('eigenrays') or {}
print('$w', round(d['ms_per_step'],3), 'ms', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'frac', round(d['roofline']['frac'],4), 'eig ms', e.get('ms'), (e.get('ztol_1m') or {}).get('ms'), e.get('parity_vs_cpu'))"
done
