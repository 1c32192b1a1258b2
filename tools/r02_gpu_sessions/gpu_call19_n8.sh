#!/bin/bash
# GPU session 19: final code (refinement park_wait 128) -- parity suite + the driver's launch lines at N = 1, 2, 4, 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
run() { n=$1; port=$2; stem=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/$stem.json 2> gpurun_out/$stem.err
  python -c "
import json
d=json.loads(open('gpurun_out/$stem.json').read().strip().splitlines()[-1]); e=d.get('eigenrays') or {}
print('$stem', round(d['ms_per_step'],2), 'ms', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'), 'eig', e.get('ms'), (e.get('ztol_1m') or {}).get('ms'), e.get('parity_vs_cpu'))"
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/$stem.err | tail -2
}
run 8 29821 r02_bench_cfg3_n8 --steps 5 --warmup 3
run 4 29822 r02_bench_cfg3_n4 --steps 5 --warmup 3 --no-python-ref
run 2 29823 r02_bench_cfg3_n2 --steps 5 --warmup 3 --no-python-ref
run 1 29824 r02_bench_cfg3_n1_on8box --steps 5 --warmup 3 --no-python-ref
