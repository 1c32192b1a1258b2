#!/bin/bash
# GPU session 13: final code on 8 GPUs -- the driver's launch lines at N = 8, 4, 2 (config 3 strong scaling), config 5 at 8
mkdir -p gpurun_out
run() { n=$1; port=$2; stem=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/$stem.json 2> gpurun_out/$stem.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/$stem.json').read().strip().splitlines()[-1]); e=d.get('eigenrays') or {}
print('$stem', round(d['ms_per_step'],2), 'ms', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'), 'eig', e.get('ms'), (e.get('ztol_1m') or {}).get('ms'))"
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/$stem.err | tail -2
}
run 8 29721 r02_bench_cfg3_n8_reference_arm --impl reference --steps 3 --warmup 1
run 8 29722 r02_bench_cfg3_n8 --steps 5 --warmup 3
run 4 29723 r02_bench_cfg3_n4 --steps 5 --warmup 3
run 2 29724 r02_bench_cfg3_n2 --steps 5 --warmup 3
run 1 29725 r02_bench_cfg3_n1_on8box --steps 5 --warmup 3 --no-python-ref
run 8 29726 r02_bench_cfg5_n8 --workload cfg5 --steps 5 --warmup 3 --no-python-ref
