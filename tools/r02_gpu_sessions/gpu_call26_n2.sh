#!/bin/bash
# GPU session 26 (2 GPUs): stdout of the multi-rank bench is exactly one JSON line (NCCL banner routed to stderr)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-python-ref > gpurun_out/r02_bench_cfg3_n2_final.json 2> gpurun_out/r02_bench_cfg3_n2_final.err
wc -l gpurun_out/r02_bench_cfg3_n2_final.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg3_n2_final.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'], d['e2e']['value'], d['eigenrays']['ms'], d['eigenrays']['ztol_1m']['ms'], d['eigenrays'].get('parity_vs_cpu'))"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r02_bench_cfg3_n2_final.err | tail -4
