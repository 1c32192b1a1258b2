#!/bin/bash
# GPU session 25 (2 GPUs): final tree under the driver's torchrun line, both arms (config 3, strong scaling)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_cfg3_n2_reference_arm_final.json 2> gpurun_out/r02_bench_cfg3_n2_reference_arm_final.err
tail -c 300 gpurun_out/r02_bench_cfg3_n2_reference_arm_final.json; tail -2 gpurun_out/r02_bench_cfg3_n2_reference_arm_final.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n2_final.json 2> gpurun_out/r02_bench_cfg3_n2_final.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg3_n2_final.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'], d['e2e']['value'], d['eigenrays']['ms'], d['eigenrays']['ztol_1m']['ms'], d['eigenrays'].get('parity_vs_cpu'))"; tail -3 gpurun_out/r02_bench_cfg3_n2_final.err
