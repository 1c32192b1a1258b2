#!/bin/bash
# GPU session 5: two ranks -- the driver's launch line for both arms of bench.py (config 3, strong scaling)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg3_n2_reference_arm.json 2> gpurun_out/r02_bench_cfg3_n2_reference_arm.err
tail -c 400 gpurun_out/r02_bench_cfg3_n2_reference_arm.json; tail -3 gpurun_out/r02_bench_cfg3_n2_reference_arm.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n2.json 2> gpurun_out/r02_bench_cfg3_n2.err
tail -c 2500 gpurun_out/r02_bench_cfg3_n2.json; tail -5 gpurun_out/r02_bench_cfg3_n2.err
