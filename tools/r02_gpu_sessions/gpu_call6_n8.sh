#!/bin/bash
# GPU session 6: 8 GPUs -- the driver's launch lines (config 3 strong scaling) at N = 8 and 4, config 5 at 8, PCIe table
mkdir -p gpurun_out
run() { # N, port, extra args..., output stem
  n=$1; port=$2; stem=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/$stem.json 2> gpurun_out/$stem.err
  tail -c 300 gpurun_out/$stem.json; echo; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/$stem.err | tail -3
}
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1; nproc >> gpurun_out/r02_topo_n8.txt
run 8 29621 r02_bench_cfg3_n8_reference_arm --impl reference --steps 3 --warmup 1
run 8 29622 r02_bench_cfg3_n8 --steps 5 --warmup 3
run 4 29623 r02_bench_cfg3_n4 --steps 5 --warmup 3
run 8 29624 r02_bench_cfg5_n8 --workload cfg5 --steps 5 --warmup 3 --no-python-ref
run 8 29625 r02_bench_cfg2_n8 --workload cfg2 --steps 5 --warmup 3 --no-python-ref --no-eigen
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29630 + n)) tools/exp_pcie.py 2>/dev/null | grep "^{" 
done > gpurun_out/r02_pcie_ranks.jsonl
cat gpurun_out/r02_pcie_ranks.jsonl
