#!/bin/bash
# GPU session 23: rays per warp chosen to fill whole rounds of resident warps (strong-scaling shares: 131 072 rays = 1.73
# rounds of 2368 warps at 32 rays per warp, 262 144 = 3.46 rounds)
mkdir -p gpurun_out
for rep in 1 2; do
for ln in default 31 30 28 26 24 16; do
  if [ $ln = default ]; then unset PRT_LANES; else export PRT_LANES=$ln; fi
  for n in 131072 262144; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 4 2>&1 | tail -1 | sed "s/^/lanes=$ln /"; done
done
done > gpurun_out/r02_variants_p.log 2>&1
cat gpurun_out/r02_variants_p.log
