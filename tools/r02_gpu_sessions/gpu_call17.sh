#!/bin/bash
# GPU session 17: park_wait (iterations a lane that hit a boundary waits for its neighbours) on config 3 and on K2
mkdir -p gpurun_out
for pw in 16 4 8 32 64 0; do
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_refine.py 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
done > gpurun_out/r02_variants_k.log 2>&1
cat gpurun_out/r02_variants_k.log
