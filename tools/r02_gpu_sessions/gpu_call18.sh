#!/bin/bash
# GPU session 18: park_wait for the eigenray refinement, repeated and interleaved (full 22-round refinement)
mkdir -p gpurun_out
for pw in 16 64 16 64 32 128 256 16; do
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_refine.py --max-iter 20 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
done > gpurun_out/r02_variants_l.log 2>&1
for pw in 16 64 128; do
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_fan.py --workload cfg3 --n 131072 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
  PRT_PARK_WAIT=$pw timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/park_wait=$pw /"
done >> gpurun_out/r02_variants_l.log 2>&1
cat gpurun_out/r02_variants_l.log
