#!/bin/bash
# GPU session 20: adaptive wait of the fan kernel (long wait only for warps whose lanes reflect together)
mkdir -p gpurun_out
for lw in off 64 128; do
  if [ $lw = off ]; then unset PRT_PARK_WAIT_LONG; else export PRT_PARK_WAIT_LONG=$lw; fi
  for n in 1048576 131072; do timeout 300 python tools/prof_fan.py --workload cfg3 --n $n --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/long=$lw /"; done
  timeout 300 python tools/prof_fan.py --workload cfg2 --n 1048576 --save 0 --reps 3 2>&1 | tail -1 | sed "s/^/long=$lw /"
  timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed "s/^/long=$lw /"
  timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 0 --lim 40 --reps 4 2>&1 | tail -1 | sed "s/^/long=$lw steep /"
done > gpurun_out/r02_variants_m.log 2>&1
cat gpurun_out/r02_variants_m.log
unset PRT_PARK_WAIT_LONG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
