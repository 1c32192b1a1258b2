#!/bin/bash
# dev: time kernel build variants under build/variants on the GPU box
for v in "$@"; do
  echo "== variant $v"
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so python tools/prof_fan.py --n 10000 --save 1001 --reps 4 2>&1 | tail -1
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so python tools/prof_fan.py --n 262144 --save 0 --reps 3 2>&1 | tail -1
  PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so python tools/prof_fan.py --n 1048576 --save 0 --reps 2 2>&1 | tail -1
done
