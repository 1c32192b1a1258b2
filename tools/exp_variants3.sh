#!/bin/bash
# dev: cfg3 + cfg2 timings of build variants (build/variants/libprt_<v>.so); "base" = the shipped library
for v in "$@"; do
  echo "== variant $v"
  if [ "$v" = base ]; then unset PRT_LIB_PATH; else export PRT_LIB_PATH=$PWD/build/variants/libprt_$v.so; fi
  python tools/prof_fan.py --workload cfg3 --n 262144 --save 0 --reps 4 | tail -1
  python tools/prof_fan.py --workload cfg3 --n 1048576 --save 0 --reps 2 | tail -1
  python tools/prof_fan.py --n 1048576 --save 0 --reps 3 | tail -1
  python tools/prof_fan.py --n 10000 --save 1001 --reps 4 | tail -1
done
