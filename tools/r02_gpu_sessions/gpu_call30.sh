#!/bin/bash
# GPU session 30: adaptive wait of the fan kernel -- threshold (quarters of the lanes in play that took part in the last
# root-search pass) x long wait, on config 5 (where the shipped 2/4 x 128 costs 2.5 %), the steep fan (where it gains
# 32 %), config 2 with save points and the config-3 shares
mkdir -p gpurun_out
run() { lw=$1; num=$2
  export PRT_PARK_WAIT_LONG=$lw PRT_PARK_LONG_NUM=$num
  c5=$(timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-python-ref --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['roofline']['kernel_ms'],2))")
  st=$(timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 0 --lim 40 --reps 4 2>&1 | tail -1 | sed 's/.*= //')
  c2=$(timeout 300 python tools/prof_fan.py --workload cfg2 --n 10000 --save 1001 --reps 4 2>&1 | tail -1 | sed 's/.*= //')
  c3=$(timeout 300 python tools/prof_fan.py --workload cfg3 --n 131072 --save 0 --reps 3 2>&1 | tail -1 | sed 's/.*= //')
  echo "long=$lw num=$num/4 | cfg5 kernel $c5 ms | steep +-40deg $st | cfg2 10k+save $c2 | cfg3 131072 $c3"
}
{ run 16 2; run 128 2; run 128 3; run 128 4; run 64 3; run 64 4; run 32 2; } > gpurun_out/r02_variants_q.log 2>&1
cat gpurun_out/r02_variants_q.log
