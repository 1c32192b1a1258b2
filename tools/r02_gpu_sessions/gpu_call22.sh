#!/bin/bash
# GPU session 22: eigenray refinement with its exact field lookups through one out-of-line copy (kernel code 92 -> 79 KB)
mkdir -p gpurun_out
for rep in 1 2; do
  for v in outline inline; do
    if [ $v = outline ]; then unset PRT_LIB_PATH; else export PRT_LIB_PATH=$PWD/build/variants/libprt_refine_inline.so; fi
    timeout 300 python tools/prof_refine.py --max-iter 2 2>&1 | tail -1 | sed "s/^/$v /"
    timeout 300 python tools/prof_refine.py --max-iter 20 2>&1 | tail -1 | sed "s/^/$v /"
    timeout 300 python tools/prof_refine.py --max-iter 20 --ztol 1.0 2>&1 | tail -1 | sed "s/^/$v ztol1 /"
  done
done > gpurun_out/r02_variants_o.log 2>&1
cat gpurun_out/r02_variants_o.log
unset PRT_LIB_PATH
timeout 900 python -m pytest tests -m gpu -q -k "refine or eigen or cfg4 or config4 or smoke" 2>&1 | tail -2
