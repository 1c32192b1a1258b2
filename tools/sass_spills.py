#!/usr/bin/env python
"""dev: where do the spill loads/stores of a kernel sit?  usage: sass_spills.py lib.so <substring of kernel name>
Prints, per basic-block-ish region between branches, the number of FP64 instructions and of STL/LDL."""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)
    print(name, len(ins), "instructions")
    blk_start, fp, st, ld, n = 0, 0, 0, 0, 0
    for k, (addr, t) in enumerate(ins):
        n += 1
        if re.search(r"\bD(FMA|MUL|ADD|SETP)\b", t):
            fp += 1
        if "STL" in t:
            st += 1
        if "LDL" in t:
            ld += 1
        if re.search(r"\b(BRA|BSSY|BSYNC|EXIT|CALL|RET|WARPSYNC|BREAK)\b", t):
            if n >= 40 or st or ld:
                print(f"  [{ins[blk_start][0]}..{addr}] {n:5d} instr, {fp:4d} fp64, {st:3d} STL, {ld:3d} LDL   ends: {t[:60]}")
            blk_start, fp, st, ld, n = k + 1, 0, 0, 0, 0
