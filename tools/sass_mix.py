#!/usr/bin/env python
"""dev: instruction mix of the largest straight-line block (the RK45 trial step) of a kernel.
usage: sass_mix.py lib.so <substring of kernel name>"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = [t for _, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)]
    blocks, cur = [], []
    for t in ins:
        cur.append(t)
        if re.search(r"\b(BRA|BSSY|BSYNC|EXIT|CALL|RET|WARPSYNC|BREAK)\b", t):
            blocks.append(cur)
            cur = []
    big = max(blocks, key=len)
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in big)
    fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(name[:60], "| largest block:", len(big), "instr, FP64-pipe", fp64)
    print("  ", ", ".join(f"{k} {v}" for k, v in mix.most_common(18)))
