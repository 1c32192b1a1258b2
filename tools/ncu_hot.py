"""Dev: per-instruction view of an ncu source page: executed count, stall samples, dominant stall; and region totals."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in data); tot_e = sum(f(r, "Instructions Executed") for r in data)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
base = None
for r in data:
    a = int(r[ix["Address"]], 16)
    if base is None: base = a
    off = a - base
    if off < lo or off > hi: continue
    s = f(r, "# Samples"); e = f(r, "Instructions Executed")
    top = max(stalls, key=lambda k: f(r, k))
    print(f"{off:6x} {e/1e3:9.0f}k {100*s/tot_s:5.2f}% {top[6:]:>16s} {f(r, top)/max(s,1)*100:4.0f}%  {r[ix['Source']][:90]}")
print("total samples", tot_s, "executed", tot_e)
