"""Summarise an .ncu-rep (raw + source pages) into markdown; run here (no GPU needed)."""
import collections, csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))

def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[1], rows[2:]

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps/SM)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active (%)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (%)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction (of 32)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate (%)"), ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_bytes.sum.per_second", "L2 throughput"), ("l1tex__t_bytes.sum.per_second", "L1 throughput"),
    ("smsp__inst_executed.sum", "warp instructions"),
]

def main(rep, title):
    v, u = raw(rep)
    print(f"### {title}\n\n| metric | value |\n|---|---|")
    for k, name in KEYS:
        if k in v:
            print(f"| {name} (`{k}`) | {v[k]} {u[k]} |")
    hdr, data = source(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]].replace(",", ""))
        except Exception: return 0.0
    tot = sum(f(r, "# Samples") for r in data) or 1.0
    ex = collections.Counter()
    for r in data:
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        ex[op] += f(r, "Instructions Executed")
    te = sum(ex.values()) or 1.0
    print("\nInstruction mix (executed warp instructions): " + ", ".join(f"{op} {100*c/te:.1f}%" for op, c in ex.most_common(14)))
    st = {k: 100 * sum(f(r, k) for r in data) / tot for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    print("\nWarp-stall samples: " + ", ".join(f"{k[6:]} {p:.1f}%" for k, p in sorted(st.items(), key=lambda x: -x[1])[:9]) + "\n")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
