"""Dev: pinned D2H / H2D bandwidth of the box (what bounds the end-to-end trajectory output)."""
import torch, time
for mb in (8, 32, 240):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, (src, dst) in (("D2H", (d, h)), ("H2D", (h, d))):
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} {mb} MiB: {best:.2f} ms  {n/best/1e6:.1f} GB/s")
# two concurrent D2H streams
n = 120 << 20
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"2 streams x 120 MiB D2H: {1e3*dt:.2f} ms  {2*n/dt/1e9:.1f} GB/s")
