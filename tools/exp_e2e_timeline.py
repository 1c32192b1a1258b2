"""Dev: timeline of the pipelined shoot_rays (kernel end / copy end per chunk, ms since call start)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
from pygenray_b200 import launch_rays as lr
import bench
w = bench.workload("cfg2")
env = bench.make_env(pr, w["arrays"], False)
ang = -w["angles"]
for i in range(6):
    fan = pr.shoot_rays(1000.0, 0.0, ang, 500e3, 1001, env, rtol=1e-9, debug=False, flatearth=False)
# monkeypatch: wrap torch.nonzero and Tensor.cpu to log host times
t00 = [0.0]; log = []
_nz = torch.nonzero
def nz(x):
    r = _nz(x); torch.cuda.current_stream().synchronize(); log.append(("kernel done", 1e3 * (time.perf_counter() - t00[0]))); return r
torch.nonzero = nz
_cpu = torch.Tensor.cpu
def cpu(self, *a, **k):
    r = _cpu(self, *a, **k); log.append(("copies done", 1e3 * (time.perf_counter() - t00[0]))); return r
torch.Tensor.cpu = cpu
for rep in range(2):
    log.clear(); torch.cuda.synchronize(); t00[0] = time.perf_counter()
    fan = pr.shoot_rays(1000.0, 0.0, ang, 500e3, 1001, env, rtol=1e-9, debug=False, flatearth=False)
    t_end = 1e3 * (time.perf_counter() - t00[0])
    print(" ".join(f"{n[0]}{t:.1f}" for n, t in log), f"| end {t_end:.1f} ms")
