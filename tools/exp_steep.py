"""Dev: kernel time of the steepest launch angles of the bench fan alone, for different rays-per-warp (PRT_LANES)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench
lo, hi, n, save = float(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
w = bench.workload("cfg2")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=0)
ang = np.linspace(lo, hi, n)
best = 1e9
for _ in range(5):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 500e3, angles=ang, num_range_save=save, counters=True, to_host=False)
    e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
c = res["counters"].cpu().numpy()
print(f"lanes={os.environ.get('PRT_LANES','auto')} [{lo},{hi}] n={n} save={save}: {best:.2f} ms  attempts max {c[1].max()} bisect max {c[4].max()}")
