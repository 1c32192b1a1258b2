"""Dev/profiling driver: run the cfg2 fan kernel a few times (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--save", type=int, default=1001)
ap.add_argument("--lim", type=float, default=15.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--counters", type=int, default=0)
a = ap.parse_args()
w = bench.workload(a.workload)
env = bench.make_env(pr, w["arrays"], w["flatearth"])
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=w["flatearth"]), device=0)
ang = np.linspace(-a.lim, a.lim, a.n)
best = 1e9
for _ in range(a.reps):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = pr.shoot_fan_arrays(denv, w["source_depth"], 0.0, w["receiver_range"], angles=ang, num_range_save=a.save, counters=bool(a.counters), to_host=False)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
    print(f"n={a.n} save={a.save}: {e0.elapsed_time(e1):.2f} ms")
print(f"{a.workload} n={a.n} save={a.save} lim={a.lim}: best of {a.reps} = {best:.2f} ms")
