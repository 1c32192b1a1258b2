"""Dev/profiling driver: BASELINE config 4 on the config-3 fan -- brackets for 64 receiver depths, then a few rounds of
regula-falsi trial shots (for ncu on prt_refine_round_kernel).  Prints timings."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--max-iter", type=int, default=2)
ap.add_argument("--ztol", type=float, default=1e-3)
a = ap.parse_args()
w = bench.workload("cfg3")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=0)
ang = np.linspace(-15, 15, a.n)
res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 1000e3, angles=ang, to_host=False)
keep = torch.nonzero(res["status"] == 0).squeeze(1)
z_end = -res["final"][1].index_select(0, keep)
th = torch.from_numpy(-ang).cuda().index_select(0, keep)
rds = np.linspace(200.0, 4000.0, 64)
torch.cuda.synchronize(); t0 = time.perf_counter()
br, counts = pr.find_brackets(z_end, th, rds)
torch.cuda.synchronize(); t1 = time.perf_counter()
ref = pr.refine_brackets(denv, br, 1000.0, 0.0, 1000e3, ztol=a.ztol, max_iter=a.max_iter)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"brackets {int(counts.sum())} in {1e3 * (t1 - t0):.2f} ms; {a.max_iter + 2} rounds in {1e3 * (t2 - t1):.1f} ms; "
      f"shots {int(ref['shots'].sum().item())}; converged {int((ref['status'] == 0).sum().item())}")
