"""Dev: where the end-to-end time of shoot_rays goes (bench workload)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench
w = bench.workload("cfg2")
env = bench.make_env(pr, w["arrays"], False)
ang = -w["angles"]
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fan = pr.shoot_rays(1000.0, 0.0, ang, 500e3, 1001, env, rtol=1e-9, debug=False, flatearth=False)
    torch.cuda.synchronize(); print(f"e2e {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
pr_ = cProfile.Profile(); pr_.enable()
for i in range(5):
    fan = pr.shoot_rays(1000.0, 0.0, ang, 500e3, 1001, env, rtol=1e-9, debug=False, flatearth=False)
pr_.disable()
pstats.Stats(pr_).sort_stats("cumulative").print_stats(28)
