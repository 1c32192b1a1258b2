"""BASELINE config 3 + 4 end to end on one GPU: 1M-ray fan to 1000 km through the internal-wave grid, then the
eigenray table for 64 receiver depths (bracketing + regula falsi).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ztol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
w = bench.workload("cfg3")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=0)
ang = np.linspace(-15, 15, n_rays)          # internal angles; stored (user) angle = -internal
R = w["receiver_range"]
out = {}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, R, angles=ang, counters=True)
    torch.cuda.synchronize(); out["fan_s"] = time.perf_counter() - t0
ok = res["status"] == 0
out.update(n_rays=n_rays, rays_ok=int(ok.sum()), ray_km_per_s=float(ok.sum()) * R / 1e3 / out["fan_s"],
           bounces=int(res["n_bottom"][ok].sum() + res["n_surface"][ok].sum()),
           attempts_per_ray_km=float(res["counters"][1].astype(np.int64).sum()) / (float(ok.sum()) * R / 1e3))
zs_end = -res["final"][1][ok]
thetas = -ang[ok]
rds = np.linspace(200.0, 4000.0, 64)
torch.cuda.synchronize(); t0 = time.perf_counter()
br, counts = pr.find_brackets(zs_end, thetas, rds, device=0)
torch.cuda.synchronize(); out["bracket_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
ref = pr.refine_brackets(denv, br, 1000.0, 0.0, R, ztol=ztol, max_iter=20)
torch.cuda.synchronize(); out["refine_s"] = time.perf_counter() - t0
st = ref["status"].cpu().numpy()
out.update(ztol=ztol, brackets=int(counts.sum()), converged=int((st == 0).sum()), dropped=int((st == 1).sum()),
           max_iter=int((st == 2).sum()), shots_per_bracket=float(ref["shots"].float().mean().item()),
           eigenrays_per_s=float((st == 0).sum()) / (out["bracket_s"] + out["refine_s"]))
# arrival-time table: travel time of every converged eigenray, per receiver
T = ref["final"][0].cpu().numpy(); rcv = br["rcv"].cpu().numpy()
out["arrivals_per_receiver_min_max"] = [int(np.bincount(rcv[st == 0], minlength=64).min()),
                                        int(np.bincount(rcv[st == 0], minlength=64).max())]
out["first_arrival_s"] = float(np.nanmin(T[st == 0])) if (st == 0).any() else None
print(json.dumps(out))
