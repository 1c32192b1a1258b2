"""Dev experiment: fan kernel time vs fan size / angle range (arrival table only)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import workloads as cases
import bench

def run(denv, angles, R, n_save=0, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, R, angles=angles, num_range_save=n_save, counters=True, to_host=False)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cnt = res["counters"].cpu().numpy().astype(np.int64)
    ok = int((res["status"] == 0).sum().item())
    return best, cnt, ok

w = bench.workload("cfg2")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=0)
for lim, n in [(10, 10000), (15, 10000), (10, 100000), (15, 100000), (10, 1000000), (15, 1000000)]:
    for ns in (0, 1001) if n <= 10000 else (0,):
        ang = np.linspace(-lim, lim, n)
        ms, cnt, ok = run(denv, ang, 500e3, ns)
        att = cnt[1]; bis = cnt[4]
        print(f"cfg2 +-{lim} n={n} n_save={ns}: {ms:9.2f} ms  {ok*500/ms*1e3/1e6:9.1f} M ray-km/s  attempts mean {att.mean():.0f} max {att.max()}  bisect mean {bis.mean():.0f} max {bis.max()}  flops {bench.algorithmic_flops(cnt)/ms/1e9:.3f} TF/s", flush=True)
