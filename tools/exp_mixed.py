"""Dev: the bench fan as several concurrent launches with different rays-per-warp for steep and shallow angles."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
from pygenray_b200 import _lib
import bench
L = _lib.require_gpu()
w = bench.workload("cfg2")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=0)
ang = np.linspace(-15, 15, 10000)
n_save = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
streams = [torch.cuda.Stream() for _ in range(6)]
def run(parts):
    """parts: list of (mask, lanes)"""
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream(); e0.record()
        keep = []
        L.prt_set_concurrency(1)
        for k, (m, lanes) in enumerate(parts):
            st = streams[k]; st.wait_stream(main)
            L.prt_set_lanes(lanes)
            with torch.cuda.stream(st):
                keep.append(pr.shoot_fan_arrays(denv, 1000.0, 0.0, 500e3, angles=ang[m], num_range_save=n_save, to_host=False))
        for k in range(len(parts)):
            main.wait_stream(streams[k])
        L.prt_set_lanes(0)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
a = np.abs(ang)
print("uniform 8:", run([(a >= 0, 8)]))
for th in (11.0, 12.0, 13.0):
    for ls, lm in ((2, 8), (4, 8), (2, 16), (1, 8)):
        print(f"steep>|{th}| lanes {ls}, rest lanes {lm}:", run([(a > th, ls), (a <= th, lm)]))
for t1, t2 in ((11.0, 13.5), (10.0, 13.0)):
    print(f"three classes >{t2}: 2, >{t1}: 4, rest 8:", run([(a > t2, 2), ((a > t1) & (a <= t2), 4), (a <= t1, 8)]))
    print(f"three classes >{t2}: 1, >{t1}: 2, rest 8:", run([(a > t2, 1), ((a > t1) & (a <= t2), 2), (a <= t1, 8)]))
print("single part on a side stream, lanes 8:", run([(a >= 0, 8)]))
print("two halves (neg, pos) lanes 8:", run([(ang < 0, 8), (ang >= 0, 8)]))
print("steep only lanes 2 (a>12):", run([(a > 12, 2)]))
print("steep only lanes 8 (a>12):", run([(a > 12, 8)]))
print("rest only lanes 8 (a<=12):", run([(a <= 12, 8)]))
