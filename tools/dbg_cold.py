"""Dev: how often the exact (cold) trial step runs on the device (library built with -DPRT_COUNT_COLD)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
from pygenray_b200 import _lib
import bench
for wl, n, lim in (("cfg2", 10000, 10.0), ("cfg1", 101, 12.0), ("cfg3", 8192, 15.0)):
    w = bench.workload(wl)
    env = bench.make_env(pr, w["arrays"], w["flatearth"])
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=w["flatearth"]), device=0)
    L = _lib.load(); L.prt_debug_cold.restype = C.c_longlong
    c0 = L.prt_debug_cold()
    res = pr.shoot_fan_arrays(denv, w["source_depth"], 0.0, w["receiver_range"], angles=np.linspace(-lim, lim, n), counters=True)
    torch.cuda.synchronize()
    print(wl, "attempts", int(res["counters"][1].sum()), "cold", L.prt_debug_cold() - c0, flush=True)
