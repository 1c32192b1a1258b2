"""Time the UNMODIFIED Python reference (oracle/_ref/pygenray, see install_ref.py) on this host: BASELINE config 1
in full -- the default OceanEnvironment2D, source at 1000 m, 101 launch angles +-12 deg, 100 km, rtol 1e-9
(/root/reference/src/pygenray/launch_rays.py:11-200, 325-484).  Prints one JSON line.  bench.py runs this file in a
subprocess (PYTHONPATH = oracle/stubs : oracle/_ref) for its `cpu_baseline.python_reference` entry.

  sequential   _shoot_ray_array + _interpolate_ray ray by ray in this process, Numba JIT warmed first: the per-core rate
  pool         shoot_rays(n_processes = os.cpu_count()) wall time including the spawn of the process pool -- what a
               user of the reference actually waits for
"""
import json
import os
import sys
import time

import numpy as np

import pygenray as pr
from pygenray.environment import OceanEnvironment2D
from pygenray.launch_rays import _interpolate_ray, _shoot_ray_array, _unpack_envi


def main():
    n_save = 1001
    env = OceanEnvironment2D()
    cin, cpin, rin, zin, depths, dr, ba = _unpack_envi(env, flatearth=True)
    angles = np.linspace(-12, 12, 101)
    c0 = pr.bilinear_interp(0.0, 1000.0, rin, zin, cin)
    rs = np.linspace(0.0, 100e3, n_save)

    def one(a):
        y0 = np.array([0, 1000.0, np.sin(np.radians(a)) / c0])
        sols, full, b, s = _shoot_ray_array(y0, 1000.0, 0.0, 100e3, cin, cpin, rin, zin, depths, dr, ba, 1e-9, True, False)
        return None if sols is None else _interpolate_ray(sols, rs)

    one(3.0)  # JIT warm-up (compiles / loads the cached machine code)
    t0 = time.perf_counter()
    ok = sum(one(-a) is not None for a in angles)
    t_seq = time.perf_counter() - t0
    out = {"config": "cfg1: default OceanEnvironment2D (flat earth), 101 rays +-12 deg, 100 km, 1001 save points",
           "rays_ok": int(ok), "sequential_s": t_seq, "sequential_ray_km_per_s_per_core": ok * 100.0 / t_seq,
           "cores": os.cpu_count()}
    if "--no-pool" not in sys.argv:
        t0 = time.perf_counter()
        fan = pr.shoot_rays(1000.0, 0.0, angles, 100e3, n_save, env, rtol=1e-9, n_processes=os.cpu_count(), debug=False)
        t_pool = time.perf_counter() - t0
        out.update(pool_s=t_pool, pool_rays_ok=int(len(fan.thetas)), pool_ray_km_per_s=len(fan.thetas) * 100.0 / t_pool)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
