"""Dev: staging time of a tomography ensemble, host route vs device route (config 5: 128 x 200 x 500 grids)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pygenray_b200 as pr
import bench
w = bench.workload("cfg5")
env = bench.make_env(pr, w["arrays"], False)
arrs = pr._unpack_envi(env, flatearth=False)
stack = w["arrays"]["c_stack"]
stack_d = torch.from_numpy(stack).cuda(); torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); cp = np.gradient(stack, arrs[3], axis=2); t1 = time.perf_counter()
    h = pr.DeviceEnsemble(stack, cp, *arrs[2:]); torch.cuda.synchronize(); t2 = time.perf_counter()
    d = pr.DeviceEnsemble.from_device(stack_d, *arrs[2:]); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"rep {rep}: np.gradient {1e3*(t1-t0):.1f} ms, host staging {1e3*(t2-t1):.1f} ms, device route {1e3*(t3-t2):.1f} ms")
    h.close(); d.close()
