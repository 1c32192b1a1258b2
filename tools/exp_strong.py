"""Strong scaling of BASELINE config 3 (one 1,048,576-ray fan, 1000 km, sharded over the ranks) and the eigenray
table built on the gathered arrival table.  Launch with torchrun; rank 0 prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pygenray_b200 as pr
from pygenray_b200 import dist as pdist
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = bench.workload("cfg3")
env = bench.make_env(pr, w["arrays"], False)
denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False), device=lr)
angles = np.linspace(-15.0, 15.0, 1 << 20)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
ms = []
for i in range(5):
    barrier(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    table, local = pdist.shoot_fan_sharded(denv, 1000.0, 0.0, 1000e3, angles, rank, world)
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if i >= 2: ms.append(float(t.item()))
ok = int((table[:, 5] == 0).sum().item())
if rank == 0:
    m = float(np.mean(ms))
    print(json.dumps({"workload": "cfg3 strong scaling: one 1,048,576-ray fan to 1000 km, launch angles in cyclic blocks of 32 over the ranks, arrival table all-gathered",
                      "n_gpus": world, "ms_per_fan": m, "ray_km_per_s": ok * 1000.0 / (m * 1e-3), "rays_ok": ok}))
if world > 1:
    dist.destroy_process_group()
