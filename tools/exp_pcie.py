"""Dev: pinned D2H bandwidth per rank with 1..N ranks copying at once (what bounds the end-to-end TRAJECTORY output of
shoot_rays at N GPUs: 241 MB per 10k-ray fan per rank).  Launch with torchrun; rank 0 prints one JSON line."""
import json, os, time
import torch, torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 240 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
out = {}
for name, (src, dst) in (("d2h", (d, h)), ("h2d", (h, d))):
    best = 0.0
    for _ in range(6):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = max(best, n / float(dt.item()) / 1e9)
    out[name + "_GBps_per_rank"] = best
    out[name + "_GBps_aggregate"] = best * world
if rank == 0:
    print(json.dumps({"ranks_copying_at_once": world, "bytes_per_rank": n, **out, "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
