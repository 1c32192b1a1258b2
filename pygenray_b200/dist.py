"""Multi-GPU fan shooting: one process per GPU (torchrun), launch angles sharded, arrival tables gathered.

The reference's only parallelism is a multiprocessing pool over launch angles
(/root/reference/src/pygenray/launch_rays.py:133-164); rays never exchange data, so the fan shards
with no data-path collective.  Launch angles are dealt to ranks in cyclic blocks of 32 (one warp of
neighbouring angles stays together, cost per ray grows with |angle| so contiguous chunks would
imbalance), every rank stages its own copy of the environment (<= tens of MB), and the only
collective is ONE all-gather of the fixed-size arrival records (T, z, p, n_bottom, n_surface, status:
6 doubles per ray) over NCCL/NVLink.  Trajectories stay sharded on the GPU that produced them.
Eigenray bracketing needs neighbouring rays in angle order, so it runs on the gathered table.
"""
import numpy as np

__all__ = ["shard_indices", "all_gather_arrivals", "shoot_fan_sharded", "bind_to_gpu_numa", "RECORD_WIDTH"]

BLOCK = 32
RECORD_WIDTH = 6  # T, z, p, n_bottom, n_surface, status


_SHARD_CACHE = {}


def shard_indices(n_rays, rank, world, block=BLOCK):
    """Indices of the rays owned by ``rank``: blocks of ``block`` consecutive angles dealt round-robin.
    (Cached per fan shape: for a 1M-ray fan the index arithmetic costs more host time than a rank's kernel.)"""
    key = (int(n_rays), int(rank), int(world), int(block))
    mine = _SHARD_CACHE.get(key)
    if mine is None:
        idx = np.arange(n_rays, dtype=np.int64)
        mine = idx[(idx // block) % world == rank]
        mine.setflags(write=False)
        if len(_SHARD_CACHE) >= 64:
            _SHARD_CACHE.clear()
        _SHARD_CACHE[key] = mine
    return mine


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU).

    Pinned host buffers are placed by first touch, so a rank that floats over both sockets ends up copying its
    trajectories (240 MB per 10k-ray fan) across the inter-socket link; with eight ranks doing so at once the
    end-to-end rate of a node is the host's, not the GPUs'.  Returns the NUMA node, or None when the topology is
    not visible (containers without /sys, single-socket hosts) -- then nothing is changed."""
    import os  # noqa: PLC0415

    try:
        import torch  # noqa: PLC0415

        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001  (best effort: never fail a run over placement)
        return None


def shard_sizes(n_rays, world, block=BLOCK):
    return [int(shard_indices(n_rays, r, world, block).shape[0]) for r in range(world)]


_GATHER_INDEX_CACHE = {}


def _gather_index(n_rays, world, block, device):
    """Row of the padded all-gather output that holds global ray i (cached per fan shape and device)."""
    key = (n_rays, world, block, str(device))
    idx = _GATHER_INDEX_CACHE.get(key)
    if idx is None:
        import torch  # noqa: PLC0415

        sizes = shard_sizes(n_rays, world, block)
        pad = max(sizes)
        src = np.empty(n_rays, dtype=np.int64)
        for r in range(world):
            mine = shard_indices(n_rays, r, world, block)
            src[mine] = r * pad + np.arange(mine.shape[0])
        idx = (torch.from_numpy(src).to(device), sizes, pad)
        _GATHER_INDEX_CACHE[key] = idx
    return idx


def all_gather_arrivals(local_records, n_rays, rank, world, block=BLOCK, group=None):
    """All-gather the per-rank arrival records and put them back in launch-angle order.

    ``local_records``: float64 tensor [n_local, RECORD_WIDTH] on the rank's device (CUDA with NCCL, CPU
    with gloo).  Returns a [n_rays, RECORD_WIDTH] tensor on the same device, identical on every rank.
    One collective + one gather by a cached index (no per-rank Python work in the step).
    """
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    if world == 1:
        assert local_records.shape[0] == n_rays
        return local_records
    src, sizes, pad = _gather_index(n_rays, world, block, local_records.device)
    assert local_records.shape[0] == sizes[rank], "local shard has the wrong number of rays"
    if sizes[rank] == pad:
        buf = local_records.contiguous()
    else:
        buf = torch.zeros((pad, RECORD_WIDTH), dtype=torch.float64, device=local_records.device)
        buf[: sizes[rank]] = local_records
    gathered = torch.empty((world * pad, RECORD_WIDTH), dtype=torch.float64, device=local_records.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    return gathered.index_select(0, src)


def records_from_result(res):
    """[n, 6] float64 arrival records from a shoot_fan_arrays(to_host=False) result (device tensors)."""
    import torch  # noqa: PLC0415

    final = res["final"]
    n = final.shape[1]
    rec = torch.empty((n, RECORD_WIDTH), dtype=torch.float64, device=final.device)
    if final.is_cuda:
        import ctypes as C  # noqa: PLC0415

        from . import _lib  # noqa: PLC0415

        with torch.cuda.device(final.device):
            rc = _lib.load().prt_pack_arrivals(
                n, C.c_void_p(final.data_ptr()), C.c_void_p(res["n_bottom"].data_ptr()),
                C.c_void_p(res["n_surface"].data_ptr()), C.c_void_p(res["status"].data_ptr()),
                C.c_void_p(rec.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "prt_pack_arrivals")
        return rec
    rec[:, :3] = final.t()
    rec[:, 3] = res["n_bottom"]
    rec[:, 4] = res["n_surface"]
    rec[:, 5] = res["status"]
    return rec


def shoot_fan_sharded(denv, source_depth, source_range, receiver_range, angles, rank, world, **kwargs):
    """Shoot this rank's share of ``angles`` (internal angles, full fan on every rank) and all-gather.

    Returns ``(table, local)``: the gathered [n_rays, 6] arrival table (device tensor, same on all ranks)
    and this rank's own shoot_fan_arrays result (trajectories stay local).  ``kwargs`` go to
    :func:`pygenray_b200.shoot_fan_arrays`.
    """
    from .launch_rays import shoot_fan_arrays  # noqa: PLC0415

    angles = np.ascontiguousarray(angles, dtype=np.float64)
    mine = shard_indices(angles.shape[0], rank, world)
    kwargs["to_host"] = False
    local = shoot_fan_arrays(denv, source_depth, source_range, receiver_range, angles=angles[mine], **kwargs)
    table = all_gather_arrivals(records_from_result(local), angles.shape[0], rank, world)
    return table, local
