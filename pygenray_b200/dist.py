"""Multi-GPU fan shooting and eigenray search: one process per GPU (torchrun), launch angles / eigenray brackets /
ensemble realisations sharded, fixed-width result records gathered.

The reference's only parallelism is a multiprocessing pool over launch angles
(/root/reference/src/pygenray/launch_rays.py:133-164); rays never exchange data, so the fan shards
with no data-path collective.  Launch angles are dealt to ranks in cyclic blocks of 32 (one warp of
neighbouring angles stays together, cost per ray grows with |angle| so contiguous chunks would
imbalance), every rank stages its own copy of the environment (<= tens of MB), and the only
collective is ONE all-gather of the fixed-size arrival records (T, z, p, n_bottom, n_surface, status:
6 doubles per ray) over NCCL/NVLink.  Trajectories stay sharded on the GPU that produced them.
Eigenray bracketing needs neighbouring rays in angle order, so it runs on the gathered table -- on every rank, it
is a sub-millisecond kernel and deterministic, so all ranks hold the same bracket list without a broadcast --
and the brackets (receiver x ray pair), whose regula-falsi trial shots are the expensive part, are dealt to the
ranks in cyclic blocks of 32 (the reference's only eigenray parallelism is a process pool over brackets,
/root/reference/src/pygenray/eigenrays.py:122-164); ONE all-gather of 10-double eigenray records returns the
table to every rank.  Tomography ensembles shard by realisation (contiguous blocks; every realisation costs the
same), again with one all-gather of the arrival records.
"""
import numpy as np

__all__ = ["shard_indices", "all_gather_rows", "all_gather_arrivals", "shoot_fan_sharded", "find_eigenrays_sharded",
           "shard_realisations", "shoot_ensemble_sharded", "bind_to_gpu_numa", "RECORD_WIDTH", "EIGEN_WIDTH",
           "EIGEN_FIELDS"]

BLOCK = 32
RECORD_WIDTH = 6  # T, z, p, n_bottom, n_surface, status
EIGEN_FIELDS = ("rcv", "k", "theta", "T", "z", "p", "n_bottom", "n_surface", "status", "shots")
EIGEN_WIDTH = len(EIGEN_FIELDS)


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
        if len(_GATHER_INDEX_CACHE) >= 16:  # bracket counts differ from fan to fan: keep the device copies bounded
            _GATHER_INDEX_CACHE.clear()
        _GATHER_INDEX_CACHE[key] = idx
    return idx


def all_gather_rows(local_rows, n_total, rank, world, block=BLOCK, group=None):
    """All-gather fixed-width rows that were dealt to the ranks by :func:`shard_indices` and put them back in order.

    ``local_rows``: float64 tensor [n_local, width] on the rank's device (CUDA with NCCL, CPU with gloo).  Returns a
    [n_total, width] tensor on the same device, identical on every rank.  One collective + one gather by a cached
    index (no per-rank Python work in the step).
    """
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    if local_rows.ndim != 2:
        raise ValueError("all_gather_rows: local_rows must be [n_local, width]")
    if n_total == 0:  # nothing to exchange (e.g. a fan without any eigenray bracket): the same on every rank, no collective
        return local_rows[:0]
    if world == 1:
        if local_rows.shape[0] != n_total:
            raise ValueError(f"all_gather_rows: {local_rows.shape[0]} local rows for {n_total} in total at world size 1")
        return local_rows
    width = local_rows.shape[1]
    src, sizes, pad = _gather_index(n_total, world, block, local_rows.device)
    if local_rows.shape[0] != sizes[rank]:
        raise ValueError(f"all_gather_rows: rank {rank} holds {local_rows.shape[0]} rows, its shard has {sizes[rank]}")
    if sizes[rank] == pad:
        buf = local_rows.contiguous()
    else:
        buf = torch.zeros((pad, width), dtype=local_rows.dtype, device=local_rows.device)
        buf[: sizes[rank]] = local_rows
    gathered = torch.empty((world * pad, width), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    return gathered.index_select(0, src)


def all_gather_arrivals(local_records, n_rays, rank, world, block=BLOCK, group=None):
    """All-gather the per-rank arrival records [n_local, RECORD_WIDTH] and put them back in launch-angle order."""
    if local_records.shape[1] != RECORD_WIDTH:
        raise ValueError(f"all_gather_arrivals: records must be [n, {RECORD_WIDTH}]")
    return all_gather_rows(local_records, n_rays, rank, world, block=block, group=group)


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


def shoot_fan_sharded(denv, source_depth, source_range, receiver_range, angles, rank, world, group=None, **kwargs):
    """Shoot this rank's share of ``angles`` (internal angles, full fan on every rank) and all-gather.

    Returns ``(table, local)``: the gathered [n_rays, 6] arrival table (device tensor, same on all ranks)
    and this rank's own shoot_fan_arrays result (trajectories stay local).  ``kwargs`` go to
    :func:`pygenray_b200.shoot_fan_arrays` (``to_host`` / ``compact`` are fixed: the table needs every ray).
    """
    from .launch_rays import shoot_fan_arrays  # noqa: PLC0415

    for fixed in ("to_host", "compact"):
        if kwargs.pop(fixed, None):
            raise ValueError(f"shoot_fan_sharded: {fixed} is not available here (records of ALL rays are gathered)")
    angles = np.ascontiguousarray(angles, dtype=np.float64)
    mine = shard_indices(angles.shape[0], rank, world)
    local = shoot_fan_arrays(denv, source_depth, source_range, receiver_range, angles=angles[mine], to_host=False,
                             **kwargs)
    table = all_gather_arrivals(records_from_result(local), angles.shape[0], rank, world, group=group)
    return table, local


# ---------------------------------------------------------------------------------------------
# eigenray search over ranks
def _gpu_bracket(z_end, thetas, receiver_depths):
    from .eigenrays import find_brackets  # noqa: PLC0415

    return find_brackets(z_end, thetas, receiver_depths, device=z_end.device.index)


def _gpu_refine(denv, br, source_depth, source_range, receiver_range, **kw):
    from .eigenrays import refine_brackets  # noqa: PLC0415

    return refine_brackets(denv, br, source_depth, source_range, receiver_range, **kw)


def find_eigenrays_sharded(denv, table, user_thetas, receiver_depths, source_depth, source_range, receiver_range,
                           rank, world, ztol=1.0, max_iter=20, rtol=1e-9, terminate_backwards=True, group=None,
                           _bracket=_gpu_bracket, _refine=_gpu_refine):
    """Eigenray table of a fan whose arrival ``table`` ([n_rays, 6], device tensor, identical on every rank --
    :func:`shoot_fan_sharded`) has been gathered.  ``user_thetas`` are the fan's launch angles in the stored
    (user) convention, one per table row.

    Every rank compacts the table to surviving rays (``RayFan`` only holds those, eigenrays.py:62 works on them),
    brackets it for all receivers, refines its cyclic blocks of 32 brackets (dealt in launch-angle order) and
    contributes their records to one all-gather.  Returns ``(records, counts)``: ``records`` [n_brackets, EIGEN_WIDTH] float64 on the device, in
    (receiver, ray) order and identical on every rank, columns ``EIGEN_FIELDS`` (``k`` indexes the COMPACTED
    fan, ``status`` is PRT_EIG_*; theta, T, z, p are NaN unless converged); ``counts`` [n_rcv] brackets per
    receiver.  ``_bracket`` / ``_refine`` are test hooks (gloo runs have no GPU).
    """
    import torch  # noqa: PLC0415

    keep = torch.nonzero(table[:, 5] == 0).squeeze(1)
    z_end = -table[:, 1].index_select(0, keep)  # stored convention: zs = -z
    th = torch.as_tensor(np.ascontiguousarray(user_thetas, dtype=np.float64)).to(table.device).index_select(0, keep)
    br, counts = _bracket(z_end, th, receiver_depths)
    n_br = int(br["z1"].shape[0])
    # deal in launch-angle order: a block of 32 then holds 32 nearly identical trial rays (one warp of the refinement)
    perm = torch.argsort(br["th1"], stable=True)
    mine = perm.index_select(0, torch.from_numpy(shard_indices(n_br, rank, world).copy()).to(table.device))
    local = {k: v.index_select(0, mine).contiguous() for k, v in br.items()}
    ref = _refine(denv, local, source_depth, source_range, receiver_range, rtol=rtol,
                  terminate_backwards=terminate_backwards, ztol=ztol, max_iter=max_iter)
    rec = torch.empty((mine.shape[0], EIGEN_WIDTH), dtype=torch.float64, device=table.device)
    rec[:, 0] = local["rcv"]
    rec[:, 1] = local["k"]
    rec[:, 2] = ref["theta"]
    rec[:, 3:6] = ref["final"].t()
    rec[:, 6] = ref["n_bottom"]
    rec[:, 7] = ref["n_surface"]
    rec[:, 8] = ref["status"]
    rec[:, 9] = ref["shots"]
    gathered = all_gather_rows(rec, n_br, rank, world, group=group)  # rows in launch-angle order
    out = torch.empty_like(gathered)
    out[perm] = gathered
    return out, counts


# ---------------------------------------------------------------------------------------------
# tomography ensembles over ranks (BASELINE config 5)
def shard_realisations(n_env_total, rank, world):
    """``(first, count)`` of the contiguous block of realisations owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(int(n_env_total), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _gpu_shoot_ensemble(dens, source_depth, source_range, receiver_range, angles, **kwargs):
    from .launch_rays import shoot_ensemble_arrays  # noqa: PLC0415

    return shoot_ensemble_arrays(dens, source_depth, source_range, receiver_range, angles, to_host=False, **kwargs)


def shoot_ensemble_sharded(dens, n_env_total, source_depth, source_range, receiver_range, angles, rank, world,
                           group=None, _shoot=_gpu_shoot_ensemble, **kwargs):
    """Shoot ``angles`` through this rank's realisations (``dens``: a DeviceEnsemble staged with the realisations
    ``shard_realisations(n_env_total, rank, world)``) and all-gather the arrival records.

    Returns ``(table, local)``: ``table`` [n_env_total, n_angles, 6] on the device, identical on every rank, and
    this rank's shoot_ensemble_arrays result (device tensors).  No exchange besides the one all-gather.
    ``_shoot`` is a test hook (gloo runs have no GPU)."""
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    first, count = shard_realisations(n_env_total, rank, world)
    if dens.n_env != count:
        raise ValueError(f"shoot_ensemble_sharded: rank {rank} staged {dens.n_env} realisations, its shard has {count}")
    kwargs.pop("to_host", None)
    local = _shoot(dens, source_depth, source_range, receiver_range, angles, **kwargs)
    k = local["final"].shape[2]
    flat = dict(final=local["final"].reshape(3, count * k), n_bottom=local["n_bottom"].reshape(-1),
                n_surface=local["n_surface"].reshape(-1), status=local["status"].reshape(-1))
    rec = records_from_result(flat)
    if world == 1:
        return rec.view(count, k, RECORD_WIDTH), local
    pad = -(-int(n_env_total) // world)  # the largest shard
    buf = rec
    if count < pad:
        buf = torch.zeros((pad * k, RECORD_WIDTH), dtype=torch.float64, device=rec.device)
        buf[: count * k] = rec
    gathered = torch.empty((world * pad * k, RECORD_WIDTH), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    gathered = gathered.view(world, pad, k, RECORD_WIDTH)
    parts = [gathered[r, : shard_realisations(n_env_total, r, world)[1]] for r in range(world)]
    return torch.cat(parts, 0), local
