"""Eigenray search: host layer over libprt.so.

Same call and result object as the reference's ``find_eigenrays``
(/root/reference/src/pygenray/eigenrays.py:11-203): brackets are adjacent surviving fan rays whose
final depths straddle a receiver depth (:62-79), each refined by regula falsi on the launch angle
(:206-268) until ``|z_end + receiver_depth| < ztol``.  Here all receivers' brackets are found on the
device (``prt_count_brackets`` / ``prt_fill_brackets``) and refined together, one trial shot per
bracket per round (``prt_refine_brackets``); converged launch angles are then shot once more with
the save grid to produce the ``Ray`` trajectories the reference returns.
"""
import ctypes as C

import numpy as np

from . import _lib
from .launch_rays import _RAISING, ATOL, _device_environment, _fan_from_device, shoot_fan_arrays
from .ray_objects import EigenRays

__all__ = ["find_eigenrays", "find_brackets", "refine_brackets", "eigenray_table"]

_POOL_THRESHOLD = 5  # eigenrays.py:122-124: more brackets than this took the Pool branch


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f64_on(x, dev):
    """float64 contiguous 1-D tensor on ``dev`` from a numpy array / sequence or a tensor that may already be there."""
    import torch  # noqa: PLC0415

    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1)).to(dev)


def find_brackets(z_end, thetas, receiver_depths, device=None):
    """Sign-change brackets of ``z_end + receiver_depth`` between adjacent rays, for every receiver.

    ``z_end`` / ``thetas`` are ``RayFan.zs[:, -1]`` / ``RayFan.thetas`` (stored conventions; numpy arrays or
    tensors already on the device).  Returns
    CUDA tensors ``z1, z2, th1, th2, rd (float64), rcv (int32), k (int64)`` ordered by
    (receiver, ray index) and the per-receiver counts as a numpy array.
    """
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        z_d, th_d, rd_d = _f64_on(z_end, dev), _f64_on(thetas, dev), _f64_on(receiver_depths, dev)
        n_rays, n_rcv = z_d.shape[0], rd_d.shape[0]
        counts_d = torch.zeros(max(n_rcv, 1), dtype=torch.int32, device=dev)
        work = torch.empty(int(L.prt_bracket_workspace_bytes(n_rays, n_rcv)), dtype=torch.uint8, device=dev)
        _lib.check(L.prt_count_brackets(n_rays, _ptr(z_d), n_rcv, _ptr(rd_d), _ptr(counts_d), _ptr(work), stream),
                   "prt_count_brackets")
        counts = counts_d[:n_rcv].cpu().numpy().astype(np.int64)  # the one read-back: output sizes
        total = int(counts.sum())
        f64 = dict(dtype=torch.float64, device=dev)
        br = dict(z1=torch.empty(total, **f64), z2=torch.empty(total, **f64), th1=torch.empty(total, **f64),
                  th2=torch.empty(total, **f64), rd=torch.empty(total, **f64),
                  rcv=torch.empty(total, dtype=torch.int32, device=dev),
                  k=torch.empty(total, dtype=torch.int64, device=dev))
        if total:
            _lib.check(
                L.prt_fill_brackets(n_rays, _ptr(z_d), _ptr(th_d), n_rcv, _ptr(rd_d), _ptr(work), _ptr(br["z1"]),
                                    _ptr(br["z2"]), _ptr(br["th1"]), _ptr(br["th2"]), _ptr(br["rd"]),
                                    _ptr(br["rcv"]), _ptr(br["k"]), stream),
                "prt_fill_brackets")
        return br, counts


def refine_brackets(denv, br, source_depth, source_range, receiver_range, rtol=1e-9, terminate_backwards=True,
                    ztol=1.0, max_iter=20, max_bounces=_lib.DEFAULT_MAX_BOUNCES, angle_order=True):
    """Regula-falsi refinement of the brackets ``br`` (CUDA tensors from :func:`find_brackets`).

    ``angle_order``: lay the brackets out over warps sorted by launch angle (an argsort of ``th1``: plumbing) so
    that a warp holds 32 nearly identical trial rays; results are per bracket and do not depend on it."""
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    dev = torch.device("cuda", denv.device)
    n = br["z1"].shape[0]
    with torch.cuda.device(dev):
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        out = dict(theta=torch.empty(n, **f64), final=torch.empty((3, n), **f64), n_bottom=torch.empty(n, **i32),
                   n_surface=torch.empty(n, **i32), status=torch.empty(n, **i32), shots=torch.empty(n, **i32),
                   ray_status=torch.empty(n, **i32))
        if n:
            work = torch.empty(int(L.prt_refine_workspace_bytes(n)), dtype=torch.uint8, device=dev)
            order = torch.argsort(br["th1"], stable=True).to(torch.int32) if angle_order else None
            rc = L.prt_refine_brackets(
                denv.handle, n, _ptr(br["z1"]), _ptr(br["z2"]), _ptr(br["th1"]), _ptr(br["th2"]), _ptr(br["rd"]),
                float(source_depth), float(source_range), float(receiver_range), float(rtol), ATOL,
                int(bool(terminate_backwards)), float(ztol), int(max_iter), _ptr(out["theta"]), _ptr(out["final"]),
                _ptr(out["n_bottom"]), _ptr(out["n_surface"]), _ptr(out["status"]), _ptr(out["shots"]),
                _ptr(out["ray_status"]), _ptr(work), _ptr(order), int(max_bounces),
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "prt_refine_brackets")
    return out


def eigenray_table(denv, br, ref, environment, source_depth, source_range, receiver_range, num_range_save=0,
                   rtol=1e-9, terminate_backwards=True):
    """Arrival ("time front") table of the converged eigenrays of a refinement, built on the device.

    ``br`` / ``ref`` are the results of :func:`find_brackets` / :func:`refine_brackets`.  Returns a dict:
    ``table`` (numpy [n_converged, EIG_TABLE_WIDTH], columns ``_lib.EIG_TABLE_FIELDS``: receiver index, arrival time,
    stored depth and ray parameter, launch angle, received angle, ray id number, bounce counts, bracket index; rows in
    (receiver, ray) order), ``offsets`` ([n_rcv + 1] row range per receiver; needs ``br['rcv']``), and, when
    ``num_range_save`` > 0, ``fan``: the converged launch angles re-shot once on the save grid
    (:func:`shoot_fan_arrays` result on the HOST, stored convention, incl. ``p_sign_changes``) -- only then is the ray
    id number defined (it counts sign changes of p on the save grid, ray_objects.py:533-535), else NaN.
    ``source_range`` / ``receiver_range`` are the INTERNAL (forward) ranges.  The reference has no array-level
    equivalent: it builds Ray objects and loops over them in EigenRays.__init__ (ray_objects.py:505-548).
    """
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    from .ray_objects import unflattened_grid_on_device  # noqa: PLC0415

    dev = torch.device("cuda", denv.device)
    n_br = int(ref["status"].shape[0])
    with torch.cuda.device(dev):
        idx = torch.nonzero(ref["status"] == _lib.EIG_CONVERGED).squeeze(1)
        n = int(idx.shape[0])
        out = dict(n_converged=n)
        fan_d = None
        if num_range_save and n:
            # shoot_ray(theta) integrates the internal angle -theta (launch_rays.py:251); deterministic, so this is
            # the converged trial ray again, now resampled on the save grid
            fan_d = shoot_fan_arrays(denv, source_depth, source_range, receiver_range,
                                     angles=-ref["theta"].index_select(0, idx), num_range_save=num_range_save, rtol=rtol,
                                     terminate_backwards=terminate_backwards, stored_convention=True,
                                     ray_id_counts=True, to_host=False)
        table_d = torch.empty((n, _lib.EIG_TABLE_WIDTH), dtype=torch.float64, device=dev)
        if n:
            grid = unflattened_grid_on_device(environment, dev)
            x_end = abs(receiver_range) if receiver_range < 0 else receiver_range  # RayFan.rs[:, -1] of the user frame
            rc = L.prt_eigen_table(
                n, n_br, _ptr(idx), _ptr(br["rcv"]), _ptr(ref["theta"]), _ptr(ref["final"]), _ptr(ref["n_bottom"]),
                _ptr(ref["n_surface"]), _ptr(fan_d["p_sign_changes"]) if fan_d is not None else None, float(x_end),
                grid["n_r"], grid["n_z"], _ptr(grid["rin"]), _ptr(grid["zin"]), _ptr(grid["cin"]), _ptr(table_d),
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "prt_eigen_table")
        table = table_d.cpu().numpy()
        out["table"] = table
        n_rcv = int(br["rcv"].max().item()) + 1 if n_br else 0
        out["offsets"] = np.searchsorted(table[:, 0], np.arange(n_rcv + 1)) if n else np.zeros(n_rcv + 1, np.int64)
        if fan_d is not None:
            host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in fan_d.items()
                    if isinstance(v, torch.Tensor)}
            for k, v in host.items():
                v.copy_(fan_d[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out["fan"] = {**{k: v.numpy() for k, v in host.items()}, "range_save": fan_d["range_save"]}
    return out


def find_eigenrays(rays, receiver_depths, source_depth, source_range, receiver_range, num_range_save,
                   environment, ztol=1, max_iter=20, num_workers=None, **kwargs):
    """Find eigenrays for ``receiver_depths`` from the fan ``rays``; same call as ``pygenray.find_eigenrays``.

    ``kwargs`` are those of ``shoot_ray`` (``rtol``, ``terminate_backwards``, ``debug``, ``flatearth``).
    ``num_workers`` is accepted and ignored.  Returns an :class:`EigenRays`.  Everything per eigenray -- brackets,
    regula falsi, the re-shot on the save grid, received angles, ray-id numbers -- is computed on the device for all
    receivers at once; the host slices the arrival table per receiver.
    """
    rtol = kwargs.pop("rtol", 1e-9)
    terminate_backwards = kwargs.pop("terminate_backwards", True)
    debug = kwargs.pop("debug", True)
    flatearth = kwargs.pop("flatearth", True)
    if kwargs:
        raise TypeError(f"shoot_ray() got an unexpected keyword argument '{next(iter(kwargs))}'")

    backwards = receiver_range < source_range
    denv = _device_environment(environment, flatearth, backwards)
    src_i, rcv_i = (-source_range, -receiver_range) if backwards else (source_range, receiver_range)

    n_rcv = len(receiver_depths)
    br, counts = find_brackets(rays.zs[:, -1], rays.thetas, receiver_depths, device=denv.device)
    ref = refine_brackets(denv, br, source_depth, src_i, rcv_i, rtol=rtol, terminate_backwards=terminate_backwards,
                          ztol=ztol, max_iter=max_iter)
    status = ref["status"].cpu().numpy()
    offsets = np.concatenate(([0], np.cumsum(counts)))
    dropped = np.flatnonzero(status == _lib.EIG_RAY_DROPPED)
    if dropped.shape[0]:
        ray_status = ref["ray_status"].cpu().numpy()
        if int(counts.max()) <= _POOL_THRESHOLD:
            # sequential branch of the reference (eigenrays.py:166-189): an exception inside a trial shot propagates
            for b in dropped:
                if int(ray_status[b]) in _RAISING:
                    exc, msg = _RAISING[int(ray_status[b])]
                    raise exc(msg)
        if debug:
            rd_all = br["rd"].cpu().numpy()
            for b in dropped[:20]:  # the reference prints one line per failure (eigenrays.py:241-245)
                print(f"Failed to find eigen ray for receiver depth {rd_all[b]} [m]: a trial ray was dropped "
                      f"({_lib.RAY_STATUS.get(int(ray_status[b]), ray_status[b])})")
            if dropped.shape[0] > 20:
                print(f"... and {dropped.shape[0] - 20} more brackets whose trial ray was dropped")

    tab = eigenray_table(denv, br, ref, environment, source_depth, src_i, rcv_i, num_range_save=num_range_save,
                         rtol=rtol, terminate_backwards=terminate_backwards)
    table, t_off, res = tab["table"], tab["offsets"], tab.get("fan")
    th1 = th2 = None
    erays_dict, num_eigenrays, num_found, failed = {}, {}, {}, {}
    for ridx, receiver_depth in enumerate(receiver_depths):
        lo, hi = int(offsets[ridx]), int(offsets[ridx + 1])
        num_eigenrays[receiver_depth] = hi - lo
        a, b = (int(t_off[ridx]), int(t_off[ridx + 1])) if ridx + 1 < len(t_off) else (0, 0)
        rows = np.arange(a, b)
        if res is not None:
            sub = {k: (v[:, rows] if k in ("final", "traj") else (v[rows] if k != "range_save" else v))
                   for k, v in res.items()}
        else:  # no converged eigenray anywhere: empty fans on the requested save grid
            sub = dict(range_save=np.linspace(src_i, rcv_i, num_range_save), traj=np.zeros((3, 0, num_range_save)),
                       n_bottom=np.zeros(0, np.int32), n_surface=np.zeros(0, np.int32))
        sub["kept"] = np.arange(rows.shape[0])
        sub.pop("p_sign_changes", None)  # ray id numbers come from the device-built table
        fan = _fan_from_device(sub, table[a:b, 4], source_depth, backwards)
        fan._eigen_table = (table[a:b, 5], table[a:b, 6])
        fan._ray_id_numbers = table[a:b, 6]
        erays_dict[ridx] = fan
        n_failed = (hi - lo) - (b - a)
        failed[ridx] = []
        if n_failed:
            if th1 is None:
                th1, th2 = br["th1"].cpu().numpy(), br["th2"].cpu().numpy()
            if (hi - lo) > _POOL_THRESHOLD:
                # the Pool branch of the reference reports every failure with the LAST bracket's angles (stale
                # loop index, eigenrays.py:160-164)
                failed[ridx] = [(th1[hi - 1], th2[hi - 1])] * n_failed
            else:
                bad = lo + np.flatnonzero(status[lo:hi] != _lib.EIG_CONVERGED)
                failed[ridx] = [(th1[j], th2[j]) for j in bad]
        num_found[ridx] = int(b - a)
    return EigenRays(receiver_depths, erays_dict, environment, num_eigenrays, num_found, failed)
