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
from .launch_rays import ATOL, _device_environment, _fan_from_device, shoot_fan_arrays
from .ray_objects import EigenRays

__all__ = ["find_eigenrays", "find_brackets", "refine_brackets"]

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
                    ztol=1.0, max_iter=20, max_bounces=_lib.DEFAULT_MAX_BOUNCES):
    """Regula-falsi refinement of the brackets ``br`` (CUDA tensors from :func:`find_brackets`)."""
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
            rc = L.prt_refine_brackets(
                denv.handle, n, _ptr(br["z1"]), _ptr(br["z2"]), _ptr(br["th1"]), _ptr(br["th2"]), _ptr(br["rd"]),
                float(source_depth), float(source_range), float(receiver_range), float(rtol), ATOL,
                int(bool(terminate_backwards)), float(ztol), int(max_iter), _ptr(out["theta"]), _ptr(out["final"]),
                _ptr(out["n_bottom"]), _ptr(out["n_surface"]), _ptr(out["status"]), _ptr(out["shots"]),
                _ptr(out["ray_status"]), _ptr(work), int(max_bounces),
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "prt_refine_brackets")
    return out


def find_eigenrays(rays, receiver_depths, source_depth, source_range, receiver_range, num_range_save,
                   environment, ztol=1, max_iter=20, num_workers=None, **kwargs):
    """Find eigenrays for ``receiver_depths`` from the fan ``rays``; same call as ``pygenray.find_eigenrays``.

    ``kwargs`` are those of ``shoot_ray`` (``rtol``, ``terminate_backwards``, ``debug``, ``flatearth``).
    ``num_workers`` is accepted and ignored.  Returns an :class:`EigenRays`.
    """
    rtol = kwargs.pop("rtol", 1e-9)
    terminate_backwards = kwargs.pop("terminate_backwards", True)
    kwargs.pop("debug", True)
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
    theta = ref["theta"].cpu().numpy()
    rd_all = br["rd"].cpu().numpy()
    th1 = br["th1"].cpu().numpy()
    th2 = br["th2"].cpu().numpy()
    trial_theta_failed = status == _lib.EIG_RAY_DROPPED

    # final shot of the converged launch angles with the save grid: shoot_ray(theta) -> internal -theta,
    # and find_eigenrays flips Ray.launch_angle back to theta (launch_rays.py:251, eigenrays.py:247-250)
    ok = status == _lib.EIG_CONVERGED
    ok_pos = np.flatnonzero(ok)  # bracket index of each re-shot ray
    res = shoot_fan_arrays(denv, source_depth, src_i, rcv_i, angles=-theta[ok], num_range_save=num_range_save,
                           rtol=rtol, terminate_backwards=terminate_backwards, stored_convention=True,
                           ray_id_counts=True)
    alive = res["status"] == 0  # deterministic re-shot of a converged trial: always OK

    erays_dict, num_eigenrays, num_found, failed = {}, {}, {}, {}
    offsets = np.concatenate(([0], np.cumsum(counts)))
    for ridx, receiver_depth in enumerate(receiver_depths):
        lo, hi = int(offsets[ridx]), int(offsets[ridx + 1])
        num_eigenrays[receiver_depth] = hi - lo
        rows = np.flatnonzero((ok_pos >= lo) & (ok_pos < hi) & alive)  # rows of `res` of this receiver
        sub = {k: (v[:, rows] if k in ("final", "traj") else (v[rows] if k != "range_save" else v))
               for k, v in res.items()}
        sub["kept"] = np.arange(rows.shape[0])
        erays_dict[ridx] = _fan_from_device(sub, theta[ok_pos[rows]], source_depth, backwards)
        failed[ridx] = []
        for b in np.arange(lo, hi)[~ok[lo:hi]]:
            if trial_theta_failed[b]:
                print(f"Failed to find eigen ray for receiver depth {rd_all[b]} [m]: a trial ray was dropped")
            # the Pool branch of the reference (> 5 brackets) reports every failure with the LAST
            # bracket's angles (stale loop index, eigenrays.py:160-164); the sequential one is exact
            j = hi - 1 if (hi - lo) > _POOL_THRESHOLD else b
            failed[ridx].append((th1[j], th2[j]))
        num_found[ridx] = int(rows.shape[0])
    return EigenRays(receiver_depths, erays_dict, environment, num_eigenrays, num_found, failed)
