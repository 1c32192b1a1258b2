"""Ray-fan shooting: host layer over libprt.so.

Public entry points keep the reference's signatures, defaults, sign conventions and error
behaviour (/root/reference/src/pygenray/launch_rays.py): :func:`shoot_rays` (:11-200),
:func:`shoot_ray` (:203-322), plus the array-level seam :func:`_unpack_envi` (:717-742) and
:func:`_mirror_envi_arrays` (:684-714).  The per-ray integration (``_shoot_ray_array`` +
``_interpolate_ray``, :325-484, 745-784) runs on the GPU for all launch angles at once
(:func:`shoot_fan_arrays` -> ``prt_shoot_fan``); the reference's multiprocessing pool and shared
memory (:133-198) have no equivalent here.  There is no CPU path: without libprt.so and a CUDA
device every call raises.
"""
import ctypes as C
import hashlib

import numpy as np
from scipy.interpolate import CubicSpline

from . import _lib, _xr
from .ray_objects import DeviceRows, Ray, RayFan

__all__ = ["shoot_rays", "shoot_ray", "shoot_fan_arrays", "shoot_ensemble_arrays", "DeviceEnvironment",
           "DeviceEnsemble", "_unpack_envi", "_mirror_envi_arrays", "bottom_angle_ppoly"]

ATOL = 1e-6  # scipy.integrate.solve_ivp default, never overridden by the reference (launch_rays.py:670-679)

# Host arrays returned by shoot_fan_arrays / shoot_rays are views of PAGE-LOCKED buffers (the D2H copies land in them
# directly: an extra host pass over 240 MB of trajectories would cost more than the kernel).  A caller that keeps many
# fans alive, or very large ones, can switch this off: the results are then copied once more into ordinary numpy
# arrays and the pinned buffers are released with the call.
PINNED_RESULTS = True

# shoot_rays() with fewer than 70 angles negates the launch angles and then calls shoot_ray(), which
# negates them again (launch_rays.py:64-67, 94-117, 251): user angle a is integrated with internal
# angle +a, whereas the >= 70 branch integrates -a.  The reference's own golden fixture pins this
# behaviour (tests/test_physics.py:313-316), so it is reproduced by default.
REPLICATE_SMALL_FAN_SIGN_QUIRK = True
_SMALL_FAN = 70


# ---------------------------------------------------------------------------------------------
# environment -> arrays -> device
def _unpack_envi(environment, flatearth=True):
    """The seven arrays the integrator consumes: cin, cpin, rin, zin, depths, depth_ranges, bottom_angles."""
    if flatearth:
        if not hasattr(environment, "sound_speed_fe"):
            raise Exception(
                "Flat earth transformation has not been applied. Set `flat_earth_transform=True` when creating "
                "the OceanEnvironment2D object."
            )
        ss, bathy = environment.sound_speed_fe, environment.bathymetry_fe
    else:
        ss, bathy = environment.sound_speed, environment.bathymetry
    cin = np.array(_xr.values_of(ss))
    cpin = np.array(_xr.differentiate(ss, "depth"))
    rin = np.array(_xr.coord_of(ss, "range"))
    zin = np.array(_xr.coord_of(ss, "depth"))
    depths = np.array(_xr.values_of(bathy))
    depth_ranges = np.array(_xr.coord_of(bathy, "range"))
    bottom_angles = np.array(environment.bottom_angle)
    return cin, cpin, rin, zin, depths, depth_ranges, bottom_angles


def _mirror_envi_arrays(cin, cpin, rin, depths, depth_ranges, bottom_angles):
    """Reflect the environment about the range axis (x' = -x) so a backwards shot runs forwards."""
    return (cin[::-1, :], cpin[::-1, :], -rin[::-1], depths[::-1], -depth_ranges[::-1], -bottom_angles[::-1])


def _check_monotonic(rin, zin, depth_ranges):
    if not np.all(np.diff(rin) >= 0):
        raise Exception("Sound speed range coordinates must be monotonically increasing.")
    if not np.all(np.diff(zin) >= 0):
        raise Exception("Sound speed depth coordinates must be monotonically increasing.")
    if not np.all(np.diff(depth_ranges) >= 0):
        raise Exception("Bathymetry range coordinates must be monotonically increasing.")


def bottom_angle_ppoly(depth_ranges, bottom_angles):
    """Not-a-knot cubic spline of the bottom angle as piecewise-polynomial coefficients.

    Stands for ``scipy.interpolate.interp1d(depth_ranges, bottom_angles, kind="cubic")`` which
    ``_shoot_ray_array`` rebuilds for every ray (launch_rays.py:397-399): the same interpolant,
    expressed in the local power basis so the device evaluates it with three FMAs.
    Returns ``(x[n], c[4][n-1])`` with the highest power first.
    """
    x = np.asarray(depth_ranges, dtype=np.float64)
    y = np.asarray(bottom_angles, dtype=np.float64)
    if x.shape[0] < 4:
        raise ValueError("x and y arrays must have at least 4 entries")  # interp1d(kind="cubic")
    cs = CubicSpline(x, y, bc_type="not-a-knot")
    return np.ascontiguousarray(cs.x), np.ascontiguousarray(cs.c)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class DeviceEnvironment:
    """One environment staged in B200 HBM (``prt_env_create``): interleaved {c, dc/dz} grid etc."""

    def __init__(self, cin, cpin, rin, zin, depths, depth_ranges, bottom_angles, device=None):
        L = _lib.require_gpu()
        import torch  # noqa: PLC0415

        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        cin, cpin, rin, zin = _as_f64(cin), _as_f64(cpin), _as_f64(rin), _as_f64(zin)
        depths, depth_ranges = _as_f64(depths), _as_f64(depth_ranges)
        if cin.ndim != 2 or cin.shape != (rin.shape[0], zin.shape[0]) or cpin.shape != cin.shape:
            raise ValueError(
                f"sound speed must be a [range][depth] array matching its coordinates, got {cin.shape} for "
                f"{rin.shape[0]} ranges x {zin.shape[0]} depths"
            )
        ba_x, ba_c = bottom_angle_ppoly(depth_ranges, bottom_angles)
        handle = C.c_void_p()
        rc = L.prt_env_create(
            self.device, rin.shape[0], zin.shape[0], rin.ctypes.data, zin.ctypes.data, cin.ctypes.data,
            cpin.ctypes.data, depths.shape[0], depth_ranges.ctypes.data, depths.ctypes.data, ba_x.shape[0],
            ba_x.ctypes.data, ba_c.ctypes.data, C.byref(handle),
        )
        _lib.check(rc, "prt_env_create")
        self._h = handle
        self.c_source_args = (rin, zin, cin)
        self.n_r, self.n_z = rin.shape[0], zin.shape[0]
        self.device_bytes = int(L.prt_env_device_bytes(handle))

    @property
    def handle(self):
        if self._h is None:
            raise _lib.PrtError("DeviceEnvironment was destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().prt_env_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class DeviceEnsemble(DeviceEnvironment):
    """``n_env`` realisations of the sound-speed field on one grid geometry, staged together
    (``prt_ensemble_create``): ``cin``/``cpin`` are ``[n_env][range][depth]``; coordinates, bathymetry and
    bottom angles are shared.  Shot with :func:`shoot_ensemble_arrays` (tomography ensembles, config 5)."""

    def __init__(self, cin, cpin, rin, zin, depths, depth_ranges, bottom_angles, device=None):
        L = _lib.require_gpu()
        import torch  # noqa: PLC0415

        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        cin, cpin, rin, zin = _as_f64(cin), _as_f64(cpin), _as_f64(rin), _as_f64(zin)
        depths, depth_ranges = _as_f64(depths), _as_f64(depth_ranges)
        if cin.ndim != 3 or cin.shape[1:] != (rin.shape[0], zin.shape[0]) or cpin.shape != cin.shape:
            raise ValueError(f"ensemble sound speed must be [n_env][range][depth], got {cin.shape}")
        ba_x, ba_c = bottom_angle_ppoly(depth_ranges, bottom_angles)
        handle = C.c_void_p()
        rc = L.prt_ensemble_create(
            self.device, cin.shape[0], rin.shape[0], zin.shape[0], rin.ctypes.data, zin.ctypes.data, cin.ctypes.data,
            cpin.ctypes.data, depths.shape[0], depth_ranges.ctypes.data, depths.ctypes.data, ba_x.shape[0],
            ba_x.ctypes.data, ba_c.ctypes.data, C.byref(handle),
        )
        _lib.check(rc, "prt_ensemble_create")
        self._h = handle
        self.n_env, self.n_r, self.n_z = cin.shape[0], rin.shape[0], zin.shape[0]
        self.device_bytes = int(L.prt_env_device_bytes(handle))


def _ensemble_from_device(cls, c_stack, rin, zin, depths, depth_ranges, bottom_angles):
    """``n_env`` realisations that already live on the GPU: ``c_stack`` is a CUDA float64 tensor
    ``[n_env][range][depth]`` (e.g. perturbations generated with torch).  The depth derivative
    (``sound_speed.differentiate("depth")`` of the reference, launch_rays.py:728-733) is formed on the device with
    numpy's operations in numpy's order -- results are bit-identical to staging ``np.gradient`` from the host --
    and only the coordinate tables cross PCIe (``prt_ensemble_create_dev``)."""
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    if not (isinstance(c_stack, torch.Tensor) and c_stack.is_cuda and c_stack.dtype == torch.float64):
        raise TypeError("c_stack must be a CUDA float64 tensor [n_env][range][depth]")
    rin, zin = _as_f64(rin), _as_f64(zin)
    depths, depth_ranges = _as_f64(depths), _as_f64(depth_ranges)
    if c_stack.ndim == 2:
        c_stack = c_stack.unsqueeze(0)
    if c_stack.ndim != 3 or tuple(c_stack.shape[1:]) != (rin.shape[0], zin.shape[0]):
        raise ValueError(f"ensemble sound speed must be [n_env][range][depth], got {tuple(c_stack.shape)}")
    c_stack = c_stack.contiguous()
    ba_x, ba_c = bottom_angle_ppoly(depth_ranges, bottom_angles)
    self = cls.__new__(cls)
    self.device = int(c_stack.device.index)
    handle = C.c_void_p()
    with torch.cuda.device(c_stack.device):
        rc = L.prt_ensemble_create_dev(
            self.device, c_stack.shape[0], rin.shape[0], zin.shape[0], rin.ctypes.data, zin.ctypes.data,
            C.c_void_p(c_stack.data_ptr()), depths.shape[0], depth_ranges.ctypes.data, depths.ctypes.data,
            ba_x.shape[0], ba_x.ctypes.data, ba_c.ctypes.data, C.c_void_p(torch.cuda.current_stream().cuda_stream),
            C.byref(handle),
        )
    _lib.check(rc, "prt_ensemble_create_dev")
    self._h = handle
    self.n_env, self.n_r, self.n_z = int(c_stack.shape[0]), rin.shape[0], zin.shape[0]
    self.device_bytes = int(L.prt_env_device_bytes(handle))
    return self


DeviceEnsemble.from_device = classmethod(_ensemble_from_device)


def shoot_ensemble_arrays(dens, source_depth, source_range, receiver_range, angles, n_env=None, num_range_save=0,
                          rtol=1e-9, terminate_backwards=True, counters=False, max_bounces=_lib.DEFAULT_MAX_BOUNCES,
                          to_host=True):
    """Shoot the same fan of INTERNAL ``angles`` through the first ``n_env`` realisations of ``dens`` in one
    launch.  Returns the dict of :func:`shoot_fan_arrays` with a leading realisation axis:
    ``final (3, n_env, n)``, ``status (n_env, n)``, ``traj (3, n_env, n, N)`` ..."""
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    if not (source_range < receiver_range):
        raise ValueError("shoot_ensemble_arrays integrates forwards: mirror the environment for backwards shots")
    n_env = dens.n_env if n_env is None else int(n_env)
    dev = torch.device("cuda", dens.device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream()
        ang_d = torch.from_numpy(_as_f64(np.atleast_1d(angles)).reshape(-1)).to(dev)
        k = ang_d.shape[0]
        n = n_env * k
        n_save = int(num_range_save) if num_range_save else 0
        out = {}
        rs_d = traj_d = None
        if n_save:
            rs_np = np.linspace(source_range, receiver_range, n_save)
            rs_d = torch.from_numpy(rs_np).to(dev)
            traj_d = torch.empty((3, n_env, k, n_save), dtype=torch.float64, device=dev)
            out["range_save"] = rs_np
        final_d = torch.empty((3, n_env, k), dtype=torch.float64, device=dev)
        nb_d = torch.empty((n_env, k), dtype=torch.int32, device=dev)
        ns_d = torch.empty((n_env, k), dtype=torch.int32, device=dev)
        st_d = torch.empty((n_env, k), dtype=torch.int32, device=dev)
        cnt_d = torch.empty((_lib.N_COUNTERS, n_env, k), dtype=torch.int32, device=dev) if counters else None

        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        if n:
            rc = L.prt_shoot_ensemble(
                dens.handle, n_env, k, ptr(ang_d), float(source_depth), float(source_range), float(receiver_range),
                float(rtol), ATOL, int(bool(terminate_backwards)), n_save, ptr(rs_d), ptr(traj_d), ptr(final_d),
                ptr(nb_d), ptr(ns_d), ptr(st_d), ptr(cnt_d), 0, None, int(max_bounces),
                C.c_void_p(stream.cuda_stream),
            )
            _lib.check(rc, "prt_shoot_ensemble")
        dev_out = dict(final=final_d, n_bottom=nb_d, n_surface=ns_d, status=st_d)
        if traj_d is not None:
            dev_out["traj"] = traj_d
        if cnt_d is not None:
            dev_out["counters"] = cnt_d
        if not to_host:
            out.update(dev_out)
            return out
        host = {kk: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for kk, v in dev_out.items()}
        for kk, v in dev_out.items():
            host[kk].copy_(v, non_blocking=True)
        stream.synchronize()
        out.update({kk: v.numpy() for kk, v in host.items()})
    return out


def _fingerprint(*arrays):
    """160-bit content digest of the environment arrays (shape-aware): the key of the device-side cache.
    (SHA-1: ~1 GB/s here, three times faster than BLAKE2b and far beyond accidental-collision range.)"""
    h = hashlib.sha1()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(repr((a.shape, a.dtype.str)).encode())
        h.update(a.view(np.uint8).reshape(-1))
    return h.digest()


def _device_environment(environment, flatearth, backwards, device=None):
    """Unpack (+ mirror) the environment and stage it once per (content, orientation, device).

    The cache key is a 160-bit digest of the arrays the environment holds (so in-place edits are seen; about 1 ms per
    MB of grid per call); the depth derivative (``np.gradient``, the expensive part of ``_unpack_envi``) is only
    formed on a miss.  Hot loops over one environment should stage it once (:class:`DeviceEnvironment`) and call
    :func:`shoot_fan_arrays`."""
    import torch  # noqa: PLC0415

    if flatearth and not hasattr(environment, "sound_speed_fe"):
        _unpack_envi(environment, flatearth=True)  # raises the reference's exception
    ss, bathy = ((environment.sound_speed_fe, environment.bathymetry_fe) if flatearth
                 else (environment.sound_speed, environment.bathymetry))
    if device is None:
        _lib.require_gpu()
        device = torch.cuda.current_device()
    key = (bool(flatearth), bool(backwards), int(device),
           _fingerprint(_xr.values_of(ss), _xr.coord_of(ss, "range"), _xr.coord_of(ss, "depth"), _xr.values_of(bathy),
                        _xr.coord_of(bathy, "range"), np.asarray(environment.bottom_angle)))
    cache = environment.__dict__.setdefault("_prt_device_cache", {})
    denv = cache.get(key)
    if denv is None:
        cin, cpin, rin, zin, depths, depth_ranges, bottom_angles = _unpack_envi(environment, flatearth=flatearth)
        _check_monotonic(rin, zin, depth_ranges)
        if backwards:
            cin, cpin, rin, depths, depth_ranges, bottom_angles = _mirror_envi_arrays(
                cin, cpin, rin, depths, depth_ranges, bottom_angles
            )
        denv = DeviceEnvironment(cin, cpin, rin, zin, depths, depth_ranges, bottom_angles, device=device)
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))  # freed by DeviceEnvironment.__del__ once no launch holds it any more
        cache[key] = denv
    return denv


# ---------------------------------------------------------------------------------------------
# the batched seam
def shoot_fan_arrays(denv, source_depth, source_range, receiver_range, angles=None, y0s=None,
                     num_range_save=0, rtol=1e-9, terminate_backwards=True, counters=False,
                     stored_convention=False, ray_id_counts=False, max_bounces=_lib.DEFAULT_MAX_BOUNCES,
                     to_host=True, compact=False):
    """Shoot ``n`` rays on the GPU.  ``angles`` are INTERNAL launch angles [deg] (positive = downwards).

    Returns a dict: ``final (3,n)`` [T, z, p], ``n_bottom``, ``n_surface``, ``status`` (n,), and, when
    ``num_range_save > 0``, ``range_save (N,)`` and ``traj (3,n,N)``; optional ``counters (6,n)`` and
    ``p_sign_changes (n,)``.  z and p are in the ODE convention (z positive down) unless
    ``stored_convention`` (then the trajectory holds -z and -p as ``Ray`` stores them).  Arrays are
    numpy (pinned-memory copies) when ``to_host`` else CUDA tensors.  ``compact`` drops the rays whose
    status is not OK on the device before anything is copied (``kept`` = their indices in the input).
    """
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    if not (source_range < receiver_range):
        raise ValueError("shoot_fan_arrays integrates forwards: mirror the environment for backwards shots")
    dev = torch.device("cuda", denv.device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream()
        if y0s is not None:
            y0_h = torch.from_numpy(_as_f64(y0s).reshape(-1, 3))
            n = y0_h.shape[0]
            y0_d = y0_h.to(dev)
            ang_d = None
        elif isinstance(angles, torch.Tensor):  # launch angles that already live on the device
            ang_d = angles.to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
            n = ang_d.shape[0]
            y0_d = None
        else:
            ang_h = torch.from_numpy(_as_f64(np.atleast_1d(angles)).reshape(-1))
            n = ang_h.shape[0]
            ang_d = ang_h.to(dev)
            y0_d = None
        n_save = int(num_range_save) if num_range_save else 0
        out = {}
        rs_d = traj_d = None
        if n_save:
            rs_np = np.linspace(source_range, receiver_range, n_save)
            rs_d = torch.from_numpy(rs_np).to(dev)
            traj_d = torch.empty((3, n, n_save), dtype=torch.float64, device=dev)
            out["range_save"] = rs_np
        final_d = torch.empty((3, n), dtype=torch.float64, device=dev)
        nb_d = torch.empty(n, dtype=torch.int32, device=dev)
        ns_d = torch.empty(n, dtype=torch.int32, device=dev)
        st_d = torch.empty(n, dtype=torch.int32, device=dev)
        cnt_d = torch.empty((_lib.N_COUNTERS, n), dtype=torch.int32, device=dev) if counters else None
        flips_d = torch.empty(n, dtype=torch.int32, device=dev) if (ray_id_counts and n_save) else None

        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        if n:
            rc = L.prt_shoot_fan(
                denv.handle, n, ptr(ang_d), ptr(y0_d), float(source_depth), float(source_range),
                float(receiver_range), float(rtol), ATOL, int(bool(terminate_backwards)), n_save, ptr(rs_d),
                ptr(traj_d), ptr(final_d), ptr(nb_d), ptr(ns_d), ptr(st_d), ptr(cnt_d),
                int(bool(stored_convention)), ptr(flips_d), int(max_bounces), C.c_void_p(stream.cuda_stream),
            )
            _lib.check(rc, "prt_shoot_fan")
        dev_out = dict(final=final_d, n_bottom=nb_d, n_surface=ns_d, status=st_d)
        if traj_d is not None:
            dev_out["traj"] = traj_d
        if cnt_d is not None:
            dev_out["counters"] = cnt_d
        if flips_d is not None:
            dev_out["p_sign_changes"] = flips_d
        if compact:
            kept = torch.nonzero(st_d == 0).squeeze(1)
            if kept.shape[0] < n:
                dev_out = {k: (v.index_select(1, kept) if k in ("final", "traj", "counters") else v[kept])
                           for k, v in dev_out.items()}
            dev_out["kept"] = kept
        if not to_host:
            out.update(dev_out)
            return out
        host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in dev_out.items()}
        for k, v in dev_out.items():
            host[k].copy_(v, non_blocking=True)
        stream.synchronize()
        out.update({k: (v.numpy() if PINNED_RESULTS else v.numpy().copy()) for k, v in host.items()})
    return out


_PIPELINE_MIN_RAYS = 4096
_PIPELINE_CHUNKS = 8
_STREAM_POOL = {}


def _streams(dev, n):
    pool = _STREAM_POOL.setdefault(dev.index, [])
    import torch  # noqa: PLC0415

    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def _shoot_fan_pipelined(denv, source_depth, source_range, receiver_range, angles, num_range_save, rtol,
                         terminate_backwards, max_bounces=_lib.DEFAULT_MAX_BOUNCES):
    """Large fan with trajectories, host arrays out: the fan is cut into contiguous chunks shot concurrently on
    separate streams, and each chunk's surviving rays are copied to the host while the slower chunks (steep,
    reflecting rays at the ends of the fan) are still integrating.  Survivors are compacted per chunk on the
    device and placed from the middle of the fan outwards, so the host arrays are contiguous without a
    host-side pass.  The host only ever waits for a chunk's kernel (to learn how many rays survived), never for
    its copies: the PCIe link stays busy from the first finished chunk to the last.  Same result dict as
    ``shoot_fan_arrays(..., stored_convention=True, ray_id_counts=True, compact=True)``."""
    L = _lib.require_gpu()
    import torch  # noqa: PLC0415

    ang = _as_f64(np.atleast_1d(angles)).reshape(-1)
    n, n_save = ang.shape[0], int(num_range_save)
    n_chunks = max(1, min(_PIPELINE_CHUNKS, n // 1024))
    per = -(-n // n_chunks)
    per = -(-per // 32) * 32
    bounds = [(a, min(a + per, n)) for a in range(0, n, per)]
    n_chunks = len(bounds)
    dev = torch.device("cuda", denv.device)
    rs_np = np.linspace(source_range, receiver_range, n_save)

    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        ang_d = torch.from_numpy(ang).pin_memory().to(dev, non_blocking=True)
        rs_d = torch.from_numpy(rs_np).to(dev)
        traj_h = torch.empty((3, n, n_save), dtype=torch.float64, pin_memory=True)
        final_h = torch.empty((3, n), dtype=torch.float64, pin_memory=True)
        i32_h = torch.empty((3, n), dtype=torch.int32, pin_memory=True)   # n_bottom, n_surface, p sign changes
        idx_h = torch.empty(n, dtype=torch.int64, pin_memory=True)
        # one device allocation per kind; chunk c owns rows [a, b) of each (its trajectory block is [3][m][N])
        traj_d = torch.empty(3 * n * n_save, dtype=torch.float64, device=dev)
        final_d = torch.empty(3 * n, dtype=torch.float64, device=dev)
        i32_d = torch.empty((4, n), dtype=torch.int32, device=dev)         # n_bottom, n_surface, status, flips
        streams = _streams(dev, n_chunks)
        # middle chunks first (they finish first: launch them first too), growing a contiguous block of host rows outwards
        centre = (n_chunks - 1) / 2.0
        order = sorted(range(n_chunks), key=lambda k: (abs(k - centre), k))
        chunks = [None] * n_chunks
        L.prt_set_concurrency(n_chunks)
        try:
            for k in order:
                a, b = bounds[k]
                m = b - a
                st = streams[k]
                st.wait_stream(main)
                c = dict(a=a, b=b, stream=st, traj=traj_d[3 * a * n_save:3 * b * n_save].view(3, m, n_save),
                         final=final_d[3 * a:3 * b].view(3, m), nb=i32_d[0, a:b], ns=i32_d[1, a:b],
                         status=i32_d[2, a:b], flips=i32_d[3, a:b])
                rc = L.prt_shoot_fan(
                    denv.handle, m, C.c_void_p(ang_d.data_ptr() + 8 * a), None, float(source_depth),
                    float(source_range), float(receiver_range), float(rtol), ATOL,
                    int(bool(terminate_backwards)), n_save, C.c_void_p(rs_d.data_ptr()),
                    C.c_void_p(c["traj"].data_ptr()), C.c_void_p(c["final"].data_ptr()),
                    C.c_void_p(c["nb"].data_ptr()), C.c_void_p(c["ns"].data_ptr()),
                    C.c_void_p(c["status"].data_ptr()), None, 1, C.c_void_p(c["flips"].data_ptr()), int(max_bounces),
                    C.c_void_p(st.cuda_stream))
                _lib.check(rc, "prt_shoot_fan")
                chunks[k] = c
        finally:
            L.prt_set_concurrency(1)

        left = right = None
        kept_slices = {}
        for k in order:
            c = chunks[k]
            with torch.cuda.stream(c["stream"]):
                idx = torch.nonzero(c["status"] == 0).squeeze(1)  # waits for this chunk's kernel only
                kc = int(idx.shape[0])
                m = c["b"] - c["a"]
                src = {key: c[key] for key in ("traj", "final", "nb", "ns", "flips")}
                if kc < m:
                    src = {key: (v.index_select(1, idx) if key in ("traj", "final") else v[idx]) for key, v in src.items()}
                if left is None:
                    lo = c["a"]
                    left, right = lo, lo + kc
                elif k < centre:
                    lo = left - kc
                    left = lo
                else:
                    lo = right
                    right = lo + kc
                if kc:
                    for q in range(3):  # plane by plane: each destination slice is contiguous (a strided
                        traj_h[q, lo:lo + kc].copy_(src["traj"][q], non_blocking=True)  # D2H would be staged)
                        final_h[q, lo:lo + kc].copy_(src["final"][q], non_blocking=True)
                    i32_h[0, lo:lo + kc].copy_(src["nb"], non_blocking=True)
                    i32_h[1, lo:lo + kc].copy_(src["ns"], non_blocking=True)
                    i32_h[2, lo:lo + kc].copy_(src["flips"], non_blocking=True)
                    idx_h[lo:lo + kc].copy_(idx, non_blocking=True)
                kept_slices[k] = (lo, kc, c["a"])
                c["keep_alive"] = (src, idx)
        for c in chunks:
            c["stream"].synchronize()
            main.wait_stream(c["stream"])
    idx_np = idx_h.numpy()
    parts = [idx_np[lo:lo + kc] + a for (lo, kc, a) in (kept_slices[k] for k in range(n_chunks))]
    kept = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    take = (lambda t: t.numpy()) if PINNED_RESULTS else (lambda t: t.numpy().copy())
    return dict(range_save=rs_np, traj=take(traj_h[:, left:right]), final=take(final_h[:, left:right]),
                n_bottom=take(i32_h[0, left:right]), n_surface=take(i32_h[1, left:right]),
                status=np.zeros(right - left, dtype=np.int32), p_sign_changes=take(i32_h[2, left:right]), kept=kept)


_RAISING = {
    5: (ValueError, "A value in x_new is outside the interpolation range of the bottom-angle spline."),
    6: (ValueError, "f(a) and f(b) must have different signs"),
    7: (RuntimeError, "Failed to converge after 100 iterations."),
    9: (RuntimeError, "ray exceeded the max_bounces safety cap (no reference equivalent: the reference would loop on)"),
}


def _debug_print(status):
    if status == 1:
        print("Integration failed with message: Required step size is less than spacing between numbers.")
    elif status == 2:
        print("ray is vertical, terminating integration")
    elif status == 3:
        print("ray left bounding box, terminating integration")
    elif status == 4:
        print("ray bounced backwards, terminating integration")


def _fan_from_device(res, stored_thetas, source_depth, backwards):
    """RayFan (stored sign convention) from a COMPACTED stored-convention result (survivors only)."""
    traj = res["traj"]
    rs = -res["range_save"] if backwards else res["range_save"]
    kept = res["kept"]
    m = traj.shape[1]
    if not isinstance(traj, np.ndarray):  # trajectories left on the device (device_trajectories=True)
        traj = [DeviceRows(traj[q]) for q in range(3)]
    fan = RayFan.__new__(RayFan)
    fan.thetas = np.asarray(stored_thetas, dtype=np.float64)[kept]
    # every ray shares the save grid: a read-only broadcast view (the reference materialises M writable copies;
    # use np.array(fan.rs) to edit)
    fan.rs = np.broadcast_to(rs, (m, rs.shape[0]))
    fan.ts, fan.zs, fan.ps = traj[0], traj[1], traj[2]
    fan.n_botts = res["n_bottom"].astype(np.int64)
    fan.n_surfs = res["n_surface"].astype(np.int64)
    fan.source_depths = np.full(m, source_depth)
    flips = res.get("p_sign_changes")
    if flips is not None:
        fan._ray_id_numbers = flips * np.sign(fan.thetas)  # strings are built on first access of .ray_ids
    return fan


def shoot_rays(source_depth, source_range, launch_angles, receiver_range, num_range_save, environment,
               rtol=1e-9, terminate_backwards=True, n_processes=None, debug=True, flatearth=True,
               device_trajectories=False):
    """Integrate a fan of rays; same call as ``pygenray.shoot_rays``.  Returns a :class:`RayFan`.

    ``n_processes`` is accepted and ignored (the fan is one GPU launch).  Dropped rays (vertical,
    left the sound-speed grid, bounced backwards, step size underflow) are absent from the fan.
    ``device_trajectories=True`` (an extension): ``ts`` / ``zs`` / ``ps`` of the returned fan stay on the GPU as
    :class:`DeviceRows` and are copied to the host only when indexed; everything else (angles, bounce counts, ray
    ids) is on the host as usual.
    """
    if type(launch_angles) is list:
        launch_angles = np.array(launch_angles)
    launch_angles = -np.asarray(launch_angles, dtype=np.float64)  # sign flip #1 (launch_rays.py:64-67)
    small = launch_angles.shape[0] < _SMALL_FAN
    backwards = receiver_range < source_range
    denv = _device_environment(environment, flatearth, backwards)
    src_i, rcv_i = (-source_range, -receiver_range) if backwards else (source_range, receiver_range)

    if small and REPLICATE_SMALL_FAN_SIGN_QUIRK:
        internal = -launch_angles  # shoot_ray() flips again (launch_rays.py:251)
        stored = internal          # Ray.launch_angle = internal angle (launch_rays.py:313-320)
    else:
        internal = launch_angles
        stored = -launch_angles    # ray.launch_angle = -launch_angles[k] (launch_rays.py:180-182)

    if device_trajectories and not small:
        dres = shoot_fan_arrays(denv, source_depth, src_i, rcv_i, angles=internal, num_range_save=num_range_save,
                                rtol=rtol, terminate_backwards=terminate_backwards, stored_convention=True,
                                ray_id_counts=True, compact=True, to_host=False)
        res = {k: (v if k in ("traj", "range_save") else v.cpu().numpy()) for k, v in dres.items()}
        return _fan_from_device(res, stored, source_depth, backwards)
    if launch_angles.shape[0] >= _PIPELINE_MIN_RAYS and num_range_save:
        res = _shoot_fan_pipelined(denv, source_depth, src_i, rcv_i, internal, num_range_save, rtol,
                                   terminate_backwards)
        return _fan_from_device(res, stored, source_depth, backwards)
    res = shoot_fan_arrays(denv, source_depth, src_i, rcv_i, angles=internal, num_range_save=num_range_save,
                           rtol=rtol, terminate_backwards=terminate_backwards, stored_convention=True,
                           ray_id_counts=True, compact=not small)
    status = res["status"]
    if small:
        for k in range(status.shape[0]):  # sequential branch: exceptions propagate, drops may print
            st = int(status[k])
            if st in _RAISING:
                exc, msg = _RAISING[st]
                raise exc(msg)
            if debug and st != 0:
                _debug_print(st)
        kept = np.flatnonzero(status == 0)
        if kept.shape[0] < status.shape[0]:
            res = {k: (v[:, kept] if k in ("final", "traj") else (v[kept] if k != "range_save" else v))
                   for k, v in res.items()}
        res["kept"] = kept
    return _fan_from_device(res, stored, source_depth, backwards)


def shoot_ray(source_depth, source_range, launch_angle, receiver_range, num_range_save, environment,
              rtol=1e-9, terminate_backwards=True, debug=True, flatearth=True):
    """Integrate one ray; same call as ``pygenray.shoot_ray``.  Returns a :class:`Ray` or ``None``.

    ``Ray.launch_angle`` holds the INTERNAL angle (= -launch_angle), as in the reference.
    """
    launch_angle = -launch_angle
    backwards = receiver_range < source_range
    denv = _device_environment(environment, flatearth, backwards)
    src_i, rcv_i = (-source_range, -receiver_range) if backwards else (source_range, receiver_range)
    res = shoot_fan_arrays(denv, source_depth, src_i, rcv_i, angles=[launch_angle],
                           num_range_save=num_range_save, rtol=rtol, terminate_backwards=terminate_backwards)
    st = int(res["status"][0])
    if st in _RAISING:
        exc, msg = _RAISING[st]
        raise exc(msg)
    if st != 0:
        if debug:
            _debug_print(st)
        return None
    r = res["range_save"]
    if backwards:
        r = -r
    return Ray(r, res["traj"][:, 0, :], int(res["n_bottom"][0]), int(res["n_surface"][0]), launch_angle,
               source_depth)
