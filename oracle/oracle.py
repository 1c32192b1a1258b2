"""ctypes binding of oracle/libprt_oracle.so (the plain-C restatement, prt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  pygenray_b200/ never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libprt_oracle.so")

STATUS_NAMES = {
    0: "OK", 1: "STEP_TOO_SMALL", 2: "VERTICAL", 3: "BBOX", 4: "BACKWARDS", 5: "BOTTOM_OOB",
    6: "BRENTQ_SIGN", 7: "BRENTQ_ITER", 8: "TS_NOT_INCREASING", 9: "MAX_BOUNCES",
}
COUNTER_NAMES = ("nfev", "attempts", "accepted", "segments", "bisect", "dense")
N_COUNTERS = len(COUNTER_NAMES)
DEFAULT_MAX_BOUNCES = 1 << 20


def build(force=False):
    """Compile prt_oracle.c with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "prt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class _Env(C.Structure):
    _fields_ = [
        ("n_r", C.c_int32), ("n_z", C.c_int32), ("n_b", C.c_int32), ("n_ba", C.c_int32),
        ("rin", C.c_void_p), ("zin", C.c_void_p), ("cin", C.c_void_p), ("cpin", C.c_void_p),
        ("depth_ranges", C.c_void_p), ("depths", C.c_void_p), ("ba_x", C.c_void_p), ("ba_c", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        d, i32, i64, vp = C.c_double, C.c_int32, C.c_int64, C.c_void_p
        L.orc_pow_m02.restype = d; L.orc_pow_m02.argtypes = [d]
        L.orc_sin.restype = d; L.orc_sin.argtypes = [d]
        L.orc_asin.restype = d; L.orc_asin.argtypes = [d]
        L.orc_bilinear.restype = d; L.orc_bilinear.argtypes = [vp, d, d, C.c_int]
        L.orc_derivsrd.restype = None; L.orc_derivsrd.argtypes = [vp, d, vp, vp]
        L.orc_event.restype = d; L.orc_event.argtypes = [vp, C.c_int, d, vp]
        L.orc_bottom_angle.restype = d; L.orc_bottom_angle.argtypes = [vp, d, vp]
        L.orc_brentq_step.restype = d; L.orc_brentq_step.argtypes = [d, d, d, vp]
        L.orc_initial_state.restype = None; L.orc_initial_state.argtypes = [vp, d, d, d, vp]
        L.orc_shoot_ray.restype = C.c_int
        L.orc_shoot_ray.argtypes = [vp, vp, d, d, d, d, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        L.orc_shoot_fan.restype = C.c_int
        L.orc_shoot_fan.argtypes = [vp, i64, vp, vp, d, d, d, d, d, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                    vp, vp, C.c_int, C.c_int]
        L.orc_refine_brackets.restype = C.c_int
        L.orc_refine_brackets.argtypes = [vp, i64, vp, vp, vp, vp, vp, d, d, d, d, d, C.c_int, d, C.c_int,
                                          vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """The seven arrays of ref:launch_rays.py:717-742 (_unpack_envi) + the bottom-angle PPoly."""

    def __init__(self, cin, cpin, rin, zin, depths, depth_ranges, ba_x, ba_c):
        self.cin, self.cpin = _f64(cin), _f64(cpin)
        self.rin, self.zin = _f64(rin), _f64(zin)
        self.depths, self.depth_ranges = _f64(depths), _f64(depth_ranges)
        self.ba_x, self.ba_c = _f64(ba_x), _f64(ba_c)
        assert self.cin.shape == (len(self.rin), len(self.zin)) == self.cpin.shape
        assert self.ba_c.shape == (4, len(self.ba_x) - 1)
        self._s = _Env(len(self.rin), len(self.zin), len(self.depths), len(self.ba_x),
                       self.rin.ctypes.data, self.zin.ctypes.data, self.cin.ctypes.data,
                       self.cpin.ctypes.data, self.depth_ranges.ctypes.data, self.depths.ctypes.data,
                       self.ba_x.ctypes.data, self.ba_c.ctypes.data)

    @property
    def ptr(self):
        return C.cast(C.pointer(self._s), C.c_void_p)


def shoot_fan(env, source_depth, source_range, receiver_range, angles_deg=None, y0s=None, rtol=1e-9,
              atol=1e-6, terminate_backwards=True, range_save=None, n_threads=0,
              max_bounces=DEFAULT_MAX_BOUNCES, counters=True):
    """Shoot a fan with INTERNAL launch angles (or explicit y0s).  Returns a dict of arrays."""
    L = lib()
    if y0s is not None:
        y0s = _f64(y0s); n = y0s.shape[0]; ang = None
    else:
        ang = _f64(np.atleast_1d(angles_deg)); n = ang.shape[0]
    traj = None; rs = None; n_save = 0
    if range_save is not None:
        rs = _f64(range_save); n_save = len(rs)
        traj = np.empty((n, 3, n_save))
    final = np.empty((n, 3)); nb = np.zeros(n, np.int32); ns = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32)
    cnt = np.zeros((n, N_COUNTERS), np.int64) if counters else None
    L.orc_shoot_fan(env.ptr, n, _p(ang), _p(y0s), source_depth, source_range, receiver_range, rtol, atol,
                    int(bool(terminate_backwards)), n_save, _p(rs), _p(traj), _p(final), _p(nb), _p(ns),
                    _p(st), _p(cnt), int(max_bounces), int(n_threads))
    return dict(traj=traj, final=final, n_bottom=nb, n_surface=ns, status=st, counters=cnt)


def refine_brackets(env, z1s, z2s, th1s, th2s, rds, source_depth, source_range, receiver_range, rtol=1e-9,
                    atol=1e-6, terminate_backwards=True, ztol=1.0, max_iter=20, n_threads=0,
                    max_bounces=DEFAULT_MAX_BOUNCES):
    L = lib()
    z1s, z2s, th1s, th2s, rds = map(_f64, (z1s, z2s, th1s, th2s, rds))
    n = len(z1s)
    theta = np.empty(n); final = np.empty((n, 3))
    nb = np.zeros(n, np.int32); ns = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
    shots = np.zeros(n, np.int32); rst = np.zeros(n, np.int32)
    L.orc_refine_brackets(env.ptr, n, _p(z1s), _p(z2s), _p(th1s), _p(th2s), _p(rds), source_depth,
                          source_range, receiver_range, rtol, atol, int(bool(terminate_backwards)), ztol,
                          int(max_iter), _p(theta), _p(final), _p(nb), _p(ns), _p(st), _p(shots), _p(rst),
                          int(max_bounces), int(n_threads))
    return dict(theta=theta, final=final, n_bottom=nb, n_surface=ns, status=st, shots=shots,
                ray_status=rst)


def num_threads():
    return lib().orc_num_threads()
