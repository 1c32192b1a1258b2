"""Shared fixtures.  `-m gpu` tests need a B200; everything else runs on CPU (oracle, host logic,
C-ABI symbol checks, and the host build of the per-ray device routine)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libprt.so (nvcc cross-compiles without a GPU), the C oracle and the host harness once."""
    import __graft_entry__ as ge

    ge.build()


def pytest_sessionfinish(session, exitstatus):
    """Bounds-check builds of libprt.so (-DPRT_BOUNDS_CHECK, loaded through PRT_LIB_PATH) count out-of-range table
    indices on the device; report the count of the whole session (profiles/r02_bounds_check.md)."""
    try:
        from pygenray_b200 import _lib

        lib = _lib._lib
        if lib is not None and hasattr(lib, "prt_debug_oob"):
            lib.prt_debug_oob.restype = C.c_longlong
            n = int(lib.prt_debug_oob())
            line = f"PRT_BOUNDS_CHECK: {n} out-of-range table indices in this session (exit status {int(exitstatus)})"
            print("\n" + line)
            out = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(out):
                with open(os.path.join(out, "r02_bounds_check.txt"), "a") as f:
                    f.write(line + "\n")
    except Exception:  # noqa: BLE001
        pass


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def flat_earth_c_numpy(c, depth, lat):
    """flat_earth_c (ref:environment.py:239-303) restated: per range column eflat with its own latitude, then xarray's
    linear interp back onto the original depth grid (scipy interp1d, bounds_error=False -> NaN outside).  Pinned
    against the reference's own output by tests/test_oracle_vs_reference.py::test_flat_earth_c_restatement."""
    from scipy.interpolate import interp1d

    import pygenray_b200 as pr

    out = np.empty_like(c)
    for r in range(c.shape[0]):
        depf, cf = pr.eflat(depth, lat[r], c[r])
        out[r] = interp1d(depf, cf, kind="linear", bounds_error=False)(depth)
    return out


def make_env(arrays, flat=False, lat=35):
    import pygenray_b200 as pr

    ssp = pr.DataArray(arrays["c"], dims=["range", "depth"], coords={"range": arrays["r"], "depth": arrays["z"]})
    bathy = pr.DataArray(arrays["bathy"], dims=["range"], coords={"range": arrays["bathy_r"]})
    return pr.OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=lat, flat_earth_transform=flat)


def oracle_env_from(env, flatearth, mirror=False):
    """OracleEnv from an OceanEnvironment2D via the product's host-side unpacking."""
    import pygenray_b200 as pr
    from oracle import oracle as orc

    cin, cpin, rin, zin, depths, dr, ba = pr._unpack_envi(env, flatearth=flatearth)
    if mirror:
        cin, cpin, rin, depths, dr, ba = pr._mirror_envi_arrays(cin, cpin, rin, depths, dr, ba)
    ba_x, ba_c = pr.bottom_angle_ppoly(dr, ba)
    return orc.OracleEnv(cin, cpin, rin, zin, depths, dr, ba_x, ba_c)


class HostHarness:
    """g++ build of pygenray_b200/csrc/prt_core.h (tests/host_harness) -- CPU-side bit parity only."""

    def __init__(self):
        import __graft_entry__ as ge

        self.lib = C.CDLL(ge.build_host_harness())
        vp, d = C.c_void_p, C.c_double
        self.lib.hh_env_create.restype = vp
        self.lib.hh_env_create.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]
        self.lib.hh_env_destroy.argtypes = [vp]
        self.lib.hh_env_flags.argtypes = [vp, vp]
        self.lib.hh_gradient.argtypes = [vp, C.c_int, vp, vp]
        self.lib.hh_shoot_fan.argtypes = [vp, C.c_int64, vp, vp, d, d, d, d, d, C.c_int, C.c_int, vp, vp, vp, vp,
                                          vp, vp, vp, C.c_int]

    def env_flags(self, oenv):
        """(z_cubic, spec_off): how the staging code rated the cell guesses of this environment."""
        e = oenv
        h = self.lib.hh_env_create(len(e.rin), len(e.zin), e.rin.ctypes.data, e.zin.ctypes.data, e.cin.ctypes.data,
                                   e.cpin.ctypes.data, len(e.depths), e.depth_ranges.ctypes.data,
                                   e.depths.ctypes.data, len(e.ba_x), e.ba_x.ctypes.data, e.ba_c.ctypes.data)
        out = (C.c_int * 3)()
        self.lib.hh_env_flags(h, out)
        self.lib.hh_env_destroy(h)
        self.last_z_uniform = int(out[2]) & 1   # exactly uniform depth grid: cell edges computed instead of loaded
        self.last_r_uniform = int(out[2]) >> 1  # ... and exactly uniform range grid
        return int(out[0]), int(out[1])

    def shoot_fan(self, oenv, source_depth, source_range, receiver_range, angles, range_save=None, rtol=1e-9,
                  atol=1e-6, terminate_backwards=True, max_bounces=1 << 20):
        """Same inputs as oracle.shoot_fan (takes the OracleEnv's arrays); outputs in oracle layout."""
        e = oenv
        h = self.lib.hh_env_create(len(e.rin), len(e.zin), e.rin.ctypes.data, e.zin.ctypes.data, e.cin.ctypes.data,
                                   e.cpin.ctypes.data, len(e.depths), e.depth_ranges.ctypes.data,
                                   e.depths.ctypes.data, len(e.ba_x), e.ba_x.ctypes.data, e.ba_c.ctypes.data)
        assert h
        ang = np.ascontiguousarray(angles, dtype=np.float64)
        n = len(ang)
        n_save = 0 if range_save is None else len(range_save)
        rs = None if range_save is None else np.ascontiguousarray(range_save, dtype=np.float64)
        traj = np.empty((3, n, n_save)) if n_save else None
        fin = np.empty((3, n)); nb = np.zeros(n, np.int32); ns = np.zeros(n, np.int32)
        st = np.zeros(n, np.int32); cnt = np.zeros((6, n), np.uint32)
        self.lib.hh_shoot_fan(h, n, ang.ctypes.data, None, source_depth, source_range, receiver_range, rtol, atol,
                              int(terminate_backwards), n_save, None if rs is None else rs.ctypes.data,
                              None if traj is None else traj.ctypes.data, fin.ctypes.data, nb.ctypes.data,
                              ns.ctypes.data, st.ctypes.data, cnt.ctypes.data, max_bounces)
        self.lib.hh_env_destroy(h)
        return dict(traj=None if traj is None else np.transpose(traj, (1, 0, 2)), final=fin.T.copy(), n_bottom=nb,
                    n_surface=ns, status=st, counters=cnt.T.astype(np.int64))


@pytest.fixture(scope="session")
def harness():
    return HostHarness()


def assert_same_fan(a, b, what="", counters=("nfev", "attempts", "accepted", "segments", "bisect")):
    """Bit-for-bit equality of two fan results (oracle-layout dicts)."""
    from oracle import oracle as orc

    assert np.array_equal(a["status"], b["status"]), f"{what}: status differs"
    assert np.array_equal(a["n_bottom"], b["n_bottom"]), f"{what}: bottom bounce counts differ"
    assert np.array_equal(a["n_surface"], b["n_surface"]), f"{what}: surface bounce counts differ"
    assert np.array_equal(a["final"], b["final"], equal_nan=True), f"{what}: arrival table differs"
    if a.get("traj") is not None and b.get("traj") is not None:
        assert np.array_equal(a["traj"], b["traj"], equal_nan=True), f"{what}: trajectories differ"
    if a.get("counters") is not None and b.get("counters") is not None:
        for name in counters:
            k = orc.COUNTER_NAMES.index(name)
            assert np.array_equal(a["counters"][:, k], b["counters"][:, k]), f"{what}: counter {name} differs"


def oracle_shoot_ray(env, source_depth, source_range, launch_angle, receiver_range, n_save, rtol=1e-9,
                     flatearth=True, terminate_backwards=True):
    """shoot_ray semantics (sign flip, backwards mirroring, stored conventions) evaluated by the oracle.

    Returns None for a dropped ray, else dict(r, t, z, p, n_bottom, n_surface, launch_angle)."""
    from oracle import oracle as orc

    internal = -launch_angle
    backwards = receiver_range < source_range
    oenv = oracle_env_from(env, flatearth, mirror=backwards)
    src, rcv = (-source_range, -receiver_range) if backwards else (source_range, receiver_range)
    rs = np.linspace(src, rcv, n_save)
    o = orc.shoot_fan(oenv, source_depth, src, rcv, angles_deg=[internal], range_save=rs, rtol=rtol,
                      terminate_backwards=terminate_backwards, n_threads=1)
    if o["status"][0] != 0:
        return None
    tr = o["traj"][0]
    return dict(r=-rs if backwards else rs, t=tr[0], z=-tr[1], p=-tr[2], n_bottom=int(o["n_bottom"][0]),
                n_surface=int(o["n_surface"][0]), launch_angle=internal)


def ray_angle_numpy(x, y, cin, rin, zin):
    """Checker for prt_received_angles: ray_angle + bilinear_interp of the reference restated in numpy scalars
    (/root/reference/src/pygenray/integration_processes.py:101-174, 306-334): cell = searchsorted - 1 clamped to
    [0, n - 2], unclamped weights, four-term sum, degrees(arcsin(p c))."""
    i = int(np.searchsorted(rin, x)) - 1
    j = int(np.searchsorted(zin, y[1])) - 1
    i = max(0, min(i, len(rin) - 2))
    j = max(0, min(j, len(zin) - 2))
    wx = (x - rin[i]) / (rin[i + 1] - rin[i])
    wy = (y[1] - zin[j]) / (zin[j + 1] - zin[j])
    c = ((1 - wx) * (1 - wy) * cin[i, j] + wx * (1 - wy) * cin[i + 1, j] + (1 - wx) * wy * cin[i, j + 1]
         + wx * wy * cin[i + 1, j + 1])
    with np.errstate(invalid="ignore"):
        return np.degrees(np.arcsin(y[2] * c)), c
