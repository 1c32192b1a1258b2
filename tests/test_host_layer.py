"""Host-side logic (no GPU): environment adapter, result containers, C-ABI surface, failure without a GPU.

Mirrors the reference's unit tests (tests/test_environment.py, tests/test_ray_objects.py) so the
drop-in objects behave the same, then checks the boundary: libprt.so exports every symbol that
include/prt.h declares and the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pygenray_b200 as pr
from pygenray_b200 import _lib
from pygenray_b200.environment import OceanEnvironment2D, eflat, munk_ssp
from pygenray_b200.ray_objects import EigenRays, Ray, RayFan

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DA = pr.DataArray


# ---- environment (reference tests/test_environment.py) -------------------------------------------
def test_munk_ssp():
    z = np.linspace(0, 5000, 200)
    c = munk_ssp(z)
    assert c.shape == z.shape
    zf = np.linspace(0, 5000, 5001)
    assert abs(zf[np.argmin(munk_ssp(zf))] - 1300.0) <= 1.0
    assert abs(munk_ssp(np.array([1300.0]))[0] - 1500.0) < 1.0
    assert np.isfinite(munk_ssp(1300.0))


def test_environment_defaults_and_flat_earth():
    env = OceanEnvironment2D()
    for name in ("sound_speed", "bathymetry", "dcdz", "bottom_angle", "bottom_angle_interp", "latitude",
                 "sound_speed_fe", "bathymetry_fe"):
        assert hasattr(env, name), name
    assert env.sound_speed.ndim == 2 and env.sound_speed.dims == ("range", "depth")
    assert env.sound_speed.values.shape == (100, 6000)
    assert np.array_equal(env.bathymetry.values, np.linspace(4500, 4900, 100))  # the reference's default
    env2 = OceanEnvironment2D(flat_earth_transform=False)
    assert not hasattr(env2, "sound_speed_fe") and not hasattr(env2, "bathymetry_fe")
    with pytest.raises(Exception, match="Flat earth"):
        pr._unpack_envi(env2, flatearth=True)
    # flat-earth values: depth and sound speed both grow; bottom angle comes from the UN-flattened bathymetry
    assert (env.sound_speed_fe.coords["depth"].values[1:] > env.sound_speed.coords["depth"].values[1:]).all()
    assert np.allclose(env.bottom_angle, np.degrees(np.arctan(400.0 / 100e3)))
    assert abs(float(env.bottom_angle_interp(5e4)) - env.bottom_angle[0]) < 1e-12


def test_environment_validation_errors():
    z = np.arange(0.0, 3000.0, 10.0)
    bathy = DA(np.ones(20) * 4000.0, dims=["range"], coords={"range": np.linspace(0, 50e3, 20)})
    env = OceanEnvironment2D(sound_speed=DA(munk_ssp(z), dims=["depth"], coords={"depth": z}), bathymetry=bathy,
                             flat_earth_transform=False)
    assert env.sound_speed.ndim == 1 and "depth" in env.sound_speed.dims
    with pytest.raises(TypeError):
        OceanEnvironment2D(sound_speed=np.ones((3, 3)))
    with pytest.raises(ValueError):
        OceanEnvironment2D(sound_speed=DA(np.ones((2, 2, 2)), dims=["range", "depth", "t"],
                                          coords={"range": [0, 1], "depth": [0, 1], "t": [0, 1]}))
    with pytest.raises(ValueError):
        OceanEnvironment2D(sound_speed=DA(np.ones(5), dims=["x"], coords={"x": np.arange(5)}))
    with pytest.raises(ValueError):
        OceanEnvironment2D(sound_speed=DA(np.ones((5, 4)), dims=["x", "depth"],
                                          coords={"x": np.arange(5), "depth": np.arange(4)}))
    with pytest.raises(TypeError):
        OceanEnvironment2D(bathymetry=np.ones(10))
    with pytest.raises(ValueError):
        OceanEnvironment2D(bathymetry=DA(np.ones(10), dims=["x"], coords={"x": np.arange(10)}))
    with pytest.raises(ValueError):
        OceanEnvironment2D(bathymetry=DA(np.ones((2, 5)), dims=["range", "x"],
                                         coords={"range": [0, 1], "x": np.arange(5)}))


def test_eflat_transform():
    dep = np.array([100.0, 500.0, 1000.0, 2000.0, 4000.0])
    cs = np.array([1500.0, 1490.0, 1480.0, 1510.0, 1520.0])
    depf, csf = eflat(dep, 35.0, cs)
    assert (depf > dep).all() and (csf > cs).all()
    re = 6371e3  # within 0.5 %
    assert np.allclose(depf, dep * (1 + dep / re * (0.5 + dep / re / 3)), rtol=1e-5)
    assert np.allclose(csf, cs * (1 + dep / re * (1 + dep / re)), rtol=1e-5)
    d0, c0 = eflat(dep, 35.0)
    assert np.array_equal(d0, depf) and (c0 == 0).all()


def test_unpack_and_mirror():
    env = OceanEnvironment2D()
    cin, cpin, rin, zin, depths, dr, ba = pr._unpack_envi(env, flatearth=True)
    assert cin.shape == cpin.shape == (100, 6000) and rin.shape == (100,) and zin.shape == (6000,)
    assert np.array_equal(cpin, np.gradient(cin, zin, axis=1))
    m = pr._mirror_envi_arrays(cin, cpin, rin, depths, dr, ba)
    assert np.array_equal(m[2], -rin[::-1]) and np.array_equal(m[0], cin[::-1]) and np.array_equal(m[5], -ba[::-1])
    assert (np.diff(m[2]) > 0).all()
    with pytest.raises(ValueError):
        pr.bottom_angle_ppoly(np.arange(3.0), np.zeros(3))  # interp1d(kind="cubic") needs 4 points


# ---- containers (reference tests/test_ray_objects.py) --------------------------------------------
def _make_ray(launch_angle, source_depth, n_bottom=0, n_surface=0, N=10, R=10000.0):
    r = np.linspace(0.0, R, N)
    y = np.vstack([r / 1500.0, np.linspace(source_depth, source_depth + R * 0.01, N),
                   np.ones(N) * np.sin(np.radians(abs(launch_angle) + 1e-3)) / 1500.0])
    return Ray(r=r, y=y, n_bottom=n_bottom, n_surface=n_surface, launch_angle=launch_angle, source_depth=source_depth)


def _fan():
    return RayFan([_make_ray(-5.0, 100.0), _make_ray(5.0, 150.0, n_bottom=1), _make_ray(-10.0, 200.0)])


def test_ray_conventions():
    ray = _make_ray(-10.0, 100.0, n_bottom=2, n_surface=3)
    assert ray.r.shape == ray.t.shape == ray.z.shape == ray.p.shape == (10,)
    assert (ray.z <= 0).all() and (ray.p < 0).all()
    assert ray.launch_angle == -10.0 and ray.source_depth == 100.0 and (ray.n_bottom, ray.n_surface) == (2, 3)
    bare = Ray(np.arange(3.0), np.zeros((3, 3)), 0, 0)
    assert not hasattr(bare, "launch_angle") and not hasattr(bare, "source_depth")


def test_rayfan_shapes_ids_indexing():
    fan = _fan()
    assert fan.thetas.shape == (3,) and fan.rs.shape == fan.ts.shape == fan.zs.shape == fan.ps.shape == (3, 10)
    assert len(fan) == 3 and len(fan.ray_ids) == 3 and all(isinstance(s, str) for s in fan.ray_ids)
    assert list(fan.ray_ids) == ["-0.0", "0.0b", "-0.0"]  # str(float) ids, 'b' suffix when a ray bounced
    assert isinstance(fan[0], Ray) and fan[1].launch_angle == 5.0 and fan[-1].launch_angle == -10.0
    assert np.array_equal(fan[0].z, fan.zs[0])  # __getitem__ re-negates
    with pytest.raises(IndexError):
        fan[3]
    sub = fan[0:2]
    assert isinstance(sub, RayFan) and np.array_equal(sub.thetas, [-5.0, 5.0]) and np.array_equal(sub.zs, fan.zs[:2])
    mask = fan[np.array([True, False, True])]
    assert np.array_equal(mask.thetas, [-5.0, -10.0])
    assert np.array_equal(fan[np.array([2, 0])].thetas, [-10.0, -5.0])
    with pytest.raises(ValueError):
        fan[np.zeros((2, 2), dtype=int)]


def test_rayfan_add_and_save_mat(tmp_path):
    a, b = _fan(), _fan()
    c = a + b
    assert len(c) == 6 and np.array_equal(c.rs[0], a.rs[0])
    assert np.array_equal(c.zs[:3], -a.zs)  # the reference does not re-negate in __add__
    with pytest.raises(TypeError):
        a + 1
    other = RayFan([_make_ray(-5.0, 100.0, R=20000.0)])
    with pytest.raises(ValueError):
        a + other
    from scipy import io

    path = str(tmp_path / "fan.mat")
    a.save_mat(path)
    d = io.loadmat(path)["rayfan"]
    for key in ("thetas", "xs", "ts", "zs", "ps", "n_botts", "n_surfs", "source_depths"):
        assert key in d.dtype.names
    assert np.allclose(d["ts"][0, 0], a.ts)


def test_eigenrays_container_without_rays_and_no_cpu_path():
    """Receivers without eigenrays give empty arrays (the reference raises an AxisError there); received angles are a
    device computation, so a container WITH rays cannot be built on a machine without a GPU."""
    env = OceanEnvironment2D(flat_earth_transform=False)
    er = EigenRays([1000.0], {0: []}, env, {1000.0: 0}, {0: 0}, {0: []})
    assert er.received_angles[0].shape == (0,) and er.ray_id[0].shape == (0,) and er.ray_id_int[0].shape == (0,)
    if _lib.load().prt_device_count() == 0:
        with pytest.raises(_lib.PrtError, match="no CUDA device"):
            EigenRays([1000.0], {0: [_make_ray(-5.0, 100.0)]}, env, {1000.0: 1}, {0: 1}, {0: []})


# ---- C ABI + no CPU fallback -----------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "prt.h")).read()
    declared = sorted(set(re.findall(r"PRT_API[^;(]*?\b(prt_\w+)\s*\(", header)))
    assert len(declared) >= 14
    assert sorted(_lib.EXPORTS) == declared, "pygenray_b200/_lib.py and include/prt.h disagree"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libprt.so does not export {name}"
    assert _lib.load().prt_abi_version() == 2
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln}
    assert {s for s in exported if s.startswith("prt_")} == set(declared)  # nothing else leaks out


def test_library_is_sm100a_and_has_no_oracle_symbols():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    nm = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in nm
    for mod in ("launch_rays.py", "eigenrays.py", "environment.py", "ray_objects.py", "_lib.py", "dist.py", "__init__.py"):
        src = open(os.path.join(ROOT, "pygenray_b200", mod)).read()
        assert "import oracle" not in src and "from oracle" not in src, mod


@pytest.mark.skipif(_lib.load().prt_device_count() > 0, reason="a GPU is present")
def test_compute_calls_fail_loudly_without_gpu():
    env = OceanEnvironment2D()
    with pytest.raises(_lib.PrtError, match="no CUDA device"):
        pr.shoot_ray(1000.0, 0.0, 5.0, 1e4, 11, env)
    with pytest.raises(_lib.PrtError):
        pr.shoot_rays(1000.0, 0.0, np.linspace(-5, 5, 80), 1e4, 11, env)
    with pytest.raises(_lib.PrtError):
        pr.DeviceEnvironment(*pr._unpack_envi(env))
    h = C.c_void_p()
    x = np.zeros(4)
    rc = _lib.load().prt_env_create(0, 2, 2, x.ctypes.data, x.ctypes.data, x.ctypes.data, x.ctypes.data, 2,
                                    x.ctypes.data, x.ctypes.data, 2, x.ctypes.data, x.ctypes.data, C.byref(h))
    assert rc == 3 and b"no CPU path" in _lib.load().prt_last_error()


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-rays", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ray_km_per_s" and line["unit"] == "ray-km/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_numa_binding_is_best_effort():
    """dist.bind_to_gpu_numa never fails a run: cpulist parsing, and None when the topology is not visible."""
    from pygenray_b200 import dist

    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("") == set()
    assert dist.bind_to_gpu_numa(0) is None or isinstance(dist.bind_to_gpu_numa(0), int)


def test_eflat_and_eflatinv_against_reference_values():
    """eflat bit-identical to the reference; eflatinv (Newton on the cubic instead of the reference's Ridder search
    to 1 mm, environment.py:404-535) within the reference's own accuracy of its output and exact against the truth."""
    from pygenray_b200.environment import eflatinv

    g = golden("ref_flatearth_jit.npz")
    dep = g["dep"]
    for lat in g["lats"]:
        depf, csf = eflat(dep, lat, 1480.0 + 0.017 * dep)
        assert np.array_equal(depf, g[f"depf_{lat}"]) and np.array_equal(csf, g[f"csf_{lat}"])
        d, c = eflatinv(depf, lat, csf)
        assert np.abs(d - g[f"dinv_{lat}"]).max() < 2e-3 and np.abs(c - g[f"cinv_{lat}"]).max() < 1e-6
        assert np.abs(d - dep).max() < 1e-9 and np.abs(c - (1480.0 + 0.017 * dep)).max() < 1e-9
    d, c = eflatinv(np.array([100.0, 2000.0]), np.array([10.0, 50.0]))
    assert d.shape == (2,) and (c == 0).all()


def test_flat_earth_c_restatement_against_reference_output():
    """The numpy restatement the device kernel is compared with (conftest.flat_earth_c_numpy) reproduces the
    reference's own flat_earth_c (ref:environment.py:239-303; golden made by oracle/make_golden.py: case_flatearth)
    on a 37 x 301 grid with per-column latitudes and on an irregular depth grid that starts below the surface."""
    from conftest import flat_earth_c_numpy
    import workloads as cases

    g = golden("ref_flatearth_jit.npz")
    a = cases.config2_arrays(n_r=37, n_z=301)
    assert hashlib.sha256(np.ascontiguousarray(a["c"]).tobytes()).hexdigest() == str(g["fec_c_sha"])
    mine = flat_earth_c_numpy(a["c"], a["z"], g["fec_lat"])
    assert not np.isnan(g["fec_out"]).any() and np.array_equal(mine, g["fec_out"])
    mine2 = flat_earth_c_numpy(g["fec2_c"], g["fec2_z"], g["fec2_lat"])
    assert np.isnan(g["fec2_out"][:, 0]).all() and np.isnan(g["fec2_out"]).sum() == 9  # above the first flattened node
    assert np.array_equal(mine2, g["fec2_out"], equal_nan=True)


def test_device_rows_indexing_semantics():
    """DeviceRows behaves like the numpy array it stands for (exercised here on a CPU tensor)."""
    import torch

    from pygenray_b200.ray_objects import DeviceRows

    a = np.arange(60, dtype=np.float64).reshape(6, 10)
    d = DeviceRows(torch.from_numpy(a.copy()))
    assert d.shape == (6, 10) and d.ndim == 2 and len(d) == 6 and d.dtype == np.float64 and d.nbytes == 480
    assert np.array_equal(d[:, -1], a[:, -1]) and np.array_equal(d[2], a[2]) and d[3, 4] == a[3, 4]
    assert np.array_equal(d[1:5, ::3], a[1:5, ::3]) and np.array_equal(d[np.array([4, 0])], a[[4, 0]])
    assert np.array_equal(d[np.array([True, False, True, False, False, True])], a[[0, 2, 5]])
    assert d._host is None and d.min() == 0.0 and d.max() == 59.0
    assert np.array_equal(d[::-1], a[::-1]) and d._host is not None  # negative step: full fetch, then cached
    assert np.array_equal(np.asarray(d), a) and np.array_equal(d[:, 0], a[:, 0])


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """include/prt.h compiles as C99 with -Wall -Werror and a C program links against libprt.so (the drop-in boundary
    is a C ABI: plain pointers and sizes, no C++ or torch types)."""
    exe = str(tmp_path / "abi_check")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_harness", "abi_check.c"), "-o", exe, "-L", lib_dir, "-l:libprt.so",
           "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert p.stdout.startswith("abi 2 ok")
