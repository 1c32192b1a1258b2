"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference (/root/reference) here.

Run in the build container only (the reference is not shipped to the GPU box):

    python oracle/make_golden.py            # all cases, JIT variant (+ no-JIT where noted)

The reference needs xarray + matplotlib, which this image lacks; oracle/stubs/ provides
minimal stand-ins.  Nothing here is product code.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SRC = "/root/reference/src"

if __name__ == "__main__" and os.environ.get("_PRT_GOLDEN_CHILD") != "1":
    # re-exec with the right PYTHONPATH / numba settings, once with JIT and once without
    for variant in ("jit", "nojit"):
        env = dict(os.environ)
        env["_PRT_GOLDEN_CHILD"] = "1"
        env["_PRT_GOLDEN_VARIANT"] = variant
        env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REF_SRC, ROOT])
        env["NUMBA_CACHE_DIR"] = "/tmp/numba_cache_prt"
        if variant == "nojit":
            env["NUMBA_DISABLE_JIT"] = "1"
        subprocess.check_call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env)
    sys.exit(0)

if sys.path and os.path.abspath(sys.path[0]) == HERE:
    sys.path[0] = ROOT  # `import oracle` must find the package, not oracle/oracle.py

import numpy as np  # noqa: E402
import xarray as xr  # noqa: E402  (the stub)
import pygenray as pr  # noqa: E402  (the reference)
from pygenray.environment import OceanEnvironment2D, munk_ssp  # noqa: E402
from pygenray.launch_rays import _interpolate_ray, _shoot_ray_array, _unpack_envi  # noqa: E402

import workloads as cases  # noqa: E402

VARIANT = os.environ.get("_PRT_GOLDEN_VARIANT", "jit")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_env(a, flat=False, lat=35):
    ssp = xr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
    bathy = xr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
    return OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=lat, flat_earth_transform=flat)


def array_level(env, flat, z_src, r_src, r_rcv, internal_angles, n_save, rtol=1e-9, terminate_backwards=True):
    """Drive ref _shoot_ray_array + _interpolate_ray directly (forward shots), recording solver stats."""
    cin, cpin, rin, zin, depths, dr, ba = _unpack_envi(env, flatearth=flat)
    c0 = pr.bilinear_interp(r_src, z_src, rin, zin, cin)
    rs = np.linspace(r_src, r_rcv, n_save)
    n = len(internal_angles)
    traj = np.full((n, 3, n_save), np.nan)
    final = np.full((n, 3), np.nan)
    nb = np.zeros(n, np.int64); ns = np.zeros(n, np.int64)
    ok = np.zeros(n, bool); nfev = np.zeros(n, np.int64); nsteps = np.zeros(n, np.int64)
    nseg = np.zeros(n, np.int64); raised = np.zeros(n, bool)
    for k, a in enumerate(internal_angles):
        y0 = np.array([0, z_src, np.sin(np.radians(a)) / c0])
        try:
            sols, full, b, s = _shoot_ray_array(y0, z_src, r_src, r_rcv, cin, cpin, rin, zin, depths, dr, ba,
                                                rtol, terminate_backwards, False)
        except Exception:
            sols = None
            raised[k] = True
        if sols is None:
            continue
        ok[k] = True
        fr = _interpolate_ray(sols, rs)
        traj[k] = fr[1:]
        final[k] = sols[-1].y[:, -1]
        nb[k], ns[k] = b, s
        nfev[k] = sum(x.nfev for x in sols)
        nsteps[k] = sum(len(x.sol.interpolants) for x in sols)
        nseg[k] = len(sols)
    return dict(traj=traj, final=final, n_bottom=nb, n_surface=ns, ok=ok, nfev=nfev, nsteps=nsteps,
                nseg=nseg, raised=raised, range_save=rs, internal_angles=np.asarray(internal_angles, float))


def fan_dict(rf, prefix="fan_"):
    return {prefix + "thetas": rf.thetas, prefix + "rs": rf.rs, prefix + "ts": rf.ts, prefix + "zs": rf.zs,
            prefix + "ps": rf.ps, prefix + "n_botts": rf.n_botts, prefix + "n_surfs": rf.n_surfs,
            prefix + "ray_ids": rf.ray_ids, prefix + "source_depths": rf.source_depths}


def save(name, **kw):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, f"{name}_{VARIANT}.npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def case_cfg1():
    """BASELINE config 1: default env (flat-earth ON), src 1000 m, 101 angles +-12 deg, 100 km."""
    env = OceanEnvironment2D()
    cin, cpin, rin, zin, depths, dr, ba = _unpack_envi(env, flatearth=True)
    angles = np.linspace(-12, 12, 101)
    out = {}
    if VARIANT == "jit":
        rf = pr.shoot_rays(1000.0, 0.0, angles, 100e3, 101, env, rtol=1e-9, n_processes=8, debug=False)
        out.update(fan_dict(rf))
    al = array_level(env, True, 1000.0, 0.0, 100e3, -angles, 101)
    out.update({"al_" + k: v for k, v in al.items()})
    # env pins: one column of c / dcdz (range independent), coordinates, bathymetry, hashes of the full arrays
    out.update(cin_col=cin[0], cpin_col=cpin[0], rin=rin, zin=zin, depths=depths, depth_ranges=dr,
               bottom_angles=ba, cin_sha=sha(cin), cpin_sha=sha(cpin), user_angles=angles)
    save("ref_cfg1", **out)


def case_fixture():
    """The reference's own golden test E (tests/test_physics.py:305-386): 5 rays, double-flip branch."""
    z = np.linspace(0.0, 6000.0, 400); r = np.linspace(0.0, 50e3, 30)
    a = dict(c=np.outer(np.ones(30), munk_ssp(z)), r=r, z=z, bathy=np.full(30, 5000.0), bathy_r=r)
    env = make_env(a)
    angles = [-8.0, -4.0, 0.0, 4.0, 8.0]
    rf = pr.shoot_rays(1300.0, 0.0, angles, 50e3, 50, env, n_processes=1, debug=False, flatearth=False)
    fx = np.load("/root/reference/tests/fixtures/munk_regression.npz")
    out = fan_dict(rf)
    out.update({"fixture_" + k: fx[k] for k in fx.files})
    save("ref_fixture", **out)


def case_cfg2_sub():
    """BASELINE config 2 sub-sample: 200x500 range-dependent grid + seamount, 31 rays +-15 deg, 500 km."""
    a = cases.config2_arrays()
    env = make_env(a)
    angles = np.linspace(-15, 15, 31)
    al = array_level(env, False, 1000.0, 0.0, 500e3, -angles, 201)
    cin, cpin, rin, zin, depths, dr, ba = _unpack_envi(env, flatearth=False)
    save("ref_cfg2_sub", cin_sha=sha(cin), cpin_sha=sha(cpin), bottom_angles=ba, user_angles=angles,
         **{"al_" + k: v for k, v in al.items()})


def case_cfg5_sub():
    """BASELINE config 5 sub-sample: two members of the tomography ensemble (config-2 grid + internal-wave perturbation,
    seeds 1000 + i), 21 of the 4096 launch angles each, 500 km."""
    out = {}
    angles = np.linspace(-15, 15, 4096)[::204][:21]
    for i in (0, 7):
        a = cases.config5_realisation(i)
        env = make_env(a)
        al = array_level(env, False, 1000.0, 0.0, 500e3, -angles, 101)
        out.update({f"m{i}_al_" + k: v for k, v in al.items()})
        out[f"m{i}_c_sha"] = sha(a["c"])
    save("ref_cfg5_sub", user_angles=angles, **out)


def case_fuzz_short():
    """40 random environments of tests/fuzz_cases.py: short_case (2-30 km): irregular / regular grids, noisy sound speeds,
    flat-earth on / off, offset sources, rtol 1e-6 .. 1e-11 -- the reference ray by ray."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fuzz_cases import short_case

    out = {}
    for seed in range(40):
        k = short_case(seed)
        a = k["arrays"]
        ssp = xr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
        bathy = xr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
        env = OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=k["lat"], flat_earth_transform=k["flat"])
        al = array_level(env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                         terminate_backwards=k["terminate_backwards"])
        out.update({f"s{seed}_" + kk: v for kk, v in al.items()})
    save("ref_fuzz_short", **out)


def case_fuzz_smooth():
    """24 random smooth environments on 1-2 m depth grids (tests/fuzz_cases.py: smooth_case): the regime in which the
    reference reproduces itself, with reflections off sloping / undulating bottoms."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fuzz_cases import smooth_case

    out = {}
    for seed in range(24):
        k = smooth_case(seed)
        a = k["arrays"]
        ssp = xr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
        bathy = xr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
        env = OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=k["lat"], flat_earth_transform=k["flat"])
        al = array_level(env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                         terminate_backwards=k["terminate_backwards"])
        out.update({f"s{seed}_" + kk: v for kk, v in al.items()})
    save("ref_fuzz_smooth", **out)


def case_bounce_fine():
    """Default 1 m grid env, steep rays that bounce (stable-step regime with reflections)."""
    env = OceanEnvironment2D()
    internal = np.array([-30, -25, -20, -16, -14, 14, 16, 20, 25, 30.0])
    al = array_level(env, True, 1000.0, 0.0, 100e3, internal, 401)
    save("ref_bounce_fine", **{"al_" + k: v for k, v in al.items()})


def case_physics():
    """Reference tests A-D, F, G inputs run through the public API (tests/test_physics.py)."""
    out = {}

    def const_env(c0=1500.0, z_max=5000.0, r_max=100e3, bathy_depth=4500.0, nz=200, nr=20):
        z = np.linspace(0.0, z_max, nz); r = np.linspace(0.0, r_max, nr)
        return make_env(dict(c=np.full((nr, nz), c0), r=r, z=z, bathy=np.full(nr, bathy_depth), bathy_r=r))

    def lin_env(c0=1500.0, g=0.05, z_max=5000.0, r_max=100e3, bathy_depth=4500.0, nz=500, nr=50):
        z = np.linspace(0.0, z_max, nz); r = np.linspace(0.0, r_max, nr)
        return make_env(dict(c=np.outer(np.ones(nr), c0 + g * z), r=r, z=z, bathy=np.full(nr, bathy_depth),
                             bathy_r=r))

    def munk_env(r_max=100e3, nr=50, nz=600, bathy_depth=5000.0):
        z = np.linspace(0.0, 6000.0, nz); r = np.linspace(0.0, r_max, nr)
        return make_env(dict(c=np.outer(np.ones(nr), munk_ssp(z)), r=r, z=z, bathy=np.full(nr, bathy_depth),
                             bathy_r=r))

    def put(tag, ray):
        out[tag + "_none"] = np.array(ray is None)
        if ray is not None:
            out[tag + "_r"] = ray.r; out[tag + "_t"] = ray.t; out[tag + "_z"] = ray.z; out[tag + "_p"] = ray.p
            out[tag + "_nb"] = np.array(ray.n_bottom); out[tag + "_ns"] = np.array(ray.n_surface)
            out[tag + "_la"] = np.array(ray.launch_angle)

    for ang in (-5.0, -10.0, -15.0):  # A
        put(f"A_{int(-ang)}", pr.shoot_ray(200.0, 0.0, ang, 30e3, 60, const_env(), rtol=1e-9, flatearth=False,
                                           debug=False))
    put("B", pr.shoot_ray(200.0, 0.0, -10.0, 20e3, 50, const_env(r_max=21e3), rtol=1e-9, flatearth=False,
                          debug=False))
    put("C", pr.shoot_ray(200.0, 0.0, -20.0, 80e3, 400, lin_env(), rtol=1e-9, flatearth=False, debug=False))
    for ang in (-5.0, -10.0, -15.0):  # D
        put(f"D_{int(-ang)}", pr.shoot_ray(1000.0, 0.0, ang, 100e3, 200, munk_env(), rtol=1e-9, flatearth=False,
                                           debug=False))
    for ang in (-85.0, -87.0, -89.0, -89.9, -90.0):  # F
        put(f"F_munk_{ang}", pr.shoot_ray(1000.0, 0.0, ang, 10e3, 50, munk_env(r_max=50e3), rtol=1e-6,
                                          flatearth=False, debug=False))
    for ang in (-89.9, -90.0):
        put(f"F_const_{ang}", pr.shoot_ray(200.0, 0.0, ang, 10e3, 50, const_env(), rtol=1e-6, flatearth=False,
                                           debug=False))
    # G backwards
    put("G_const", pr.shoot_ray(200.0, 30e3, -10.0, 0.0, 60, const_env(), rtol=1e-9, flatearth=False, debug=False))
    z = np.linspace(0.0, 6000.0, 400); r = np.linspace(0.0, 100e3, 80)
    c2 = np.array([munk_ssp(z, sofar_depth=1300 + 0.01 * ri) for ri in r])
    bv = np.linspace(4500.0, 4900.0, len(r))
    envg = make_env(dict(c=c2, r=r, z=z, bathy=bv, bathy_r=r))
    put("G_bwd", pr.shoot_ray(200.0, 60e3, -15.0, 10e3, 80, envg, rtol=1e-9, flatearth=False, debug=False))
    if VARIANT == "jit":
        angles = np.linspace(-15.0, 15.0, 80)
        rf = pr.shoot_rays(200.0, 40e3, angles, 5e3, 60, munk_env(r_max=50e3), rtol=1e-9, flatearth=False,
                           debug=False)
        out.update(fan_dict(rf, "G_fan_"))
    save("ref_physics", **out)


def case_eigen():
    """find_eigenrays on the default env: 101-ray fan, 100 km, 2 receiver depths, ztol 1e-3 (untested by
    the reference itself -- these vectors are the only pin of the eigenray path)."""
    env = OceanEnvironment2D()
    angles = np.linspace(-12, 12, 101)
    rf = pr.shoot_rays(1000.0, 0.0, angles, 100e3, 101, env, rtol=1e-9, n_processes=8, debug=False)
    rds = [1000.0, 2500.0]
    er = pr.find_eigenrays(rf, rds, 1000.0, 0.0, 100e3, 101, env, ztol=1e-3, max_iter=20, num_workers=8,
                           debug=False)
    out = fan_dict(rf)
    out["receiver_depths"] = np.array(rds)
    for i in range(len(rds)):
        out[f"e{i}_ts"] = er.ts[i]; out[f"e{i}_zs"] = er.zs[i]; out[f"e{i}_ps"] = er.ps[i]
        out[f"e{i}_rs"] = er.rs[i]
        out[f"e{i}_launch"] = er.launch_angles[i]; out[f"e{i}_recv"] = er.received_angles[i]
        out[f"e{i}_ray_id"] = er.ray_id[i]; out[f"e{i}_ray_id_int"] = er.ray_id_int[i]
        out[f"e{i}_nb"] = er.n_botts[i]; out[f"e{i}_ns"] = er.n_surfs[i]
        out[f"e{i}_num"] = np.array(er.num_eigenrays[rds[i]]); out[f"e{i}_found"] = np.array(er.num_eigenrays_found[i])
    save("ref_eigen", **out)


def case_eigen_smooth():
    """The reference's public shoot_rays + find_eigenrays on 6 random smooth environments (tests/fuzz_cases.py:
    smooth_case): 81-ray fan +-20 deg, two receiver depths, ztol 1e-3 -- eigenrays with 0-5 reflections off sloping bottoms,
    flat-earth on / off."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fuzz_cases import smooth_case

    out = {}
    for seed in range(6):
        k = smooth_case(seed)
        a = k["arrays"]
        ssp = xr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
        bathy = xr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
        env = OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=k["lat"], flat_earth_transform=k["flat"])
        angles = np.linspace(-20, 20, 81)
        rf = pr.shoot_rays(k["zs"], 0.0, angles, k["r1"], 33, env, rtol=k["rtol"], n_processes=8, debug=False,
                           flatearth=k["flat"])
        d0 = float(a["bathy"].min())
        rds = [round(0.3 * d0, 1), round(0.6 * d0, 1)]
        er = pr.find_eigenrays(rf, rds, k["zs"], 0.0, k["r1"], 33, env, ztol=1e-3, max_iter=20, num_workers=8, debug=False,
                               rtol=k["rtol"], flatearth=k["flat"])
        P = f"s{seed}_"
        out.update({P + kk: v for kk, v in fan_dict(rf).items()})
        out[P + "receiver_depths"] = np.array(rds)
        for i in range(len(rds)):
            out[P + f"e{i}_ts"] = er.ts[i]; out[P + f"e{i}_zs"] = er.zs[i]; out[P + f"e{i}_rs"] = er.rs[i]
            out[P + f"e{i}_launch"] = er.launch_angles[i]; out[P + f"e{i}_recv"] = er.received_angles[i]
            out[P + f"e{i}_ray_id_int"] = er.ray_id_int[i]
            out[P + f"e{i}_nb"] = er.n_botts[i]; out[P + f"e{i}_ns"] = er.n_surfs[i]
            out[P + f"e{i}_num"] = np.array(er.num_eigenrays[rds[i]]); out[P + f"e{i}_found"] = np.array(er.num_eigenrays_found[i])
    save("ref_eigen_smooth", **out)


def case_backwards_smooth():
    """The reference's public shoot_ray (single-ray branch: sign flip, backwards mirroring, stored conventions) on 6
    random smooth environments (tests/fuzz_cases.py: smooth_case, seeds 6-11): 5 launch angles each, BACKWARDS from the far
    end of the track to 10 % of it, with flat-earth on / off; the first and last environment also with
    terminate_backwards=False."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fuzz_cases import smooth_case

    out = {}
    for seed in range(6, 12):
        k = smooth_case(seed)
        a = k["arrays"]
        ssp = xr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"]})
        bathy = xr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
        env = OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, lat=k["lat"], flat_earth_transform=k["flat"])
        tb = seed not in (6, 11)
        for j, ang in enumerate((-15.0, -7.0, 0.0, 7.0, 15.0)):
            ray = pr.shoot_ray(k["zs"], k["r1"], ang, 0.1 * k["r1"], 25, env, rtol=k["rtol"], terminate_backwards=tb,
                               flatearth=k["flat"], debug=False)
            tag = f"s{seed}_a{j}"
            out[tag + "_none"] = np.array(ray is None)
            if ray is not None:
                out[tag + "_r"] = ray.r; out[tag + "_t"] = ray.t; out[tag + "_z"] = ray.z; out[tag + "_p"] = ray.p
                out[tag + "_nb"] = np.array(ray.n_bottom); out[tag + "_ns"] = np.array(ray.n_surface)
                out[tag + "_la"] = np.array(ray.launch_angle)
    save("ref_backwards_smooth", **out)


def case_cfg3_sub():
    """BASELINE config 3 sub-sample: Munk + internal waves on the 2001 x 1201 grid, 1000 km; 64 rays of the
    1 048 576-ray fan linspace(-15, 15, 1 << 20) (evenly spread ray indices, both ends included)."""
    a = cases.config3_arrays()
    env = make_env(a)
    n_fan = 1 << 20
    idx = np.round(np.linspace(0, n_fan - 1, 64)).astype(np.int64)
    internal = np.linspace(-15.0, 15.0, n_fan)[idx]
    al = array_level(env, False, 1000.0, 0.0, 1000e3, internal, 101)
    cin, cpin, rin, zin, depths, dr, ba = _unpack_envi(env, flatearth=False)
    save("ref_cfg3_sub", cin_sha=sha(cin), cpin_sha=sha(cpin), bottom_angles=ba, ray_index=idx,
         **{"al_" + k: v for k, v in al.items()})


def case_cfg4_sub():
    """BASELINE config 4 in miniature: the reference's own shoot_rays + find_eigenrays through 1000 km of the
    config-3 internal-wave field -- an 81-ray fan around -9 deg (a caustic: 2-3 brackets per receiver depth),
    3 receiver depths, ztol = 1e-3 (config 4), 0.05 and 1 (the reference's default)."""
    a = cases.config3_arrays()
    env = make_env(a)
    user = np.linspace(-9.02, -8.98, 81)
    rf = pr.shoot_rays(1000.0, 0.0, user, 1000e3, 11, env, rtol=1e-9, n_processes=8, debug=False, flatearth=False)
    rds = [1130.0, 1160.0, 1190.0]
    out = fan_dict(rf)
    out["receiver_depths"] = np.array(rds)
    out["user_angles"] = user
    for tag, ztol in (("z3", 1e-3), ("z05", 0.05), ("z1", 1.0)):
        try:
            er = pr.find_eigenrays(rf, rds, 1000.0, 0.0, 1000e3, 11, env, ztol=ztol, max_iter=20, num_workers=8,
                                   debug=False, flatearth=False)
        except Exception as e:  # noqa: BLE001
            # the reference cannot build an EigenRays when a receiver depth has no converged eigenray
            # (RayFan([]) -> numpy AxisError, ray_objects.py:135-144): recorded, not papered over
            out[tag + "_raised"] = np.array(type(e).__name__)
            continue
        for i in range(len(rds)):
            pre = f"{tag}_e{i}_"
            out[pre + "num"] = np.array(er.num_eigenrays[rds[i]])
            out[pre + "found"] = np.array(er.num_eigenrays_found[i])
            if er.num_eigenrays_found[i] == 0:
                continue
            out[pre + "ts_end"] = er.ts[i][:, -1]; out[pre + "zs_end"] = er.zs[i][:, -1]
            out[pre + "ps_end"] = er.ps[i][:, -1]
            out[pre + "launch"] = er.launch_angles[i]; out[pre + "recv"] = er.received_angles[i]
            out[pre + "ray_id_int"] = er.ray_id_int[i]
            out[pre + "nb"] = er.n_botts[i]; out[pre + "ns"] = er.n_surfs[i]
    save("ref_cfg4_sub", **out)


def case_flatearth():
    """eflat / eflatinv of the reference (environment.py:371-466) on a spread of depths and latitudes, and the
    reference's range-dependent flat_earth_c (environment.py:239-303, sequential branch) on a 37 x 301 config-2 grid
    with one latitude per range column and on an irregular depth grid that does not start at the surface.  xarray's
    1-D linear DataArray.interp is the stand-in's scipy interp1d (see oracle/stubs/xarray)."""
    from pygenray.environment import eflat, eflatinv, flat_earth_c

    dep = np.concatenate(([1.0], np.linspace(10.0, 6000.0, 41)))
    out = {"dep": dep, "lats": np.array([-60.0, 0.0, 35.0, 80.0])}
    for lat in out["lats"]:
        depf, csf = eflat(dep, lat, 1480.0 + 0.017 * dep)
        d_inv, c_inv = eflatinv(depf, lat, csf)
        out[f"depf_{lat}"] = depf; out[f"csf_{lat}"] = csf; out[f"dinv_{lat}"] = d_inv; out[f"cinv_{lat}"] = c_inv
    a = cases.config2_arrays(n_r=37, n_z=301)
    lat = np.linspace(20.0, 47.0, 37)
    c = xr.DataArray(a["c"].T, dims=["depth", "range"], coords={"depth": a["z"], "range": a["r"], "lat": ("range", lat)})
    fe = flat_earth_c(c, verbose=False, n_cpus=1)
    assert fe.dims == ("range", "depth")
    out["fec_lat"] = lat; out["fec_out"] = fe.values; out["fec_c_sha"] = sha(a["c"])
    z = np.sort(np.concatenate(([5.0, 5400.0], np.random.default_rng(2).uniform(5.0, 5400.0, 120))))
    c2 = np.array([munk_ssp(z, 1300.0 + 10 * i, 0.00737) for i in range(9)])
    lat2 = np.linspace(-70, 70, 9)
    fe2 = flat_earth_c(xr.DataArray(c2.T, dims=["depth", "range"],
                                    coords={"depth": z, "range": np.arange(9.0), "lat": ("range", lat2)}),
                       verbose=False, n_cpus=1)
    out["fec2_z"] = z; out["fec2_c"] = c2; out["fec2_lat"] = lat2; out["fec2_out"] = fe2.values
    save("ref_flatearth", **out)


CASES = dict(cfg1=case_cfg1, fixture=case_fixture, cfg2_sub=case_cfg2_sub, bounce_fine=case_bounce_fine,
             physics=case_physics, eigen=case_eigen, cfg3_sub=case_cfg3_sub,
             cfg4_sub=case_cfg4_sub, flatearth=case_flatearth, cfg5_sub=case_cfg5_sub,
             fuzz_short=case_fuzz_short, fuzz_smooth=case_fuzz_smooth,
             eigen_smooth=case_eigen_smooth, backwards_smooth=case_backwards_smooth)
# cases re-run without the JIT to measure the reference's own self-consistency envelope
NOJIT_CASES = ("cfg1", "cfg2_sub", "bounce_fine", "cfg3_sub", "cfg4_sub", "cfg5_sub", "fuzz_short", "fuzz_smooth", "backwards_smooth")

if __name__ == "__main__":
    want = sys.argv[1:] or list(CASES)
    for name in want:
        if VARIANT == "nojit" and name not in NOJIT_CASES:
            continue
        print(f"[{VARIANT}] {name}")
        CASES[name]()
