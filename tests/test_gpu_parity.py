"""GPU parity: libprt.so (sm_100a kernels, called through the C ABI) against the CPU oracle, bit for bit,
on the same seeded inputs; then the public API against the reference's golden vectors, and
size-independent properties at BASELINE.json's full sizes.  Run with `-m gpu` on a B200."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_same_fan, flat_earth_c_numpy, golden, make_env, oracle_env_from, ray_angle_numpy
import workloads as cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pr():
    import pygenray_b200 as pr_

    pr_._lib.require_gpu()
    return pr_


def _gpu_fan(pr, env, flatearth, zs, r0, R, angles, n_save, mirror=False, **kw):
    arrs = pr._unpack_envi(env, flatearth=flatearth)
    if mirror:
        cin, cpin, rin, depths, dr, ba = pr._mirror_envi_arrays(arrs[0], arrs[1], arrs[2], arrs[4], arrs[5], arrs[6])
        arrs = (cin, cpin, rin, arrs[3], depths, dr, ba)
    denv = pr.DeviceEnvironment(*arrs)
    res = pr.shoot_fan_arrays(denv, zs, r0, R, angles=angles, num_range_save=n_save, counters=True, **kw)
    out = dict(final=res["final"].T.copy(), n_bottom=res["n_bottom"], n_surface=res["n_surface"],
               status=res["status"], counters=res["counters"].T.astype(np.int64),
               traj=np.transpose(res["traj"], (1, 0, 2)) if n_save else None)
    denv.close()
    return out


def _oracle_fan(env, flatearth, zs, r0, R, angles, n_save, mirror=False, **kw):
    oenv = oracle_env_from(env, flatearth, mirror=mirror)
    rs = np.linspace(r0, R, n_save) if n_save else None
    return orc.shoot_fan(oenv, zs, r0, R, angles_deg=angles, range_save=rs, **kw)


CASES = {
    "config1_munk": dict(env="default", flat=True, zs=1000.0, r0=0.0, R=100e3, ang=np.linspace(-12, 12, 101), n=1001),
    "steep_bouncing_1m_grid": dict(env="default", flat=True, zs=1000.0, r0=0.0, R=100e3, ang=np.linspace(-40, 40, 321), n=401),
    "config2_range_dependent": dict(env="cfg2", flat=False, zs=1000.0, r0=0.0, R=500e3, ang=np.linspace(-15, 15, 1000), n=201),
    "status_codes": dict(env="cfg2", flat=False, zs=50.0, r0=0.0, R=500e3, ang=np.linspace(-89.5, 89.5, 180), n=301),
    "arrival_only": dict(env="cfg2", flat=False, zs=1000.0, r0=0.0, R=500e3, ang=np.linspace(-15, 15, 777), n=0),
    "mirrored": dict(env="cfg2", flat=False, zs=300.0, r0=-4e5, R=-1e5, ang=np.linspace(-15, 15, 97), n=120, mirror=True),
    "single_ray": dict(env="cfg2", flat=False, zs=1000.0, r0=0.0, R=200e3, ang=np.array([7.5]), n=33),
    "ragged_33_rays": dict(env="cfg2", flat=False, zs=1000.0, r0=0.0, R=200e3, ang=np.linspace(-14, 14, 33), n=2),
    # source depth, receiver range and save points exactly on grid nodes: weights of exactly 0 / 1 force the exact re-run
    "node_aligned": dict(env="round", flat=False, zs=1000.0, r0=0.0, R=50e3,
                         ang=np.array([-9.0, -3.0, -1e-9, 0.0, 1e-9, 3.0, 9.0, 14.0]), n=26),
    "source_on_surface": dict(env="round", flat=False, zs=0.0, r0=2e3, R=8e4, ang=np.linspace(-5, 12, 9), n=40),
    # irregular depth grid: the staging code switches the speculative trial step off
    "irregular_depths": dict(env="irregular", flat=False, zs=700.0, r0=0.0, R=1e5, ang=np.linspace(-20, 20, 65), n=90),
}


def _env(pr, name):
    if name == "default":
        return pr.OceanEnvironment2D()
    if name == "round":
        return make_env(cases.config2_arrays(n_r=51, n_z=551, R=1e5))  # 2 km x 10 m grid, nodes at round numbers
    if name == "irregular":
        rng = np.random.default_rng(3)
        a = cases.config2_arrays(n_r=40, n_z=90, R=1e5)
        z = np.sort(np.concatenate(([0.0, 5500.0], rng.uniform(0, 5500.0, 88))))
        a.update(z=z, c=np.array([cases.munk_ssp(z, 1300.0) for _ in a["r"]]))
        return make_env(a)
    return make_env(cases.config2_arrays())


@pytest.mark.parametrize("name", list(CASES))
def test_fan_bit_identical_to_oracle(pr, name):
    c = CASES[name]
    env = _env(pr, c["env"])
    g = _gpu_fan(pr, env, c["flat"], c["zs"], c["r0"], c["R"], c["ang"], c["n"], mirror=c.get("mirror", False))
    o = _oracle_fan(env, c["flat"], c["zs"], c["r0"], c["R"], c["ang"], c["n"], mirror=c.get("mirror", False))
    assert_same_fan(o, g, name)


@pytest.mark.parametrize("rtol,tb", [(1e-6, True), (1e-12, True), (1e-9, False)])
def test_fan_options_bit_identical(pr, rtol, tb):
    env = make_env(cases.config2_arrays())
    ang = np.linspace(-80, 80, 65)
    g = _gpu_fan(pr, env, False, 50.0, 0.0, 150e3, ang, 64, rtol=rtol, terminate_backwards=tb)
    o = _oracle_fan(env, False, 50.0, 0.0, 150e3, ang, 64, rtol=rtol, terminate_backwards=tb)
    assert_same_fan(o, g, f"rtol={rtol} terminate_backwards={tb}")


def test_launch_hints_do_not_change_results(pr):
    """Rays per warp (prt_set_lanes) and the concurrency hint only lay rays out over warps differently: the
    arrival table, trajectories and work counters are bit-identical for every setting."""
    env = make_env(cases.config2_arrays())
    ang = np.linspace(-15, 15, 203)
    L = pr._lib.load()
    ref = _gpu_fan(pr, env, False, 1000.0, 0.0, 300e3, ang, 61)
    try:
        for lanes in (1, 2, 4, 8, 16, 32):
            assert L.prt_set_lanes(lanes) == 0
            assert_same_fan(ref, _gpu_fan(pr, env, False, 1000.0, 0.0, 300e3, ang, 61), f"lanes={lanes}")
        L.prt_set_lanes(0)
        L.prt_set_concurrency(8)
        assert_same_fan(ref, _gpu_fan(pr, env, False, 1000.0, 0.0, 300e3, ang, 61), "concurrency hint")
    finally:
        L.prt_set_lanes(0)
        L.prt_set_concurrency(1)


def test_empty_fan_and_argument_errors(pr):
    env = make_env(cases.config2_arrays())
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 1e5, angles=np.array([]), num_range_save=11)
    assert res["final"].shape == (3, 0) and res["traj"].shape == (3, 0, 11)
    with pytest.raises(ValueError):
        pr.shoot_fan_arrays(denv, 1000.0, 1e5, 0.0, angles=[1.0])
    L = pr._lib.load()
    assert L.prt_shoot_fan(denv.handle, 4, None, None, 0.0, 0.0, 1.0, 1e-9, 1e-6, 1, 0, None, None, None, None, None,
                           None, None, 0, None, 0, None) == 1  # PRT_ERR_ARG
    assert b"angles_dev" in L.prt_last_error()


def test_host_pointer_entry_point(pr):
    """prt_shoot_fan_host: the call a non-Python host binds (host arrays in/out)."""
    env = make_env(cases.config2_arrays())
    arrs = pr._unpack_envi(env, flatearth=False)
    denv = pr.DeviceEnvironment(*arrs)
    L = pr._lib.load()
    ang = np.linspace(-15, 15, 257)
    n, n_save = len(ang), 50
    rs = np.linspace(0, 3e5, n_save)
    traj = np.empty((3, n, n_save)); fin = np.empty((3, n)); nb = np.zeros(n, np.int32); ns = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32); cnt = np.zeros((6, n), np.uint32); flips = np.zeros(n, np.int32)
    rc = L.prt_shoot_fan_host(denv.handle, n, ang.ctypes.data, None, 1000.0, 0.0, 3e5, 1e-9, 1e-6, 1, n_save,
                              rs.ctypes.data, traj.ctypes.data, fin.ctypes.data, nb.ctypes.data, ns.ctypes.data,
                              st.ctypes.data, cnt.ctypes.data, 0, flips.ctypes.data, 0)
    assert rc == 0, L.prt_last_error()
    o = _oracle_fan(env, False, 1000.0, 0.0, 3e5, ang, n_save)
    g = dict(final=fin.T.copy(), n_bottom=nb, n_surface=ns, status=st, counters=cnt.T.astype(np.int64),
             traj=np.transpose(traj, (1, 0, 2)))
    assert_same_fan(o, g, "prt_shoot_fan_host")
    ok = st == 0
    assert np.array_equal(flips[ok], np.sum(np.diff(np.sign(traj[2][ok])) != 0, axis=1))


# ---- public API vs the reference's golden vectors ------------------------------------------------
def test_shoot_rays_config1_against_reference_fan(pr):
    g = golden("ref_cfg1_jit.npz")
    fan = pr.shoot_rays(1000.0, 0.0, g["user_angles"], 100e3, 101, pr.OceanEnvironment2D(), rtol=1e-9, debug=False)
    assert np.array_equal(fan.thetas, g["fan_thetas"]) and np.array_equal(fan.rs, g["fan_rs"])
    assert np.array_equal(fan.n_botts, g["fan_n_botts"]) and np.array_equal(fan.n_surfs, g["fan_n_surfs"])
    assert np.array_equal(fan.ray_ids, g["fan_ray_ids"])
    assert np.abs(fan.ts - g["fan_ts"]).max() < 1e-9      # north star: 1e-6 s
    assert np.abs(fan.zs - g["fan_zs"]).max() < 1e-5      # north star: 1e-3 m
    assert np.abs(fan.ps - g["fan_ps"]).max() * 1500 < 1e-9  # north star: 1e-6 rad
    assert np.array_equal(fan.source_depths, g["fan_source_depths"])


def test_shoot_rays_small_fan_reference_fixture(pr):
    """< 70 angles: the double sign flip of the reference, pinned by its own regression fixture."""
    g = golden("ref_fixture_jit.npz")
    z = np.linspace(0.0, 6000.0, 400); r = np.linspace(0.0, 50e3, 30)
    env = make_env(dict(c=np.outer(np.ones(30), cases.munk_ssp(z)), r=r, z=z, bathy=np.full(30, 5000.0), bathy_r=r))
    fan = pr.shoot_rays(1300.0, 0.0, [-8.0, -4.0, 0.0, 4.0, 8.0], 50e3, 50, env, n_processes=1, debug=False,
                        flatearth=False)
    np.testing.assert_allclose(fan.ts, g["fixture_ts"], atol=1e-6)  # the reference's own window (test_physics.py:360-386)
    np.testing.assert_allclose(fan.zs, g["fixture_zs"], atol=0.1)
    np.testing.assert_allclose(fan.ps, g["fixture_ps"], atol=0.1)
    np.testing.assert_array_equal(fan.n_botts, g["fixture_n_botts"])
    np.testing.assert_array_equal(fan.n_surfs, g["fixture_n_surfs"])
    assert np.array_equal(fan.thetas, g["fan_thetas"]) and np.array_equal(fan.ray_ids, g["fan_ray_ids"])


def test_shoot_ray_conventions_and_drops(pr):
    g = golden("ref_physics_jit.npz")
    z = np.linspace(0.0, 6000.0, 600); r = np.linspace(0.0, 100e3, 50)
    env = make_env(dict(c=np.outer(np.ones(50), cases.munk_ssp(z)), r=r, z=z, bathy=np.full(50, 5000.0), bathy_r=r))
    ray = pr.shoot_ray(1000.0, 0.0, -5.0, 100e3, 200, env, rtol=1e-9, flatearth=False, debug=False)
    assert ray.launch_angle == float(g["D_5_la"]) == 5.0
    assert np.array_equal(ray.r, g["D_5_r"]) and ray.n_bottom == 0 and ray.n_surface == 0
    assert np.abs(ray.t - g["D_5_t"]).max() < 1e-4 and np.abs(ray.z - g["D_5_z"]).max() < 0.5
    z5 = np.linspace(0.0, 6000.0, 600); r5 = np.linspace(0.0, 50e3, 50)
    env5 = make_env(dict(c=np.outer(np.ones(50), cases.munk_ssp(z5)), r=r5, z=z5, bathy=np.full(50, 5000.0), bathy_r=r5))
    assert pr.shoot_ray(1000.0, 0.0, -90.0, 10e3, 50, env5, rtol=1e-6, flatearth=False, debug=False) is None
    assert bool(g["F_munk_-90.0_none"])
    # backwards shot: end points exact (reference test G)
    zc = np.linspace(0.0, 5000.0, 200); rc = np.linspace(0.0, 100e3, 20)
    envc = make_env(dict(c=np.full((20, 200), 1500.0), r=rc, z=zc, bathy=np.full(20, 4500.0), bathy_r=rc))
    ray = pr.shoot_ray(200.0, 30e3, -10.0, 0.0, 60, envc, rtol=1e-9, flatearth=False, debug=False)
    assert ray.r[0] == 30e3 and ray.r[-1] == 0.0
    assert np.abs(ray.t - g["G_const_t"]).max() < 1e-9 and np.abs(ray.z - g["G_const_z"]).max() < 1e-6
    with pytest.raises(Exception, match="monotonically"):
        bad = make_env(dict(c=np.full((20, 200), 1500.0), r=rc[::-1], z=zc, bathy=np.full(20, 4500.0), bathy_r=rc))
        pr.shoot_ray(200.0, 0.0, -10.0, 1e4, 10, bad, flatearth=False)


def test_backwards_shoot_ray_on_random_smooth_environments_against_reference(pr):
    """The public shoot_ray of the product, BACKWARDS, against the reference's on 6 random smooth environments
    (oracle/make_golden.py: case_backwards_smooth; flat-earth on / off, terminate_backwards on / off): stored launch
    angle, range grid and bounce counts exact, arrival within the north-star tolerances."""
    from fuzz_cases import smooth_case

    g = golden("ref_backwards_smooth_jit.npz")
    n = 0
    for seed in range(6, 12):
        k = smooth_case(seed)
        env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
        for j, ang in enumerate((-15.0, -7.0, 0.0, 7.0, 15.0)):
            ray = pr.shoot_ray(k["zs"], k["r1"], ang, 0.1 * k["r1"], 25, env, rtol=k["rtol"], flatearth=k["flat"],
                               terminate_backwards=seed not in (6, 11), debug=False)
            tag = f"s{seed}_a{j}"
            assert (ray is None) == bool(g[tag + "_none"]), tag
            if ray is None:
                continue
            n += 1
            assert ray.launch_angle == float(g[tag + "_la"]) and np.array_equal(ray.r, g[tag + "_r"]), tag
            assert ray.n_bottom == int(g[tag + "_nb"]) and ray.n_surface == int(g[tag + "_ns"]), tag
            assert abs(ray.t[-1] - g[tag + "_t"][-1]) < 1e-6 and abs(ray.z[-1] - g[tag + "_z"][-1]) < 1e-3, tag
            assert abs(ray.p[-1] - g[tag + "_p"][-1]) * 1500.0 < 1e-6, tag
    assert n == 30


def test_find_eigenrays_against_reference(pr):
    g = golden("ref_eigen_jit.npz")
    env = pr.OceanEnvironment2D()
    fan = pr.shoot_rays(1000.0, 0.0, np.linspace(-12, 12, 101), 100e3, 101, env, rtol=1e-9, debug=False)
    rds = [float(v) for v in g["receiver_depths"]]
    er = pr.find_eigenrays(fan, rds, 1000.0, 0.0, 100e3, 101, env, ztol=1e-3, max_iter=20, debug=False)
    for i, rd in enumerate(rds):
        assert er.num_eigenrays[rd] == int(g[f"e{i}_num"]) and er.num_eigenrays_found[i] == int(g[f"e{i}_found"])
        assert np.abs(np.radians(er.launch_angles[i] - g[f"e{i}_launch"])).max() < 1e-6
        assert np.abs(er.ts[i][:, -1] - g[f"e{i}_ts"][:, -1]).max() < 1e-6
        assert np.abs(er.zs[i][:, -1] + rd).max() < 1e-3
        assert np.array_equal(er.ray_id[i], g[f"e{i}_ray_id"]) and np.array_equal(er.ray_id_int[i], g[f"e{i}_ray_id_int"])
        assert np.array_equal(er.n_botts[i], g[f"e{i}_nb"]) and np.array_equal(er.n_surfs[i], g[f"e{i}_ns"])
        assert np.abs(np.radians(er.received_angles[i] - g[f"e{i}_recv"])).max() < 1e-6
        assert np.array_equal(er.rs[i], g[f"e{i}_rs"])


def test_find_eigenrays_on_random_smooth_environments_against_reference(pr):
    """The public shoot_rays + find_eigenrays of the product against the reference's on 6 random smooth environments
    (oracle/make_golden.py: case_eigen_smooth; sloping bottoms, flat-earth on / off, eigenrays with 0-4 reflections):
    same fan, same brackets, same converged set, launch / received angles within 1e-6 rad, arrival times within 1e-6 s,
    ray ids and bounce counts identical."""
    from fuzz_cases import smooth_case

    g = golden("ref_eigen_smooth_jit.npz")
    n_found = 0
    for seed in range(6):
        k = smooth_case(seed)
        env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
        P = f"s{seed}_"
        fan = pr.shoot_rays(k["zs"], 0.0, np.linspace(-20, 20, 81), k["r1"], 33, env, rtol=k["rtol"], debug=False,
                            flatearth=k["flat"])
        assert np.array_equal(fan.thetas, g[P + "fan_thetas"]) and np.array_equal(fan.n_botts, g[P + "fan_n_botts"])
        assert np.array_equal(fan.n_surfs, g[P + "fan_n_surfs"]) and np.array_equal(fan.ray_ids, g[P + "fan_ray_ids"])
        rds = [float(v) for v in g[P + "receiver_depths"]]
        er = pr.find_eigenrays(fan, rds, k["zs"], 0.0, k["r1"], 33, env, ztol=1e-3, max_iter=20, debug=False,
                               rtol=k["rtol"], flatearth=k["flat"])
        for i, rd in enumerate(rds):
            e = P + f"e{i}_"
            assert er.num_eigenrays[rd] == int(g[e + "num"]) and er.num_eigenrays_found[i] == int(g[e + "found"]), (seed, rd)
            n_found += er.num_eigenrays_found[i]
            assert np.abs(np.radians(er.launch_angles[i] - g[e + "launch"])).max() < 1e-6
            assert np.abs(np.radians(er.received_angles[i] - g[e + "recv"])).max() < 1e-6
            assert np.abs(er.ts[i][:, -1] - g[e + "ts"][:, -1]).max() < 1e-6
            assert np.abs(er.zs[i][:, -1] + rd).max() < 1e-3
            assert np.array_equal(er.ray_id_int[i], g[e + "ray_id_int"])
            assert np.array_equal(er.n_botts[i], g[e + "nb"]) and np.array_equal(er.n_surfs[i], g[e + "ns"])
            assert np.array_equal(er.rs[i], g[e + "rs"])
    assert n_found == 54


def test_eigenray_refinement_bit_identical_to_oracle(pr):
    """Brackets over 16 receiver depths on the range-dependent grid; device regula falsi == oracle."""
    env = make_env(cases.config2_arrays())
    fan = pr.shoot_rays(1000.0, 0.0, np.linspace(-14, 14, 400), 300e3, 31, env, rtol=1e-9, debug=False, flatearth=False)
    rds = np.linspace(300.0, 4000.0, 16)
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    br, counts = pr.find_brackets(fan.zs[:, -1], fan.thetas, rds)
    # bracketing == numpy restatement of eigenrays.py:62-79
    exp = [np.where(np.diff(np.sign(fan.zs[:, -1] + rd)))[0] for rd in rds]
    assert np.array_equal(counts, [len(k) for k in exp])
    assert np.array_equal(br["k"].cpu().numpy(), np.concatenate(exp))
    ref = pr.refine_brackets(denv, br, 1000.0, 0.0, 300e3, ztol=1e-3, max_iter=20)
    host = {k: v.cpu().numpy() for k, v in br.items()}
    o = orc.refine_brackets(oracle_env_from(env, False), host["z1"], host["z2"], host["th1"], host["th2"], host["rd"],
                            1000.0, 0.0, 300e3, ztol=1e-3, max_iter=20)
    assert np.array_equal(ref["status"].cpu().numpy(), o["status"])
    assert np.array_equal(ref["shots"].cpu().numpy(), o["shots"])
    assert np.array_equal(ref["theta"].cpu().numpy(), o["theta"], equal_nan=True)
    assert np.array_equal(ref["final"].cpu().numpy().T, o["final"], equal_nan=True)
    assert (o["status"] == 0).sum() > 5 and (o["status"] == 2).sum() > 5  # both outcomes exercised


# ---- full-size properties (BASELINE.json sizes; no oracle pass over the whole set) -----------------
def test_config2_full_size_properties(pr):
    env = make_env(cases.config2_arrays())
    arrs = pr._unpack_envi(env, flatearth=False)
    denv = pr.DeviceEnvironment(*arrs)
    ang = np.linspace(-15, 15, 10000)
    a = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 500e3, angles=ang, num_range_save=1001, counters=True)
    b = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 500e3, angles=ang[::-1].copy(), num_range_save=0)
    # determinism + independence of rays: reversing the fan order permutes the arrival table
    assert np.array_equal(a["final"], b["final"][:, ::-1], equal_nan=True)
    assert np.array_equal(a["status"], b["status"][::-1])
    ok = a["status"] == 0
    assert ok.mean() > 0.98
    # last save column is the exact final state; travel time increases along every surviving ray
    assert np.array_equal(a["traj"][:, ok, -1], a["final"][:, ok])
    assert (np.diff(a["traj"][0][ok], axis=1) > 0).all()
    assert not np.isnan(a["traj"][:, ok]).any() and np.isnan(a["traj"][:, ~ok]).all()
    # a 200-ray subsample against the oracle, bit for bit
    sub = np.arange(0, 10000, 50)
    o = _oracle_fan(env, False, 1000.0, 0.0, 500e3, ang[sub], 1001)
    assert np.array_equal(np.transpose(a["traj"][:, sub], (1, 0, 2)), o["traj"], equal_nan=True)
    assert np.array_equal(a["n_bottom"][sub], o["n_bottom"]) and np.array_equal(a["n_surface"][sub], o["n_surface"])


def test_ensemble_launch_bit_identical_per_realisation(pr):
    """BASELINE config 5 in miniature: 6 perturbed realisations x 96 rays in ONE launch == the oracle run
    realisation by realisation."""
    envs = [make_env(cases.config5_realisation(i)) for i in range(6)]
    arrs = [pr._unpack_envi(e, flatearth=False) for e in envs]
    cin = np.stack([a[0] for a in arrs]); cpin = np.stack([a[1] for a in arrs])
    dens = pr.DeviceEnsemble(cin, cpin, *arrs[0][2:])
    ang = np.linspace(-14, 14, 96)
    g = pr.shoot_ensemble_arrays(dens, 1000.0, 0.0, 200e3, ang, num_range_save=41, counters=True)
    assert g["final"].shape == (3, 6, 96) and g["traj"].shape == (3, 6, 96, 41)
    for i, env in enumerate(envs):
        o = _oracle_fan(env, False, 1000.0, 0.0, 200e3, ang, 41)
        gi = dict(final=g["final"][:, i].T.copy(), n_bottom=g["n_bottom"][i], n_surface=g["n_surface"][i],
                  status=g["status"][i], counters=g["counters"][:, i].T.astype(np.int64),
                  traj=np.transpose(g["traj"][:, i], (1, 0, 2)))
        assert_same_fan(o, gi, f"realisation {i}")
    assert not np.array_equal(g["final"][:, 0], g["final"][:, 1])  # the realisations differ
    with pytest.raises(pr._lib.PrtError):
        pr.shoot_fan_arrays(dens, 1000.0, 0.0, 1e5, angles=ang)  # an ensemble handle needs the ensemble call


@pytest.mark.parametrize("grid", ["linspace", "flat_earth", "irregular"])
def test_ensemble_staged_on_device_equals_host_staging(pr, grid):
    """prt_ensemble_create_dev: sound-speed realisations that already live in HBM, depth derivative formed by a
    kernel == the host route (np.gradient, interleaving, one big H2D copy): bit-identical fans, also when the depth
    grid is not uniform (np.gradient switches to its three-point non-uniform formula)."""
    import torch

    a = cases.config2_arrays()
    if grid == "flat_earth":
        a["z"] = a["z"] * (1 + a["z"] / 6.371e6 * (0.5 + a["z"] / 6.371e6 / 3))
    elif grid == "irregular":
        rng = np.random.default_rng(5)
        a["z"] = np.sort(np.concatenate(([0.0, 5500.0], rng.uniform(0, 5500.0, len(a["z"]) - 2))))
    stack = np.stack([cases.config5_realisation(i)["c"] for i in range(4)])
    env = make_env(dict(a, c=stack[0]))
    arrs = pr._unpack_envi(env, flatearth=False)
    cp_stack = np.gradient(stack, arrs[3], axis=2, edge_order=1)
    host = pr.DeviceEnsemble(stack, cp_stack, *arrs[2:])
    dev = pr.DeviceEnsemble.from_device(torch.from_numpy(stack).cuda(), *arrs[2:])
    assert dev.n_env == 4 and dev.device_bytes == host.device_bytes
    ang = np.linspace(-13, 13, 64)
    gh = pr.shoot_ensemble_arrays(host, 1000.0, 0.0, 150e3, ang, num_range_save=31, counters=True)
    gd = pr.shoot_ensemble_arrays(dev, 1000.0, 0.0, 150e3, ang, num_range_save=31, counters=True)
    for k in ("final", "traj", "n_bottom", "n_surface", "status", "counters"):
        assert np.array_equal(gh[k], gd[k], equal_nan=(k in ("final", "traj"))), (grid, k)
    assert (gh["status"] == 0).any()
    host.close(); dev.close()


@pytest.mark.parametrize("seed", range(100, 130))
def test_fuzz_random_environments_gpu(pr, seed):
    """Seeded random environments / shots (tests/fuzz_cases.py): libprt.so == oracle, bit for bit."""
    from fuzz_cases import random_case

    k = random_case(seed)
    env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
    g = _gpu_fan(pr, env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                 terminate_backwards=k["terminate_backwards"])
    o = _oracle_fan(env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                    terminate_backwards=k["terminate_backwards"])
    assert_same_fan(o, g, f"fuzz seed {seed}")


def test_reference_pinned_random_environments_gpu(pr):
    """The 40 short and 24 smooth random environments on which the oracle is compared with the unmodified reference ray
    by ray (tests/test_oracle_vs_reference.py::test_random_*): libprt.so == oracle, bit for bit, on exactly those
    inputs -- and kept / dropped and bounce counts == the reference's."""
    from fuzz_cases import short_case, smooth_case

    for tag, gen, n in (("short", short_case, 40), ("smooth", smooth_case, 24)):
        _reference_pinned_cases(pr, golden(f"ref_fuzz_{tag}_jit.npz"), gen, n)


def _reference_pinned_cases(pr, g, gen, n_seeds):
    for seed in range(n_seeds):
        k = gen(seed)
        env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
        d = _gpu_fan(pr, env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                     terminate_backwards=k["terminate_backwards"])
        o = _oracle_fan(env, k["flat"], k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                        terminate_backwards=k["terminate_backwards"])
        assert_same_fan(o, d, f"reference-pinned fuzz seed {seed}")
        P = f"s{seed}_"
        m = g[P + "ok"]
        assert np.array_equal(d["status"] == 0, m), seed
        assert np.array_equal(d["n_bottom"][m], g[P + "n_bottom"][m]) and np.array_equal(d["n_surface"][m], g[P + "n_surface"][m])


def test_sharded_fan_single_rank_and_records(pr):
    """dist.shoot_fan_sharded at world 1 == plain fan; prt_pack_arrivals lays out [n][6] records."""
    from pygenray_b200 import dist as pdist

    env = make_env(cases.config2_arrays())
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    ang = np.linspace(-15, 15, 130)
    table, local = pdist.shoot_fan_sharded(denv, 1000.0, 0.0, 1e5, ang, 0, 1)
    ref = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 1e5, angles=ang)
    t = table.cpu().numpy()
    assert np.array_equal(t[:, :3], ref["final"].T, equal_nan=True)
    assert np.array_equal(t[:, 3], ref["n_bottom"]) and np.array_equal(t[:, 4], ref["n_surface"])
    assert np.array_equal(t[:, 5], ref["status"])


def test_pipelined_large_fan_equals_single_launch(pr):
    """shoot_rays on >= 4096 rays (chunked launches, copies overlapped, survivors placed centre-out) returns
    exactly what one launch + device compaction returns, including dropped rays at both ends of the fan."""
    env = make_env(cases.config2_arrays())
    user = np.linspace(-16, 16, 5000)
    fan = pr.shoot_rays(1000.0, 0.0, user, 300e3, 37, env, rtol=1e-9, debug=False, flatearth=False)
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    ref = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 300e3, angles=-user, num_range_save=37, stored_convention=True,
                              ray_id_counts=True, compact=True)
    assert len(fan) == ref["traj"].shape[1] < 5000  # some steep rays are dropped
    assert np.array_equal(fan.thetas, user[ref["kept"]])
    assert np.array_equal(fan.ts, ref["traj"][0]) and np.array_equal(fan.zs, ref["traj"][1])
    assert np.array_equal(fan.ps, ref["traj"][2])
    assert np.array_equal(fan.n_botts, ref["n_bottom"]) and np.array_equal(fan.n_surfs, ref["n_surface"])
    ids = ref["p_sign_changes"] * np.sign(fan.thetas)
    free = (fan.n_botts == 0) & (fan.n_surfs == 0)
    assert list(fan.ray_ids[:50]) == [str(v) if f else f"{v}b" for v, f in zip(ids[:50], free[:50])]


# ---- BASELINE configs 3 and 4 at full size: ONE 1 048 576-ray launch through 1000 km of the internal-wave field ---
@pytest.fixture(scope="module")
def cfg3_fan(pr):
    """The north-star fan: arrival table + work counters of all 1 048 576 rays (device tensors) and the pieces
    the config-3/4 tests share."""
    import torch

    env = make_env(cases.config3_arrays())
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    ang = np.linspace(-15.0, 15.0, 1 << 20)
    res = pr.shoot_fan_arrays(denv, 1000.0, 0.0, 1000e3, angles=ang, counters=True, to_host=False)
    torch.cuda.synchronize()
    return dict(env=env, denv=denv, ang=ang, res=res)


def test_config3_full_size_subsample_bit_identical(pr, cfg3_fan):
    """Every 1024th ray of the full-size launch (1024 rays, both fan ends included) == the oracle, bit for bit,
    including the work counters; plus size-independent properties of the whole table."""
    res, ang = cfg3_fan["res"], cfg3_fan["ang"]
    sub = np.unique(np.concatenate((np.arange(0, 1 << 20, 1024), [(1 << 20) - 1])))
    o = _oracle_fan(cfg3_fan["env"], False, 1000.0, 0.0, 1000e3, ang[sub], 0)
    final = res["final"].cpu().numpy()
    status, nb, ns = res["status"].cpu().numpy(), res["n_bottom"].cpu().numpy(), res["n_surface"].cpu().numpy()
    cnt = res["counters"].cpu().numpy()
    g = dict(final=final[:, sub].T.copy(), n_bottom=nb[sub], n_surface=ns[sub], status=status[sub],
             counters=cnt[:, sub].T.astype(np.int64), traj=None)
    assert_same_fan(o, g, "config 3 sub-sample of the 1M-ray launch")
    ok = status == 0
    assert ok.mean() > 0.97 and np.isnan(final[:, ~ok]).all() and not np.isnan(final[:, ok]).any()
    assert (final[0, ok] > 1000e3 / 1560.0).all() and (final[0, ok] < 1000e3 / 1400.0).all()  # travel times
    free = ok & (nb + ns == 0)
    assert free.mean() > 0.8 and (np.abs(ang[free]) < 14.0).all()
    # a second launch of a shuffled slice of the fan returns the same bits (rays are independent; lane/warp
    # placement must not matter)
    rng = np.random.default_rng(7)
    pick = rng.choice(1 << 20, 4096, replace=False)
    b = pr.shoot_fan_arrays(cfg3_fan["denv"], 1000.0, 0.0, 1000e3, angles=ang[pick], counters=True)
    assert np.array_equal(b["final"], final[:, pick], equal_nan=True) and np.array_equal(b["status"], status[pick])
    assert np.array_equal(b["counters"], cnt[:, pick])


def test_config4_brackets_and_refinement_on_the_full_fan(pr, cfg3_fan):
    """Config 4 on the full fan: the arrival table is compacted to surviving rays (as RayFan would), brackets for
    4 of the 64 receiver depths are found on the device (== numpy's restatement of eigenrays.py:62-79 over the
    ~1M-ray table) and a 1024-bracket sample is refined at ztol 1e-3 and 1.0 == the oracle's regula falsi (status,
    shot counts, converged angles, arrival states)."""
    import torch

    res, ang, denv = cfg3_fan["res"], cfg3_fan["ang"], cfg3_fan["denv"]
    keep = torch.nonzero(res["status"] == 0).squeeze(1)
    z_end_d = -res["final"][1].index_select(0, keep)            # stored convention: zs = -z
    th_d = torch.from_numpy(-ang).to(z_end_d.device).index_select(0, keep)  # user angle = -internal
    rds = np.linspace(200.0, 4000.0, 64)[[3, 21, 40, 62]]
    br, counts = pr.find_brackets(z_end_d, th_d, rds)
    z_end, th = z_end_d.cpu().numpy(), th_d.cpu().numpy()
    exp = [np.where(np.diff(np.sign(z_end + rd)))[0] for rd in rds]
    assert np.array_equal(counts, [len(k) for k in exp]) and counts.min() > 1000
    k_all = np.concatenate(exp)
    assert np.array_equal(br["k"].cpu().numpy(), k_all)
    assert np.array_equal(br["rcv"].cpu().numpy(), np.repeat(np.arange(4), counts))
    assert np.array_equal(br["z1"].cpu().numpy(), z_end[k_all]) and np.array_equal(br["z2"].cpu().numpy(), z_end[k_all + 1])
    assert np.array_equal(br["th1"].cpu().numpy(), th[k_all]) and np.array_equal(br["th2"].cpu().numpy(), th[k_all + 1])
    assert np.array_equal(br["rd"].cpu().numpy(), np.repeat(rds, counts))
    # 256 brackets per receiver, spread over the fan
    offs = np.concatenate(([0], np.cumsum(counts)))
    pick = np.concatenate([offs[i] + np.round(np.linspace(0, counts[i] - 1, 256)).astype(np.int64) for i in range(4)])
    pick_d = torch.from_numpy(pick).to(z_end_d.device)
    sub = {k: v.index_select(0, pick_d).contiguous() for k, v in br.items()}
    host = {k: v.cpu().numpy() for k, v in sub.items()}
    oenv = oracle_env_from(cfg3_fan["env"], False)
    for ztol in (1e-3, 1.0):
        ref = pr.refine_brackets(denv, sub, 1000.0, 0.0, 1000e3, ztol=ztol, max_iter=20)
        o = orc.refine_brackets(oenv, host["z1"], host["z2"], host["th1"], host["th2"], host["rd"], 1000.0, 0.0,
                                1000e3, ztol=ztol, max_iter=20)
        assert np.array_equal(ref["status"].cpu().numpy(), o["status"]), ztol
        assert np.array_equal(ref["shots"].cpu().numpy(), o["shots"]), ztol
        assert np.array_equal(ref["theta"].cpu().numpy(), o["theta"], equal_nan=True), ztol
        assert np.array_equal(ref["final"].cpu().numpy().T, o["final"], equal_nan=True), ztol
        assert np.array_equal(ref["n_bottom"].cpu().numpy(), o["n_bottom"]), ztol
        conv = o["status"] == 0
        assert (np.abs(-o["final"][conv, 1] + host["rd"][conv]) < ztol).all()
        # through 1000 km of internal waves the arrival depth is a chaotic function of the launch angle: regula
        # falsi converges on a few per cent of the brackets at 1 m and on almost none at 1 mm (oracle and GPU alike)
        if ztol == 1.0:
            assert 0.02 < conv.mean() < 0.5
        else:
            assert (o["status"] != 0).mean() > 0.9 and (o["shots"][o["status"] == 2] == 22).all()
            assert (o["status"] == 1).any()  # steep brackets: a trial ray dropped (bounced backwards / left the grid)


def test_bracket_fill_edge_cases(pr):
    """Ragged tiles, NaN neighbours, exact zeros, fewer than two rays, more receivers than rays."""
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 7, 2048, 2049, 4097, 10000):
        z = rng.normal(0.0, 900.0, n) - 1500.0
        if n > 10:
            z[rng.choice(n, 5, replace=False)] = np.nan
            z[3] = -1000.0  # exactly on a receiver depth: sign 0
        th = np.linspace(-10, 10, n)
        rds = np.array([1000.0, 1500.0, 250.0, 1e5, -5.0])
        br, counts = pr.find_brackets(z, th, rds)
        with np.errstate(invalid="ignore"):
            exp = [np.where(np.diff(np.sign(z + rd)))[0] for rd in rds] if n >= 2 else [np.zeros(0, np.int64)] * 5
        assert np.array_equal(counts, [len(k) for k in exp]), n
        assert np.array_equal(br["k"].cpu().numpy(), np.concatenate(exp)), n
        assert np.array_equal(br["rcv"].cpu().numpy(), np.repeat(np.arange(5), counts)), n


# ---- device-side eigenray post-processing (EigenRays.__init__, ray_objects.py:505-548) ------------------------------
def test_eigenrays_container_from_ray_lists(pr, tmp_path):
    """EigenRays built the reference's way (lists of Ray): received angles from prt_received_angles == the numpy
    restatement of ray_angle on the stored end state and the un-flattened sound speed (quirk Q10)."""
    from pygenray_b200 import EigenRays, Ray, RayFan

    def make_ray(launch_angle, source_depth, n_bottom=0, n_surface=0, N=10, R=10000.0):
        r = np.linspace(0.0, R, N)
        y = np.vstack([r / 1500.0, np.linspace(source_depth, source_depth + R * 0.01, N),
                       np.cos(np.linspace(0, 7, N)) * np.sin(np.radians(abs(launch_angle) + 1e-3)) / 1500.0])
        return Ray(r=r, y=y, n_bottom=n_bottom, n_surface=n_surface, launch_angle=launch_angle, source_depth=source_depth)

    env = pr.OceanEnvironment2D(flat_earth_transform=False)
    fans = {0: [make_ray(-5.0, 100.0), make_ray(5.0, 100.0, n_surface=1)],
            1: RayFan([make_ray(-5.0, 100.0), make_ray(5.0, 150.0, n_bottom=1), make_ray(-10.0, 200.0)]), 2: []}
    er = EigenRays([1000.0, 2000.0, 3000.0], fans, env, {1000.0: 2, 2000.0: 3, 3000.0: 0}, {0: 2, 1: 3, 2: 0},
                   {0: [], 1: [], 2: []})
    assert er.ts[0].shape == (2, 10) and er.zs[1].shape == (3, 10) and er.received_angles[2].shape == (0,)
    assert list(er.ray_id[0]) == ["-2.0", "2.0b"] and list(er.ray_id_int[0]) == [-2, 2]
    assert np.array_equal(er.launch_angles[1], [-5.0, 5.0, -10.0])
    ss = env.sound_speed
    for ridx in (0, 1):
        exp = [ray_angle_numpy(er.rs[ridx][k, -1], (er.ts[ridx][k, -1], er.zs[ridx][k, -1], er.ps[ridx][k, -1]),
                               np.asarray(ss.values), np.asarray(ss.coords["range"].values),
                               np.asarray(ss.coords["depth"].values))[0] for k in range(er.rs[ridx].shape[0])]
        assert np.abs(er.received_angles[ridx] - np.array(exp)).max() < 1e-11  # degrees; asin differs by ulps
    er.save_mat(str(tmp_path / "e.mat"))


def test_eigenray_table_matches_host_restatement(pr):
    """prt_eigen_table on a config-2 refinement: rows == (receiver, T, -z, -p, theta, received angle, ray id number,
    bounces, bracket) assembled on the host from the same refinement + re-shot; find_eigenrays serves the same
    numbers through the reference's EigenRays object."""
    env = make_env(cases.config2_arrays())
    fan = pr.shoot_rays(1000.0, 0.0, np.linspace(-14, 14, 400), 300e3, 31, env, rtol=1e-9, debug=False, flatearth=False)
    rds = [float(v) for v in np.linspace(300.0, 4000.0, 6)]
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    br, counts = pr.find_brackets(fan.zs[:, -1], fan.thetas, rds)
    ref = pr.refine_brackets(denv, br, 1000.0, 0.0, 300e3, ztol=1.0, max_iter=20)
    tab = pr.eigenray_table(denv, br, ref, env, 1000.0, 0.0, 300e3, num_range_save=31)
    st = ref["status"].cpu().numpy()
    conv = np.flatnonzero(st == 0)
    t = tab["table"]
    assert t.shape == (len(conv), 10) and len(conv) > 20 and tab["n_converged"] == len(conv)
    fin = ref["final"].cpu().numpy()
    assert np.array_equal(t[:, 9], conv) and np.array_equal(t[:, 0], br["rcv"].cpu().numpy()[conv])
    assert np.array_equal(t[:, 1], fin[0, conv]) and np.array_equal(t[:, 2], -fin[1, conv]) and np.array_equal(t[:, 3], -fin[2, conv])
    assert np.array_equal(t[:, 4], ref["theta"].cpu().numpy()[conv])
    assert np.array_equal(t[:, 7], ref["n_bottom"].cpu().numpy()[conv]) and np.array_equal(t[:, 8], ref["n_surface"].cpu().numpy()[conv])
    ss = env.sound_speed
    cin, rin, zin = np.asarray(ss.values), np.asarray(ss.coords["range"].values), np.asarray(ss.coords["depth"].values)
    exp = np.array([ray_angle_numpy(300e3, (t[k, 1], t[k, 2], t[k, 3]), cin, rin, zin)[0] for k in range(len(conv))])
    assert np.abs(t[:, 5] - exp).max() < 1e-11
    traj = tab["fan"]["traj"]  # re-shot of the converged angles, stored convention
    assert np.array_equal(traj[:, :, -1], np.stack((t[:, 1], t[:, 2], t[:, 3])))  # same ray as the converged trial
    ids = np.sum(np.diff(np.sign(traj[2]), axis=1) != 0, axis=1) * np.sign(t[:, 4])
    assert np.array_equal(t[:, 6], ids)
    assert np.array_equal(tab["offsets"], np.searchsorted(t[:, 0], np.arange(7)))
    er = pr.find_eigenrays(fan, rds, 1000.0, 0.0, 300e3, 31, env, ztol=1.0, max_iter=20, flatearth=False, debug=False)
    for i in range(6):
        a, b = tab["offsets"][i], tab["offsets"][i + 1]
        assert er.num_eigenrays_found[i] == b - a and er.num_eigenrays[rds[i]] == counts[i]
        assert np.array_equal(er.received_angles[i], t[a:b, 5]) and np.array_equal(er.launch_angles[i], t[a:b, 4])
        assert np.array_equal(er.ray_id_int[i], t[a:b, 6].astype(np.int64))
        assert np.array_equal(er.ts[i][:, -1], t[a:b, 1]) and np.array_equal(er.zs[i][:, -1], t[a:b, 2])
        free = (t[a:b, 7] == 0) & (t[a:b, 8] == 0)
        assert list(er.ray_id[i]) == [f"{v}{'' if f else 'b'}" for v, f in zip(t[a:b, 6], free)]
        assert len(er.failed_eray_theta_brackets[i]) == counts[i] - (b - a)


def test_find_eigenrays_when_nothing_converges(pr):
    """ztol far below the step-sequence noise: no bracket converges; the product returns empty per-receiver arrays
    (the reference raises an AxisError on RayFan([]), recorded in tests/golden/ref_cfg4_sub_*.npz)."""
    env = make_env(cases.config2_arrays())
    fan = pr.shoot_rays(1000.0, 0.0, np.linspace(-14, 14, 100), 300e3, 11, env, rtol=1e-9, debug=False, flatearth=False)
    er = pr.find_eigenrays(fan, [1500.0], 1000.0, 0.0, 300e3, 11, env, ztol=1e-12, max_iter=3, flatearth=False, debug=False)
    assert er.num_eigenrays_found[0] == 0 and er.num_eigenrays[1500.0] > 0
    assert er.ts[0].shape == (0, 11) and er.received_angles[0].shape == (0,)
    assert len(er.failed_eray_theta_brackets[0]) == er.num_eigenrays[1500.0]


# ---- earth flattening on the device (environment.py:121-173, 239-401) -----------------------------------------------
def test_flat_earth_c_and_transform_rd(pr):
    import torch
    from pygenray_b200.environment import flat_earth_c, flat_earth_device

    a = cases.config2_arrays(n_r=37, n_z=301)
    lat = np.linspace(20.0, 47.0, 37)
    c = pr.DataArray(a["c"].T, dims=["depth", "range"], coords={"depth": a["z"], "range": a["r"], "lat": lat})
    out = flat_earth_c(c, verbose=False, n_cpus=3)
    assert out.dims == ("range", "depth") and np.array_equal(out.coords["lat"].values, lat)
    exp = flat_earth_c_numpy(a["c"], a["z"], lat)
    assert np.array_equal(np.isnan(out.values), np.isnan(exp))
    np.testing.assert_allclose(out.values, exp, rtol=1e-13, atol=0, equal_nan=True)
    assert np.abs(out.values - a["c"]).max() > 0.1  # the transform does something (c_f - c ~ c z / R_e)
    # the reference's own flat_earth_c on the same input (oracle/make_golden.py: case_flatearth)
    g = golden("ref_flatearth_jit.npz")
    assert np.array_equal(g["fec_lat"], lat)
    np.testing.assert_allclose(out.values, g["fec_out"], rtol=1e-13, atol=0, equal_nan=True)
    # irregular, non-zero-based depth grid: generic interval search, NaN above the first flattened node
    z = np.sort(np.concatenate(([5.0, 5400.0], np.random.default_rng(2).uniform(5.0, 5400.0, 120))))
    c2 = np.array([cases.munk_ssp(z, 1300.0 + 10 * i) for i in range(9)])
    o2 = flat_earth_device(torch.from_numpy(c2).cuda(), z, np.linspace(-70, 70, 9), resample=True).cpu().numpy()
    e2 = flat_earth_c_numpy(c2, z, np.linspace(-70, 70, 9))
    assert np.array_equal(g["fec2_z"], z) and np.array_equal(g["fec2_c"], c2)
    assert np.array_equal(np.isnan(o2), np.isnan(g["fec2_out"]))
    np.testing.assert_allclose(o2, g["fec2_out"], rtol=1e-13, atol=0, equal_nan=True)
    assert np.isnan(e2[:, 0]).all() and np.array_equal(np.isnan(o2), np.isnan(e2))
    np.testing.assert_allclose(o2, e2, rtol=1e-13, atol=0, equal_nan=True)
    # environment method: attributes as the reference sets them (bathymetry copied, dcdz on the flattened field)
    ssp = pr.DataArray(a["c"], dims=["range", "depth"], coords={"range": a["r"], "depth": a["z"], "lat": lat})
    bathy = pr.DataArray(a["bathy"], dims=["range"], coords={"range": a["bathy_r"]})
    env = pr.OceanEnvironment2D(sound_speed=ssp, bathymetry=bathy, flat_earth_transform=False)
    env.flat_earth_transform_rd()
    np.testing.assert_allclose(env.sound_speed_fe.values, exp, rtol=1e-13, equal_nan=True)
    assert np.array_equal(env.bathymetry_fe.values, a["bathy"])
    assert np.array_equal(env.dcdz.values, np.gradient(env.sound_speed_fe.values, a["z"], axis=1))
    fan = pr.shoot_rays(1000.0, 0.0, np.linspace(-8, 8, 9), 100e3, 21, env, debug=False, flatearth=True)
    assert len(fan) == 9
    # single-latitude transform of an ensemble on the device == OceanEnvironment2D.flat_earth_transform per member
    stack = np.stack([cases.config5_realisation(i)["c"] for i in range(3)])
    dev = flat_earth_device(torch.from_numpy(stack).cuda(), cases.config2_arrays()["z"], 35.0, resample=False).cpu().numpy()
    for i in range(3):
        e = make_env(dict(cases.config2_arrays(), c=stack[i]), flat=True, lat=35)
        assert np.array_equal(dev[i], e.sound_speed_fe.values)


# ---- tomography-ensemble driver --------------------------------------------------------
def test_iw_realisations_generated_on_device(pr):
    """prt_iw_perturb == the numpy definition of the config-5 realisations (workloads.config5_realisation) to rounding
    (device cos and a different summation order: not bit-identical, 1e-11 m/s), and the ensemble staged from it shoots
    the same fans as the host-generated one to well below the integration tolerance."""
    from pygenray_b200 import ensemble as ens

    a = cases.config2_arrays()
    seeds = [1000 + i for i in range(5)]
    dev = ens.iw_realisations_device(a["c"], a["r"], a["z"], ens.iw_phases(seeds, J=8, K=8))
    host = np.stack([cases.config5_realisation(i)["c"] for i in range(5)])
    assert dev.shape == host.shape and np.abs(dev.cpu().numpy() - host).max() < 1e-11
    iax = np.argmin(np.abs(a["z"] - 1300.0))
    rms = np.sqrt(np.mean((dev.cpu().numpy()[:, :, iax] - a["c"][None, :, iax]) ** 2, axis=1))
    assert np.abs(rms - 0.5).max() < 1e-12
    env = make_env(a)
    arrs = pr._unpack_envi(env, flatearth=False)
    dens = pr.DeviceEnsemble.from_device(dev, *arrs[2:])
    ang = np.linspace(-8, 8, 33)
    g = pr.shoot_ensemble_arrays(dens, 1000.0, 0.0, 100e3, ang)
    for i in range(5):
        o = _oracle_fan(make_env(cases.config5_realisation(i)), False, 1000.0, 0.0, 100e3, ang, 0)
        assert np.array_equal(g["status"][i], o["status"])
        # a 1e-11 m/s difference of the grids changes the step sequence on this 11 m grid: arrival times agree to the
        # step-sequence noise of the integrator (~1e-6 s), not to 1e-11 / c
        # (reflections amplify it: rays that bounced get a looser bound)
        free = (o["n_bottom"] + o["n_surface"]) == 0
        assert free.sum() > 10 and np.array_equal(g["n_bottom"][i], o["n_bottom"])
        dT = np.abs(g["final"][0, i] - o["final"][:, 0])
        assert np.median(dT[free]) < 5e-6 and dT[free].max() < 2e-4 and dT.max() < 1e-3
        assert np.median(np.abs(g["final"][1, i] - o["final"][:, 1])[free]) < 0.05


def test_travel_time_sensitivity_central_differences(pr):
    from pygenray_b200 import ensemble as ens

    a = cases.config2_arrays()
    env = make_env(a)
    arrs = pr._unpack_envi(env, flatearth=False)
    r, z = a["r"], a["z"]
    modes = np.stack([np.exp(-((z[None, :] - 1000.0) / 400.0) ** 2) * np.cos(2 * np.pi * r[:, None] / 2e5),
                      np.exp(-z[None, :] / 1500.0) * np.ones_like(r)[:, None]])
    ang = np.linspace(-9, 9, 41)
    eps = 0.1  # m/s: T+ - T- ~ 5e-3 s, three orders above the ~1e-6 s step-sequence noise of the integrator
    s1 = ens.travel_time_sensitivity(a["c"], modes, eps, *arrs[2:], 1000.0, 0.0, 200e3, ang)
    assert s1["dT"].shape == (2, 41) and s1["valid"].all()
    # == explicit host-staged runs of the two perturbed environments, bit for bit
    for m in range(2):
        fin = []
        for sgn in (1.0, -1.0):
            e = make_env(dict(a, c=a["c"] + sgn * eps * modes[m]))
            d = pr.DeviceEnvironment(*pr._unpack_envi(e, flatearth=False))
            fin.append(pr.shoot_fan_arrays(d, 1000.0, 0.0, 200e3, angles=ang)["final"])
        assert np.array_equal(s1["dT"][m], (fin[0][0] - fin[1][0]) / (2 * eps))
        assert np.array_equal(s1["dz"][m], (fin[0][1] - fin[1][1]) / (2 * eps))
    # a uniform speed-up shortens every travel time: dT/dalpha ~ -integral(mode / c^2) ds < 0, of the expected size
    assert (s1["dT"][1] < 0).all() and np.abs(s1["dT"][1] / (200e3 / 1500.0 ** 2) + np.exp(-1000.0 / 1500.0)).max() < 0.25
    # second-order accuracy: halving eps changes the estimate by O(eps^2)
    s2 = ens.travel_time_sensitivity(a["c"], modes, eps / 2, *arrs[2:], 1000.0, 0.0, 200e3, ang)
    assert np.abs(s1["dT"] - s2["dT"]).max() < 2e-2 * np.abs(s2["dT"]).max()
    assert np.array_equal(s1["T0"], s2["T0"])


def test_sharded_eigenray_search_single_rank(pr):
    """dist.find_eigenrays_sharded at world size 1 == find_brackets + refine_brackets on the compacted table (the
    world-2 deal / gather logic is covered on the CPU by tests/test_dist_gloo.py)."""
    from pygenray_b200 import dist as pdist

    env = make_env(cases.config2_arrays())
    denv = pr.DeviceEnvironment(*pr._unpack_envi(env, flatearth=False))
    internal = np.linspace(-17, 17, 700)  # the steepest rays are dropped: the table is compacted before bracketing
    table, _ = pdist.shoot_fan_sharded(denv, 1000.0, 0.0, 300e3, internal, 0, 1)
    rds = np.linspace(300.0, 4000.0, 12)
    rec, counts = pdist.find_eigenrays_sharded(denv, table, -internal, rds, 1000.0, 0.0, 300e3, 0, 1, ztol=0.5, max_iter=20)
    t = table.cpu().numpy()
    keep = np.flatnonzero(t[:, 5] == 0)
    assert len(keep) < 700
    br, counts2 = pr.find_brackets(-t[keep, 1], -internal[keep], rds)
    ref = pr.refine_brackets(denv, br, 1000.0, 0.0, 300e3, ztol=0.5, max_iter=20, angle_order=False)
    r = rec.cpu().numpy()
    assert np.array_equal(counts, counts2) and r.shape == (int(counts.sum()), pdist.EIGEN_WIDTH)
    assert np.array_equal(r[:, 0], br["rcv"].cpu().numpy()) and np.array_equal(r[:, 1], br["k"].cpu().numpy())
    assert np.array_equal(r[:, 2], ref["theta"].cpu().numpy(), equal_nan=True)
    assert np.array_equal(r[:, 3:6], ref["final"].cpu().numpy().T, equal_nan=True)
    assert np.array_equal(r[:, 8], ref["status"].cpu().numpy()) and np.array_equal(r[:, 9], ref["shots"].cpu().numpy())
    assert (r[:, 8] == 0).sum() > 20


def test_single_save_point_like_the_reference(pr):
    """num_range_save = 1: np.linspace(source, receiver, 1) = [source] and the only column holds the exact final state
    (launch_rays.py:772-775 overwrites the last column) -- accepted like the reference, == the oracle."""
    env = make_env(cases.config2_arrays())
    ang = np.linspace(-12, 12, 9)
    g = _gpu_fan(pr, env, False, 1000.0, 0.0, 80e3, ang, 1)
    o = _oracle_fan(env, False, 1000.0, 0.0, 80e3, ang, 1)
    assert_same_fan(o, g, "n_save = 1")
    assert np.array_equal(g["traj"][:, :, 0], g["final"])
    ray = pr.shoot_ray(1000.0, 0.0, -5.0, 80e3, 1, env, flatearth=False, debug=False)
    assert ray.r.shape == (1,) and ray.r[0] == 0.0 and ray.t.shape == (1,) and ray.t[0] > 50.0


def test_device_trajectories_are_fetched_lazily(pr):
    """shoot_rays(device_trajectories=True): ts/zs/ps stay on the GPU (DeviceRows); rows, columns and slices come to the
    host on demand and equal the ordinary fan; find_eigenrays works on such a fan (it reads one column)."""
    env = make_env(cases.config2_arrays())
    user = np.linspace(-17, 17, 300)
    host = pr.shoot_rays(1000.0, 0.0, user, 300e3, 41, env, debug=False, flatearth=False)
    dev = pr.shoot_rays(1000.0, 0.0, user, 300e3, 41, env, debug=False, flatearth=False, device_trajectories=True)
    assert isinstance(dev.zs, pr.DeviceRows) and dev.zs.shape == host.zs.shape and len(dev) == len(host) <= 300
    assert np.array_equal(dev.thetas, host.thetas) and np.array_equal(dev.n_botts, host.n_botts)
    assert list(dev.ray_ids) == list(host.ray_ids)
    assert np.array_equal(dev.zs[:, -1], host.zs[:, -1]) and np.array_equal(dev.ts[7], host.ts[7])
    assert np.array_equal(dev.ps[3:9, ::4], host.ps[3:9, ::4]) and dev.ts[5, 40] == host.ts[5, 40]
    assert np.array_equal(dev.zs[np.array([2, 0, 11])], host.zs[[2, 0, 11]])
    assert np.array_equal(dev.zs[::-1][0], host.zs[-1])  # an index form torch does not take: falls back to a full fetch
    ray = dev[4]
    assert np.array_equal(ray.z, host[4].z) and np.array_equal(ray.t, host[4].t)
    e_h = pr.find_eigenrays(host, [1500.0], 1000.0, 0.0, 300e3, 21, env, ztol=0.5, flatearth=False, debug=False)
    dev2 = pr.shoot_rays(1000.0, 0.0, user, 300e3, 41, env, debug=False, flatearth=False, device_trajectories=True)
    e_d = pr.find_eigenrays(dev2, [1500.0], 1000.0, 0.0, 300e3, 21, env, ztol=0.5, flatearth=False, debug=False)
    assert dev2.zs._host is None  # the eigenray search did not pull the trajectories
    assert np.array_equal(e_h.launch_angles[0], e_d.launch_angles[0]) and np.array_equal(e_h.ts[0], e_d.ts[0])
    assert np.array_equal(np.asarray(dev.zs), host.zs)
