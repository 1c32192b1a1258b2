"""CPU-side bit parity: the product's per-ray routine (prt_core.h, compiled for the host by
tests/host_harness) against the independent plain-C oracle, on the same seeded inputs.

These run without a GPU; the `-m gpu` tests repeat them through libprt.so on the device.
"""
import numpy as np
import pytest

from conftest import assert_same_fan, make_env, oracle_env_from
import workloads as cases
from oracle import oracle as orc


def _both(harness, oenv, zs, r0, R, angles, n_save, **kw):
    rs = None if not n_save else np.linspace(r0, R, n_save)
    o = orc.shoot_fan(oenv, zs, r0, R, angles_deg=angles, range_save=rs, **kw)
    h = harness.shoot_fan(oenv, zs, r0, R, angles, range_save=rs, **kw)
    return o, h


def test_config1_munk_flat_earth(harness):
    import pygenray_b200 as pr

    oenv = oracle_env_from(pr.OceanEnvironment2D(), True)
    o, h = _both(harness, oenv, 1000.0, 0.0, 100e3, np.linspace(-12, 12, 101), 1001)
    assert (o["status"] == 0).all() and o["n_bottom"].sum() == 0
    assert_same_fan(o, h, "config 1")


def test_bounces_fine_grid(harness):
    import pygenray_b200 as pr

    oenv = oracle_env_from(pr.OceanEnvironment2D(), True)
    o, h = _both(harness, oenv, 1000.0, 0.0, 100e3, np.linspace(-40, 40, 161), 401)
    assert o["n_bottom"].sum() > 100 and o["n_surface"].sum() > 100
    assert_same_fan(o, h, "steep rays on the 1 m grid")


def test_config2_range_dependent(harness):
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    o, h = _both(harness, oenv, 1000.0, 0.0, 500e3, np.linspace(-15, 15, 151), 201)
    assert_same_fan(o, h, "config 2")


def test_dropped_rays_and_status_codes(harness):
    """Near-vertical, bbox-leaving and backwards-bouncing rays: statuses and NaN fills agree."""
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    o, h = _both(harness, oenv, 50.0, 0.0, 500e3, np.linspace(-89.5, 89.5, 90), 301)
    assert set(np.unique(o["status"])) >= {0, 2, 4}
    assert_same_fan(o, h, "status codes")


def test_arrival_only_equals_saved_final(harness):
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    ang = np.linspace(-15, 15, 41)
    a = harness.shoot_fan(oenv, 1000.0, 0.0, 300e3, ang)
    b = harness.shoot_fan(oenv, 1000.0, 0.0, 300e3, ang, range_save=np.linspace(0, 300e3, 77))
    assert np.array_equal(a["final"], b["final"], equal_nan=True)
    assert np.array_equal(b["traj"][:, :, -1], b["final"], equal_nan=True)


@pytest.mark.parametrize("rtol", [1e-6, 1e-12, 1e-15])
def test_tolerances(harness, rtol):
    """rtol below 100 eps is clamped (validate_tol); loose and tight tolerances walk the same code."""
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    o, h = _both(harness, oenv, 800.0, 0.0, 80e3, np.linspace(-14, 14, 15), 50, rtol=rtol)
    assert_same_fan(o, h, f"rtol={rtol}")


def test_terminate_backwards_false(harness):
    oenv = oracle_env_from(make_env(cases.config2_arrays()), False)
    o, h = _both(harness, oenv, 50.0, 0.0, 200e3, np.linspace(-80, 80, 33), 101, terminate_backwards=False)
    assert_same_fan(o, h, "terminate_backwards=False")


def test_nonuniform_grid_and_short_bathymetry(harness):
    """Irregular range/depth grids exercise the hunt + bisection cell search; a bathymetry grid that
    ends before the receiver makes bottom bounces leave the spline (BOTTOM_OOB)."""
    rng = np.random.default_rng(7)
    a = cases.config2_arrays(n_r=60, n_z=120, R=2e5)
    r = np.sort(np.concatenate(([0.0, 2e5], rng.uniform(0, 2e5, 58))))
    z = np.sort(np.concatenate(([0.0, 5500.0], rng.uniform(0, 5500.0, 118))))
    c = np.array([cases.munk_ssp(z, 1300 - 200 * ri / 2e5) for ri in r])
    a.update(r=r, z=z, c=c, bathy_r=np.linspace(0, 1.2e5, 30), bathy=np.linspace(4000, 4300, 30))
    oenv = oracle_env_from(make_env(a), False)
    o, h = _both(harness, oenv, 700.0, 0.0, 2e5, np.linspace(-20, 20, 81), 150)
    assert (o["status"] == 5).any()
    assert_same_fan(o, h, "non-uniform grid")


def test_cell_guess_rating(harness):
    """Uniform grids keep the linear guess, the flat-earth depth grid gets the refined one, irregular grids
    switch the speculative trial step off (results are exact in every case; this only decides speed)."""
    assert harness.env_flags(oracle_env_from(make_env(cases.config2_arrays()), False)) == (0, 0)
    assert harness.env_flags(oracle_env_from(make_env(cases.config1_arrays(), flat=True), True)) == (1, 0)
    rng = np.random.default_rng(3)
    a = cases.config2_arrays(n_r=40, n_z=90, R=1e5)
    z = np.sort(np.concatenate(([0.0, 5500.0], rng.uniform(0, 5500.0, 88))))
    a.update(z=z, c=np.array([cases.munk_ssp(z, 1300.0) for _ in a["r"]]))
    assert harness.env_flags(oracle_env_from(make_env(a), False))[1] == 1


def test_exactly_uniform_depth_grid_computes_its_cell_edges(harness):
    """np.arange / np.linspace grids with a representable step (config 3: 5 m; the reference's default 1 m grid before
    flat-earth) are detected at staging time: the trial step computes {zin[j], 1/dz} instead of loading them -- same
    bits as the table, hence the same results as the oracle; a linspace with an inexact step keeps the table."""
    a3 = cases.config2_arrays(n_r=41, n_z=1201, R=2e5)
    a3["z"] = np.linspace(0, 6000, 1201)
    a3["c"] = np.array([cases.munk_ssp(a3["z"], 1300.0 - 1e-3 * r) for r in a3["r"]])
    oenv = oracle_env_from(make_env(a3), False)
    assert harness.env_flags(oenv) == (0, 0) and harness.last_z_uniform == 1 and harness.last_r_uniform == 1  # 5 km x 5 m
    o, h = _both(harness, oenv, 1000.0, 0.0, 2e5, np.linspace(-16, 16, 65), 80)
    assert_same_fan(o, h, "exactly uniform depth and range grids")
    assert (o["n_bottom"] + o["n_surface"]).sum() > 0
    a4 = dict(a3, r=np.linspace(0, 2e5, 44))  # range step 4651.16...: not representable -> range edges stay in the table
    a4["c"] = np.array([cases.munk_ssp(a3["z"], 1300.0 - 1e-3 * r) for r in a4["r"]])
    a4["bathy_r"], a4["bathy"] = a4["r"], np.interp(a4["r"], a3["bathy_r"], a3["bathy"])
    oenv4 = oracle_env_from(make_env(a4), False)
    assert harness.env_flags(oenv4) == (0, 0) and harness.last_z_uniform == 1 and harness.last_r_uniform == 0
    o, h = _both(harness, oenv4, 1000.0, 0.0, 2e5, np.linspace(-16, 16, 33), 0)
    assert_same_fan(o, h, "exactly uniform depth grid, inexact range step")
    harness.env_flags(oracle_env_from(make_env(cases.config2_arrays()), False))
    assert harness.last_z_uniform == 0  # linspace(0, 5500, 500): the reciprocals of the cell widths differ by rounding
    harness.env_flags(oracle_env_from(make_env(cases.config1_arrays(), flat=False), False))
    assert harness.last_z_uniform == 1  # np.arange(0, 6000, 1)
    harness.env_flags(oracle_env_from(make_env(cases.config1_arrays(), flat=True), True))
    assert harness.last_z_uniform == 0  # flat-earth stretching: refined guess + table


def test_depth_derivative_tables_equal_numpy_gradient(harness):
    """The device-side staging of GPU-resident sound-speed grids (prt_ensemble_create_dev) forms dc/dz from
    coefficient tables; they must reproduce np.gradient(c, z, edge_order=1) -- what _unpack_envi hands to the
    integrator -- bit for bit, on uniform, linspace, stretched and irregular depth grids."""
    rng = np.random.default_rng(11)
    grids = [np.arange(0.0, 6000.0, 1.0), np.linspace(0, 5500, 500), np.linspace(0, 5500, 500) * (1 + 1e-4 * np.linspace(0, 1, 500)),
             np.sort(np.concatenate(([0.0, 4000.0], rng.uniform(0, 4000, 77)))), np.array([0.0, 10.0]), np.array([0.0, 3.0, 11.0])]
    for z in grids:
        z = np.ascontiguousarray(z, dtype=np.float64)
        f = np.ascontiguousarray(cases.munk_ssp(z, 1300.0) + rng.normal(0, 0.3, z.shape))
        out = np.empty_like(f)
        harness.lib.hh_gradient(z.ctypes.data, len(z), f.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, np.gradient(f, z, edge_order=1)), len(z)


def test_abscissae_on_grid_nodes(harness):
    """Source depth, receiver range and save points exactly on grid nodes, launch angles including 0: weights
    of exactly 0 and 1 cannot be proven by the speculative lookup, so these steps take the exact re-run."""
    a = cases.config2_arrays(n_r=51, n_z=551, R=1e5)  # 2 km x 10 m grid: nodes at round numbers
    assert a["z"][100] == 1000.0 and a["r"][25] == 50e3
    oenv = oracle_env_from(make_env(a), False)
    ang = np.array([-9.0, -3.0, -1e-9, 0.0, 1e-9, 3.0, 9.0, 14.0])
    for R, n_save in ((50e3, 26), (1e5, 51)):
        o, h = _both(harness, oenv, 1000.0, 0.0, R, ang, n_save)
        assert_same_fan(o, h, f"node-aligned, R={R}")
    o, h = _both(harness, oenv, 0.0, 2e3, 8e4, np.linspace(-5, 12, 9), 40)  # source on the surface node and a range node
    assert_same_fan(o, h, "source on the surface")


def test_backwards_mirrored_environment(harness):
    a = cases.config2_arrays()
    oenv = oracle_env_from(make_env(a), False, mirror=True)
    o, h = _both(harness, oenv, 300.0, -4e5, -1e5, np.linspace(-15, 15, 31), 120)
    assert (o["status"] == 0).any()
    assert_same_fan(o, h, "mirrored environment, negative ranges")


def test_reference_pinned_random_environments(harness):
    """The 40 short and 24 smooth random environments of the reference comparison (fuzz_cases.short_case /
    smooth_case): product core == oracle."""
    from fuzz_cases import short_case, smooth_case

    todo = [("short", s, short_case(s)) for s in range(40)] + [("smooth", s, smooth_case(s)) for s in range(24)]
    for tag, seed, k in todo:
        env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
        oenv = oracle_env_from(env, k["flat"])
        o, h = _both(harness, oenv, k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                     terminate_backwards=k["terminate_backwards"])
        assert_same_fan(o, h, f"{tag} fuzz seed {seed}")


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_random_environments(harness, seed):
    """Random grids (uniform and irregular, noisy sound speed), bathymetry shorter or longer than the
    track, random sources, angle ranges up to +-89 deg, save grids from 2 points up, tolerances."""
    from fuzz_cases import random_case

    k = random_case(seed)
    env = make_env(k["arrays"], flat=k["flat"], lat=k["lat"])
    oenv = oracle_env_from(env, k["flat"])
    o, h = _both(harness, oenv, k["zs"], k["r0"], k["r1"], k["angles"], k["n_save"], rtol=k["rtol"],
                 terminate_backwards=k["terminate_backwards"])
    assert_same_fan(o, h, f"fuzz seed {seed}")
