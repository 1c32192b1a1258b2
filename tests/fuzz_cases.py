"""Seeded random environments / shots for parity fuzzing (used by CPU and GPU parity tests)."""
import numpy as np

import workloads as cases


def random_case(seed):
    rng = np.random.default_rng(seed)
    n_r = int(rng.integers(4, 60))
    n_z = int(rng.integers(8, 200))
    R = float(rng.uniform(2e4, 3e5))
    zmax = float(rng.uniform(2000, 6500))
    if rng.random() < 0.5:
        r = np.linspace(0, R, n_r)
    else:
        r = np.sort(np.concatenate(([0.0, R], rng.uniform(0, R, n_r - 2))))
    if rng.random() < 0.5:
        z = np.linspace(0, zmax, n_z)
    else:
        z = np.sort(np.concatenate(([0.0, zmax], rng.uniform(0, zmax, n_z - 2))))
    # keep grid points distinct (duplicate coordinates make the reference divide by zero)
    r = np.unique(np.round(r, 3)); z = np.unique(np.round(z, 3))
    n_r, n_z = len(r), len(z)
    axis = rng.uniform(600, 1600)
    c = np.array([cases.munk_ssp(z, axis + rng.uniform(-200, 200) * ri / R) for ri in r])
    c += rng.normal(0, rng.choice([0.0, 0.05, 0.5]), size=c.shape)
    n_b = int(rng.integers(4, 40))
    b_end = R * float(rng.choice([1.0, 1.0, 1.0, 0.6, 1.3]))
    bathy_r = np.linspace(0, b_end, n_b)
    depth0 = rng.uniform(0.4, 1.1) * zmax
    bathy = depth0 + rng.uniform(-0.2, 0.2) * zmax * np.sin(np.linspace(0, rng.uniform(1, 9), n_b)) \
        + rng.normal(0, 10, n_b)
    arrays = dict(c=c, r=r, z=z, bathy=bathy, bathy_r=bathy_r)
    n_rays = int(rng.integers(1, 48))
    lim = float(rng.choice([8.0, 15.0, 30.0, 60.0, 89.0]))
    angles = np.sort(rng.uniform(-lim, lim, n_rays))
    zs = float(rng.uniform(5, min(depth0, zmax) * 0.9))
    r0 = float(rng.choice([0.0, rng.uniform(0, 0.3 * R)]))
    r1 = float(rng.uniform(r0 + 0.2 * (R - r0), R * float(rng.choice([1.0, 1.0, 0.9]))))
    n_save = int(rng.choice([0, 2, 3, 17, 120, 400]))
    rtol = float(rng.choice([1e-9, 1e-9, 1e-6, 1e-11]))
    tb = bool(rng.random() < 0.8)
    return dict(arrays=arrays, flat=bool(rng.random() < 0.3), zs=zs, r0=r0, r1=r1, angles=angles, n_save=n_save,
                rtol=rtol, terminate_backwards=tb, lat=float(rng.uniform(-60, 60)))


def short_case(seed):
    """The same generator scaled to short ranges (2-30 km, a few dozen steps per ray): there the reference reproduces
    itself to ~1e-9, so the oracle can be compared with it ray by ray on random environments (irregular grids, noisy
    sound speeds, flat-earth on / off, offset sources); used by oracle/make_golden.py: case_fuzz_short and
    tests/test_oracle_vs_reference.py."""
    k = random_case(seed)
    f = 0.1
    a = k["arrays"]
    k["arrays"] = dict(c=a["c"], r=a["r"] * f, z=a["z"], bathy=a["bathy"], bathy_r=a["bathy_r"] * f)
    k["r0"], k["r1"] = k["r0"] * f, k["r1"] * f
    ang = k["angles"]
    k["angles"] = ang[np.abs(ang) <= 30.0][:8] if (np.abs(ang) <= 30.0).any() else np.clip(ang[:4], -30.0, 30.0)
    k["n_save"] = 17
    return k


def smooth_case(seed):
    """Random SMOOTH environments on 1-2 m depth grids (the regime in which the reference reproduces itself to ~1e-9, so
    the north-star tolerances can be checked directly): range-dependent Munk axis, a smooth range-dependent perturbation,
    sloping / undulating bathymetry (exercises the bottom-angle spline and the reflection law), flat-earth on / off,
    10-40 km, launch angles up to +-25 deg.  Used by oracle/make_golden.py: case_fuzz_smooth and the tests."""
    rng = np.random.default_rng(5000 + seed)
    n_r = int(rng.integers(5, 40))
    R = float(rng.uniform(1e4, 4e4))
    zmax = float(rng.uniform(3000, 6000))
    dz = float(rng.choice([1.0, 2.0]))
    z = np.arange(0.0, zmax, dz)
    r = np.linspace(0.0, R, n_r)
    axis, drift = rng.uniform(700, 1500), rng.uniform(-150, 150)
    amp, L, zs_scale = rng.uniform(0.0, 1.5), rng.uniform(0.5, 3.0) * R, rng.uniform(400, 1200)
    c = np.array([cases.munk_ssp(z, axis + drift * ri / R) + amp * np.sin(2 * np.pi * ri / L) * np.exp(-z / zs_scale)
                  for ri in r])
    n_b = int(rng.integers(5, 30))
    bathy_r = np.linspace(0.0, R, n_b)
    d0 = rng.uniform(0.45, 0.9) * z[-1]
    bathy = d0 + rng.uniform(-0.08, 0.08) * z[-1] * bathy_r / R \
        + rng.uniform(0, 0.04) * z[-1] * np.sin(np.linspace(0, rng.uniform(1, 7), n_b))
    bathy = np.minimum(bathy, z[-1] - 10.0)
    arrays = dict(c=c, r=r, z=z, bathy=bathy, bathy_r=bathy_r)
    zs = float(rng.uniform(50, 0.8 * bathy.min()))
    angles = np.sort(rng.uniform(-25.0, 25.0, 8))
    return dict(arrays=arrays, flat=bool(rng.random() < 0.4), zs=zs, r0=0.0, r1=R, angles=angles, n_save=17,
                rtol=float(rng.choice([1e-9, 1e-10])), terminate_backwards=True, lat=float(rng.uniform(-60, 60)))
