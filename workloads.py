"""Synthetic environments / workloads named in BASELINE.json `configs`.

Pure numpy, no dependency on the product or on the oracle; used by oracle/make_golden.py (to feed the reference),
by tests/ and by bench.py so all three see bit-identical inputs.  TEST/BENCH INFRASTRUCTURE -- the product package
does not import it.
"""
import numpy as np


def munk_ssp(z, sofar_depth=1300.0, eps=0.00737):
    """Munk profile, restated from ref:environment.py:218-236."""
    zh = 2 * (z - sofar_depth) / sofar_depth
    return 1500 * (1 + eps * (zh - 1 + np.exp(-zh)))


def config1_arrays():
    """Default OceanEnvironment2D() inputs (ref:environment.py:62-90), before flat-earth."""
    z = np.arange(0, 6000, 1)
    r = np.linspace(0, 100e3, 100)
    c = np.array([munk_ssp(z)] * 100)
    bathy_r = np.linspace(0, 100e3, 100)
    bathy = np.linspace(4500, 4900, 100)
    return dict(c=c, r=r, z=z, bathy=bathy, bathy_r=bathy_r)


def config2_arrays(n_r=200, n_z=500, R=5e5):
    """Range-dependent gridded SSP (200 x 500) with a seamount."""
    r = np.linspace(0, R, n_r)
    z = np.linspace(0, 5500, n_z)
    c = np.empty((n_r, n_z))
    for i, ri in enumerate(r):
        c[i] = munk_ssp(z, 1300.0 - 300.0 * ri / R) + 2.0 * np.sin(2 * np.pi * ri / 80e3) * np.exp(-z / 800.0)
    bathy_r = r.copy()
    bathy = 4800.0 - 1500.0 * np.exp(-(((r - 250e3) / 40e3) ** 2))
    return dict(c=c, r=r, z=z, bathy=bathy, bathy_r=bathy_r)


def internal_wave_dc(r, z, J=20, K=8, seed=20260101, rms_axis=0.5):
    """Builder-defined synthetic internal-wave sound-speed perturbation (config 3/5)."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi, size=(J, K))
    kappa = 2 * np.pi / (np.geomspace(5e3, 100e3, K))
    xi = (1.0 - np.exp(-z / 1000.0)) / (1.0 - np.exp(-z[-1] / 1000.0))
    dc = np.zeros((len(r), len(z)))
    for j in range(1, J + 1):
        vert = np.sin(j * np.pi * xi) * np.exp(-z / 1000.0)
        a = 1.0 / (j * j + 9.0)
        for k in range(K):
            dc += a * np.cos(kappa[k] * r + phi[j - 1, k])[:, None] * vert[None, :]
    iax = np.argmin(np.abs(z - 1300.0))
    dc *= rms_axis / np.sqrt(np.mean(dc[:, iax] ** 2))
    return dc


def config3_arrays(n_r=2001, n_z=1201, R=1e6, seed=20260101):
    """Munk + internal-wave-perturbed SSP, 1000 km basin scale (config 3)."""
    r = np.linspace(0, R, n_r)
    z = np.linspace(0, 6000, n_z)
    c = munk_ssp(z)[None, :] + internal_wave_dc(r, z, seed=seed)
    bathy_r = np.linspace(0, R, 201)
    bathy = 5200.0 + 300.0 * np.sin(2 * np.pi * bathy_r / 3.3e5)
    return dict(c=c, r=r, z=z, bathy=bathy, bathy_r=bathy_r)


def config5_realisation(i, n_r=200, n_z=500, R=5e5):
    """Tomography ensemble member i: config-2 grid + IW perturbation with seed 1000+i (config 5)."""
    a = config2_arrays(n_r, n_z, R)
    a["c"] = a["c"] + internal_wave_dc(a["r"], a["z"], J=8, seed=1000 + i)
    return a
