"""Ocean environment: the host-side input adapter of the ray path.

Mirrors the constructor, attributes and error behaviour of the reference's ``OceanEnvironment2D``
(/root/reference/src/pygenray/environment.py:14-119) plus ``munk_ssp`` (:218-236), ``eflat``
(:371-401) and ``flat_earth_transform`` (:121-151) -- the parts that produce the seven arrays the
ray integrator consumes.  Works without xarray (see :mod:`pygenray_b200._xr`).  The range-dependent
transform (``flat_earth_c`` :239-368, ``flat_earth_transform_rd`` :153-173) runs on the GPU for whole grids and
ensembles of grids (``prt_flat_earth``); ``eflatinv`` / ``_ridder`` (:404-535) are host conveniences.  Plotting is out
of scope.
"""
import numpy as np
import scipy.interpolate

from . import _xr
from ._xr import DataArray

__all__ = ["OceanEnvironment2D", "munk_ssp", "eflat", "eflatinv", "flat_earth_c", "flat_earth_device", "DataArray"]

_WGS84_A = 6378137.0
_WGS84_B = 6356752.314


def munk_ssp(z, sofar_depth=1300, eps=0.00737):
    """Munk canonical sound-speed profile c(z) [m/s] (environment.py:218-236)."""
    zeta = 2 * (z - sofar_depth) / sofar_depth
    return 1500 * (1 + eps * (zeta - 1 + np.exp(-zeta)))


def _earth_radius(lat):
    """Local WGS-84 radius used by the earth-flattening transform (environment.py:382-395)."""
    a2 = _WGS84_A * _WGS84_A
    b2 = _WGS84_B * _WGS84_B
    fact = (_WGS84_B / _WGS84_A) ** 4
    ll = np.pi * lat / 180.0
    ree1 = a2 / np.sqrt(a2 * np.cos(ll) * np.cos(ll) + b2 * np.sin(ll) * np.sin(ll))
    return ree1 * np.sqrt(np.cos(ll) * np.cos(ll) + fact * np.sin(ll) * np.sin(ll))


def eflat(dep, lat, cs=None):
    """Earth-flattening transform of depths (and sound speeds): returns ``(depf, csf)``."""
    if cs is None:
        cs = np.zeros_like(dep)
    re = _earth_radius(lat)
    E = dep / re
    depf = dep * (1.0 + E * (0.50 + E / 3.0))
    csf = cs * (1.0 + E * (1.0 + E))
    return depf, csf


class OceanEnvironment2D:
    """2-D ocean acoustic environment (sound speed on a range x depth grid + bathymetry).

    Parameters and attributes follow the reference class: ``sound_speed``, ``bathymetry``,
    ``dcdz``, ``latitude``, ``sound_speed_fe`` / ``bathymetry_fe`` (when
    ``flat_earth_transform=True``), ``bottom_angle`` [deg] and ``bottom_angle_interp``.
    """

    def __init__(self, sound_speed=None, bathymetry=None, lat=35, flat_earth_transform=True, verbose=False):
        self.latitude = lat

        if sound_speed is None:
            z = np.arange(0, 6000, 1)
            sound_speed = DataArray(
                np.array([munk_ssp(z)] * 100),
                dims=["range", "depth"],
                coords={"depth": z, "range": np.linspace(0, 100e3, 100)},
            )
        else:
            if not _xr.is_dataarray(sound_speed):
                raise TypeError("sound_speed must be an xarray DataArray.")
            if sound_speed.ndim not in [1, 2]:
                raise ValueError("sound_speed must be 1D or 2D.")
            if "depth" not in sound_speed.dims:
                raise ValueError("sound_speed must have a 'depth' dimension.")
            if sound_speed.ndim == 2 and "range" not in sound_speed.dims:
                raise ValueError("2D sound_speed must have a 'range' dimension.")

        if bathymetry is None:
            bathymetry = DataArray(
                np.linspace(4500, 4900, 100), dims=["range"], coords={"range": np.linspace(0, 100e3, 100)}
            )
        else:
            if not _xr.is_dataarray(bathymetry):
                raise TypeError("bathymetry must be an xarray DataArray.")
            if bathymetry.ndim != 1:
                raise ValueError("bathymetry must be 1D.")
            if "range" not in bathymetry.dims:
                raise ValueError("bathymetry must have a 'range' dimension.")

        self.sound_speed = sound_speed
        self.dcdz = _xr.differentiate(sound_speed, "depth")
        self.bathymetry = bathymetry

        if flat_earth_transform:
            self.flat_earth_transform(lat=lat)

        # bottom slope from the UN-flattened bathymetry (environment.py:111-119)
        bathy_r = _xr.coord_of(bathymetry, "range")
        slope = np.gradient(_xr.values_of(bathymetry), bathy_r)
        self.bottom_angle = np.degrees(np.arctan(slope))
        self.bottom_angle_interp = scipy.interpolate.interp1d(bathy_r, self.bottom_angle, kind="cubic")

    def flat_earth_transform(self, lat):
        """Single-latitude earth flattening of the sound speed and bathymetry (WGS-84)."""
        ss = self.sound_speed
        depth = _xr.coord_of(ss, "depth")
        vals = _xr.values_of(ss)
        dims = _xr.dims_of(ss)
        if "range" not in dims:
            raise KeyError("range")
        if dims.index("range") != 0:
            vals = np.moveaxis(vals, dims.index("range"), 0)
        depf, _ = eflat(depth, lat)
        cf = np.stack([eflat(depth, lat, row)[1] for row in vals]) if vals.ndim == 2 else eflat(depth, lat, vals)[1]
        self.sound_speed_fe = DataArray(
            cf, dims=["range", "depth"], coords={"range": _xr.coord_of(ss, "range"), "depth": depf}
        )
        bathy_flat, _ = eflat(_xr.values_of(self.bathymetry), lat)
        self.bathymetry_fe = DataArray(
            bathy_flat, dims=["range"], coords={"range": _xr.coord_of(self.bathymetry, "range")}
        )

    def flat_earth_transform_rd(self):
        """Range-dependent earth flattening (environment.py:153-173): every range column with its own latitude
        (``sound_speed`` needs a ``lat`` coordinate along range).  As in the reference the flattened field is put
        back on the original depth grid and the bathymetry is copied unchanged."""
        c_fe = flat_earth_c(self.sound_speed, verbose=False)
        self.sound_speed_fe = c_fe
        self.dcdz = c_fe.differentiate("depth")
        self.bathymetry_fe = self.bathymetry.copy(deep=True)


def flat_earth_device(c_dev, depth, lat, resample):
    """Earth flattening of sound-speed grids that live on the GPU: ``c_dev`` CUDA float64 ``[..., n_r, n_z]``
    (any number of leading realisation axes), ``depth`` [n_z] un-flattened, ``lat`` scalar or [n_r] degrees.
    ``resample=False``: values of the single-latitude transform (on the stretched grid); ``True``: ``flat_earth_c``
    (back on the original grid).  Returns a new tensor of the same shape."""
    import ctypes as C  # noqa: PLC0415

    import torch  # noqa: PLC0415

    from . import _lib  # noqa: PLC0415

    L = _lib.require_gpu()
    if not (isinstance(c_dev, torch.Tensor) and c_dev.is_cuda and c_dev.dtype == torch.float64 and c_dev.ndim >= 2):
        raise TypeError("c_dev must be a CUDA float64 tensor [..., n_r, n_z]")
    c_dev = c_dev.contiguous()
    n_r, n_z = int(c_dev.shape[-2]), int(c_dev.shape[-1])
    n_env = int(c_dev.numel() // (n_r * n_z))
    depth = np.ascontiguousarray(depth, dtype=np.float64)
    re = np.ascontiguousarray(np.broadcast_to(_earth_radius(np.asarray(lat, dtype=np.float64)), (n_r,)), dtype=np.float64)
    if depth.shape != (n_z,):
        raise ValueError("depth must have one entry per depth node")
    out = torch.empty_like(c_dev)
    with torch.cuda.device(c_dev.device):
        d_d, re_d = torch.from_numpy(depth).to(c_dev.device), torch.from_numpy(re).to(c_dev.device)
        rc = L.prt_flat_earth(n_env, n_r, n_z, C.c_void_p(d_d.data_ptr()), C.c_void_p(re_d.data_ptr()),
                              C.c_void_p(c_dev.data_ptr()), int(bool(resample)), C.c_void_p(out.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "prt_flat_earth")
    return out


def flat_earth_c(c, verbose=False, n_cpus=None, chunk_size=None):
    """Range-dependent flat-earth transform of a 2-D sound-speed slice with coordinates ``depth``, ``range`` and
    ``lat`` (one latitude per range); same call as the reference (environment.py:239-305).  ``n_cpus`` and
    ``chunk_size`` are accepted and ignored: the whole grid is one kernel launch.  Returns a DataArray with
    dimensions (range, depth) and coordinates range, depth, lat."""
    import torch  # noqa: PLC0415

    from . import _lib  # noqa: PLC0415

    _lib.require_gpu()
    dims = _xr.dims_of(c)
    vals = _xr.values_of(c).astype(np.float64)
    if vals.ndim != 2 or "range" not in dims or "depth" not in dims:
        raise ValueError("c must be 2-D with dimensions 'depth' and 'range'")
    if dims.index("range") != 0:
        vals = vals.T
    depth, rng, lat = _xr.coord_of(c, "depth"), _xr.coord_of(c, "range"), _xr.coord_of(c, "lat")
    if verbose:
        print(f"Processing {vals.shape[0]} range points in one GPU launch")
    out = flat_earth_device(torch.from_numpy(np.ascontiguousarray(vals)).cuda(), depth, lat, resample=True)
    return DataArray(out.cpu().numpy(), dims=["range", "depth"], coords={"range": rng, "depth": depth, "lat": lat})


def eflatinv(depf, lat, csf=None):
    """Inverse earth flattening (environment.py:404-466): true depth (and sound speed) from flat-earth values.
    The reference brackets the root and runs Ridder's method to 1 mm (:468-535); z_f = z (1 + E / 2 + E^2 / 3) is a
    cubic in z with a single positive root, solved here by Newton's method from z = z_f (converges quadratically,
    well below the reference's 1 mm)."""
    depf = np.reshape(np.asarray(depf, dtype=np.float64), (-1,))
    lat = np.reshape(np.asarray(lat, dtype=np.float64), (-1,))
    csf = np.zeros(depf.shape) if csf is None else np.reshape(np.asarray(csf, dtype=np.float64), (-1,))
    re = _earth_radius(lat)
    dep = depf.copy()
    for _ in range(8):
        E = dep / re
        f = dep * (1.0 + E * (0.5 + E / 3.0)) - depf
        df = 1.0 + E + E * E
        dep = dep - f / df
    E = dep / re
    return dep, csf / (1.0 + E * (1.0 + E))
