"""Ocean environment: the host-side input adapter of the ray path.

Mirrors the constructor, attributes and error behaviour of the reference's ``OceanEnvironment2D``
(/root/reference/src/pygenray/environment.py:14-119) plus ``munk_ssp`` (:218-236), ``eflat``
(:371-401) and ``flat_earth_transform`` (:121-151) -- the parts that produce the seven arrays the
ray integrator consumes.  Works without xarray (see :mod:`pygenray_b200._xr`).  One-off
preprocessing that is not on the path (``flat_earth_transform_rd``, ``flat_earth_c``,
``eflatinv``, plotting) is out of scope.
"""
import numpy as np
import scipy.interpolate

from . import _xr
from ._xr import DataArray

__all__ = ["OceanEnvironment2D", "munk_ssp", "eflat", "DataArray"]

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
