"""pygenray_b200 -- B200-native ray-fan shooting and eigenray search (drop-in for that path of pygenray).

    import pygenray_b200 as pr
    env  = pr.OceanEnvironment2D()                       # Munk profile, 100 km
    fan  = pr.shoot_rays(1000.0, 0.0, angles, 100e3, 1001, env)
    eigs = pr.find_eigenrays(fan, [1000.0, 2500.0], 1000.0, 0.0, 100e3, 1001, env, ztol=1e-3)

Everything numerical runs in libprt.so (hand-written sm_100a CUDA, C ABI in include/prt.h).
There is no CPU fallback: without the built library and a CUDA device the compute calls raise.
"""
from . import _lib, dist, ensemble
from .eigenrays import eigenray_table, find_brackets, find_eigenrays, refine_brackets
from .environment import DataArray, OceanEnvironment2D, eflat, munk_ssp
from .launch_rays import (DeviceEnsemble, DeviceEnvironment, _mirror_envi_arrays, _unpack_envi,
                          bottom_angle_ppoly, shoot_ensemble_arrays, shoot_fan_arrays, shoot_ray, shoot_rays)
from .ray_objects import DeviceRows, EigenRays, Ray, RayFan

__all__ = [
    "OceanEnvironment2D", "DataArray", "munk_ssp", "eflat", "shoot_rays", "shoot_ray", "find_eigenrays",
    "Ray", "RayFan", "EigenRays", "DeviceRows", "shoot_fan_arrays", "shoot_ensemble_arrays", "find_brackets", "refine_brackets", "eigenray_table",
    "DeviceEnvironment",
    "DeviceEnsemble", "dist", "ensemble",
    "bottom_angle_ppoly", "_unpack_envi", "_mirror_envi_arrays",
]
__version__ = "0.1.0"
