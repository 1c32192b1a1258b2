"""Test-only alias: lets the reference's OWN test files (`from pygenray.launch_rays import shoot_rays`, ...) run
against pygenray_b200 unchanged.  Used by tests/test_reference_suite.py; not product code."""
import sys

import pygenray_b200
from pygenray_b200 import *  # noqa: F401, F403
from pygenray_b200 import eigenrays, environment, launch_rays, ray_objects

for _name, _mod in dict(environment=environment, launch_rays=launch_rays, ray_objects=ray_objects,
                        eigenrays=eigenrays).items():
    sys.modules[__name__ + "." + _name] = _mod
__version__ = getattr(pygenray_b200, "__version__", "0")
