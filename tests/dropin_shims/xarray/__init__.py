"""Test-only alias: `xr.DataArray` of the reference's test files -> the product's own minimal DataArray
(xarray is not in the image).  Used by tests/test_reference_suite.py."""
from pygenray_b200 import DataArray  # noqa: F401
