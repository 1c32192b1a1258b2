"""Labelled-array shim so the environment object works with or without xarray installed.

The reference's ``OceanEnvironment2D`` is built on ``xarray.DataArray``
(/root/reference/src/pygenray/environment.py:6, 66-99).  The B200 path only needs the numbers, the
dimension names and the coordinate vectors, so anything that quacks like a DataArray
(``.values``, ``.dims``, ``.coords[name].values``) is accepted; :class:`DataArray` below is the
minimal concrete type used when xarray is absent (it is in this image).
"""
import numpy as np

try:  # pragma: no cover - xarray is optional
    import xarray as _xarray
except Exception:  # noqa: BLE001
    _xarray = None


class _Coordinate:
    __slots__ = ("values", "name")

    def __init__(self, values, name):
        self.values = np.asarray(values)
        self.name = name

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def __getitem__(self, k):
        return self.values[k]

    def max(self):
        return self.values.max()

    def min(self):
        return self.values.min()


class DataArray:
    """``DataArray(data, dims=[...], coords={...})`` with the handful of members the path uses."""

    def __init__(self, data, dims=None, coords=None, name=None):
        self.values = np.asarray(data)
        if isinstance(dims, str):
            dims = (dims,)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError("number of dims does not match data.ndim")
        self.coords = {}
        for key, val in dict(coords or {}).items():
            val = getattr(val, "values", val)
            self.coords[key] = _Coordinate(np.asarray(val), key)
        self.name = name

    ndim = property(lambda self: self.values.ndim)
    shape = property(lambda self: self.values.shape)
    sizes = property(lambda self: dict(zip(self.dims, self.values.shape)))

    def __getattr__(self, item):
        coords = self.__dict__.get("coords")
        if coords is not None and item in coords:
            return coords[item]
        raise AttributeError(item)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def _coord_dict(self):
        return {k: v.values for k, v in self.coords.items()}

    def differentiate(self, dim, edge_order=1):
        """xarray semantics: ``np.gradient`` along ``dim`` against its coordinate."""
        axis = self.dims.index(dim)
        grad = np.gradient(self.values, self.coords[dim].values, axis=axis, edge_order=edge_order)
        return DataArray(grad, dims=self.dims, coords=self._coord_dict())

    def copy(self, deep=True):
        coords = {k: (v.copy() if deep else v) for k, v in self._coord_dict().items()}
        return DataArray(self.values.copy() if deep else self.values, dims=self.dims, coords=coords)


def is_dataarray(obj):
    if isinstance(obj, DataArray):
        return True
    if _xarray is not None and isinstance(obj, _xarray.DataArray):
        return True
    return False


def dims_of(obj):
    return tuple(obj.dims)


def values_of(obj):
    return np.asarray(obj.values)


def coord_of(obj, name):
    """1-D coordinate vector ``name`` of a DataArray-like object."""
    c = obj.coords[name]
    return np.asarray(getattr(c, "values", c))


def differentiate(obj, dim):
    """``obj.differentiate(dim).values`` without needing xarray."""
    axis = dims_of(obj).index(dim)
    return np.gradient(values_of(obj), coord_of(obj, dim), axis=axis, edge_order=1)
