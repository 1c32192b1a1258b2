"""Minimal stand-in for xarray so the UNMODIFIED reference package can be imported
in a container without xarray.  Test infrastructure only.

Implements exactly the DataArray surface that /root/reference/src/pygenray/environment.py
and launch_rays.py touch: construction with dims/coords, .values/.dims/.ndim/.sizes,
coordinate attribute access (.range/.depth -> object with .values), .differentiate
(== numpy.gradient along the named dim, edge_order=1: xarray's real semantics),
.isel, concat, assign_coords.
"""
import numpy as np


class _Coord:
    def __init__(self, values, name):
        self.values = np.asarray(values)
        self.name = name

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def max(self):
        return self.values.max()

    def min(self):
        return self.values.min()


class DataArray:
    def __init__(self, data, dims=None, coords=None, name=None):
        self.values = np.asarray(data)
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims) if dims is not None else tuple(
            f"dim_{i}" for i in range(self.values.ndim))
        self.coords = {}
        if coords is not None:
            for k, v in dict(coords).items():
                if isinstance(v, (DataArray, _Coord)):
                    v = v.values
                self.coords[k] = _Coord(np.asarray(v), k)
        self.name = name

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def shape(self):
        return self.values.shape

    @property
    def sizes(self):
        return dict(zip(self.dims, self.values.shape))

    def __getattr__(self, item):
        coords = self.__dict__.get("coords", {})
        if item in coords:
            return coords[item]
        raise AttributeError(item)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def differentiate(self, dim, edge_order=1):
        ax = self.dims.index(dim)
        g = np.gradient(self.values, self.coords[dim].values, axis=ax,
                        edge_order=edge_order)
        return DataArray(g, dims=self.dims, coords={k: v.values for k, v in self.coords.items()})

    def isel(self, indexers):
        out = self.values
        dims = list(self.dims)
        coords = {k: v.values for k, v in self.coords.items()}
        for d, idx in indexers.items():
            ax = dims.index(d)
            out = np.take(out, idx, axis=ax)
            dims.pop(ax)
            coords.pop(d, None)
        return DataArray(out, dims=dims, coords=coords)

    def assign_coords(self, coords):
        new = {k: v.values for k, v in self.coords.items()}
        for k, v in coords.items():
            new[k] = v.values if isinstance(v, (DataArray, _Coord)) else np.asarray(v)
        return DataArray(self.values, dims=self.dims, coords=new)

    def copy(self, deep=True):
        return DataArray(self.values.copy(), dims=self.dims,
                         coords={k: v.values.copy() for k, v in self.coords.items()})


def concat(objs, dim):
    vals = np.stack([o.values for o in objs], axis=0)
    first = objs[0]
    coords = {k: v.values for k, v in first.coords.items()}
    return DataArray(vals, dims=(dim,) + tuple(first.dims), coords=coords)
