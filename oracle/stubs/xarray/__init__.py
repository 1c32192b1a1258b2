"""Minimal stand-in for xarray so the UNMODIFIED reference package can be imported
in a container without xarray.  Test infrastructure only.

Implements exactly the DataArray surface that /root/reference/src/pygenray/environment.py
and launch_rays.py touch: construction with dims/coords, .values/.dims/.ndim/.sizes,
coordinate attribute access (.range/.depth -> object with .values), .differentiate
(== numpy.gradient along the named dim, edge_order=1: xarray's real semantics),
.isel, concat, assign_coords; and, for flat_earth_c (environment.py:239-368): numpy arithmetic on
DataArrays / coordinates, coordinates that live on a named dimension ("lat" along "range", sliced by
.isel), and .interp(dim=new) for 1-D arrays -- scipy.interpolate.interp1d(kind="linear",
bounds_error=False, fill_value=nan), which is what xarray's own interp() calls for a 1-D linear
interpolation (xarray/core/missing.py: interp_func -> _interp1d -> ScipyInterpolator).
"""
import numpy as np


class _ArrayOps:
    """numpy semantics for + - * / ** and unary minus: operands are unwrapped to ndarrays."""
    __array_priority__ = 1000

    @staticmethod
    def _v(x):
        return x.values if isinstance(x, _ArrayOps) else x

    def __add__(self, o): return self.values + self._v(o)
    def __radd__(self, o): return self._v(o) + self.values
    def __sub__(self, o): return self.values - self._v(o)
    def __rsub__(self, o): return self._v(o) - self.values
    def __mul__(self, o): return self.values * self._v(o)
    def __rmul__(self, o): return self._v(o) * self.values
    def __truediv__(self, o): return self.values / self._v(o)
    def __rtruediv__(self, o): return self._v(o) / self.values
    def __pow__(self, o): return self.values ** self._v(o)
    def __neg__(self): return -self.values

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return getattr(ufunc, method)(*[self._v(i) for i in inputs], **kwargs)


class _Coord(_ArrayOps):
    def __init__(self, values, name, dim=None):
        self.values = np.asarray(values)
        self.name = name
        self.dim = dim  # the dimension a 1-D coordinate lives on (None: scalar)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def max(self):
        return self.values.max()

    def min(self):
        return self.values.min()


class DataArray(_ArrayOps):
    def __init__(self, data, dims=None, coords=None, name=None):
        self.values = np.asarray(data)
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims) if dims is not None else tuple(
            f"dim_{i}" for i in range(self.values.ndim))
        self.coords = {}
        if coords is not None:
            for k, v in dict(coords).items():
                dim = None
                if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], str):
                    dim, v = v  # xarray's (dim, values) form
                elif isinstance(v, _Coord):
                    dim = v.dim
                if isinstance(v, (DataArray, _Coord)):
                    v = v.values
                v = np.asarray(v)
                if dim is None and v.ndim == 1:
                    if k in self.dims:
                        dim = k
                    else:  # a non-index coordinate given as a bare array: the one dimension of that length
                        fits = [d for d, n in zip(self.dims, self.values.shape) if n == v.shape[0]]
                        dim = fits[0] if len(fits) == 1 else None
                self.coords[k] = _Coord(v, k, dim)
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
        return DataArray(g, dims=self.dims, coords=self._coord_items())

    def _coord_items(self):
        return {k: ((v.dim, v.values) if v.dim is not None else v.values) for k, v in self.coords.items()}

    def isel(self, indexers):
        out = self.values
        dims = list(self.dims)
        coords = self._coord_items()
        for d, idx in indexers.items():
            ax = dims.index(d)
            out = np.take(out, idx, axis=ax)
            dims.pop(ax)
            for k, v in list(coords.items()):  # coordinates on the indexed dimension become scalars
                if isinstance(v, tuple) and v[0] == d:
                    coords[k] = np.take(v[1], idx)
        return DataArray(out, dims=dims, coords=coords)

    def interp(self, **new):
        (dim, x_new), = new.items()
        assert self.dims == (dim,), "stand-in: 1-D linear interpolation only"
        from scipy.interpolate import interp1d

        x_new = np.asarray(x_new.values if isinstance(x_new, _ArrayOps) else x_new)
        f = interp1d(self.coords[dim].values, self.values, kind="linear", bounds_error=False, fill_value=np.nan)
        coords = {k: v for k, v in self._coord_items().items() if k != dim}
        coords[dim] = x_new
        return DataArray(f(x_new), dims=(dim,), coords=coords)

    def assign_coords(self, coords):
        new = self._coord_items()
        for k, v in coords.items():
            new[k] = v if isinstance(v, (_Coord, tuple)) else (v.values if isinstance(v, DataArray) else np.asarray(v))
        return DataArray(self.values, dims=self.dims, coords=new)

    def copy(self, deep=True):
        return DataArray(self.values.copy(), dims=self.dims,
                         coords={k: ((v[0], v[1].copy()) if isinstance(v, tuple) else np.copy(v))
                                 for k, v in self._coord_items().items()})


def concat(objs, dim):
    vals = np.stack([o.values for o in objs], axis=0)
    first = objs[0]
    coords = {k: v for k, v in first._coord_items().items() if isinstance(v, tuple)}  # scalar coords differ per slice
    return DataArray(vals, dims=(dim,) + tuple(first.dims), coords=coords)
