"""A small stand-in for the part of xarray's public contract that the reference's matrix-build chain uses - GOLDEN-VECTOR
INFRASTRUCTURE ONLY (xarray is not installable in this image; the reference's functions are executed UNCHANGED on these
objects by make_golden_compute_phase.py).  Every method states the documented xarray behaviour it reproduces:

* ``Dataset[name]`` / ``Dataset[[names]]``, ``ds.time`` / ``ds.level`` attribute access to coordinates, ``bool(ds)`` =
  "has data variables", ``ds.attrs`` / ``da.attrs`` (copied on construction);
* ``sel(time=slice(a, b), level=[...])``: label based, both slice ends INCLUSIVE, list labels in request order,
  ``KeyError`` for a missing label (pandas ``Index.slice_indexer`` / ``get_indexer``);
* ``resample(time=delta).nearest()``: a reindex onto the labels of the resampling bins (pandas ``resample(delta)``,
  default ``origin="start_day"``) with ``method="nearest"`` - carried out by pandas itself;
* ``mean(dim)`` / ``std(dim)``: NaN-skipping for floating point (``np.nanmean`` / ``np.nanstd``, ddof = 0, what xarray
  calls when bottleneck is absent - the reference does not depend on it), the reduced dimension's coordinate is dropped,
  attributes are dropped (``keep_attrs`` default);
* ``ds - other`` / ``ds / other``: variable by variable, broadcasting BY DIMENSION NAME, the first operand's dimension
  order is kept, attributes are dropped;
* ``stack(space=[a, b, c])``: the new dimension is appended LAST, its positions run in C order over (a, b, c) in the order
  GIVEN (whatever the variable's own axis order), its coordinate values are the (a, b, c) label tuples;
* ``xr.DataArray(data, dims=, coords=, attrs=)`` with coords given as arrays, ``(dim, array)`` pairs or DataArrays;
  ``xr.Dataset(data_vars, coords=)`` merges the variables' own coordinates with ``coords``;
* ``xr.concat(objs, dim=)`` along an existing dimension: data and every coordinate on that dimension are concatenated,
  attributes of the first object are kept;
* ``ds.coords[name] = values``, ``ds.assign_coords(name=(dim, values))`` (returns a new object).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _as_dims(dims):
    return (dims,) if isinstance(dims, str) else tuple(dims)


class Coords(dict):
    """name -> DataArray; ``keys()`` order = insertion order, like xarray's coordinate mapping.  Assigning a bare array
    means "index coordinate of the dimension of that name" (slice_tools.py:408), a ``(dim, array)`` pair names its
    dimension, a DataArray contributes its variable (dims + values)."""

    def __setitem__(self, key, value):
        if isinstance(value, DataArray):
            value = DataArray(value.values, value.dims)
        elif isinstance(value, tuple) and len(value) == 2 and isinstance(value[0], (str, tuple, list)):
            value = DataArray(value[1], _as_dims(value[0]))
        else:
            value = DataArray(value, (key,))
        dict.__setitem__(self, key, value)

    set = __setitem__


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self._data = data.values if isinstance(data, DataArray) else np.asarray(data)
        if self._data.dtype.kind == "M":                 # xarray holds times as datetime64[ns] (the reference relies on it:
            self._data = self._data.astype("datetime64[ns]")     # ``time.values[0].astype(int) * 1e-9``, slice_tools.py:121)
        if dims is None:
            raise TypeError("the stand-in needs explicit dims")
        self.dims = _as_dims(dims)
        assert len(self.dims) == self._data.ndim, (self.dims, self._data.shape)
        self.coords = Coords()
        for k, v in (coords.items() if coords is not None else ()):
            self.coords.set(k, v)
        for k, c in self.coords.items():
            for dm, n in zip(c.dims, c.shape):
                assert dm in self.dims and self._data.shape[self.dims.index(dm)] == n, f"coordinate {k} does not fit"
        self.attrs = dict(attrs) if attrs else {}
        self.name = name

    values = property(lambda self: self._data)
    data = property(lambda self: self._data)
    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __getitem__(self, key):
        """positional indexing of a 1-D coordinate (``X.coords["time"][d - 1:]``, slice_tools.py:260)"""
        assert len(self.dims) == 1 and isinstance(key, slice)
        return DataArray(self._data[key], self.dims, None, self.attrs)

    # -- what the reference's own TESTS additionally use (run_reference_tests_on_standin.py) --------------------
    ndim = property(lambda self: self._data.ndim)
    sizes = property(lambda self: dict(zip(self.dims, self._data.shape)))

    def __len__(self):
        return self._data.shape[0]

    def __getattr__(self, name):
        coords = self.__dict__.get("coords", {})
        if name in coords:
            c = coords[name]
            return DataArray(c.values, c.dims, {k: v for k, v in coords.items() if set(v.dims) <= set(c.dims)})
        raise AttributeError(name)

    def __eq__(self, other):
        return DataArray(self._data == (other.values if isinstance(other, DataArray) else other), self.dims)

    __hash__ = None

    def all(self):
        return bool(np.all(self._data))

    def astype(self, dtype):
        return DataArray(self._data.astype(dtype), self.dims, None, self.attrs)

    def _reduce(self, fn, dim=None):
        if dim is None:
            return DataArray(fn(self._data), ())
        ax = self.dims.index(dim)
        return DataArray(fn(self._data, axis=ax), tuple(d for d in self.dims if d != dim),
                         {k: c for k, c in self.coords.items() if dim not in c.dims})

    def mean(self, dim=None):
        return self._reduce(np.nanmean if self._data.dtype.kind == "f" else np.mean, dim)

    def std(self, dim=None):
        return self._reduce(np.nanstd if self._data.dtype.kind == "f" else np.std, dim)

    def min(self, dim=None):
        return self._reduce(np.min, dim)

    def max(self, dim=None):
        return self._reduce(np.max, dim)

    def diff(self, dim):
        return DataArray(np.diff(self._data, axis=self.dims.index(dim)), self.dims)

    def to_dataset(self, name):
        return Dataset({name: self})

    def sel(self, **indexers):
        """Label selection: a tuple is ONE label (xarray keeps tuples whole); a label that occurs several times keeps the
        dimension (pandas ``get_loc`` on a non-unique index gives a mask), a unique scalar label drops it; a boolean
        array is a mask."""
        out = self
        for dim, lab in indexers.items():
            ax = out.dims.index(dim)
            labels = out.coords[dim].values
            mask = None
            if isinstance(lab, (DataArray, np.ndarray)) and np.asarray(lab).dtype.kind == "b":
                mask = np.asarray(lab)
            else:
                hits = np.array([x == lab for x in labels.tolist()], dtype=bool)
                if not hits.any():
                    raise KeyError(lab)
                if hits.sum() > 1:
                    mask = hits
                else:                                                # unique scalar label: the dimension goes
                    i = int(np.argmax(hits))
                    co = {k: c for k, c in out.coords.items() if dim not in c.dims}
                    out = DataArray(np.take(out._data, i, axis=ax), tuple(d for d in out.dims if d != dim), co, out.attrs)
                    continue
            co = {k: (DataArray(np.compress(mask, c.values, axis=c.dims.index(dim)), c.dims) if dim in c.dims else c)
                  for k, c in out.coords.items()}
            out = DataArray(np.compress(mask, out._data, axis=ax), out.dims, co, out.attrs)
        return out

    def transpose(self, *dims):
        assert sorted(dims) == sorted(self.dims)
        out = DataArray(np.transpose(self._data, [self.dims.index(d) for d in dims]), dims, None, self.attrs)
        out.coords = self.coords
        return out


class _Resampler:
    def __init__(self, ds, dim, delta):
        self.ds, self.dim, self.delta = ds, dim, delta

    def nearest(self):
        old = pd.DatetimeIndex(self.ds.coords[self.dim].values)
        full = pd.Series(np.arange(len(old)), index=old).resample(self.delta).asfreq().index
        idx = old.get_indexer(full, method="nearest")
        return self.ds._take(self.dim, idx, full.values)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars: dict[str, DataArray] = {}
        self.coords = Coords()
        for k, v in (coords.items() if coords is not None else ()):
            self.coords.set(k, v)
        for k, v in (data_vars or {}).items():
            assert isinstance(v, DataArray)
            for ck, cv in v.coords.items():             # the variables' own coordinates are merged in
                if ck in self.coords:
                    a, b = self.coords[ck].values, cv.values
                    assert a.shape == b.shape and all(x == y for x, y in zip(a.tolist(), b.tolist())), f"conflicting {ck}"
                else:
                    self.coords[ck] = cv
            var = DataArray(v.values, v.dims, None, v.attrs)
            self.data_vars[k] = var
        self.attrs = dict(attrs) if attrs else {}

    # -- access ------------------------------------------------------------------------------------------------
    def __bool__(self):
        return bool(self.data_vars)

    def __getattr__(self, name):
        coords = self.__dict__.get("coords", {})
        if name in coords:
            c = coords[name]
            return DataArray(c.values, c.dims, {name: c} if c.dims == (name,) else None)
        if name in self.__dict__.get("data_vars", {}):
            return self[name]
        raise AttributeError(name)

    @property
    def sizes(self):
        out = {}
        for v in list(self.data_vars.values()) + list(self.coords.values()):
            out.update(zip(v.dims, v.shape))
        return out

    dims = sizes

    def _coords_for(self, dims):
        return {k: c for k, c in self.coords.items() if all(d in dims for d in c.dims)}

    def __getitem__(self, key):
        if isinstance(key, str):
            v = self.data_vars[key]
            return DataArray(v.values, v.dims, self._coords_for(v.dims), v.attrs, name=key)
        out = Dataset(None, None, self.attrs)
        out.coords = Coords(self.coords)
        out.data_vars = {k: self.data_vars[k] for k in key}      # KeyError for an unknown variable, like xarray
        return out

    def _new(self, data_vars, coords, attrs=None):
        out = Dataset(None, None, attrs)
        out.coords = coords
        out.data_vars = data_vars
        return out

    def _take(self, dim, idx, new_labels=None):
        dv = {k: DataArray(np.take(v.values, idx, axis=v.dims.index(dim)), v.dims, None, v.attrs) if dim in v.dims else v
              for k, v in self.data_vars.items()}
        co = Coords()
        for k, c in self.coords.items():
            if dim in c.dims:
                vals = np.take(c.values, idx, axis=c.dims.index(dim)) if (new_labels is None or k != dim) else new_labels
                co[k] = DataArray(vals, c.dims)
            else:
                co[k] = c
        return self._new(dv, co, self.attrs)

    # -- selection ---------------------------------------------------------------------------------------------
    def sel(self, **indexers):
        out = self
        for dim, lab in indexers.items():
            index = pd.Index(out.coords[dim].values)
            if isinstance(lab, slice):
                sl = index.slice_indexer(lab.start, lab.stop)          # both ends inclusive
                idx = np.arange(len(index))[sl]
            else:
                idx = index.get_indexer(list(lab))
                if (idx < 0).any():
                    raise KeyError(f"not all values found in index {dim!r}")
            out = out._take(dim, idx)
        return out

    def resample(self, **kw):
        (dim, delta), = kw.items()
        return _Resampler(self, dim, delta)

    # -- reductions and arithmetic -----------------------------------------------------------------------------
    def _reduce(self, fn, dim):
        dv = {}
        for k, v in self.data_vars.items():
            ax = v.dims.index(dim)
            dv[k] = DataArray(fn(v.values, axis=ax), tuple(d for d in v.dims if d != dim))
        co = Coords({k: c for k, c in self.coords.items() if dim not in c.dims})
        return self._new(dv, co, None)

    def mean(self, dim):
        return self._reduce(np.nanmean, dim)

    def std(self, dim):
        return self._reduce(np.nanstd, dim)

    def _binary(self, other, op):
        assert list(self.data_vars) == list(other.data_vars)
        dv = {}
        for k, a in self.data_vars.items():
            b = other.data_vars[k]
            assert all(d in a.dims for d in b.dims)
            bt = np.transpose(b.values, [b.dims.index(d) for d in a.dims if d in b.dims])
            shape = [a.shape[i] if d in b.dims else 1 for i, d in enumerate(a.dims)]
            dv[k] = DataArray(op(a.values, bt.reshape(shape)), a.dims)
        return self._new(dv, Coords(self.coords), None)

    def __sub__(self, other):
        return self._binary(other, np.subtract)

    def __truediv__(self, other):
        with np.errstate(divide="ignore", invalid="ignore"):
            return self._binary(other, np.true_divide)

    # -- reshaping ---------------------------------------------------------------------------------------------
    def stack(self, **kw):
        (new, dims), = kw.items()
        dims = list(dims)
        dv = {}
        for k, v in self.data_vars.items():
            others = [d for d in v.dims if d not in dims]
            t = np.transpose(v.values, [v.dims.index(d) for d in others + dims])
            dv[k] = DataArray(t.reshape(t.shape[: len(others)] + (-1,)), (*others, new), None, v.attrs)
        labels = [self.coords[d].values for d in dims]
        grids = np.meshgrid(*labels, indexing="ij")
        tuples = np.empty(grids[0].size, dtype=object)
        tuples[:] = list(zip(*(g.reshape(-1).tolist() for g in grids)))
        co = Coords({k: c for k, c in self.coords.items() if k not in dims})
        co[new] = DataArray(tuples, (new,))
        for d, g in zip(dims, grids):                      # the levels of the stacked index stay available on ``new``
            co[d] = DataArray(g.reshape(-1), (new,))
        return self._new(dv, co, self.attrs)

    def assign_coords(self, **kw):
        out = self._new(dict(self.data_vars), Coords(self.coords), self.attrs)
        for k, v in kw.items():
            out.coords.set(k, v)
        return out


def concat(objs, dim):
    first = objs[0]
    ax = first.dims.index(dim)
    out = DataArray(np.concatenate([o.values for o in objs], axis=ax), first.dims, None, first.attrs, first.name)
    for k, c in first.coords.items():
        if dim in c.dims:
            out.coords[k] = DataArray(np.concatenate([o.coords[k].values for o in objs], axis=c.dims.index(dim)), c.dims)
        else:
            out.coords[k] = c
    return out


Coordinates = Coords        # only named in the reference's type annotations (era5_svd.py:270)


# -- file round trip of the stand-in (tests of the product's xarray BRANCHES only: dataset.write_netcdf / read_netcdf take
#    them when xarray + netCDF4 are importable; the container format here is a pickle, the call signatures are xarray's) --
def _to_netcdf(self, path, format=None, **kw):
    import pickle

    assert format in (None, "NETCDF4", "NETCDF4_CLASSIC", "NETCDF3_64BIT", "NETCDF3_CLASSIC"), format
    with open(path, "wb") as f:
        pickle.dump({"format": format,
                     "data_vars": {k: (v.dims, v.values, v.attrs) for k, v in self.data_vars.items()},
                     "coords": {k: (c.dims, c.values) for k, c in self.coords.items()}, "attrs": self.attrs}, f)


Dataset.to_netcdf = _to_netcdf


def _as_netcdf4_returns(v):
    """How an attribute comes back through netCDF4 / xarray: a one-element sequence as a SCALAR (str or NumPy scalar), a
    longer list of strings as a list, a longer numeric sequence as an ndarray, Python numbers as NumPy scalars - what the
    reference's readers cope with (``str_to_list`` / ``int_to_list`` / ``.tolist()``, era5_svd.py:86-99, :175-183)."""
    if isinstance(v, (list, tuple)):
        if len(v) == 1:
            return _as_netcdf4_returns(v[0])
        return list(v) if all(isinstance(x, str) for x in v) else np.asarray(v)
    if isinstance(v, bool):
        return np.int64(int(v))
    if isinstance(v, int):
        return np.int64(v)
    if isinstance(v, float):
        return np.float64(v)
    return v


def open_dataset(path, **kw):
    import pickle

    with open(path, "rb") as f:
        d = pickle.load(f)
    dec = lambda a: {k: _as_netcdf4_returns(v) for k, v in a.items()}      # noqa: E731
    ds = Dataset({k: DataArray(v, dims, None, dec(attrs)) for k, (dims, v, attrs) in d["data_vars"].items()},
                 {k: (dims, v) for k, (dims, v) in d["coords"].items()}, dec(d["attrs"]))
    ds.file_format = d["format"]
    return ds
