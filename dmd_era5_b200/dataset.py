"""Minimal labelled-array containers + NetCDF I/O for the SVD stage's inputs and outputs.

The reference packages everything in xarray objects (src/dmd_era5/era5_svd/era5_svd.py:266-333) and
writes ``to_netcdf(..., format="NETCDF4")`` (:434).  xarray / netCDF4 / h5py are not installable in
this image, so the stage carries its own small containers with the same vocabulary
(dims, coords, attrs, data_vars) and converts to real xarray objects when xarray is importable
(``Dataset.to_xarray()``, ``write_netcdf`` then uses xarray's NETCDF4 writer = the reference's file
layout).  Without xarray, files are written / read as NetCDF-3 64-bit-offset through
``scipy.io.netcdf_file`` with the same variable / dimension / attribute schema; NetCDF-3 cannot hold
int64, string arrays or list-of-string attributes, so: int64 -> int32, str coords -> char arrays,
list[str] attrs -> one comma-joined string (split again on read), datetime64 -> CF "seconds since".
"""
from __future__ import annotations

import numpy as np


class LazyTake:
    """A pending ``np.take`` along one or more axes of a host array.  ``slice_era5_dataset`` / ``resample_era5_dataset``
    (index selections along time and level, slice_tools.py:20-141) return these instead of copies: the SVD stage
    gathers the selected planes straight into its pinned staging buffers (stage.stage_blocks), so a slice is read once
    on its way to the device instead of being copied once per selection step.  Anything else that touches
    ``DataArray.values`` gets the materialised ndarray (same values as the eager ``np.take`` chain)."""

    def __init__(self, base: np.ndarray, index: dict | None = None):
        self.base = base
        self.index = dict(index or {})          # axis -> 1-D integer index array (absent = the whole axis)

    @property
    def shape(self):
        return tuple(len(self.index[ax]) if ax in self.index else n for ax, n in enumerate(self.base.shape))

    @property
    def ndim(self):
        return self.base.ndim

    @property
    def dtype(self):
        return self.base.dtype

    def take(self, idx, axis: int) -> "LazyTake":
        idx = np.asarray(idx, dtype=np.int64)
        n = self.shape[axis]
        if idx.size and (idx.min() < -n or idx.max() >= n):
            raise IndexError(f"index out of bounds for axis {axis} with size {n}")
        idx = np.where(idx < 0, idx + n, idx)
        new = dict(self.index)
        new[axis] = new[axis][idx] if axis in new else idx
        if len(new[axis]) == self.base.shape[axis] and np.array_equal(new[axis], np.arange(self.base.shape[axis])):
            del new[axis]                       # identity selection: nothing pending
        return LazyTake(self.base, new)

    def materialise(self) -> np.ndarray:
        a = self.base
        for ax in sorted(self.index):
            a = np.take(a, self.index[ax], axis=ax)
        return a

    def __array__(self, dtype=None, copy=None):
        a = self.materialise()
        return a if dtype is None else a.astype(dtype, copy=False)


class DataArray:
    def __init__(self, values, dims, coords: dict | None = None, attrs: dict | None = None, name: str | None = None):
        self._values = values if hasattr(values, "shape") else np.asarray(values)
        self.dims = tuple(dims)
        if len(self.dims) != self._values.ndim:
            raise ValueError(f"dims {self.dims} do not match a {self._values.ndim}-D array")
        # coords: name -> (dims tuple, 1-D ndarray)
        self.coords: dict[str, tuple[tuple[str, ...], np.ndarray]] = {}
        for k, v in (coords or {}).items():
            self.coords[k] = (tuple(v[0]) if isinstance(v[0], (tuple, list)) else (v[0],), np.asarray(v[1])) \
                if isinstance(v, tuple) else ((k,), np.asarray(v))
        self.attrs = dict(attrs or {})
        self.name = name

    @property
    def values(self):
        """The data as an ndarray (a pending selection is carried out on first access and kept)."""
        if isinstance(self._values, LazyTake):
            self._values = self._values.materialise()
        return self._values

    @values.setter
    def values(self, v):
        self._values = v if hasattr(v, "shape") else np.asarray(v)

    def lazy(self) -> LazyTake:
        """The data as a pending selection on its host array, without materialising it."""
        v = self._values
        if isinstance(v, LazyTake):
            return v
        return LazyTake(v if isinstance(v, np.ndarray) else np.asarray(v))

    @property
    def shape(self):
        return tuple(self._values.shape)

    @property
    def sizes(self):
        return dict(zip(self.dims, self.shape))

    def coord(self, name: str) -> np.ndarray:
        return self.coords[name][1]


class Dataset:
    def __init__(self, data_vars: dict | None = None, coords: dict | None = None, attrs: dict | None = None):
        self.data_vars: dict[str, DataArray] = dict(data_vars or {})
        self.coords: dict[str, tuple[tuple[str, ...], np.ndarray]] = {}
        for k, v in (coords or {}).items():
            self.coords[k] = (tuple(v[0]) if isinstance(v[0], (tuple, list)) else (v[0],), np.asarray(v[1])) \
                if isinstance(v, tuple) else ((k,), np.asarray(v))
        for da in self.data_vars.values():
            for k, v in da.coords.items():
                self.coords.setdefault(k, v)
        self.attrs = dict(attrs or {})

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self.data_vars[k] for k in key}, self.coords, self.attrs)
        return self.data_vars[key]

    def __contains__(self, key):
        return key in self.data_vars

    @property
    def sizes(self):
        out = {}
        for da in self.data_vars.values():
            out.update(da.sizes)
        for k, (dims, v) in self.coords.items():
            if dims == (k,):
                out.setdefault(k, v.shape[0])
        return out

    def coord(self, name: str) -> np.ndarray:
        return self.coords[name][1]

    def to_xarray(self):
        import xarray as xr  # only when available

        dv = {k: xr.DataArray(np.asarray(v.values), dims=v.dims, attrs=v.attrs) for k, v in self.data_vars.items()}
        return xr.Dataset(dv, coords={k: (d, v) for k, (d, v) in self.coords.items()}, attrs=self.attrs)


# ------------------------------------------------------------------------------------------------
# NetCDF I/O
# ------------------------------------------------------------------------------------------------
_EPOCH = np.datetime64("1970-01-01T00:00:00", "s")
LAZY_READ_BYTES = 32 << 20      # read_netcdf: numeric variables at least this large stay memory-mapped (NetCDF-3 path)


def _have_xarray() -> bool:
    try:
        import xarray  # noqa: F401
        import netCDF4  # noqa: F401
        return True
    except Exception:
        return False


def _encode_attr(v):
    if isinstance(v, bool):
        return int(v)
    if isinstance(v, (list, tuple)):
        if all(isinstance(x, str) for x in v):
            return ",".join(v)
        return np.asarray(v, dtype=np.int32 if all(isinstance(x, (int, np.integer)) for x in v) else np.float64)
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def _netcdf3(path: str, mode: str, **kw):
    """scipy's netcdf_file mirrors every global attribute onto the instance, so an attribute named
    "variables" (the reference's schema has one) would shadow ``netcdf_file.variables``; keep such
    names in the attribute table only."""
    from scipy.io import netcdf_file

    class _File(netcdf_file):
        def __setattr__(self, attr, value):
            if attr in ("variables", "dimensions") and not isinstance(value, dict) and "_attributes" in self.__dict__:
                self._attributes[attr] = value
                return
            super().__setattr__(attr, value)

    return _File(path, mode, **kw)


NETCDF3_VAR_LIMIT = 2 ** 31 - 4      # largest variable the NetCDF-3 (scipy) writer takes


def _flatten_for_classic(ds: Dataset, allow_int64: bool):
    """Dataset -> (dims, [(name, dims, array, attrs)], global attrs) in the classic data model: datetimes as seconds
    since the epoch (CF units), strings as fixed-width char arrays, bool as int8; int64 kept only when the target
    format has it (CDF-5), else int32 like the NetCDF-3 path."""
    dims = {name: int(size) for name, size in ds.sizes.items()}
    out = []

    def put(name, vdims, arr, attrs=None):
        arr = arr if isinstance(arr, np.memmap) else np.asarray(arr)
        extra = {}
        if np.issubdtype(arr.dtype, np.datetime64):
            arr = (arr.astype("datetime64[s]") - _EPOCH).astype(np.float64)
            extra = {"units": "seconds since 1970-01-01 00:00:00", "calendar": "proleptic_gregorian"}
        if arr.dtype.kind in ("U", "O", "S"):
            strs = np.asarray([str(x) for x in arr.ravel()], dtype="S")
            n = strs.dtype.itemsize
            dname = f"string{n}"
            dims.setdefault(dname, n)
            arr = strs.view("S1").reshape(arr.shape + (n,))
            vdims = tuple(vdims) + (dname,)
        elif arr.dtype == np.int64 and not allow_int64:
            arr = arr.astype(np.int32)
        elif arr.dtype == np.bool_:
            arr = arr.astype(np.int8)
        out.append((name, tuple(vdims), arr, {k: _encode_attr(v) for k, v in {**(attrs or {}), **extra}.items()}))

    for name, (vdims, arr) in ds.coords.items():
        put(name, vdims, arr)
    for name, da in ds.data_vars.items():
        put(name, da.dims, da.values, da.attrs)
    gattrs = {k: _encode_attr(v) for k, v in ds.attrs.items()}
    gattrs["coordinates_hint"] = " ".join(ds.coords)      # which variables are coordinates (for read_netcdf)
    return dims, out, gattrs


def write_netcdf(ds: Dataset, path: str, format: str | None = None) -> str:
    """Write ``ds``; returns the format used:
      "NETCDF4"            through xarray + netCDF4 when importable - the reference's call (era5_svd.py:434);
      "NETCDF3_64BIT"      scipy's NetCDF-3 writer (CDF-2) while every variable fits its 2 GiB limit;
      "NETCDF3_64BIT_DATA" the CDF-5 writer of cdf5.py otherwise (no size limit: c2's X, c3's U), readable by
                           netCDF-C >= 4.4 / xarray's netcdf4 engine / ncdump.
    ``format`` forces one of the three (tests)."""
    import os

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if format not in (None, "NETCDF4", "NETCDF3_64BIT", "NETCDF3_64BIT_DATA"):
        raise ValueError(f"format {format} is not supported.")
    if format == "NETCDF4" or (format is None and _have_xarray()):
        ds.to_xarray().to_netcdf(path, format="NETCDF4")     # the reference's call, era5_svd.py:434
        return "NETCDF4"
    biggest = max([int(np.asarray(a).nbytes) if not isinstance(a, np.memmap) else int(a.nbytes)
                   for _, a in ds.coords.values()] + [int(da.values.nbytes) for da in ds.data_vars.values()] + [0])
    if format == "NETCDF3_64BIT_DATA" or (format is None and biggest > NETCDF3_VAR_LIMIT):
        from .cdf5 import write_classic

        dims, variables, gattrs = _flatten_for_classic(ds, allow_int64=True)
        write_classic(path, dims, variables, gattrs, version=5)
        return "NETCDF3_64BIT_DATA"
    if biggest > NETCDF3_VAR_LIMIT:
        raise ValueError(f"a variable of {biggest / 2 ** 30:.1f} GiB exceeds the 2 GiB per-variable limit of NetCDF-3; "
                         "use format=None / 'NETCDF3_64BIT_DATA' (CDF-5) or install xarray + netCDF4")
    dims, variables, gattrs = _flatten_for_classic(ds, allow_int64=False)
    with _netcdf3(path, "w", version=2) as f:
        for name, size in dims.items():
            f.createDimension(name, int(size))
        for name, vdims, arr, attrs in variables:
            var = f.createVariable(name, "c" if arr.dtype.kind == "S" else arr.dtype, vdims)
            var[:] = arr
            for k, v in attrs.items():
                var._attributes[k] = v
        # straight into the attribute table: an attribute called "variables" (the reference's schema has
        # one) must not shadow netcdf_file.variables
        for k, v in gattrs.items():
            f._attributes[k] = v
    return "NETCDF3_64BIT"


def _decode_attr(k, v):
    if isinstance(v, bytes):
        v = v.decode()
    if isinstance(v, np.ndarray) and v.ndim == 0:
        v = v.item()
    return v


def read_netcdf(path: str, lazy: bool = False) -> Dataset:
    """Read a file written by ``write_netcdf`` (or any NetCDF the available backend can open).
    lazy (NetCDF-3 path): large numeric variables come back memory-mapped (read-only, on-disk byte order) instead of
    being read; for INPUT files only - the mapping must not outlive a rewrite of the file."""
    if _have_xarray():
        import xarray as xr

        x = xr.open_dataset(path)
        dv = {k: DataArray(v.values, v.dims, attrs=dict(v.attrs)) for k, v in x.data_vars.items()}
        co = {k: (tuple(v.dims), v.values) for k, v in x.coords.items()}
        return Dataset(dv, co, dict(x.attrs))
    with open(path, "rb") as fh:
        magic = fh.read(4)
    if magic == b"CDF\x05":
        return _read_cdf5(path, lazy)
    # NetCDF-3 through scipy.  Large numeric variables (the slice itself: GBs) are NOT read here: they come back as
    # read-only np.memmap views of the file in its on-disk (big-endian) byte order, so that the only pass over the
    # data is the staging copy into pinned memory (stage.stage_blocks), which converts the byte order on the way.
    # Everything small (coordinates, attributes, results) is decoded eagerly to native arrays as before.
    import warnings

    big = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        f = _netcdf3(path, "r", mmap=True)
    try:
        attrs = {k: _decode_attr(k, v) for k, v in f._attributes.items()}
        coord_names = set(str(attrs.pop("coordinates_hint", "")).split())
        dv, co = {}, {}
        base_addr = f._mm_buf.__array_interface__["data"][0] if getattr(f, "_mm_buf", None) is not None else None
        for name in list(f.variables):
            var = f.variables[name]
            dims = tuple(var.dimensions)
            vattrs = {k: _decode_attr(k, v) for k, v in var._attributes.items()}
            data = var.data
            lazy_ok = (lazy and base_addr is not None and not var.isrec and data.dtype.kind in "fiu" and
                       data.nbytes >= LAZY_READ_BYTES and not str(vattrs.get("units", "")).startswith("seconds since"))
            if lazy_ok:
                big[name] = (data.__array_interface__["data"][0] - base_addr, data.dtype, tuple(data.shape), dims, vattrs)
                del data, var
                continue
            arr = np.array(data)
            del data, var
            if arr.dtype.byteorder == ">":                      # NetCDF-3 is big-endian on disk
                arr = arr.astype(arr.dtype.newbyteorder("="))
            if arr.dtype.kind == "S" and dims and dims[-1].startswith("string"):
                arr = np.array([b"".join(row).decode().rstrip("\x00") for row in arr.reshape(-1, arr.shape[-1])],
                               dtype=object).reshape(arr.shape[:-1])
                arr = arr.astype(str)
                dims = dims[:-1]
            units = str(vattrs.get("units", ""))
            if units.startswith("seconds since 1970-01-01"):
                arr = (_EPOCH + arr.astype(np.int64).astype("timedelta64[s]")).astype("datetime64[ns]")
                vattrs = {k: v for k, v in vattrs.items() if k not in ("units", "calendar")}
            if name in coord_names or dims == (name,):
                co[name] = (dims, arr)
            else:
                dv[name] = DataArray(arr, dims, attrs=vattrs)
    finally:
        f.variables = {}                                        # no views of the mapping are left: close() is clean
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            f.close()
    for name, (offset, dt, shape, dims, vattrs) in big.items():
        arr = np.memmap(path, dtype=dt, mode="r", offset=offset, shape=shape)
        if name in coord_names or dims == (name,):
            co[name] = (dims, np.asarray(arr).astype(dt.newbyteorder("=")))
        else:
            dv[name] = DataArray(arr, dims, attrs=vattrs)
    return Dataset(dv, co, attrs)


def _read_cdf5(path: str, lazy: bool) -> Dataset:
    """A CDF-5 file (cdf5.py): small variables decoded eagerly; with ``lazy`` numeric variables of at least
    LAZY_READ_BYTES stay memory-mapped in their on-disk (big-endian) byte order, exactly like the NetCDF-3 path."""
    from .cdf5 import read_classic

    dims, variables, gattrs = read_classic(path)
    attrs = dict(gattrs)
    coord_names = set(str(attrs.pop("coordinates_hint", "")).split())
    dv, co = {}, {}
    for name, (vdims, data, vattrs) in variables.items():
        vattrs = dict(vattrs)
        units = str(vattrs.get("units", ""))
        is_coord = name in coord_names or vdims == (name,)
        if (lazy and not is_coord and data.dtype.kind in "fiu" and data.nbytes >= LAZY_READ_BYTES
                and not units.startswith("seconds since")):
            dv[name] = DataArray(data, vdims, attrs=vattrs)
            continue
        arr = np.array(data)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        if arr.dtype.kind == "S" and vdims and vdims[-1].startswith("string"):
            arr = np.array([b"".join(row).decode().rstrip("\x00") for row in arr.reshape(-1, arr.shape[-1])],
                           dtype=object).reshape(arr.shape[:-1]).astype(str)
            vdims = vdims[:-1]
        if units.startswith("seconds since 1970-01-01"):
            arr = (_EPOCH + arr.astype(np.int64).astype("timedelta64[s]")).astype("datetime64[ns]")
            vattrs = {k: v for k, v in vattrs.items() if k not in ("units", "calendar")}
        if is_coord:
            co[name] = (vdims, arr)
        else:
            dv[name] = DataArray(arr, vdims, attrs=vattrs)
    return Dataset(dv, co, attrs)
