"""Reference-named matrix-build helpers (src/dmd_era5/slice_tools/slice_tools.py) on the stage's own
containers (``dataset.Dataset`` / ``DataArray``).

``era5_svd.main`` does NOT go through these one by one: it runs the fused device pipeline
(``pipeline.build_matrix_device``: stack + centre + scale + transpose [+ tf32 split] in one kernel
sequence, delay embedding kept virtual).  The functions here keep the reference's call signatures and
error texts for callers that use them individually; selection / resampling / flattening / delay
embedding are pure index work (host), ``standardize_data`` runs the CUDA build kernel.

Deviation (documented in DESIGN.md): the reference's ``space`` coordinate is an object array of
(level, latitude, longitude) tuples (slice_tools.py:346) - O(m) Python objects, unusable at 40 M rows.
Here ``space`` is ``arange(m)`` from the start and the three per-row coordinates ``level``,
``latitude``, ``longitude`` are carried in closed form - exactly what the reference's
``space_coord_to_level_lat_lon`` (slice_tools.py:368-414) produces at the end of ``main``.
"""
from __future__ import annotations

import logging
from datetime import datetime, timedelta

import numpy as np

from .dataset import DataArray, Dataset

SPATIAL_ORDER = ("level", "latitude", "longitude")
logger = logging.getLogger("ERA5Processing")            # the reference's logger name (slice_tools.py:12)


def log_and_print(lg: logging.Logger, msg: str, level: str = "info") -> None:
    """src/dmd_era5/logger.py:42-46 (kept local: importing era5_svd here would load torch for pure index work)."""
    getattr(lg, level.lower())(msg)
    print(msg)


def _get_dataset_time_bounds(ds: Dataset) -> dict:
    """slice_tools.py:106-123 (quirk Q1: local-time conversion via datetime.fromtimestamp)."""
    t = ds.coord("time").astype("datetime64[ns]").astype(np.int64)
    return {"first": datetime.fromtimestamp(t[0] * 1e-9), "last": datetime.fromtimestamp(t[-1] * 1e-9)}


def _subset(ds: Dataset, dim: str, index: np.ndarray, new_coord: np.ndarray | None = None) -> Dataset:
    dv = {}
    for name, da in ds.data_vars.items():
        ax = da.dims.index(dim)
        # pending selection (dataset.LazyTake): carried out when .values is read, or fused into the staging copy
        dv[name] = DataArray(da.lazy().take(index, ax), da.dims, attrs=da.attrs)
    coords = dict(ds.coords)
    coords[dim] = ((dim,), ds.coord(dim)[index] if new_coord is None else new_coord)
    return Dataset(dv, coords, ds.attrs)


def slice_era5_dataset(ds: Dataset, start_datetime=None, end_datetime=None, levels: list | None = None) -> Dataset:
    """slice_tools.py:20-103: time range (inclusive labels) and pressure levels IN THE ORDER GIVEN."""
    start_dt = datetime.fromisoformat(start_datetime) if isinstance(start_datetime, str) else start_datetime
    end_dt = datetime.fromisoformat(end_datetime) if isinstance(end_datetime, str) else end_datetime
    bounds = _get_dataset_time_bounds(ds)
    start_dt = start_dt or bounds["first"]
    end_dt = end_dt or bounds["last"]
    if start_dt < bounds["first"] or end_dt > bounds["last"]:
        msg = f"Time range ({start_dt} to {end_dt}) is outside dataset"
        msg += f"bounds ({bounds['first']} to {bounds['last']})."
        log_and_print(logger, msg, "error")
        raise ValueError(msg)
    if start_dt >= end_dt:
        msg = "Start datetime must be before end datetime."
        log_and_print(logger, msg, "error")
        raise ValueError(msg)
    all_levels = list(ds.coord("level"))
    levels = levels or all_levels
    times = ds.coord("time").astype("datetime64[ns]")
    keep = np.nonzero((times >= np.datetime64(start_dt, "ns")) & (times <= np.datetime64(end_dt, "ns")))[0]
    out = _subset(ds, "time", keep)
    try:
        idx = np.array([all_levels.index(lv) for lv in levels], dtype=np.int64)
    except ValueError as e:
        msg = "Requested level is not available in the dataset."
        msg += f"Available levels: {all_levels}"
        log_and_print(logger, msg, "error")
        raise ValueError(msg) from e
    log_and_print(logger, f"Dataset slicing completed successfully using {start_dt}to {end_dt} and levels {levels}")
    return _subset(out, "level", idx)


def resample_nearest_index(times_ns: np.ndarray, delta_ns: int) -> tuple[np.ndarray, np.ndarray]:
    """Index form of ``ds.resample(time=delta).nearest()`` (slice_tools.py:139): labels on a grid of
    width delta anchored at midnight of the first day, each taking the nearest original sample."""
    times_ns = np.asarray(times_ns, dtype=np.int64)
    day = 86400 * 10**9
    origin = (times_ns[0] // day) * day
    first = origin + ((times_ns[0] - origin) // delta_ns) * delta_ns
    last = origin + ((times_ns[-1] - origin) // delta_ns) * delta_ns
    labels = np.arange(first, last + 1, delta_ns, dtype=np.int64)
    pos = np.searchsorted(times_ns, labels)
    lo = np.clip(pos - 1, 0, len(times_ns) - 1)
    hi = np.clip(pos, 0, len(times_ns) - 1)
    idx = np.where(np.abs(times_ns[hi] - labels) <= np.abs(labels - times_ns[lo]), hi, lo)
    return labels, idx


def resample_era5_dataset(ds: Dataset, delta_time: timedelta) -> Dataset:
    """slice_tools.py:126-141."""
    t = ds.coord("time").astype("datetime64[ns]").astype(np.int64)
    labels, idx = resample_nearest_index(t, int(delta_time.total_seconds()) * 10**9)
    out = _subset(ds, "time", idx, labels.astype("datetime64[ns]"))
    log_and_print(logger, f"Resampled the dataset with time delta: {delta_time}")
    return out


def log_standardize(dim: str, scale: bool) -> None:
    """The reference's progress lines of standardize_data (slice_tools.py:164-176); also emitted by the fused device
    build of the stage, which replaces that call (stage._compute)."""
    log_and_print(logger, f"Standardizing data along {dim} dimension...")
    log_and_print(logger, f"Removing mean along {dim} dimension...")
    if scale:
        log_and_print(logger, f"Scaling to unit variance along {dim} dimension...")


def standardize_data(data: Dataset, dim: str = "time", scale: bool = True):
    """slice_tools.py:144-179 on the GPU: mean over ``dim`` (NaN skipping), centre, std (ddof = 0) of the
    centred data, divide.  Returns (data, mean, std | None) like the reference."""
    import torch

    from ._cabi import BUILD_MEAN_CENTER, BUILD_SCALE
    from .era5_svd import get_ops

    ops = get_ops()
    log_standardize(dim, scale)
    flags = BUILD_MEAN_CENTER | (BUILD_SCALE if scale else 0)
    out, means, stds = {}, {}, {}
    with torch.cuda.device(ops.device):
        for name, da in data.data_vars.items():
            ax = da.dims.index(dim)
            a = np.moveaxis(np.asarray(da.values), ax, 0)
            rest_dims = tuple(d for d in da.dims if d != dim)
            T, rest = a.shape[0], a.shape[1:]
            src = torch.from_numpy(np.ascontiguousarray(a.reshape(T, -1), dtype=a.dtype.newbyteorder("="))).to(ops.device)
            P = src.shape[1]
            X = ops.empty((P, T), src.dtype)
            mu = ops.empty((P,), src.dtype)
            sd = ops.empty((P,), src.dtype) if scale else None
            ops.build_rows(src, X, mu, sd, None, flags)
            arr = np.moveaxis(X.t().contiguous().cpu().numpy().reshape((T,) + rest), 0, ax)
            out[name] = DataArray(arr, da.dims, attrs=da.attrs)
            means[name] = DataArray(mu.cpu().numpy().reshape(rest), rest_dims)
            if scale:
                stds[name] = DataArray(sd.cpu().numpy().reshape(rest), rest_dims)
    rest_coords = {k: v for k, v in data.coords.items() if k != dim}
    ds_out = Dataset(out, data.coords, data.attrs)
    return ds_out, Dataset(means, rest_coords), (Dataset(stds, rest_coords) if scale else None)


def _apply_delay_embedding_np(X: np.ndarray, d: int) -> np.ndarray:
    """slice_tools.py:182-211 (same result as the reference's sliding_window_view expression)."""
    if X.ndim != 2:
        raise ValueError("Input array must be 2D.")
    if not isinstance(d, int) or isinstance(d, bool) or d <= 0:
        raise ValueError("Delay must be an integer greater than 0.")
    n = X.shape[1] - d + 1
    if n < 1:       # the reference's sliding_window_view raises here (numpy's text), slice_tools.py:207
        raise ValueError("window shape cannot be larger than input array shape")
    return np.concatenate([X[:, j : j + n] for j in range(d)], axis=0)


def space_coords(levels, lats, lons, n_vars: int, d: int = 1):
    """Closed form of the tiled (level, latitude, longitude) row labels (slice_tools.py:346, :259)."""
    levels, lats, lons = np.asarray(levels), np.asarray(lats), np.asarray(lons)
    L, A, O = len(levels), len(lats), len(lons)
    reps = n_vars * d
    return (np.tile(np.repeat(levels, A * O), reps), np.tile(np.tile(np.repeat(lats, O), L), reps),
            np.tile(np.tile(lons, L * A), reps))


def flatten_era5_variables(era5_ds: Dataset) -> DataArray:
    """slice_tools.py:277-365: (time, level, lat, lon) per variable -> (space, time), variables stacked
    along space in dataset order; fields without a time axis flatten to 1-D."""
    variables = list(era5_ds.data_vars.keys())
    coords = sorted(era5_ds.coords.keys())
    space3 = sorted(["latitude", "longitude", "level"])
    if coords != sorted(space3 + ["time"]) and coords != space3:
        msg = """
        Input dataset must have coordinates ('latitude', 'longitude', 'level')
        or ('latitude', 'longitude', 'level', 'time').
        """
        raise ValueError(msg)
    has_time = "time" in era5_ds.coords
    parts = []
    for v in variables:
        da = era5_ds[v]
        order = (["time"] if has_time else []) + list(SPATIAL_ORDER)
        a = np.transpose(np.asarray(da.values), [da.dims.index(dm) for dm in order])
        parts.append(a.reshape(a.shape[0], -1).T if has_time else a.reshape(-1))
    data = np.concatenate(parts, axis=0)
    lev, lat, lon = space_coords(era5_ds.coord("level"), era5_ds.coord("latitude"), era5_ds.coord("longitude"), len(variables))
    S = lev.shape[0] // len(variables)
    co = {"space": (("space",), np.arange(data.shape[0])), "original_variable": (("space",), np.repeat(variables, S)),
          "level": (("space",), lev), "latitude": (("space",), lat), "longitude": (("space",), lon)}
    if has_time:
        co["time"] = (("time",), era5_ds.coord("time"))
    attrs = dict(era5_ds.attrs)
    attrs["original_variables"] = variables
    attrs["space_coords"] = list(SPATIAL_ORDER)
    return DataArray(data, ("space", "time") if has_time else ("space",), co, attrs)


def apply_delay_embedding(X: DataArray, d: int) -> DataArray:
    """slice_tools.py:214-274: (m, T) -> (m d, T - d + 1); block j carries delay d - 1 - j (:265-268)."""
    if not isinstance(X, DataArray):
        raise ValueError("Input data must be a xr.DataArray")
    if sorted(X.dims) != ["space", "time"]:
        raise ValueError("Input data must have dimensions ('space', 'time').")
    for need in ("original_variable", "space", "time"):
        if need not in X.coords:
            raise ValueError("Input data must have coordinates ('space', 'time', 'original_variable').")
    result = _apply_delay_embedding_np(np.asarray(X.values), d)
    m0 = X.shape[0]
    co = {"space": (("space",), np.arange(m0 * d)), "time": (("time",), X.coord("time")[d - 1:]),
          "original_variable": (("space",), np.tile(X.coord("original_variable"), d)),
          "delay": (("space",), np.repeat(np.flip(np.arange(d)), m0))}
    for k in ("level", "latitude", "longitude"):
        if k in X.coords:
            co[k] = (("space",), np.tile(X.coord(k), d))
    attrs = dict(X.attrs)
    attrs["delay_embedding"] = d
    return DataArray(result, ("space", "time"), co, attrs)


def space_coord_to_level_lat_lon(ds: Dataset) -> Dataset:
    """slice_tools.py:368-414.  The per-row level / latitude / longitude coordinates already exist in
    closed form here; only ``space`` is (re)set to arange, as the reference does."""
    if "space" not in ds.coords:
        raise ValueError("Input dataset must have a 'space' coordinate.")
    m = ds.coords["space"][1].shape[0]
    ds.coords["space"] = (("space",), np.arange(m))
    return ds
