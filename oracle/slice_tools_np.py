"""NumPy restatement of the reference's matrix-build semantics (test oracle only).

Every function cites the reference file:line it follows.  Paths are relative to the
reference repository root (ClimeTrend/DMD-ERA5).  A "dataset" here is a plain dict
``{"vars": {name: ndarray (T, L, A, O)}, "time": (T,), "level": (L,),
"latitude": (A,), "longitude": (O,)}`` because xarray is not installable in this image.
"""
from __future__ import annotations

import numpy as np


def standardize_np(arr: np.ndarray, scale: bool = True, axis: int = 0):
    """src/dmd_era5/slice_tools/slice_tools.py:144-179 (standardize_data).

    mean = data.mean(dim) (NaN skipping, xarray default); data = data - mean;
    std = data.std(dim) of the *centred* data, ddof=0; data = data / std.
    No epsilon guard: std == 0 gives inf/nan exactly like the reference.
    Arithmetic stays in the input dtype (xarray/numpy semantics).
    """
    mean = np.nanmean(arr, axis=axis, keepdims=True).astype(arr.dtype)
    out = arr - mean
    if not scale:
        return out, np.squeeze(mean, axis=axis), None
    std = np.nanstd(out, axis=axis, keepdims=True).astype(arr.dtype)  # ddof=0
    with np.errstate(divide="ignore", invalid="ignore"):
        out = out / std
    return out, np.squeeze(mean, axis=axis), np.squeeze(std, axis=axis)


def flatten_np(var_arrays: list[np.ndarray]) -> np.ndarray:
    """src/dmd_era5/slice_tools/slice_tools.py:323-336 (flatten_era5_variables).

    stack(space=[level, latitude, longitude]) then transpose to (space, time) per
    variable, concatenated along space in dataset-variable order:
    row r = v*S + (l*A + a)*O + o,  X[r, t] = var_v[t, l, a, o].
    Fields without a time axis (L, A, O) flatten to 1-D (:330-332).
    """
    out = []
    for a in var_arrays:
        if a.ndim == 4:
            T = a.shape[0]
            out.append(a.reshape(T, -1).T)
        elif a.ndim == 3:
            out.append(a.reshape(-1))
        else:
            raise ValueError("variable arrays must be (T,L,A,O) or (L,A,O)")
    return np.ascontiguousarray(np.concatenate(out, axis=0))


def delay_embed_np(X: np.ndarray, d: int) -> np.ndarray:
    """src/dmd_era5/slice_tools/slice_tools.py:182-211 (_apply_delay_embedding_np).

    Block j in [0, d) holds X[:, j : j + n], n = T - d + 1, stacked along space.
    Same error strings as the reference (:199-205).
    """
    if X.ndim != 2:
        raise ValueError("Input array must be 2D.")
    if not isinstance(d, int) or isinstance(d, bool) or d <= 0:
        raise ValueError("Delay must be an integer greater than 0.")
    n = X.shape[1] - d + 1
    return np.concatenate([X[:, j : j + n] for j in range(d)], axis=0)


def delay_coord_np(m0: int, d: int) -> np.ndarray:
    """slice_tools.py:265-268: delay = repeat(flip(arange(d)), m0) -> block j has d-1-j."""
    return np.repeat(np.flip(np.arange(d)), m0)


def space_coords_np(levels, lats, lons, n_vars: int, d: int = 1):
    """Closed form of slice_tools.py:346 (tile of stacked (level,lat,lon) tuples),
    :259 (tiling by d) and :402-414 (space_coord_to_level_lat_lon)."""
    levels = np.asarray(levels); lats = np.asarray(lats); lons = np.asarray(lons)
    L, A, O = len(levels), len(lats), len(lons)
    lev = np.repeat(levels, A * O)
    lat = np.tile(np.repeat(lats, O), L)
    lon = np.tile(lons, L * A)
    reps = n_vars * d
    return np.tile(lev, reps), np.tile(lat, reps), np.tile(lon, reps)


def original_variable_np(variables: list[str], S: int, d: int = 1) -> np.ndarray:
    """slice_tools.py:338 (np.repeat(variables, S)) then :263 (tile by d)."""
    return np.tile(np.repeat(np.asarray(variables), S), d)


def build_matrix_np(var_arrays: list[np.ndarray], mean_center: bool, scale: bool, d: int):
    """src/dmd_era5/era5_svd/era5_svd.py:389-414: order of operations of the build.

    standardize (only when mean_center; scale without mean_center is ignored, Q4)
    -> flatten -> delay-embed; X_mean/X_std are returned only when a mean exists AND
    d > 1 (quirk Q3, era5_svd.py:400-414), replicated d times along space.
    """
    means, stds, data = [], [], []
    for a in var_arrays:
        if mean_center and scale:
            x, mu, sd = standardize_np(a, scale=True)
        elif mean_center:
            x, mu, sd = standardize_np(a, scale=False)
        else:
            x, mu, sd = a, None, None
        data.append(x); means.append(mu); stds.append(sd)
    X = delay_embed_np(flatten_np(data), d)
    X_mean = X_std = None
    if mean_center and d > 1:
        X_mean = np.concatenate([flatten_np(means)] * d)
        if scale:
            X_std = np.concatenate([flatten_np(stds)] * d)
    return X, X_mean, X_std


def resample_nearest_index(times_ns: np.ndarray, delta_ns: int) -> tuple[np.ndarray, np.ndarray]:
    """slice_tools.py:126-141: ds.resample(time=delta).nearest().

    pandas/xarray semantics (xarray Resample.nearest = reindex(full_index, method="nearest"); pandas
    Index.get_indexer(method="nearest")): bins of width delta anchored at midnight of the first day ("start_day"
    origin); the label grid runs from floor(first) to floor(last); every label takes the nearest original sample and
    a label exactly half way between two samples takes the LATER one (pandas breaks ties towards the larger index
    value).  Pinned against pandas itself in tests/test_oracle.py (regular, tied, irregular, up- and down-sampling
    cases), beyond the reference's own 25 hourly -> 5 six-hourly case (tests/test_02_slice_tools.py:85-101).
    Returns (label_times_ns, source_index).
    """
    times_ns = np.asarray(times_ns, dtype=np.int64)
    day = 86400 * 10**9
    origin = (times_ns[0] // day) * day
    first = origin + ((times_ns[0] - origin) // delta_ns) * delta_ns
    last = origin + ((times_ns[-1] - origin) // delta_ns) * delta_ns
    labels = np.arange(first, last + 1, delta_ns, dtype=np.int64)
    # argmin over the reversed distances: the LAST minimum, i.e. the later sample on an exact tie
    n = len(times_ns)
    idx = np.array([n - 1 - int(np.argmin(np.abs(times_ns - t)[::-1])) for t in labels])
    return labels, idx
