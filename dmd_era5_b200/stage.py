"""Orchestration and packaging of the SVD stage - the reference's ``era5_svd.main`` and helpers
(src/dmd_era5/era5_svd/era5_svd.py:42-227, 266-453) with the compute phase replaced by the fused
device pipeline.  Same signatures, return tuple, phase-wrapped error messages, attribute schema and
output layout (README.md:97-119): data_vars U (space, components), s (components), V (components,
time), optional X, X_mean, X_std; coords space, components, time, original_variable, delay, level,
latitude, longitude.

DVC plumbing: ``dvc_tools.py`` of this package (same calls as the reference's); where ``dvc`` / ``GitPython`` are not
installed a retrieval behaves like the reference's failed retrieval (warning, falls through to computing) and adding
results raises the reference's "Error adding SVD results to DVC" error.
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np
import torch

from .config_parser import config_parser
from .dataset import DataArray, Dataset, read_netcdf, write_netcdf
from .era5_svd import get_ops, log_and_print, logger
from .pipeline import build_matrix_device, svd_device
from .slice_tools import (log_standardize, resample_era5_dataset, slice_era5_dataset, space_coord_to_level_lat_lon,
                          space_coords)


def add_config_attributes(ds: Dataset, parsed_config: dict) -> Dataset:
    """era5_svd.py:42-66."""
    a = ds.attrs
    a["source_path"] = parsed_config["source_path"]
    a["n_components"] = parsed_config["n_components"]
    a["variables"] = parsed_config["variables"]
    a["levels"] = parsed_config["levels"]
    a["mean_center"] = int(parsed_config["mean_center"])
    a["scale"] = int(parsed_config["scale"])
    a["delay_embedding"] = parsed_config["delay_embedding"]
    a["svd_type"] = parsed_config["svd_type"]
    a["era5_slice_path"] = parsed_config["era5_slice_path"]
    a["date_processed"] = datetime.now().isoformat()
    a["save_data_matrix"] = int(parsed_config["save_data_matrix"])
    return ds


def _as_str_list(obj) -> list[str]:
    if isinstance(obj, (list, tuple, np.ndarray)):
        return [str(x) for x in obj]
    return [s for s in str(obj).split(",")]


def _as_int_list(obj) -> list[int]:
    """era5_svd.py:90-99: levels stored as a (possibly length-1) integer array."""
    if isinstance(obj, (int, np.integer)):
        return [int(obj)]
    arr = np.asarray(obj)
    if not np.issubdtype(arr.dtype, np.integer):
        raise ValueError("Levels must be integers.")
    return [int(x) for x in arr.ravel().tolist()]


def _retrieve_from_dvc(parsed_config: dict, data_type: str, what: str, path_key: str, lazy: bool):
    """The ``retrieve_from_dvc`` closures of era5_svd.py:116-129 / 190-203: check the matching version out of DVC and
    open it; a failed retrieval (nothing logged, nothing matching, no cache / remote, or no dvc installation) is a
    warning and ``None``, never an error."""
    from . import dvc_tools

    log_and_print(logger, f"Attempting to retrieve {what} from DVC...")
    try:
        dvc_tools.retrieve_data_from_dvc(parsed_config, data_type=data_type)
        log_and_print(logger, f"{what} retrieved from DVC: {parsed_config[path_key]}")
        return read_netcdf(parsed_config[path_key], lazy=lazy)
    except (FileNotFoundError, ValueError) as e:
        log_and_print(logger, f"Could not retrieve {what} from DVC: {e}", "warning")
        return None


def retrieve_era5_slice(parsed_config: dict, use_dvc: bool = False):
    """era5_svd.py:69-154: slice in the working directory whose attrs cover the request
    (superset match on variables / levels, equal source_path)."""
    path = parsed_config["era5_slice_path"]
    if os.path.exists(path):
        log_and_print(logger, "ERA5 slice found in working directory.")
        ds = read_netcdf(path, lazy=True)        # input file: the variables stay memory-mapped until they are staged
        a = ds.attrs
        ok = (sorted(parsed_config["variables"]) == sorted(set(_as_str_list(a["variables"])) & set(parsed_config["variables"]))
              and sorted(parsed_config["levels"]) == sorted(set(_as_int_list(a["levels"])) & set(parsed_config["levels"]))
              and parsed_config["source_path"] == a["source_path"])
        if ok:
            log_and_print(logger, "ERA5 slice matches configuration.")
            return ds, False
        log_and_print(logger, "ERA5 slice does not match configuration.")
        if use_dvc:
            ds = _retrieve_from_dvc(parsed_config, "era5_slice", "ERA5 slice", "era5_slice_path", lazy=True)
            return (ds, True) if ds is not None else (None, False)
        log_and_print(logger, "ERA5 slice in working directory does not match configuration.", "warning")
        return None, False
    log_and_print(logger, "ERA5 slice not found in working directory.", "warning")
    if use_dvc:
        ds = _retrieve_from_dvc(parsed_config, "era5_slice", "ERA5 slice", "era5_slice_path", lazy=True)
        return (ds, True) if ds is not None else (None, False)
    return None, False


def retrieve_svd_results(parsed_config: dict, use_dvc: bool = False):
    """era5_svd.py:157-227: result-level memoisation; matches on the same seven attributes as the
    reference (svd_type and save_data_matrix are NOT compared - quirk Q5)."""
    path = parsed_config["save_path"]
    if os.path.exists(path):
        log_and_print(logger, "SVD results found in working directory.")
        ds = read_netcdf(path)
        a = ds.attrs
        ok = (parsed_config["source_path"] == a["source_path"] and parsed_config["n_components"] == a["n_components"]
              and parsed_config["variables"] == _as_str_list(a["variables"])
              and parsed_config["levels"] == _as_int_list(a["levels"])
              and parsed_config["mean_center"] == a["mean_center"] and parsed_config["scale"] == a["scale"]
              and parsed_config["delay_embedding"] == a["delay_embedding"])
        if ok:
            log_and_print(logger, "SVD results match configuration.")
            return ds, False
        log_and_print(logger, "SVD results do not match configuration.")
        if use_dvc:
            ds = _retrieve_from_dvc(parsed_config, "era5_svd", "SVD results", "save_path", lazy=False)
            return (ds, True) if ds is not None else (None, False)
        log_and_print(logger, "SVD results in working directory do not match configuration.", "warning")
        return None, False
    log_and_print(logger, "SVD results not found in working directory.", "warning")
    if use_dvc:
        ds = _retrieve_from_dvc(parsed_config, "era5_svd", "SVD results", "save_path", lazy=False)
        return (ds, True) if ds is not None else (None, False)
    return None, False


def combine_svd_results(U, s, V, coords: dict, **kwargs) -> Dataset:
    """era5_svd.py:266-333.  ``coords`` = coords of the (delay-embedded) matrix: space, time,
    original_variable, delay (+ level / latitude / longitude carried along)."""
    k = U.shape[1]
    space_co = {name: coords[name] for name in ("space", "original_variable", "delay", "level", "latitude", "longitude")
                if name in coords}
    dv = {
        "U": DataArray(U, ("space", "components"), {**space_co, "components": np.arange(k)}),
        "s": DataArray(s, ("components",), {"components": np.arange(s.shape[0])}),
        "V": DataArray(V, ("components", "time"), {"components": np.arange(V.shape[0]), "time": coords["time"]}),
    }
    for name, dims in (("X", ("space", "time")), ("X_mean", ("space",)), ("X_std", ("space",))):
        val = kwargs.get(name)
        if val is not None:
            dv[name] = val if isinstance(val, DataArray) else DataArray(val, dims)
    all_coords = dict(coords)
    all_coords["components"] = (("components",), np.arange(k))
    return Dataset(dv, all_coords)


_NATIVE_DIMS = ("time", "level", "latitude", "longitude")
_STAGING: dict = {}            # (device, bytes) -> two pinned host buffers, reused across calls


def _pinned_ring(device, nbytes: int):
    key = (str(device), int(nbytes))
    if key not in _STAGING:
        _STAGING.clear()       # one ring per process is enough; do not accumulate pinned memory across sizes
        _STAGING[key] = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    return _STAGING[key]


def _gather_chunk(base: np.ndarray, t_idx, l_idx, c0: int, c1: int, host: np.ndarray) -> None:
    """host[...] = base[time selection c0:c1][:, level selection] in ONE pass over the source where NumPy allows it
    (contiguous time range or a native-byte-order gather); byte-order conversion rides on the same copy."""
    native = base.dtype.isnative
    if l_idx is None:
        if t_idx is None:
            np.copyto(host, base[c0:c1])                               # view -> pinned buffer
        elif native:
            np.take(base, t_idx[c0:c1], axis=0, out=host, mode="clip")  # indices validated by LazyTake.take
        else:
            np.copyto(host, base[t_idx[c0:c1]])
        return
    rows = base[c0:c1] if t_idx is None else base[t_idx[c0:c1]]
    if native:
        np.take(rows, l_idx, axis=1, out=host, mode="clip")
    else:
        np.copyto(host, np.take(rows, l_idx, axis=1))


_POOL = None


def _gather_chunk_mt(base: np.ndarray, t_idx, l_idx, c0: int, c1: int, host: np.ndarray) -> None:
    """_gather_chunk over a few host threads (NumPy releases the GIL inside large copies): the host-side gather, not
    PCIe, bounds the upload of a slice that is not pinned yet."""
    global _POOL
    n = c1 - c0
    threads = min(8, len(os.sched_getaffinity(0)), n)
    if threads <= 1 or host.nbytes < (8 << 20):
        return _gather_chunk(base, t_idx, l_idx, c0, c1, host)
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=8, thread_name_prefix="era5svd-stage")
    step = -(-n // threads)
    futs = [_POOL.submit(_gather_chunk, base, t_idx, l_idx, c0 + a, min(c1, c0 + a + step), host[a : a + step])
            for a in range(0, n, step)]
    for f in futs:
        f.result()


def shard_pieces(S: int, n_vars: int, r0: int, r1: int) -> list[tuple[int, int, int]]:
    """Base rows [r0, r1) of the stacked matrix (row r = v * S + p, slice_tools.py:323-336) as per-variable point
    ranges [(v, p0, p1), ...] - what one rank of a row-sharded run has to stage."""
    out = []
    for v in range(n_vars):
        a, b = max(r0, v * S), min(r1, (v + 1) * S)
        if a < b:
            out.append((v, a - v * S, b - v * S))
    return out


def stage_blocks(ds: Dataset, variables: list[str], ops, chunk_bytes: int | None = None,
                 pieces: list[tuple[int, int, int]] | None = None) -> tuple[list[torch.Tensor], int]:
    """Host slice -> device blocks in the native (T, S = L*A*O) time-major layout, one per variable
    (what era5_svd.py:384-388 hands to the matrix build, era5_download.py:36-42 being the file layout).

    The pending time / level selections of ``slice_era5_dataset`` and ``resample_era5_dataset`` (dataset.LazyTake) are
    carried out HERE, while the data is copied into a ring of two pinned staging buffers (``chunk_bytes`` each,
    default 128 MiB, $ERA5SVD_STAGE_CHUNK_MB): chunk i + 1 is gathered on the host while chunk i crosses PCIe on a
    copy stream.  A slice is therefore read ONCE on its way to the device (any byte-order or dtype-preserving layout
    fix-up happens in that same pass) instead of once per selection step plus a pinning copy."""
    if chunk_bytes is None:
        chunk_bytes = int(os.environ.get("ERA5SVD_STAGE_CHUNK_MB", "128")) << 20
    dev = ops.device
    copy_stream = torch.cuda.Stream(device=dev)
    copy_stream.wait_stream(torch.cuda.current_stream(dev))
    blocks = []
    ring = None
    ring_events = [None, None]
    slot = 0
    S_full = None
    todo = [(vi, None, None) for vi in range(len(variables))] if pieces is None else list(pieces)
    for vi, p0, p1 in todo:
        v = variables[vi]
        da = ds[v]
        lz = da.lazy()
        if da.dims != _NATIVE_DIMS:
            # any other dimension order: materialise and transpose on the host (not what era5_download writes)
            a = np.transpose(np.asarray(da.values), [da.dims.index(dm) for dm in _NATIVE_DIMS])
            lz = type(lz)(np.ascontiguousarray(a))
        base = lz.base
        T, L, A, O = lz.shape
        S_full = L * A * O
        t_idx, l_idx = lz.index.get(0), lz.index.get(1)
        col0, ncols = 0, S_full
        if p0 is not None:
            # the levels that hold points [p0, p1): gather only those, keep the column window inside them
            l_lo, l_hi = p0 // (A * O), (p1 - 1) // (A * O) + 1
            l_idx = (np.arange(L) if l_idx is None else np.asarray(l_idx))[l_lo:l_hi]
            col0, ncols, L = p0 - l_lo * A * O, p1 - p0, l_hi - l_lo
        if 2 in lz.index or 3 in lz.index:
            base = np.take(np.take(base, lz.index.get(2, np.arange(base.shape[2])), axis=2),
                           lz.index.get(3, np.arange(base.shape[3])), axis=3)
        tdt = torch.from_numpy(np.empty(0, dtype=base.dtype.newbyteorder("="))).dtype
        row_bytes = L * A * O * base.dtype.itemsize
        block = torch.empty((T, ncols), dtype=tdt, device=dev)
        if T * row_bytes < (1 << 16):
            # tiny slice (the mock configuration): one pageable copy
            full = np.ascontiguousarray(np.asarray(lz), dtype=base.dtype.newbyteorder("=")).reshape(T, -1)
            if p0 is not None:
                full = np.ascontiguousarray(full[:, p0:p1])
            block.copy_(torch.from_numpy(full))
            blocks.append(block)
            continue
        rows = max(1, min(T, chunk_bytes // row_bytes))
        need = rows * row_bytes
        if ring is None or ring[0].numel() < need:
            ring = _pinned_ring(dev, max(need, chunk_bytes if row_bytes <= chunk_bytes else need))
            ring_events = [None, None]
        for c0 in range(0, T, rows):
            c1 = min(T, c0 + rows)
            buf = ring[slot]
            if ring_events[slot] is not None:
                ring_events[slot].synchronize()          # the previous upload from this buffer has finished
            host = buf[: (c1 - c0) * row_bytes].view(tdt).view(c1 - c0, L, A, O).numpy()
            _gather_chunk_mt(base, t_idx, l_idx, c0, c1, host)
            with torch.cuda.stream(copy_stream):
                block[c0:c1].copy_(buf[: (c1 - c0) * row_bytes].view(tdt).view(c1 - c0, L * A * O)[:, col0 : col0 + ncols],
                                   non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            ring_events[slot] = ev
            slot ^= 1
        blocks.append(block)
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    for b in blocks:
        b.record_stream(copy_stream)
    return blocks, (S_full if pieces is not None else blocks[0].shape[1])


def _to_blocks(ds: Dataset, variables: list[str], ops) -> tuple[list[torch.Tensor], int]:
    return stage_blocks(ds, variables, ops)


def main(config: dict | None = None, write_to_netcdf: bool = False, use_dvc: bool = False):
    """era5_svd.py:336-453.  Returns (svd_results, added_to_dvc, retrieved_from_dvc)."""
    if config is None:
        from .config_parser import config_reader

        config = config_reader("era5-svd")        # the reference reads config.ini at import (quirk Q8)
    added_to_dvc = False
    parsed_config = config_parser(config, "era5-svd")
    try:
        svd_results, retrieved_from_dvc = retrieve_svd_results(parsed_config, use_dvc)
    except Exception as e:
        msg = f"Error retrieving SVD results: {e}"
        log_and_print(logger, msg, "error")
        raise Exception(msg) from e

    if svd_results is None:
        try:
            ds, _ = retrieve_era5_slice(parsed_config, use_dvc)
            if ds is None:
                msg = ("Could not retrieve ERA5 slice from working directory or DVC." if use_dvc else
                       "\n                    Could not retrieve ERA5 slice from working directory.\n"
                       "                    Consider using DVC to retrieve the ERA5 slice, if available.\n                    ")
                log_and_print(logger, msg, "error")
                raise FileNotFoundError(msg)
        except Exception as e:
            msg = f"Error retrieving ERA5 slice: {e}"
            log_and_print(logger, msg, "error")
            raise Exception(msg) from e
        try:
            svd_results = _compute(ds, parsed_config)
        except Exception as e:
            msg = f"Error in the SVD on ERA5 process: {e}"
            log_and_print(logger, msg, "error")
            raise Exception(msg) from e

        if write_to_netcdf:
            try:
                log_and_print(logger, "Writing SVD results to NetCDF...")
                write_netcdf(svd_results, parsed_config["save_path"])
                log_and_print(logger, f"SVD results written to {parsed_config['save_path']}")
            except Exception as e:
                msg = f"Error writing SVD results to NetCDF: {e}"
                log_and_print(logger, msg, "error")
                raise Exception(msg) from e
            if use_dvc:                                                  # era5_svd.py:442-451
                try:
                    from . import dvc_tools

                    log_and_print(logger, "Adding SVD results to DVC...")
                    dvc_tools.add_data_to_dvc(parsed_config["save_path"], svd_results.attrs)
                    log_and_print(logger, "SVD results added to DVC.")
                    added_to_dvc = True
                except Exception as e:
                    msg = f"Error adding SVD results to DVC: {e}"
                    log_and_print(logger, msg, "error")
                    raise Exception(msg) from e
    return svd_results, added_to_dvc, retrieved_from_dvc


def _prepare(ds: Dataset, parsed_config: dict) -> Dataset:
    """Selections of the compute phase (era5_svd.py:384-388): host index work, lazy on file-backed data."""
    ds = ds[parsed_config["variables"]]                                  # config order (:384)
    ds = slice_era5_dataset(ds, levels=parsed_config["levels"])          # (:385) no start/end: quirk Q7
    return resample_era5_dataset(ds, parsed_config["delta_time"])        # (:388)


def _row_weights(ds: Dataset, n_vars: int) -> np.ndarray:
    """sqrt(cos(latitude)) per base row - opt-in extension (absent in the reference)."""
    lat = np.deg2rad(np.asarray(ds.coord("latitude"), dtype=np.float64))
    w = np.sqrt(np.clip(np.cos(lat), 0.0, None))
    L, O = len(ds.coord("level")), len(ds.coord("longitude"))
    return np.tile(np.tile(np.repeat(w, O), L), n_vars)


def _device_arrays(ds: Dataset, parsed_config: dict, ops, comm=None, rank: int = 0, world: int = 1) -> dict:
    """Build + SVD of this rank's rows on ``ops.device`` (era5_svd.py:389-415).  Returns host arrays: U (rows of this
    rank, block-major over the delay blocks), s, V, optional X / mean / std of this rank, and the row range."""
    from .dist import shard_rows

    variables, d = parsed_config["variables"], parsed_config["delay_embedding"]
    mean_center = bool(parsed_config["mean_center"])
    scale = bool(parsed_config["scale"]) and mean_center                 # quirk Q4 (:389-395)
    precision = parsed_config.get("precision", "auto")
    S = int(np.prod([len(ds.coord(c)) for c in ("level", "latitude", "longitude")]))
    m0 = len(variables) * S
    r0, r1 = shard_rows(m0, world, rank)
    with torch.cuda.device(ops.device):
        if world > 1:
            blocks, _ = stage_blocks(ds, variables, ops, pieces=shard_pieces(S, len(variables), r0, r1))
        else:
            blocks, _ = _to_blocks(ds, variables, ops)
        weights = None
        if parsed_config.get("area_weighting"):                          # opt-in extension (absent in the reference)
            w_rows = _row_weights(ds, len(variables))[r0:r1]
            weights = torch.from_numpy(w_rows.astype(np.float32 if blocks[0].dtype == torch.float32 else np.float64)).to(ops.device)
        # opt-in extension (north_star "float cast"; absent in the reference): matrix_dtype = "float32" stores the
        # snapshot matrix in float32 whatever the slice's dtype (the cast happens inside the build kernel, statistics
        # still accumulate in float64), which halves the memory and enables the tensor-core passes for float64 slices
        mdt = parsed_config.get("matrix_dtype")
        if mdt not in (None, "float32", "float64"):
            raise ValueError(f"matrix_dtype {mdt} is not supported.")
        xdtype = None if mdt is None else (torch.float32 if mdt == "float32" else torch.float64)
        if weights is not None and xdtype is not None:
            weights = weights.to(xdtype)
        built = build_matrix_device(ops, blocks, mean_center=mean_center, scale=scale, weights=weights,
                                    check_finite=True, dtype=xdtype)
        del blocks
        label = "standard" if parsed_config["svd_type"] == "standard" else "randomized"
        if rank == 0:
            log_and_print(logger, f"Performing {label} SVD...")
        U, s, V = svd_device(ops, built.X, svd_type=parsed_config["svd_type"], n_components=parsed_config["n_components"],
                             delay=d, seed=parsed_config.get("random_seed"), precision=precision, comm=comm,
                             row_offset=r0, m0_global=m0, centred=mean_center)
        if rank == 0:
            log_and_print(logger, f"{label.capitalize()} SVD complete.")
        bad = built.nonfinite.to(torch.float64)
        if comm is not None:
            comm.allreduce_sum_(bad)
        if float(bad.item()) > 0:
            raise ValueError("Input contains NaN or infinity.")          # sklearn check_array (extmath.py:546)
        dt = built.X.dtype
        out = {"U": U.to(dt).cpu().numpy(), "s": s.to(dt).cpu().numpy(), "V": V.to(dt).cpu().numpy(),
               "X": built.X.cpu().numpy() if parsed_config["save_data_matrix"] else None,
               "mean": built.mean.cpu().numpy() if built.mean is not None else None,
               "std": built.std.cpu().numpy() if built.std is not None else None, "rows": (r0, r1), "m0": m0, "S": S}
    return out


def _compute(ds: Dataset, parsed_config: dict) -> Dataset:
    """The compute phase of main (era5_svd.py:384-425) on the device(s).  ``n_gpus`` > 1 (opt-in key)
    row-shards the stage over that many GPUs of this node (stage_multi.py); the result is the same Dataset."""
    variables, d = parsed_config["variables"], parsed_config["delay_embedding"]
    n_gpus = int(parsed_config.get("n_gpus", 1) or 1)
    dsp = _prepare(ds, parsed_config)
    if parsed_config["mean_center"]:                 # standardize_data's progress lines (:389-392); Q4: scale needs centring
        log_standardize("time", bool(parsed_config["scale"]))
    if n_gpus > 1:
        # a rank needs at least one 128-row tile of the base matrix (dist.shard_rows): small inputs use fewer GPUs
        S_ = int(np.prod([len(dsp.coord(c)) for c in ("level", "latitude", "longitude")]))
        tiles = -(-(len(variables) * S_) // 128)
        if tiles < n_gpus:
            log_and_print(logger, f"n_gpus = {n_gpus}, but the matrix has only {tiles} row tile(s) of 128: using {tiles} GPU(s).",
                          "warning")
            n_gpus = tiles
    if n_gpus > 1:
        from .stage_multi import compute_multi

        arr = compute_multi(parsed_config, n_gpus)
    else:
        arr = _device_arrays(dsp, parsed_config, get_ops(parsed_config.get("device", "cuda:0")))
    ds = dsp
    U_h, s_h, V_h, X_h, mean_h, std_h, S = (arr[x] for x in ("U", "s", "V", "X", "mean", "std", "S"))

    m0 = len(variables) * S
    times = ds.coord("time")
    lev, lat, lon = space_coords(ds.coord("level"), ds.coord("latitude"), ds.coord("longitude"), len(variables), d)
    coords = {
        "space": (("space",), np.arange(m0 * d)),
        "time": (("time",), times[d - 1:]),
        "original_variable": (("space",), np.tile(np.repeat(variables, S), d)),
        "delay": (("space",), np.repeat(np.flip(np.arange(d)), m0)),
        "level": (("space",), lev), "latitude": (("space",), lat), "longitude": (("space",), lon),
    }
    extra = {}
    if X_h is not None:
        n = X_h.shape[1] - d + 1
        # the reference hands its DataArray to the Dataset (era5_svd.py:416-418), so X carries the attributes that
        # flatten_era5_variables / apply_delay_embedding set on it (slice_tools.py:362-363, :271), after the slice's own
        # attributes when nothing was centred (xarray's arithmetic in standardize_data drops them otherwise)
        x_attrs = {} if parsed_config["mean_center"] else dict(ds.attrs)
        x_attrs.update(original_variables=list(variables), space_coords=["level", "latitude", "longitude"], delay_embedding=d)
        extra["X"] = DataArray(np.concatenate([X_h[:, j : j + n] for j in range(d)], axis=0), ("space", "time"), attrs=x_attrs)
    if mean_h is not None and d > 1:                                     # quirk Q3 (:400-414): dropped when d == 1
        # flattened by the same function as X (:401, :406), hence with its two attributes; xr.concat keeps them
        stat_attrs = {"original_variables": list(variables), "space_coords": ["level", "latitude", "longitude"]}
        extra["X_mean"] = DataArray(np.concatenate([mean_h] * d), ("space",), attrs=stat_attrs)
        if std_h is not None:
            extra["X_std"] = DataArray(np.concatenate([std_h] * d), ("space",), attrs=stat_attrs)
    out = combine_svd_results(U_h, s_h, V_h, coords, **extra)
    out = add_config_attributes(out, parsed_config)
    return space_coord_to_level_lat_lon(out)
