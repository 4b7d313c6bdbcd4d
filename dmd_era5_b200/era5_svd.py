"""Drop-in for the numerical core of ``dmd_era5.era5_svd`` (reference:
src/dmd_era5/era5_svd/era5_svd.py).

``svd_on_era5(da, parsed_config)`` keeps the reference's signature, return contract
(U (m, k), s (k,), V (k, n) as NumPy arrays in the dtype of X), log messages and error text
(era5_svd.py:230-263), but factorises on the GPU through the C ABI.  There is no CPU fallback.

Opt-in extension keys read from ``parsed_config`` (all default to the reference's behaviour;
the reference's parser ignores unknown keys):
    random_seed : int | None   - None draws Omega from NumPy's global RandomState exactly like the
                                 unseeded reference call (quirk Q6)
    precision   : "auto" (default: "tf32x3" for float32 X when the sketch fits one tensor-core tile, else "native") |
                  "native" (FP64 DMMA for float64 X, FP32 FMA for float32 X) | "tf32x3"
    device      : torch device (default "cuda:0")
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .device_ops import CudaOps
from .pipeline import padded_ld, svd_device

logger = logging.getLogger("ERA5-SVD")


def log_and_print(lg: logging.Logger, msg: str, level: str = "info") -> None:
    """src/dmd_era5/logger.py:42-46: log and print."""
    getattr(lg, level.lower())(msg)
    print(msg)


def setup_logger(name: str, log_file: str, level=logging.INFO) -> logging.Logger:
    """src/dmd_era5/logger.py:7-39: logger ``name`` writing ``<project root>/logs/<log_file>`` in the reference's format;
    existing handlers are replaced.  The reference does this when its modules are IMPORTED (quirk Q8: importing
    era5_svd.py creates ``logs/``); here importing has no side effects and the module entry (``run_module``) sets the two
    loggers of the path up."""
    import os

    from .config_parser import project_root

    formatter = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    log_path = os.path.join(project_root(), "logs")
    if not os.path.exists(log_path):
        os.makedirs(log_path)
    file_handler = logging.FileHandler(os.path.join(log_path, log_file))
    file_handler.setFormatter(formatter)
    lg = logging.getLogger(name)
    lg.setLevel(level)
    for handler in lg.handlers[:]:
        lg.removeHandler(handler)
    lg.addHandler(file_handler)
    return lg


_OPS: dict[str, CudaOps] = {}


def get_ops(device="cuda:0") -> CudaOps:
    key = str(device)
    if key not in _OPS:
        if not torch.cuda.is_available():
            raise RuntimeError("dmd_era5_b200 needs a CUDA device: there is no CPU fallback for the SVD stage")
        _OPS[key] = CudaOps(device)
    return _OPS[key]


def _values(da) -> np.ndarray:
    """``da.values`` for an xarray.DataArray / our DataArray shim, or the array itself."""
    X = da.values if hasattr(da, "values") else da
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError("Input array must be 2D.")
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    return X


def host_to_device_matrix(ops: CudaOps, X: np.ndarray) -> torch.Tensor:
    """Copy a host (m, n) matrix into a padded device buffer (rows 32-byte aligned); returns the
    (m, n) view.  Pinned staging + async copy on the current stream."""
    m, n = X.shape
    t = torch.from_numpy(np.ascontiguousarray(X, dtype=X.dtype.newbyteorder("=")))     # memory-mapped NetCDF-3 data is big-endian
    ld = padded_ld(n, t.dtype)
    buf = ops.empty((m, ld), t.dtype)
    view = buf[:, :n]
    view.copy_(t.pin_memory() if m * n >= 1 << 16 else t, non_blocking=True)
    return view


def looks_centred(X: np.ndarray, rows: int = 1024, ratio: float = 8.0) -> bool:
    """Does the matrix look time-mean-centred?  Median over a sample of rows of |row mean| / row std (a centred or
    anomaly field: << 1; temperature with its 250 K mean left in: 30 - 50).  svd_on_era5 receives a bare array, so this is
    how precision "auto" decides between the mixed schedule and 3xTF32 in every pass (pipeline.svd_device, ``centred``):
    the single-product passes truncate X relative to its VALUES."""
    m, n = X.shape
    if n < 2 or m == 0:
        return True
    idx = np.unique(np.linspace(0, m - 1, min(m, rows)).astype(np.int64))
    S = np.asarray(X[idx], dtype=np.float64)
    mu, sd = S.mean(axis=1), S.std(axis=1)
    ok = np.isfinite(mu) & np.isfinite(sd) & (sd > 0)
    if not ok.any():
        return True
    return float(np.median(np.abs(mu[ok]) / sd[ok])) <= ratio


def svd_on_era5(da, parsed_config: dict) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Perform SVD on the pre-processed ERA5 slice (era5_svd.py:230-263).

    Returns U (n_samples, n_components), s (n_components,), V (n_components, n_features).
    """
    X = _values(da)
    svd_type = parsed_config["svd_type"]
    n_components = parsed_config["n_components"]
    if svd_type not in ("standard", "randomized"):
        msg = f"SVD type {svd_type} is not supported."
        raise ValueError(msg)
    ops = get_ops(parsed_config.get("device", "cuda:0"))
    label = "standard" if svd_type == "standard" else "randomized"
    log_and_print(logger, f"Performing {label} SVD...")
    with torch.cuda.device(ops.device):
        Xd = host_to_device_matrix(ops, X)
        nonfinite = ops.check_finite(Xd)        # asynchronous; read after the SVD has been queued
        U, s, V = svd_device(ops, Xd, svd_type=svd_type, n_components=n_components,
                             seed=parsed_config.get("random_seed"),
                             precision=parsed_config.get("precision", "auto"),
                             centred=X.dtype != np.float32 or looks_centred(X))
        if int(nonfinite.item()):
            # sklearn check_array (extmath.py:546) / LAPACK on the reference side; main() wraps it as
            # "Error in the SVD on ERA5 process: ..." (era5_svd.py:426-429)
            raise ValueError("Input contains NaN or infinity.")
        out_dtype = Xd.dtype
        U_h = U.to(out_dtype).cpu().numpy()
        s_h = s.to(out_dtype).cpu().numpy()
        V_h = V.to(out_dtype).cpu().numpy()
    log_and_print(logger, f"{label.capitalize()} SVD complete.")
    return U_h, s_h, V_h


# Orchestration / packaging names of the reference's era5_svd module live in stage.py (which imports
# from this module), so they are re-exported lazily to keep the reference's import surface:
#   from dmd_era5_b200.era5_svd import main, combine_svd_results, add_config_attributes, ...
_STAGE_NAMES = ("main", "combine_svd_results", "add_config_attributes", "retrieve_era5_slice", "retrieve_svd_results")


def __getattr__(name):
    if name in _STAGE_NAMES:
        from . import stage

        return getattr(stage, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def check_if_dvc_repo() -> bool:
    """era5_svd.py:457-464: is the project root a DVC repository?  (``dvc`` imported lazily; absent = no repository.)"""
    try:
        from dvc.repo import Repo as DvcRepo

        from .config_parser import project_root

        with DvcRepo(project_root()) as _:
            return True
    except Exception:
        return False


def run_module() -> None:
    """``python -m dmd_era5_b200.era5_svd`` = the reference's ``python -m dmd_era5.era5_svd.era5_svd`` (era5_svd.py:455-478):
    the ``[era5-svd]`` section of config.ini, results written to NetCDF, DVC used when the project is a DVC repository."""
    from .stage import main

    setup_logger("ERA5-SVD", "era5_svd.log")                    # era5_svd.py:33
    setup_logger("ERA5Processing", "era5_processing.log")        # slice_tools.py:12
    if not check_if_dvc_repo():
        log_and_print(logger, "Not a Data Version Control (DVC) repository. Will not use DVC.", level="warning")
        log_and_print(logger, "To initialize a DVC repository, run `dvc init`.", level="warning")
        main(write_to_netcdf=True)
    else:
        main(write_to_netcdf=True, use_dvc=True)


if __name__ == "__main__":
    run_module()
