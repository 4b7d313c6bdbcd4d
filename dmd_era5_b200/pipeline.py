"""Device pipeline of the SVD stage: native ERA5 arrays -> snapshot matrix -> U, s, V.

Host-side mirror of the compute part of ``era5_svd.main``
(src/dmd_era5/era5_svd/era5_svd.py:384-415): variable / level stacking, time-mean removal,
optional scaling, flattening, (virtual) delay embedding and the SVD, all on the GPU.
Row layout of the snapshot matrix (slice_tools.py:323-336): r = v * S + (l * A + a) * O + o.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import torch

from ._cabi import BUILD_CHECK_FINITE, BUILD_MEAN_CENTER, BUILD_SCALE, PREC_NATIVE
from .device_ops import CudaOps
from .dist import LocalComm, shard_rows
from .rsvd import PRECISIONS, draw_omega, n_iter_auto, nvtx_range, randomized_svd_device
from .standard import standard_svd_device


def padded_ld(T: int, dtype: torch.dtype) -> int:
    """Leading dimension of X: rows start on 32-byte boundaries (TMA needs 16)."""
    q = 8 if dtype == torch.float32 else 4
    return -(-T // q) * q


@dataclass
class BuiltMatrix:
    X: torch.Tensor | None          # (m0_local, T) view into a padded buffer, tall dtype (None if only split)
    Xhi: torch.Tensor | None = None # tf32 hi / lo images (precision "tf32x3"): Xhi + Xlo == X exactly
    Xlo: torch.Tensor | None = None
    mean: torch.Tensor | None = None    # (m0_local,) or None
    std: torch.Tensor | None = None
    row_offset: int = 0             # global index of the first local base row
    m0_global: int = 0
    nonfinite: torch.Tensor | None = None
    buffers: list = field(default_factory=list)


def build_matrix_device(ops: CudaOps, var_blocks: list[torch.Tensor], *, mean_center: bool, scale: bool,
                        dtype: torch.dtype | None = None, weights: torch.Tensor | None = None,
                        check_finite: bool = False, comm=None, split: bool = False,
                        keep_x: bool = True) -> BuiltMatrix:
    """Stack (variable, level) blocks into this rank's rows of the snapshot matrix.

    var_blocks : DEVICE tensors (T, P_i) in the native time-major layout, in row order: one per
                 variable (its selected levels flattened to S = L*A*O points) on a single rank, or
                 the pieces of those blocks that fall into this rank's row range (``shard_rows``
                 over the concatenated row index) when the matrix is row-sharded.
    split      : also write the tf32 hi / lo images for precision "tf32x3" (same pass, float32 only);
                 keep_x=False then skips X itself (Xhi + Xlo == X exactly, so nothing is lost).
    Reference quirk Q4 (era5_svd.py:389-395): ``scale`` without ``mean_center`` does nothing.
    """
    T = var_blocks[0].shape[0]
    dtype = dtype or var_blocks[0].dtype
    m0 = sum(int(b.shape[1]) for b in var_blocks)
    ld = padded_ld(T, dtype)
    if split and dtype != torch.float32:
        raise TypeError("split=True needs a float32 snapshot matrix")
    X = ops.empty((m0, ld), dtype)[:, :T] if (keep_x or not split) else None
    Xhi = ops.empty((m0, ld), dtype)[:, :T] if split else None
    Xlo = ops.empty((m0, ld), dtype)[:, :T] if split else None
    do_center = bool(mean_center)
    do_scale = bool(mean_center and scale)
    mean = ops.empty((m0,), dtype) if do_center else None
    std = ops.empty((m0,), dtype) if do_scale else None
    flag = ops.zeros((1,), torch.int32) if check_finite else None
    flags = (BUILD_MEAN_CENTER if do_center else 0) | (BUILD_SCALE if do_scale else 0) | (BUILD_CHECK_FINITE if check_finite else 0)
    r = 0
    with nvtx_range("era5svd.build_matrix"):
        for b in var_blocks:
            P = int(b.shape[1])
            mu = mean[r : r + P] if do_center else None
            sd = std[r : r + P] if do_scale else None
            w = weights[r : r + P] if weights is not None else None
            if split:
                ops.build_rows_split(b, X[r : r + P] if X is not None else None, Xhi[r : r + P], Xlo[r : r + P], mu, sd, w,
                                     flags, flag)
            else:
                ops.build_rows(b, X[r : r + P], mu, sd, w, flags, flag)
            r += P
    return BuiltMatrix(X=X, Xhi=Xhi, Xlo=Xlo, mean=mean, std=std, row_offset=0, m0_global=m0, nonfinite=flag)


_OMEGA_CACHE: dict = {}


def _omega_on_device(ops: CudaOps, n: int, k: int, seed: int | None, dtype: torch.dtype):
    """Test matrix for the randomized SVD.  Unseeded (the reference's behaviour, quirk Q6) it is drawn from NumPy's
    global RandomState at every call.  With a seed it is a pure function of (n, k, seed, dtype), so the device copy
    is cached: repeated calls (SvdStageStream) then issue no small host -> device copy, which would otherwise queue
    behind the multi-GB slice transfer on the same copy engine and stall the compute stream."""
    if seed is None:
        return draw_omega(n, k, None, dtype)
    key = (str(ops.device), n, k, int(seed), dtype)
    om = _OMEGA_CACHE.get(key)
    if om is None:
        om = ops.to_device(torch.from_numpy(draw_omega(n, k, seed, dtype)), non_blocking=False)
        if len(_OMEGA_CACHE) > 16:
            _OMEGA_CACHE.clear()
        _OMEGA_CACHE[key] = om
    return om


def svd_device(ops: CudaOps, X: torch.Tensor | None, *, svd_type: str, n_components: int, delay: int = 1,
               seed: int | None = None, precision: str = "auto", comm=None, row_offset: int = 0,
               m0_global: int | None = None, n_iter: int | None = None, stats: dict | None = None,
               split: tuple[torch.Tensor, torch.Tensor] | None = None, full_iters: int | None = None,
               centred: bool = True):
    """SVD of the (virtual) delay-embedded matrix whose base rows are X (device, tall dtype).
    ``split`` = (Xhi, Xlo) passes pre-split tf32 images (precision "tf32x3"); X may then be None.
    ``centred`` = False (the stage passes ``mean_center``): precision "auto" then keeps 3xTF32 in EVERY power iteration -
    the single-product iterations truncate X to tf32 relative to its VALUES, and a field that still carries its time mean
    (temperature: 250 K with anomalies of a few K) loses 2^-11 * 250 K = 0.12 K there (sigma error 7e-5 emulated against
    5e-6 for centred data).
    Dispatch and error text follow svd_on_era5 (era5_svd.py:247-262)."""
    if X is not None and X.dim() == 2 and (X.stride(1) != 1 or X.stride(0) % (4 if X.dtype == torch.float32 else 2) != 0
                                          or X.data_ptr() % 16 != 0):
        # the TMA-fed kernels need rows that start on 16-byte boundaries: a caller's tightly packed (or transposed) device
        # matrix is repacked once into the padded layout the build kernel produces (one extra copy of X)
        buf = ops.empty((X.shape[0], padded_ld(X.shape[1], X.dtype)), X.dtype)
        buf[:, : X.shape[1]].copy_(X)
        X = buf[:, : X.shape[1]]
    ref = X if X is not None else split[0]
    n = ref.shape[1] - delay + 1
    if precision != "auto" and precision not in PRECISIONS:
        raise ValueError(f"precision {precision} is not supported. Supported values are auto, {', '.join(PRECISIONS)}.")
    if precision in ("tf32x3", "tf32mix") and ref.dtype != torch.float32:
        raise ValueError(f"precision {precision} needs a float32 snapshot matrix (got {ref.dtype}); use 'native' or 'auto'.")
    if precision == "auto":
        # float32 data: tensor-core 3xTF32 passes whenever the sketch / component count fits one MMA tile (<= 128);
        # float64 data (and anything wider): the native path (FP64 DMMA / FP32 FMA)
        width = min(int(n_components) + 10, n) if svd_type == "randomized" else min(int(n_components), n)
        precision = ("tf32mix" if centred else "tf32x3") if (ref.dtype == torch.float32 and width <= 128) else "native"
    if svd_type == "standard":
        if X is None:
            X = split[0] + split[1]
        with nvtx_range("era5svd.standard_svd"):
            return standard_svd_device(ops, X, n_components, delay=delay, comm=comm, precision=PRECISIONS[precision],
                                       stats=stats)
    if svd_type == "randomized":
        omega0 = _omega_on_device(ops, n, n_components, seed, ref.dtype)
        with nvtx_range("era5svd.randomized_svd"):
            return randomized_svd_device(ops, X, n_components, omega0, n_iter=n_iter, delay=delay,
                                     precision=PRECISIONS[precision], comm=comm, row_offset=row_offset,
                                     m0_global=m0_global, stats=stats, split=split, full_iters=full_iters)
    msg = f"SVD type {svd_type} is not supported."
    raise ValueError(msg)
