"""Standard (full) SVD on the device by the Gram route.

Replaces ``np.linalg.svd(X, full_matrices=False)`` + truncation
(src/dmd_era5/era5_svd/era5_svd.py:249-254; LAPACK gesdd, O(m n^2) on the host and an m x n
temporary U).  For the tall-skinny snapshot matrix:

    G = X^T X          (n x n, float64; tall pass, all-reduced over row shards)
    G = V L V^T        (Jacobi eigensolver)         s = sqrt(L)
    U_k = X V_k S_k^-1 (tall pass)                  Vt_k = V_k^T

Only k columns of U are ever formed.  Like the reference, no sign normalisation is applied
(LAPACK's signs are arbitrary too); parity is up to a per-pair sign.  The Gram matrix squares the
condition number: sigma_i is accurate to ~ eps * (sigma_1 / sigma_i)^2, i.e. 1e-6 down to
sigma_i ~ 1e-5 sigma_1 in float64.
"""
from __future__ import annotations

import torch

from ._cabi import PREC_NATIVE
from .dist import LocalComm


def standard_svd_device(ops, X: torch.Tensor, n_components: int, *, delay: int = 1, comm=None):
    """X: this rank's base rows (m0_local, T).  Returns (U_local (m0_local*delay, k), s (k,), Vt (k, n))."""
    comm = comm or LocalComm()
    m0, T = X.shape
    d = delay
    n = T - d + 1
    if n < 1:
        raise ValueError("delay embedding larger than the number of snapshots")
    k = min(int(n_components), n)
    G = None
    for j in range(d):
        Xj = X[:, j : j + n]
        G = ops.project(Xj, Xj, G, accumulate=j > 0, precision=PREC_NATIVE)
    comm.allreduce_sum_(G)
    lam, V = ops.syevj(G)
    s, inv_s = ops.sigma_from_eig(lam)
    Vk = V[:, :k].t().contiguous()              # (k, n): rows = right singular vectors, descending
    M = Vk.clone()
    ops.scale_rows(M, inv_s[:k].contiguous())   # S_k^-1 V_k^T
    Mt = ops.convert(M.t().contiguous(), X.dtype)   # (n, k) = V_k S_k^-1 in the tall dtype
    U = ops.empty((m0 * d, k), X.dtype)
    for j in range(d):
        ops.sketch(X[:, j : j + n], Mt, U[j * m0 : (j + 1) * m0], PREC_NATIVE)
    return U, s[:k], Vk
