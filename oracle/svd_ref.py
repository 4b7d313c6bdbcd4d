"""SVD oracle: the reference's own library calls + a line-by-line restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's numerics for this path are two calls
(src/dmd_era5/era5_svd/era5_svd.py:246-259):
    np.linalg.svd(X, full_matrices=False)            -> truncate to k       (:251-254)
    sklearn.utils.extmath.randomized_svd(X, n_components=k)                 (:258)
The second lives in scikit-learn (unpinned by the reference's pyproject.toml:40; this
image has 1.9.0): randomized_svd extmath.py:402-557, _randomized_svd :560-633,
_randomized_range_finder :313-385, svd_flip :924-982.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


def standard_svd_ref(X: np.ndarray, k: int):
    """era5_svd.py:249-254 verbatim: full thin SVD then slice; no sign normalisation."""
    U, s, V = np.linalg.svd(X, full_matrices=False)
    return U[:, :k], s[:k], V[:k, :]


def randomized_svd_ref(X: np.ndarray, k: int, seed: int):
    """era5_svd.py:258 made deterministic: the reference passes no
    random_state, so sklearn draws from NumPy's global RandomState; seeding it
    immediately before the call reproduces the run bit for bit."""
    from sklearn.utils.extmath import randomized_svd

    np.random.seed(seed)
    return randomized_svd(X, n_components=k)


def omega_ref(n_features: int, k: int, seed: int, dtype) -> np.ndarray:
    """The test matrix the seeded reference call draws: extmath.py:323
    ``random_state.normal(size=(A.shape[1], size))`` with size = k + 10 (:582), cast to
    float32 when A is float32 (:324-334)."""
    Q = np.random.RandomState(seed).normal(size=(n_features, k + 10))
    if np.dtype(dtype) == np.float32:
        Q = Q.astype(np.float32, copy=False)
    return Q


def n_iter_auto(m: int, n: int, k: int) -> int:
    """extmath.py:586-589."""
    return 7 if k < 0.1 * min(m, n) else 4


def svd_flip_u_np(u: np.ndarray, v: np.ndarray):
    """extmath.py:964-972 (u_based_decision=True): sign of the FIRST max-|.| entry of
    each column of u, applied to the column of u and the row of v."""
    idx = np.argmax(np.abs(u.T), axis=1)
    signs = np.sign(u[idx, np.arange(u.shape[1])])
    return u * signs[np.newaxis, :], v * signs[:, np.newaxis]


def randomized_svd_restated(X: np.ndarray, k: int, seed: int):
    """Restatement of _randomized_svd / _randomized_range_finder for tall X
    (transpose='auto' is False when m >= n, extmath.py:591-595) with all defaults the
    reference uses: n_oversamples=10, n_iter='auto', LU normaliser when n_iter > 2
    (:342-354), QR for the last range sample (:383), gesdd for the small SVD (:615),
    u-based sign flip (:623), truncation (:633)."""
    m, n = X.shape
    if m < n:
        raise ValueError("restatement covers the tall (space >= time) case only")
    Q = omega_ref(n, k, seed, X.dtype)
    q = n_iter_auto(m, n, k)
    lu = lambda A: linalg.lu(A, permute_l=True, check_finite=False)[0]
    if q <= 2:
        lu = lambda A: A
    for _ in range(q):
        Q = lu(X @ Q)
        Q = lu(X.T @ Q)
    Q, _ = linalg.qr(X @ Q, mode="economic", check_finite=False)
    B = Q.T @ X
    Uhat, s, Vt = linalg.svd(B, full_matrices=False, lapack_driver="gesdd")
    U = Q @ Uhat
    U, Vt = svd_flip_u_np(U, Vt)
    return U[:, :k], s[:k], Vt[:k, :]
