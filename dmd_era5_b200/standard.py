"""Standard (full) SVD on the device by the Gram route.

Replaces ``np.linalg.svd(X, full_matrices=False)`` + truncation
(src/dmd_era5/era5_svd/era5_svd.py:249-254; LAPACK gesdd, O(m n^2) on the host and an m x n
temporary U).  For the tall-skinny snapshot matrix:

    G = X^T X          (n x n, float64; tall passes, all-reduced over row shards)
    G = V L V^T        k largest eigenpairs only (the reference truncates to n_components):
                         n <= 118 : one-CTA Jacobi (small_f64.cu)
                         larger n : Householder tridiagonalisation on all SMs + bisection / inverse iteration
                                    + one Rayleigh-Ritz step + back-transformation (eig_tridiag.cu)
    s = sqrt(L);  U_k = X V_k S_k^-1 (tall pass);  Vt_k = V_k^T

Only k columns of U are ever formed.  Like the reference, no sign normalisation is promised
(LAPACK's signs are arbitrary too); parity is up to a per-pair sign.  The Gram matrix squares the
condition number: sigma_i is accurate to ~ eps * (sigma_1 / sigma_i)^2, i.e. 1e-6 down to
sigma_i ~ 1e-5 sigma_1 for float64 data.  Components whose eigenvalue comes out <= 0 (below that
floor, e.g. the null direction of mean-centred data when n_components ~ n) get sigma = 0 and a ZERO column of U -
the reference returns an arbitrary orthonormal completion there.

float32 data (the real ERA5 dtype): the tensor-core Gram accumulates in fp32, eps ~ 1e-6, which alone would
leave sigma_i at 1e-2 relative for sigma_i = 0.01 sigma_1 where LAPACK's sgesdd is backward stable.  The Gram route
therefore only supplies the SUBSPACE: the k + 10 leading eigenvectors start REFINE_ITERS power iterations and the
final Rayleigh-Ritz stage of the randomized driver (rsvd.py) on X itself - Y = X V, Q = orth(Y), B = Q^T X, small SVD -
whose singular values carry the accuracy of the tall passes (3xTF32: 1e-7 relative to sigma_i, not to sigma_1^2) and
whose error in the subspace enters only squared.  Cost: 2 * REFINE_ITERS + 2 tall passes of width k + 10 against the
n / 2 column blocks of the Gram matrix (c4: +15 %).
"""
from __future__ import annotations

import math

import torch

from ._cabi import PREC_NATIVE, PREC_TF32X3
from .dist import LocalComm

REFINE_ITERS = 2        # power iterations of the float32 refinement (each contracts the leak by (sigma_{k+11}/sigma_i)^2)
JACOBI_MAX_N = 118      # largest n whose working copies fit one CTA's shared memory (syevj_kernel)
TC_BLOCK = 112          # sketch-width block of the tensor-core project kernel
TC_BLOCK_X1 = 256       # column block of the single-product Gram (two M = 128 operand tiles per CTA: half the reads of X)
PREC_TF32MIX = 2        # = rsvd.PREC_TF32MIX (driver-level value, never crosses the C ABI)


def _orth2(ops, Z: torch.Tensor) -> torch.Tensor:
    """Two CholeskyQR sweeps: orthonormal basis of span(Z) for a well-conditioned n x k float64 Z."""
    for _ in range(2):
        G = ops.gemm(Z, Z, transA=True)
        _, Rinv = ops.chol_inv(G, 1e-14)
        Z = ops.gemm(Z, Rinv)
    return Z


SUBSPACE_MIN_N = 1024        # below this the tridiagonal route is already milliseconds
SUBSPACE_MAX_ITERS = 60
_START: dict = {}


def _subspace_start(ops, n: int, b: int) -> torch.Tensor:
    """Deterministic start block for the subspace iteration (host NumPy RandomState(0); cached per device and shape)."""
    import numpy as np

    key = (str(ops.device), n, b)
    if key not in _START:
        if len(_START) > 4:
            _START.clear()
        _START[key] = ops.to_device(torch.from_numpy(np.random.RandomState(0).standard_normal((n, b))), non_blocking=False)
    return _START[key].clone()


def sym_eig_topk_subspace(ops, G: torch.Tensor, k: int, tol: float, stats: dict | None = None):
    """k largest eigenpairs of the symmetric positive semi-definite float64 G by block subspace iteration with
    Rayleigh-Ritz, or None when it has not converged within SUBSPACE_MAX_ITERS (the caller then takes the direct route).

    The standard route only ever needs the k << n leading pairs of the n x n Gram matrix (the reference truncates to
    n_components, era5_svd.py:252-254), while the Householder tridiagonalisation moves 16 n^3 / 3 bytes whatever k is
    (n = 8760: 3.6 TB, 0.9 s, replicated on every rank).  One iteration here is G V on the FP64 tensor cores
    (2 n^2 b flop: 0.85 ms at n = 8760, b = 128) + sketch-sized float64 factors.  Convergence of pair j goes like
    (lam_{b+1} / lam_j)^iterations, i.e. fast exactly when the spectrum decays (every physical snapshot matrix); a flat
    spectrum (white noise) does not converge and falls back.  Accepted only when EVERY returned pair satisfies
    ||G v - lam v|| <= tol * lam_1, so the result is an eigen-decomposition to that residual whichever route ran."""
    n = G.shape[0]
    b = min(n, 128, -(-(k + 24) // 16) * 16)    # the Rayleigh-Ritz problem must fit the one-CTA Jacobi solver (<= 128)
    if b < min(n, k + 8):                        # too few guard vectors to converge the k-th pair
        return None
    from .rsvd import _orth

    V = _orth(ops, _subspace_start(ops, n, b), 1e-14, passes=2)
    hist: list[float] = []
    for it in range(SUBSPACE_MAX_ITERS):
        W = ops.sketch(G, V, None, PREC_NATIVE)                     # G V  (n x b, FP64 DMMA)
        # Rayleigh-Ritz in every iteration: the columns of W Q are then ~ lam_j v_j - nearly orthogonal, so that after
        # column scaling the block is well conditioned for CholeskyQR whatever lam_1 / lam_b is
        H = ops.gemm(V, W, transA=True)                             # b x b Rayleigh quotient
        lam, Q = ops.syevj(0.5 * (H + H.t()))
        V = ops.gemm(V, Q)
        W = ops.gemm(W, Q)
        if it >= 2:
            resid = float(torch.linalg.vector_norm(W[:, :k] - V[:, :k] * lam[:k], dim=0).max() / lam[0])   # one host read
            if resid <= tol:
                if stats is not None:
                    stats["eig_route"] = f"subspace iteration, {it + 1} iterations, residual {resid:.1e}"
                return lam[:k].contiguous(), V[:, :k].contiguous()
            hist.append(resid)
            if len(hist) >= 6:
                # measured contraction per iteration over the last four; give up as soon as it cannot reach tol within
                # the budget (dense or flat spectrum): the direct route is then cheaper than iterating on
                rate = (hist[-1] / hist[-5]) ** 0.25
                if rate >= 1.0 or it + math.log(tol / resid) / math.log(rate) > SUBSPACE_MAX_ITERS:
                    if stats is not None:
                        stats["eig_route_note"] = f"subspace iteration abandoned after {it + 1} iterations (contraction {rate:.2f})"
                    return None
        # shifted CholeskyQR3 while the Ritz vectors are still far from eigenvectors (columns of W strongly coupled)
        V = _orth(ops, W, 1e-14, shifted=it < 3, passes=1 if it >= 3 else 2)
    return None


def sym_eig_topk(ops, G: torch.Tensor, k: int, refine: bool = True, tol: float = 1e-13,
                 stats: dict | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """k largest eigenpairs of the symmetric float64 matrix G (n x n; not modified).
    Returns (lam (k,) descending, V (n, k) orthonormal columns)."""
    n = G.shape[0]
    k = min(int(k), n)
    if n <= JACOBI_MAX_N:
        lam, V = ops.syevj(G.clone())
        return lam[:k].contiguous(), V[:, :k].contiguous()
    if refine and n >= SUBSPACE_MIN_N and 8 * k <= n:
        out = sym_eig_topk_subspace(ops, G, k, tol, stats)
        if out is not None:
            return out
    if stats is not None:
        stats["eig_route"] = "Householder tridiagonalisation + bisection / inverse iteration"
    A = (0.5 * (G + G.t())).contiguous()                 # both triangles are read; destroyed by the reduction
    d, e, tau = ops.tridiag_reduce(A)
    kk = min(n, k + min(8, n - k))                        # a few guard vectors resolve a cluster cut by the k-th value
    W, Z = ops.tridiag_eig_topk(d, e, kk)
    Z = _orth2(ops, Z)
    if refine:
        # Rayleigh-Ritz in the tridiagonal basis: decouples vectors of close eigenvalues (inverse iteration leaves
        # them mixed) and returns Ritz values consistent with the orthonormalised vectors
        H = ops.gemm(Z, ops.tridiag_apply(d, e, Z), transA=True)
        lam, Wh = sym_eig_topk(ops, H, kk, refine=False)
        Z = ops.gemm(Z, Wh)
    else:
        lam = W
    V = ops.tridiag_backtransform(A, tau, Z[:, :k].contiguous())
    return lam[:k].contiguous(), V


def gram_device(ops, X: torch.Tensor, n: int, delay: int, precision: int) -> torch.Tensor:
    """G = sum_j X_j^T X_j over the delay windows X_j = X[:, j : j + n] (float64, this rank's rows only).

    The windows overlap: (X_j^T X_j)[a, b] = F[j + a, j + b] with F = X^T X the Gram matrix of the BASE matrix (T x T,
    T = n + delay - 1), so F is formed ONCE - on aligned column blocks whatever the delay - and G is the sum of its
    ``delay`` shifted diagonal blocks: one pass set over X instead of ``delay``, and no operand ever starts at an
    unaligned column (the Y operand of the tensor-core kernels needs 16-byte alignment).

    precision PREC_TF32MIX: ONE tf32 product per k-step on the raw float32 tiles (era5svd_project_tf32x1, the HBM-bound
    kernel of the early power iterations) - for float32 data the Gram matrix only supplies the SUBSPACE that the
    full-precision refinement passes start from (standard_svd_device), exactly like the single-product power iterations
    of the randomized driver; the column block itself (a view of X with X's pitch) is the Y operand: no split images."""
    T = n + delay - 1
    X = X[:, :T]
    if precision in (PREC_TF32MIX, PREC_TF32X3):
        if X.dtype != torch.float32:
            raise TypeError("precision 'tf32mix' / 'tf32x3' needs a float32 snapshot matrix")
        # symmetric: block column [c0, c1) is computed for the time rows t >= c0 only (the column window X[:, c0:]
        # is the X operand) and mirrored, which halves the passes over X
        F = ops.zeros((T, T), torch.float64)
        if precision == PREC_TF32MIX:
            for c0 in range(0, T, TC_BLOCK_X1):
                c1 = min(T, c0 + TC_BLOCK_X1)
                ops.project_tf32x1(X[:, c0:], X[:, c0:c1], F[c0:, c0:c1])     # F[c0:, c0:c1], t >= c0 only
        else:
            for c0 in range(0, T, TC_BLOCK):
                c1 = min(T, c0 + TC_BLOCK)
                hi, lo = ops.split_tf32(X[:, c0:c1])              # the (Y_hi, Y_lo) operand of the project kernel
                F[c0:, c0:c1] = ops.project_tf32x3(X[:, c0:], None, hi, lo)   # (T - c0, c1 - c0) float64
        low = torch.tril(F, -1)
        F = torch.tril(F) + low.t()                               # the diagonal blocks' upper parts come from the mirror
    else:
        F = ops.project(X, X, None, precision=PREC_NATIVE)
    if delay == 1:
        return F
    G = F[:n, :n].clone()
    for j in range(1, delay):
        G += F[j : j + n, j : j + n]
    return G


def standard_svd_device(ops, X: torch.Tensor, n_components: int, *, delay: int = 1, comm=None,
                        precision: int = PREC_NATIVE, stats: dict | None = None):
    """X: this rank's base rows (m0_local, T).  Returns (U_local (m0_local*delay, k), s (k,), Vt (k, n))."""
    comm = comm or LocalComm()
    m0, T = X.shape
    d = delay
    n = T - d + 1
    if n < 1:
        raise ValueError("window shape cannot be larger than input array shape")   # numpy's text in the reference (sliding_window_view, slice_tools.py:207)
    k = min(int(n_components), n)
    tc_ok = precision in (PREC_TF32X3, PREC_TF32MIX) and X.dtype == torch.float32
    use_tc = tc_ok and k <= 128
    mixed = tc_ok and precision == PREC_TF32MIX and min(n, k + 10) <= 128
    G = gram_device(ops, X, n, d, (PREC_TF32MIX if mixed else PREC_TF32X3) if tc_ok else PREC_NATIVE)
    comm.allreduce_sum_(G)
    if X.dtype == torch.float32:
        # float32 data: the Gram route supplies the subspace, the randomized driver's final stage the values
        from .rsvd import randomized_svd_device

        kk = min(n, k + 10)
        _, V = sym_eig_topk(ops, G, kk, tol=1e-9, stats=stats)       # the Gram matrix itself is only ~1e-6 accurate
        prec = PREC_TF32X3 if (tc_ok and kk <= 128) else PREC_NATIVE
        if mixed:
            # every refinement iteration is a full-precision one (the tf32 Gram matrix already holds what single-product
            # iterations could deliver); their projections take Y truncated (two products, rsvd.py)
            return randomized_svd_device(ops, X, k, V, n_iter=REFINE_ITERS, delay=d, precision=PREC_TF32MIX, comm=comm,
                                         full_iters=REFINE_ITERS)
        return randomized_svd_device(ops, X, k, V, n_iter=REFINE_ITERS, delay=d, precision=prec, comm=comm)
    lam, V = sym_eig_topk(ops, G, k, stats=stats)
    s, inv_s = ops.sigma_from_eig(lam)
    Vk = V.t().contiguous()                     # (k, n): rows = right singular vectors, descending
    M = Vk.clone()
    ops.scale_rows(M, inv_s.contiguous())       # S_k^-1 V_k^T
    Mt = M.t().contiguous()                     # (n, k) = V_k S_k^-1
    if use_tc:
        U = ops.empty((m0 * d, ops.tf32_ldy(k)), X.dtype)[:, :k]
        for j in range(d):
            ops.sketch_tf32x3(X[:, j : j + n], None, Mt, U[j * m0 : (j + 1) * m0], None, None)
    else:
        Mt = ops.convert(Mt, X.dtype)
        U = ops.empty((m0 * d, k), X.dtype)
        for j in range(d):
            ops.sketch(X[:, j : j + n], Mt, U[j * m0 : (j + 1) * m0], PREC_NATIVE)
    return U, s, Vk
