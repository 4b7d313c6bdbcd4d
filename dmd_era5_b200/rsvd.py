"""Randomized SVD driver on the device - the B200 schedule of sklearn's ``_randomized_svd``.

Reference control flow being replaced (scikit-learn 1.9.0, called from
src/dmd_era5/era5_svd/era5_svd.py:258 with all defaults):
    _randomized_range_finder  sklearn/utils/extmath.py:313-385
        Q = normal(n x l)                                   :323   (host RandomState, l = k + 10)
        n_iter x { Q = lu(A @ Q); Q = lu(A.T @ Q) }         :377-379
        Q = qr(A @ Q)                                       :383
    _randomized_svd           extmath.py:560-633
        B = Q.T @ M; Uhat, s, Vt = svd(B); U = Q @ Uhat     :606-619
        svd_flip (u-based); truncate to k                   :623, :633

In exact arithmetic U, s, Vt depend only on span((A^T A)^q Omega_0); the LU / QR normalisers are
there for floating-point stability only.  The device schedule therefore keeps the
same Omega_0, q and flip rule but never factorises a TALL matrix inside the power iterations:

    Omega = orth(Omega_0)                                   (n x l, float64, CholeskyQR2)
    q x { Y = X Omega            [tall pass, not normalised; Omega held in tf32-representable values on the
                                  tensor-core path: 2 instead of 3 products per k-step, see below]
          Z = X^T Y              [tall pass -> n x l float64, all-reduce over row shards]
          last iteration only: T = Omega^T Z = Y^T Y = W L W^T (l x l Jacobi), Z <- Z W   (Rayleigh-Ritz)
          Omega = orth(Z)        }  (CholeskyQR2; shifted CholeskyQR3 after the random start)
    Y = X Omega; Z' = X^T Y; G = Y^T Y                      (last tall passes)
    R = chol(G); B = R^-T Z'^T (= Q^T X with Q = Y R^-1)    (l x n)
    eig(B B^T) -> Uhat, s;  Vt = S^-1 Uhat^T B;  U = Y (R^-1 Uhat)
    svd_flip via a MAXLOC reduction over row shards; truncate to k.

The Rayleigh-Ritz rotation makes the columns of Y = X Omega nearly orthogonal with norms ~ sigma_j,
so (i) storing Y in float32 keeps every direction at full relative precision and (ii) the Gram
matrices handed to Cholesky are well conditioned after diagonal scaling.
All small factors are float64 and replicated on every rank.
"""
from __future__ import annotations

import numpy as np
import torch

from contextlib import contextmanager

from ._cabi import PREC_NATIVE, PREC_TF32X3
from .dist import LocalComm


@contextmanager
def nvtx_range(name: str):
    """NVTX range around a phase of the device schedule (visible in nsys / ncu timelines); no-op on CPU tensors."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()

# driver-level precision: the tensor-core path with the EARLY power iterations in single-product TF32 (the C ABI only
# knows PREC_NATIVE / PREC_TF32X3; this value never crosses it)
PREC_TF32MIX = 2
PRECISIONS = {"native": PREC_NATIVE, "fp64": PREC_NATIVE, "fp32": PREC_NATIVE, "tf32x3": PREC_TF32X3,
              "tf32mix": PREC_TF32MIX}
# power iterations (counted from the end) that keep fp32-level 3xTF32 products under "tf32mix"
MIX_FULL_ITERS = 1


def n_iter_auto(m: int, n: int, k: int) -> int:
    """extmath.py:586-589: 7 power iterations if k < 0.1 * min(M.shape) else 4."""
    return 7 if k < 0.1 * min(m, n) else 4


def draw_omega(n_features: int, k: int, seed: int | None, dtype: torch.dtype, n_oversamples: int = 10) -> np.ndarray:
    """The reference's test matrix, drawn on the host exactly as extmath.py:323-334 does:
    ``random_state.normal(size=(n, k + 10))`` from NumPy's RandomState (the global one when the
    reference's call is unseeded, era5_svd.py:258), cast to float32 when X is float32.
    Returned as float64 (the cast is value-preserving)."""
    rs = np.random.mtrand._rand if seed is None else np.random.RandomState(seed)
    om = rs.normal(size=(n_features, k + n_oversamples))
    if dtype == torch.float32:
        om = om.astype(np.float32)
    return np.ascontiguousarray(om, dtype=np.float64)


class _Blocks:
    """The (virtual) delay-embedded operator: block j is the view X[:, j : j + n]
    (src/dmd_era5/slice_tools/slice_tools.py:207-211, never materialised)."""

    def __init__(self, X: torch.Tensor, d: int):
        self.X = X
        self.d = d
        self.m0 = X.shape[0]
        self.n = X.shape[1] - d + 1
        if self.n < 1:
            raise ValueError("window shape cannot be larger than input array shape")   # numpy's text in the reference (sliding_window_view, slice_tools.py:207)

    def view(self, j: int) -> torch.Tensor:
        return self.X[:, j : j + self.n]


def _orth(ops, P: torch.Tensor, rel_tol: float, shifted: bool = False, passes: int = 1) -> torch.Tensor:
    """Well-conditioned basis of span(P) for a small (n x l) float64 matrix: column scaling followed by
    ``passes`` CholeskyQR sweeps (Gram -> Cholesky -> triangular inverse -> GEMM).  The power
    iteration only needs a WELL-CONDITIONED basis, not an orthonormal one (T = Omega^T Z = Y^T Y and the
    final Q = Y R^-1 hold for any basis), so one sweep (orthonormal to ~cond(P)^2 u) is enough.
    ``shifted`` prepends one shifted CholeskyQR pass (Fukaya et al., shifted CholeskyQR3): with the
    shift s = 11 (n l + l (l + 1)) u ||P||^2 the first Cholesky cannot break down, so the procedure is
    stable for cond(P) up to ~1/u; used after the random start where cond(Z) ~ kappa(X)^2."""
    ops.col_normalize(P)
    n, l = P.shape
    if shifted:
        shift = 11.0 * (n * l + l * (l + 1)) * 1.1e-16 * l      # ||P||_2^2 <= ||P||_F^2 = l
        G = shift * torch.eye(l, dtype=torch.float64, device=P.device)
        G = ops.gemm(P, P, transA=True, beta=1.0, C=G)
        _, Rinv = ops.chol_inv(G, rel_tol)
        P = ops.gemm(P, Rinv)
    for _ in range(passes):
        G = ops.gemm(P, P, transA=True)
        _, Rinv = ops.chol_inv(G, rel_tol)
        P = ops.gemm(P, Rinv)
    return P


def randomized_svd_device(ops, X: torch.Tensor | None, n_components: int, omega0, *, n_iter: int | None = None,
                          delay: int = 1, precision: int = PREC_NATIVE, comm=None, row_offset: int = 0,
                          m0_global: int | None = None, m_global: int | None = None, stats: dict | None = None,
                          split: tuple[torch.Tensor, torch.Tensor] | None = None, full_iters: int | None = None):
    """Randomized SVD of the (row-sharded, optionally delay-embedded) snapshot matrix.

    X          : this rank's rows of the BASE matrix, (m0_local, T) tall-dtype device tensor
    omega0     : (n, l) test matrix (host ndarray or tensor), n = T - delay + 1
    row_offset : global index of this rank's first base row; m0_global: total base rows
    split      : (Xhi, Xlo) pre-split tf32 images of X for precision tf32x3 (X may then be None)
    precision  : PREC_NATIVE | PREC_TF32X3 | PREC_TF32MIX.  "tf32mix": subspace iteration is self-correcting - an error
                 made in iteration i is contracted by (sigma_{l+1} / sigma_j)^2 in every later one - so only the last
                 ``full_iters`` (default 1) power iterations and the final range / projection passes, which fix sigma and
                 the vectors, run with fp32-level 3xTF32 products; the earlier ones use ONE tf32 product per k-step on the
                 raw float32 tiles (era5svd_sketch_tf32x1 / era5svd_project_tf32x1: pure TMA -> tcgen05, HBM bound); the
                 projection of the full-precision iteration(s) takes Y truncated (era5svd_project_tf32x2, two products).
    Returns (U_local (m0_local * delay, k) tall dtype, s (k,) float64, Vt (k, n) float64).
    Rows of U_local are ordered block-major: block j holds rows [j * m0_local, (j + 1) * m0_local),
    i.e. global rows j * m0_global + row_offset + r.
    """
    comm = comm or LocalComm()
    mixed = precision == PREC_TF32MIX and split is None
    if precision == PREC_TF32MIX:
        precision = PREC_TF32X3
    if X is None:
        if split is None or precision != PREC_TF32X3:
            raise ValueError("X may only be omitted when pre-split images are given with precision tf32x3")
    blocks = _Blocks(X if X is not None else split[0], delay)
    m0, n, d = blocks.m0, blocks.n, delay
    m0_global = m0 if m0_global is None else m0_global
    m_global = m0_global * d if m_global is None else m_global
    k = int(n_components)
    om = torch.as_tensor(np.asarray(omega0) if not torch.is_tensor(omega0) else omega0, dtype=torch.float64)
    if om.shape[0] != n:
        raise ValueError(f"omega0 must have {n} rows, got {tuple(om.shape)}")
    l = min(om.shape[1], n)  # a sketch wider than the row space adds nothing (exact SVD either way)
    om = om[:, :l].contiguous()
    if n_iter is None:
        n_iter = n_iter_auto(m_global, n, k)
    tall = blocks.X.dtype
    rel_tol = 1e-13 if tall == torch.float64 else 1e-6
    # Jacobi stopping thresholds (relative off-diagonal size): full precision for float64 data; for
    # float32 data the factors carry ~1e-7 anyway, so 1e-10 changes sigma by < 1e-12 relative
    # (the Rayleigh-Ritz rotation is a choice of basis, not a result: it only has to leave the columns of Y graded and
    # the scaled Gram matrix well conditioned, for which |cos(y_p, y_q)| <= 1e-4 is ample)
    rr_tol = 0.0 if tall == torch.float64 else 1e-4
    eig_tol = 0.0 if tall == torch.float64 else 1e-10
    # A Gaussian n x l block is already well conditioned (cond ~ (1 + sqrt(l/n)) / (1 - sqrt(l/n))): like the
    # reference, Omega_0 is used as drawn (column-normalised); when l is close to n it is orthonormalised.
    Omega = ops.to_device(om, non_blocking=False).clone()
    if 4 * l > n:
        Omega = _orth(ops, Omega, 1e-13)
    else:
        ops.col_normalize(Omega)

    use_tc = precision == PREC_TF32X3
    # On-chip split path: the basis Omega is ours to choose (any well-conditioned basis of the same span serves), so
    # it is kept in tf32-representable values - every formula below uses the SAME rounded Omega, so nothing is
    # approximated - and the sketch needs two tensor-core products per k-step instead of three (Omega_lo = 0).
    om_tf32 = use_tc and split is None
    if om_tf32:
        ops.round_tf32_(Omega)
    if use_tc:
        # tensor-core path: operands pre-split into tf32 hi / lo images (see csrc/gemm_tc.cu)
        if tall != torch.float32:
            raise TypeError("precision 'tf32x3' needs a float32 snapshot matrix")
        if l > 128:
            raise ValueError(f"precision 'tf32x3' supports sketch widths up to 128, got l = {l}")
        # pre-split images if the caller has them (gemm_tc.cu), else the plain matrix, split on chip
        # (gemm_tc2.cu: X crosses HBM once per pass)
        Xhi, Xlo = split if split is not None else (X, None)
        ldy = ops.tf32_ldy(l)
        Y = ops.empty((m0 * d, ldy), tall)[:, :l]
        Yhi = ops.empty((m0 * d, ldy), tall)[:, :l]
        Ylo = ops.empty((m0 * d, ldy), tall)[:, :l]
    else:
        Y = ops.empty((m0 * d, l), tall)

    def tall_pass(Omega64: torch.Tensor, keep_y: bool = False, final: bool = False, low: bool = False,
                  ytrunc: bool = False) -> torch.Tensor:
        """Y = X_d Omega (kept in the preallocated buffers), returns Z = X_d^T Y (all-reduced).
        On the on-chip-split path the power iterations keep Y as ONE plain float32 image (split again on chip by the
        projection); only the final pass writes the hi / lo pair that the Gram and U = Y M kernels consume."""
        Z = None
        fused = False

        def last_block(j: int) -> None:
            # the projection of the last delay block finishes with the all-reduce over the row shards inside the kernel
            # that sums its partial tiles (PeerComm: era5svd_comm_fuse_next_project); other communicators return False
            nonlocal fused
            if j == d - 1:
                fused = comm.fuse_next_project(n, l)

        if use_tc:
            for j in range(d):
                rows = slice(j * m0, (j + 1) * m0)
                xh, xl = Xhi[:, j : j + n], (Xlo[:, j : j + n] if Xlo is not None else None)
                if low:
                    ops.sketch_tf32x1(xh, Omega64, Y[rows])
                    last_block(j)
                    Z = ops.project_tf32x1(xh, Y[rows], Z, accumulate=j > 0)
                    continue
                if xl is None and not final:
                    ops.sketch_tf32x3(xh, None, Omega64, Y[rows], None, None, om_tf32=om_tf32)
                    last_block(j)
                    if ytrunc:      # X exact (split on chip), Y taken as tf32(Y): two products, error filtered by X^T
                        Z = ops.project_tf32x2(xh, Y[rows], Z, accumulate=j > 0)
                    else:
                        Z = ops.project_tf32x3(xh, None, Y[rows], None, Z, accumulate=j > 0)
                    continue
                ops.sketch_tf32x3(xh, xl, Omega64, Y[rows] if keep_y else None, Yhi[rows], Ylo[rows], om_tf32=om_tf32)
                last_block(j)
                Z = ops.project_tf32x3(xh, xl, Yhi[rows], Ylo[rows], Z, accumulate=j > 0)
        else:
            Om_t = ops.convert(Omega64, tall)
            for j in range(d):
                Yj = Y[j * m0 : (j + 1) * m0]
                ops.sketch(blocks.view(j), Om_t, Yj, PREC_NATIVE)
                last_block(j)
                Z = ops.project(blocks.view(j), Yj, Z, accumulate=j > 0, precision=PREC_NATIVE)
        if not fused:
            comm.allreduce_sum_(Z)
        if stats is not None:
            stats["tall_passes"] = stats.get("tall_passes", 0) + 2
        return Z

    n_low = max(0, n_iter - (MIX_FULL_ITERS if full_iters is None else int(full_iters))) if mixed else 0
    if stats is not None:
        stats["low_precision_iters"] = n_low
    for it in range(n_iter):
        # "tf32mix": the full-precision power iterations still take Y truncated to tf32 in their PROJECTION (two products):
        # Z = X^T (Y + dY) = X^T Y + X^T dY - an error in Y is filtered by X^T exactly like the iteration filters the
        # sketch; measured / emulated effect on sigma and the vectors: none (tests/test_driver_cpu.py)
        with nvtx_range(f"era5svd.power_iteration[{it}]" + (".tf32x1" if it < n_low else "")):
            Z = tall_pass(Omega, low=it < n_low, ytrunc=mixed)
        if it == n_iter - 1:
            # Rayleigh-Ritz rotation before the final pass: T = Omega^T Z = Y^T Y (l x l) = W L W^T, so
            # the columns of X (Z W) come out nearly orthogonal and graded (~ sigma_j u_j) and the Gram
            # matrix of the stored float32 Y is well conditioned after diagonal scaling.
            T = ops.gemm(Omega, Z, transA=True)
            _, W = ops.syevj(T, tol=rr_tol)                # only has to decouple the columns
            Z = ops.gemm(Z, W)
        # cond(Z) ~ kappa(X)^2 after the random start (shifted CholeskyQR3), <~ kappa(X) afterwards.
        # Pivot threshold: 1e-13 while the columns are still coupled (a legitimate trailing direction of the random
        # start has a relative pivot ~ (sigma_l / sigma_1)^4).  After the Rayleigh-Ritz rotation the legitimate columns
        # are decoupled (pivot ~ 1), so for float32 data 1e-6 separates them from directions that only exist as
        # rounding noise of the tall passes: a sketch wider than rank(X) leaves columns Z w = X^T (noise), which lie
        # in the row space already spanned (measured pivots 1e-13 ... 1e-8: the fp32 accumulate of such a cancelling
        # column is good to ~1e-4 of its own norm) and would otherwise survive as spurious singular values
        # ~1e-5 sigma_1 (the single-product early iterations never drop them: truncated X has full rank).  A genuine
        # direction only falls below 1e-6 when sigma_j <~ 1e-5 sigma_1, under the noise floor of float32 passes.
        last = it == n_iter - 1
        Omega = _orth(ops, Z, 1e-6 if (last and tall == torch.float32) else 1e-13, shifted=(it == 0 and n_iter > 1))
        if om_tf32:
            ops.round_tf32_(Omega)

    with nvtx_range("era5svd.final_range_and_projection"):
        Zp = tall_pass(Omega, keep_y=not use_tc, final=True)           # n x l
    # l x l Gram matrix of the STORED (rounded) Y, so that Q = Y R^-1 is orthonormal for the Y we keep
    fused_g = comm.fuse_next_project(l, l)
    G = ops.project_tf32x3(Yhi, Ylo, Yhi, Ylo) if use_tc else ops.project(Y, Y, precision=PREC_NATIVE)
    if not fused_g:
        comm.allreduce_sum_(G)
    _, Rinv = ops.chol_inv(G, rel_tol)
    B = ops.gemm(Rinv, Zp, transA=True, transB=True)   # l x n  = R^-T Z'^T = Q^T X
    BBt = ops.gemm(B, B, transB=True)                  # l x l
    lam, Uh = ops.syevj(BBt, tol=eig_tol)
    s, inv_s = ops.sigma_from_eig(lam)
    Vt = ops.gemm(Uh, B, transA=True)                  # l x n
    ops.scale_rows(Vt, inv_s)
    kk = min(k, l)
    M = ops.gemm(Rinv, Uh[:, :kk])                     # l x k
    if use_tc:
        U = ops.empty((m0 * d, ops.tf32_ldy(kk)), tall)[:, :kk]
        ops.sketch_tf32x3(Yhi, Ylo, M, U, None, None)            # (m0 * d) x k, Y = Yhi + Ylo exactly
    else:
        U = ops.sketch(Y, ops.convert(M, tall), None, PREC_NATIVE)

    # svd_flip (extmath.py:964-972): first row of max |U[:, j]| over ALL rows decides the sign
    cands = [ops.col_absmax(U[j * m0 : (j + 1) * m0], j * m0_global + row_offset) for j in range(d)]
    a = torch.stack([c[0] for c in cands]); r = torch.stack([c[1] for c in cands]); sg = torch.stack([c[2] for c in cands])
    if comm.world > 1:
        # ONE all-gather for the three candidate arrays (value, global row, sign): row indices < 2^53 are exact in float64
        packed = comm.allgather(torch.stack([a, r.to(torch.float64), sg])).permute(1, 0, 2, 3).reshape(3, -1, kk)
        a, r, sg = packed[0].contiguous(), packed[1].to(torch.int64).contiguous(), packed[2].contiguous()
    sign = ops.maxloc_combine(a, r, sg)
    ops.scale_cols(U, sign)
    Vk = Vt[:kk]
    ops.scale_rows(Vk, sign)
    return U, s[:kk], Vk
