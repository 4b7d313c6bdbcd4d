"""GPU: every C-ABI kernel against a float64 NumPy statement of the same operation."""
import numpy as np
import pytest
import torch

from dmd_era5_b200._cabi import BUILD_CHECK_FINITE, BUILD_MEAN_CENTER, BUILD_SCALE, PREC_NATIVE
from oracle.slice_tools_np import standardize_np

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
@pytest.mark.parametrize("flags", [0, BUILD_MEAN_CENTER, BUILD_MEAN_CENTER | BUILD_SCALE])
def test_build_rows(ops, dtype, tol, flags):
    rng = np.random.RandomState(0)
    T, P, ld_src = 37, 301, 320
    src = (rng.rand(T, ld_src) * 30 + 250).astype(dtype)
    d_src = dev(src)[:, :P]
    ldx = 40
    buf = torch.zeros((P, ldx), dtype=d_src.dtype, device="cuda")
    X = buf[:, :T]
    mean = torch.zeros(P, dtype=d_src.dtype, device="cuda")
    std = torch.zeros(P, dtype=d_src.dtype, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.build_rows(d_src, X, mean if flags & 1 else None, std if flags & 2 else None, None,
                   flags | BUILD_CHECK_FINITE, flag)
    a = src[:, :P]
    if flags & 1:
        ref, mu, sd = standardize_np(a, scale=bool(flags & 2))
        assert np.allclose(mean.cpu().numpy(), mu, rtol=tol, atol=0)
        if flags & 2:
            assert np.allclose(std.cpu().numpy(), sd, rtol=10 * tol, atol=0)
    else:
        ref = a
    scale = 1.0 if flags & 2 else 300.0
    assert np.max(np.abs(X.cpu().numpy() - ref.T)) <= 20 * tol * scale
    assert int(flag.item()) == 0
    assert float(buf[:, T:].abs().max()) == 0.0       # padding untouched


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
@pytest.mark.parametrize("T,P,ld_src", [(744, 1000, 1000), (1460, 333, 336), (2920, 200, 200), (5000, 64, 64), (744, 130, 131),
                                        (8760, 96, 96), (3600, 40, 40)])
def test_build_rows_long_series(ops, dtype, tol, T, P, ld_src):
    """Long time axes: 16-point TMA tiles (float32, T up to ~3500), 32-point tiles split along time over a cluster of 4 / 8
    CTAs with the statistics combined through distributed shared memory (T = 3600, 5000, 8760), the register-staged tile
    kernel and the two-kernel path (float64), aligned and misaligned source pitches, NaN skipping, scaling."""
    rng = np.random.RandomState(T + P)
    src = (rng.rand(T, ld_src) * 30 + 250 + 5 * np.sin(np.arange(T))[:, None]).astype(dtype)
    src[T // 3, 7] = np.nan                      # skipped by mean / std like xarray does
    d_src = dev(src)[:, :P]
    X = torch.zeros((P, T + (-T) % 8), dtype=d_src.dtype, device="cuda")[:, :T]
    mean = torch.zeros(P, dtype=d_src.dtype, device="cuda")
    std = torch.zeros(P, dtype=d_src.dtype, device="cuda")
    ops.build_rows(d_src, X, mean, std, None, BUILD_MEAN_CENTER | BUILD_SCALE, None)
    # float64 statement of the same operations on the same (float32 or float64) data: the kernel accumulates the
    # statistics in float64 and rounds once, so it must agree with this to the rounding of the matrix dtype ...
    ref, mu, sd = standardize_np(src[:, :P].astype(np.float64), scale=True)
    assert np.allclose(mean.cpu().numpy(), mu, rtol=tol / 10, atol=0)
    assert np.allclose(std.cpu().numpy(), sd, rtol=10 * tol, atol=0)
    got = X.cpu().numpy()
    ok = ~np.isnan(ref.T)
    assert np.array_equal(np.isnan(got), ~ok)
    assert np.max(np.abs(got[ok] - ref.T[ok])) <= (1e-12 if dtype == np.float64 else 5e-6)
    # ... and with the reference-dtype oracle (NumPy sums float32 means row after row, so its own mean drifts ~T * eps)
    ref32, mu32, sd32 = standardize_np(src[:, :P], scale=True)
    assert np.allclose(mean.cpu().numpy(), mu32, rtol=max(tol, 2e-9 * T), atol=0)
    assert np.max(np.abs(got[ok] - ref32.T[ok])) <= (20 * tol if dtype == np.float64 else max(1e-4, 1e-7 * T))


@pytest.mark.parametrize("no_tma", [0, 8])
@pytest.mark.parametrize("T,P,off,xdt", [(744, 70, 1, torch.float32), (100, 1000, 3, torch.float64), (64, 33, 0, torch.float32),
                                         (1460, 97, 2, torch.float32), (1700, 40, 0, torch.float32), (9, 5, 1, torch.float32),
                                         (70, 40000, 1, torch.float32), (300, 33000, 0, torch.float32)])
def test_build_rows_tma_views(ops, T, P, off, xdt, no_tma):
    """float32 sources take the TMA kernel (whole [T x 32] tile in flight, swizzled smem tile); column-offset views
    (base not 16-byte aligned -> shifted tensor map), ragged point tails, more tiles than resident CTAs (persistent loop,
    box-by-box refill), the float64 cast, weights, no centring with
    the finiteness flag, and the same calls through the register-staged kernel (flag 8 = ERA5SVD_BUILD_NO_TMA)."""
    rng = np.random.RandomState(T * 7 + P)
    ld_src = ((off + P + 3) // 4) * 4 + 4
    src = (rng.rand(T, ld_src) * 30 + 250 + 3 * np.cos(np.arange(T) / 5.0)[:, None]).astype(np.float32)
    src[T // 2, off + 3] = np.nan
    d_src = dev(src)[:, off:off + P]
    w = (rng.rand(P) + 0.5).astype(np.float32 if xdt == torch.float32 else np.float64)
    a = src[:, off:off + P].astype(np.float64)
    ref, mu, sd = standardize_np(a, scale=True)
    ok = ~np.isnan(ref.T)
    tol = 5e-6 if xdt == torch.float32 else 1e-12
    for flags, wt in [(BUILD_MEAN_CENTER | BUILD_SCALE, None), (BUILD_MEAN_CENTER, w), (0, None)]:
        X = torch.full((P, T + (-T) % 8 + 8), -7.0, dtype=xdt, device="cuda")
        mean = torch.zeros(P, dtype=xdt, device="cuda"); std = torch.zeros(P, dtype=xdt, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.build_rows(d_src, X[:, :T], mean if flags & 1 else None, std if flags & 2 else None,
                       dev(wt) if wt is not None else None, flags | BUILD_CHECK_FINITE | no_tma, flag)
        got = X[:, :T].cpu().numpy().astype(np.float64)
        assert float((X[:, T:] + 7.0).abs().max()) == 0.0                   # padding untouched
        assert int(flag.item()) == 1                                        # the NaN is reported
        assert np.array_equal(np.isnan(got), ~ok)
        if flags == 0:
            assert np.array_equal(got[ok], a.T[ok])                         # plain transpose (+ exact cast)
            continue
        assert np.allclose(mean.cpu().numpy(), mu, rtol=1e-7 if xdt == torch.float32 else 1e-14, atol=0)
        if flags & 2:
            assert np.allclose(std.cpu().numpy(), sd, rtol=2e-5, atol=0)
            assert np.max(np.abs(got[ok] - ref.T[ok])) <= (5e-5 if xdt == torch.float32 else 1e-5)
        else:
            want = (a - mu).T * wt[:, None].astype(np.float64)
            # float32 matrix: the mean is rounded to float32 first (it is what mean_out stores): 300 K * 6e-8
            assert np.max(np.abs(got[ok] - want[ok])) <= (6e-5 if xdt == torch.float32 else 1e-10)


def test_build_rows_split_tma(ops):
    """hi / lo images from the TMA kernel: Xhi + Xlo == X exactly, Xhi is a tf32 value."""
    rng = np.random.RandomState(5)
    T, P = 200, 75
    src = (rng.rand(T, 76) * 30 + 250).astype(np.float32)
    d_src = dev(src)[:, :P]
    X, Xhi, Xlo = (torch.zeros((P, 200), dtype=torch.float32, device="cuda") for _ in range(3))
    mean = torch.zeros(P, dtype=torch.float32, device="cuda")
    ops.build_rows_split(d_src, X, Xhi, Xlo, mean, None, None, BUILD_MEAN_CENTER, None)
    assert torch.equal(Xhi + Xlo, X)
    assert int((Xhi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    ref, mu, _ = standardize_np(src[:, :P].astype(np.float64), scale=False)
    assert np.max(np.abs(X.cpu().numpy() - ref.T)) <= 6e-5


def test_build_rows_nan_and_flag(ops):
    src = np.random.RandomState(1).rand(16, 64)
    src[3, 5] = np.nan
    src[7, 9] = np.inf
    X = torch.zeros((64, 16), dtype=torch.float64, device="cuda")
    mean = torch.zeros(64, dtype=torch.float64, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.build_rows(dev(src), X, mean, None, None, BUILD_MEAN_CENTER | BUILD_CHECK_FINITE, flag)
    assert int(flag.item()) == 1
    assert np.isclose(mean[5].item(), np.nanmean(src[:, 5]))   # NaN skipped like xarray's mean


def test_build_rows_weights_and_cast(ops):
    rng = np.random.RandomState(2)
    src = rng.rand(20, 50)
    w = rng.rand(50).astype(np.float32)
    X = torch.zeros((50, 24), dtype=torch.float32, device="cuda")[:, :20]
    mean = torch.zeros(50, dtype=torch.float32, device="cuda")
    ops.build_rows(dev(src), X, mean, None, dev(w), BUILD_MEAN_CENTER, None)
    ref = ((src - src.mean(0)).T * w[:, None])
    assert np.allclose(X.cpu().numpy(), ref, atol=1e-6)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 1e-5)])
@pytest.mark.parametrize("m,n,l", [(1000, 96, 22), (257, 33, 110), (130, 200, 130), (5, 3, 2)])
def test_sketch_and_project_native(ops, dtype, tol, m, n, l):
    rng = np.random.RandomState(m + n + l)
    ldx = n + 3
    Xh = rng.standard_normal((m, ldx)).astype(dtype)
    Om = rng.standard_normal((n, l)).astype(dtype)
    X = dev(Xh)[:, :n]
    Y = ops.sketch(X, dev(Om), None, PREC_NATIVE)
    Yref = Xh[:, :n].astype(np.float64) @ Om.astype(np.float64)
    assert np.max(np.abs(Y.cpu().numpy() - Yref)) <= tol * np.sqrt(n) * 10
    Z = ops.project(X, Y, None, False, PREC_NATIVE)
    Zref = Xh[:, :n].astype(np.float64).T @ Y.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(Z.cpu().numpy() - Zref)) <= tol * np.abs(Zref).max() * 20
    Z2 = ops.project(X, Y, Z.clone(), True, PREC_NATIVE)
    assert np.allclose(Z2.cpu().numpy(), 2 * Z.cpu().numpy(), rtol=1e-12)


def test_project_many_splits_deterministic(ops):
    rng = np.random.RandomState(5)
    X = dev(rng.standard_normal((20000, 70)))
    Y = dev(rng.standard_normal((20000, 30)))
    Z1 = ops.project(X, Y).cpu().numpy()
    Z2 = ops.project(X, Y).cpu().numpy()
    assert np.array_equal(Z1, Z2)
    ref = X.cpu().numpy().T @ Y.cpu().numpy()
    assert np.max(np.abs(Z1 - ref)) < 1e-10


def test_delay_block_views(ops):
    rng = np.random.RandomState(6)
    Xh = rng.standard_normal((300, 41))
    X = dev(Xh)
    Om = rng.standard_normal((38, 9))
    for j in range(4):
        Y = ops.sketch(X[:, j : j + 38], dev(Om))
        assert np.allclose(Y.cpu().numpy(), Xh[:, j : j + 38] @ Om, atol=1e-12)


@pytest.mark.parametrize("tA,tB", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_f64(ops, tA, tB):
    rng = np.random.RandomState(7)
    M, N, K = 110, 75, 744
    A = rng.standard_normal((K, M) if tA else (M, K))
    B = rng.standard_normal((N, K) if tB else (K, N))
    C = ops.gemm(dev(A), dev(B), transA=tA, transB=tB)
    ref = (A.T if tA else A) @ (B.T if tB else B)
    assert np.max(np.abs(C.cpu().numpy() - ref)) < 1e-11
    C0 = rng.standard_normal((M, N))
    C2 = ops.gemm(dev(A), dev(B), transA=tA, transB=tB, alpha=0.5, beta=2.0, C=dev(C0))
    assert np.max(np.abs(C2.cpu().numpy() - (0.5 * ref + 2.0 * C0))) < 1e-11


@pytest.mark.parametrize("n", [1, 2, 7, 24, 110, 111, 150])
def test_syevj(ops, n):
    rng = np.random.RandomState(n)
    Q = np.linalg.qr(rng.standard_normal((n, n)))[0]
    w = np.sort(10.0 ** rng.uniform(-6, 2, size=n))[::-1]
    A = (Q * w) @ Q.T
    W, V = ops.syevj(dev(A))
    W, V = W.cpu().numpy(), V.cpu().numpy()
    assert np.all(np.diff(W) <= 0)
    assert np.max(np.abs(W - w) / w) < 1e-9 * max(1.0, w[0] / w[-1] * 1e-6) or np.max(np.abs(W - w)) < 1e-13 * w[0] * n
    assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-13 * n
    assert np.max(np.abs(A @ V - V * W)) < 1e-12 * w[0] * n


@pytest.mark.parametrize("n", [3, 25, 110, 128])
def test_syevj_fast_path_and_fallback(ops, n, monkeypatch):
    """n <= 128: Cholesky + one-sided Jacobi on the factor (positive definite input); a singular or indefinite matrix
    drops a pivot and the two-sided kernel takes over.  Both paths against numpy.linalg.eigh, and against each other."""
    rng = np.random.RandomState(100 + n)
    B = rng.standard_normal((2 * n, n)) * (0.9 ** np.arange(n))
    A = B.T @ B
    ref = np.linalg.eigvalsh(A)[::-1]
    W1, V1 = (x.cpu().numpy() for x in ops.syevj(dev(A)))
    monkeypatch.setenv("ERA5SVD_SYEVJ_TWOSIDED", "1")
    W2, V2 = (x.cpu().numpy() for x in ops.syevj(dev(A)))
    monkeypatch.delenv("ERA5SVD_SYEVJ_TWOSIDED")
    for W, V in ((W1, V1), (W2, V2)):
        assert np.max(np.abs(W - ref) / ref) < 1e-9
        assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-13 * n
        assert np.max(np.abs(A @ V - V * W)) < 1e-12 * ref[0] * n
    assert np.max(np.abs(np.abs(np.sum(V1 * V2, axis=0)) - 1.0)) < 1e-8        # same eigenvectors up to sign
    # singular (rank n - 1) and indefinite inputs: the pivot test rejects them, two-sided Jacobi answers
    Bs = B.copy(); Bs[:, -1] = Bs[:, 0]
    for M in (Bs.T @ Bs, A - 0.5 * ref[0] * np.eye(n)):
        W, V = (x.cpu().numpy() for x in ops.syevj(dev(M)))
        refm = np.linalg.eigvalsh(M)[::-1]
        assert np.max(np.abs(W - refm)) < 1e-12 * np.abs(refm).max() * n
        assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-13 * n
        assert np.max(np.abs(M @ V - V * W)) < 1e-12 * np.abs(refm).max() * n


def test_syevj_psd_relative_accuracy(ops):
    # graded PSD matrix: Jacobi keeps small eigenvalues to high relative accuracy
    rng = np.random.RandomState(3)
    n = 60
    B = rng.standard_normal((200, n)) * (0.8 ** np.arange(n))
    A = B.T @ B
    W, _ = ops.syevj(dev(A))
    ref = np.linalg.svd(B, compute_uv=False) ** 2
    assert np.max(np.abs(W.cpu().numpy() - ref) / ref) < 1e-9


@pytest.mark.parametrize("l", [1, 5, 33, 110, 118, 119, 128, 129, 200])     # register-tiled path up to 128; smem working copies up to 118
def test_chol_inv(ops, l):
    rng = np.random.RandomState(l)
    B = rng.standard_normal((3 * l + 5, l))
    G = B.T @ B
    R, Rinv = ops.chol_inv(dev(G), 1e-13)
    R, Rinv = R.cpu().numpy(), Rinv.cpu().numpy()
    assert np.allclose(np.tril(R, -1), 0) and np.allclose(np.tril(Rinv, -1), 0)
    assert np.max(np.abs(R.T @ R - G)) < 1e-11 * np.abs(G).max()
    assert np.max(np.abs(R @ Rinv - np.eye(l))) < 1e-9


def test_chol_inv_rank_deficient(ops):
    rng = np.random.RandomState(0)
    B = rng.standard_normal((50, 6))
    B[:, 4] = B[:, 1] + B[:, 2]          # dependent column
    R, Rinv = ops.chol_inv(dev(B.T @ B), 1e-10)
    Q = B @ Rinv.cpu().numpy()
    assert np.all(np.isfinite(Q))
    assert np.linalg.norm(Q[:, 4]) < 1e-100                      # dropped direction
    keep = [0, 1, 2, 3, 5]
    assert np.max(np.abs(Q[:, keep].T @ Q[:, keep] - np.eye(5))) < 1e-8


def test_small_helpers(ops):
    rng = np.random.RandomState(1)
    P = rng.standard_normal((744, 110))
    Pd = dev(P)
    nrm = ops.col_normalize(Pd)
    assert np.allclose(nrm.cpu().numpy(), np.linalg.norm(P, axis=0))
    assert np.allclose(np.linalg.norm(Pd.cpu().numpy(), axis=0), 1.0)
    s, inv = ops.sigma_from_eig(dev(np.array([4.0, 1.0, 0.0, -1e-20])))
    assert np.allclose(s.cpu().numpy(), [2, 1, 0, 0]) and np.allclose(inv.cpu().numpy(), [0.5, 1, 0, 0])
    c = ops.convert(dev(P), torch.float32)
    assert c.dtype == torch.float32 and np.array_equal(c.cpu().numpy(), P.astype(np.float32))
    V = dev(P[:5].copy())
    ops.scale_rows(V, dev(np.array([1., -1., 2., 0., 1.])))
    assert np.allclose(V.cpu().numpy(), P[:5] * np.array([1., -1., 2., 0., 1.])[:, None])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_col_absmax_first_max_and_flip(ops, dtype):
    rng = np.random.RandomState(4)
    m, k = 70001, 37
    U = rng.standard_normal((m, k)).astype(dtype)
    U[123, 0] = -9.0; U[60000, 0] = 9.0; U[124, 0] = -9.0      # tie in |.| -> first row wins
    U[69999, 1] = 50.0
    a, row, sg = ops.col_absmax(dev(U), 1000)
    ref_idx = np.argmax(np.abs(U), axis=0)
    assert np.array_equal(row.cpu().numpy(), ref_idx + 1000)
    assert np.array_equal(sg.cpu().numpy(), np.sign(U[ref_idx, np.arange(k)]))
    assert np.allclose(a.cpu().numpy(), np.abs(U).max(axis=0))
    # combine: two candidate sets, second has the same |max| at a lower row for column 0
    a2, r2, s2 = a.clone(), row.clone(), sg.clone()
    r2[0] = 5; s2[0] = 1.0
    sign = ops.maxloc_combine(torch.stack([a, a2]), torch.stack([row, r2]), torch.stack([sg, s2]))
    assert sign[0].item() == 1.0 and np.array_equal(sign.cpu().numpy()[1:], sg.cpu().numpy()[1:])
    Ud = dev(U)
    ops.scale_cols(Ud, sg)
    assert np.array_equal(Ud.cpu().numpy(), U * sg.cpu().numpy().astype(dtype))


@pytest.mark.parametrize("m,k,ld", [(70001, 100, 112), (33, 16, 16), (5, 128, 128), (1000, 100, 100)])
def test_scale_cols_vectorised(ops, m, k, ld):
    """float32 U with 16-byte aligned rows takes the float4 kernel: every entry scaled, the row padding untouched."""
    rng = np.random.RandomState(m + k)
    buf = torch.from_numpy(rng.standard_normal((m, ld)).astype(np.float32)).cuda()
    ref = buf.clone()
    sg = torch.from_numpy(rng.choice([-1.0, 1.0, 0.0], size=k)).cuda()
    ops.scale_cols(buf[:, :k], sg)
    assert torch.equal(buf[:, :k], ref[:, :k] * sg.float()) and torch.equal(buf[:, k:], ref[:, k:])


@pytest.mark.parametrize("T,P,dtype", [(25, 300, torch.float64), (744, 4096, torch.float32), (1460, 2048, torch.float32),
                                       (2000, 512, torch.float32)])
def test_nonfinite_flag_reports_the_written_value(ops, T, P, dtype):
    """a time-constant point under scale=True has std 0, so 0 / 0 = NaN is WRITTEN to X although every
    source value is finite; the reference raises "Input contains NaN" (sklearn check_array).  The flag must report it
    in every build kernel (TMA tile kernel, register-staged tile kernel, two-kernel path for long series)."""
    from dmd_era5_b200.pipeline import build_matrix_device

    g = torch.Generator(device="cuda"); g.manual_seed(T)
    src = torch.randn((T, P), generator=g, device="cuda", dtype=dtype) + 250.0
    ok = build_matrix_device(ops, [src], mean_center=True, scale=True, check_finite=True)
    assert int(ok.nonfinite.item()) == 0
    src[:, P // 3] = 7.0                                       # constant in time: std = 0
    bad = build_matrix_device(ops, [src], mean_center=True, scale=True, check_finite=True)
    assert int(bad.nonfinite.item()) == 1
    assert not bool(torch.isfinite(bad.X[P // 3]).all())
    only_centred = build_matrix_device(ops, [src], mean_center=True, scale=False, check_finite=True)
    assert int(only_centred.nonfinite.item()) == 0             # centring alone leaves zeros there
    src[T // 2, 5] = float("inf")
    assert int(build_matrix_device(ops, [src], mean_center=False, scale=False, check_finite=True).nonfinite.item()) == 1


def test_svd_on_era5_rejects_nonfinite_input():
    from dmd_era5_b200.era5_svd import svd_on_era5

    X = np.random.RandomState(0).standard_normal((500, 40))
    X[17, 3] = np.nan
    for svd_type in ("randomized", "standard"):
        with pytest.raises(ValueError, match="Input contains NaN or infinity"):
            svd_on_era5(X, {"svd_type": svd_type, "n_components": 5, "random_seed": 0})
    X[17, 3] = np.inf
    with pytest.raises(ValueError, match="Input contains NaN or infinity"):
        svd_on_era5(X.astype(np.float32), {"svd_type": "randomized", "n_components": 5, "random_seed": 0})
