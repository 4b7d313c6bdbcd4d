"""GPU: time-sized symmetric eigensolver (csrc/eig_tridiag.cu) and the Gram-route standard SVD built on it,
against numpy.linalg.eigh / the reference's np.linalg.svd call (oracle/svd_ref.standard_svd_ref)."""
import numpy as np
import pytest
import torch

from dmd_era5_b200.standard import sym_eig_topk
from oracle.compare import sigma_rel_err, vector_angles
from oracle.svd_ref import standard_svd_ref
from oracle.synthetic_np import lowrank_field_np

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sym(n, seed, decay=0.97, noise=1e-6):
    rng = np.random.RandomState(seed)
    Q = np.linalg.qr(rng.standard_normal((n, n)))[0]
    w = decay ** np.arange(n) + noise * rng.rand(n)
    return (Q * w) @ Q.T, np.sort(w)[::-1]


@pytest.mark.parametrize("n", [3, 5, 130, 257, 744])
def test_tridiag_reduce_preserves_spectrum(ops, n):
    A, w = sym(n, n)
    Ad = dev(A)
    d, e, tau = ops.tridiag_reduce(Ad)
    d, e = d.cpu().numpy(), e.cpu().numpy()
    T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    assert np.allclose(np.sort(np.linalg.eigvalsh(T))[::-1], w, rtol=0, atol=1e-13 * n)


@pytest.mark.parametrize("n,k", [(130, 130), (300, 40), (744, 100), (1460, 100)])
def test_sym_eig_topk_vs_eigh(ops, n, k):
    A, w = sym(n, 7 * n)
    lam, V = sym_eig_topk(ops, dev(A), k)
    lam, V = lam.cpu().numpy(), V.cpu().numpy()
    assert np.allclose(lam, w[:k], rtol=0, atol=1e-13 * n)
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-12
    resid = np.linalg.norm(A @ V - V * lam, axis=0)
    assert resid.max() < 1e-12 * n


def test_sym_eig_topk_clustered(ops):
    """Repeated / tightly clustered eigenvalues: the invariant subspace and the values must still be right."""
    n, k = 400, 30
    rng = np.random.RandomState(3)
    Q = np.linalg.qr(rng.standard_normal((n, n)))[0]
    w = np.concatenate([[5.0, 5.0, 5.0, 4.0, 4.0 + 1e-13, 3.0], 2.0 * 0.9 ** np.arange(n - 6)])
    A = (Q * w) @ Q.T
    lam, V = sym_eig_topk(ops, dev(A), k)
    lam, V = lam.cpu().numpy(), V.cpu().numpy()
    assert np.allclose(lam, np.sort(w)[::-1][:k], atol=1e-12)
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-11
    assert np.linalg.norm(A @ V - V * lam, axis=0).max() < 1e-11


def test_standard_svd_large_n_fp64():
    """Standard SVD with n beyond the one-CTA Jacobi limit, float64: sigma 1e-6, leading vectors 1e-5 rad."""
    from dmd_era5_b200.era5_svd import svd_on_era5

    X = lowrank_field_np(6000, 300, r=80, rho=0.9, seed=11)
    U0, s0, V0 = standard_svd_ref(X, 40)
    U, s, V = svd_on_era5(X, {"svd_type": "standard", "n_components": 40})
    assert U.shape == (6000, 40) and s.shape == (40,) and V.shape == (40, 300)
    assert sigma_rel_err(s, s0) < 1e-6
    assert vector_angles(U[:, :20], U0[:, :20]).max() < 1e-5 and vector_angles(V[:20].T, V0[:20].T).max() < 1e-5


def test_standard_svd_tf32x3_gram():
    """float32 data, Gram on the tensor cores (project kernel over 112-column blocks) + tridiagonal eigensolver."""
    from dmd_era5_b200.era5_svd import svd_on_era5

    X = lowrank_field_np(30000, 744, r=60, rho=0.9, seed=12, dtype=np.float32)
    U0, s0, V0 = standard_svd_ref(X.astype(np.float64), 20)
    U, s, V = svd_on_era5(X, {"svd_type": "standard", "n_components": 20, "precision": "tf32x3"})
    assert U.dtype == np.float32 and U.shape == (30000, 20)
    # Gram route in ~1e-6 arithmetic: sigma_i accurate to ~1e-6 (sigma_1 / sigma_i)^2 (sigma_20 = 0.135 sigma_1)
    assert sigma_rel_err(s, s0) < 1e-4
    assert vector_angles(U[:, :10], U0[:, :10]).max() < 1e-3


def test_tridiag_multi_chunk_slow_decay_and_reproducible(ops):
    """n beyond one staged column chunk (1024) with a SLOWLY decaying spectrum (every column matters), run three times:
    exact similarity transform, residuals at rounding level, bitwise identical results."""
    n, k = 1300, 60
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    B = torch.randn((n + 40, n), device="cuda", dtype=torch.float64, generator=g) * (0.999 ** torch.arange(n, device="cuda", dtype=torch.float64))
    G = B.t() @ B
    G = 0.5 * (G + G.t())
    first = None
    for _ in range(3):
        A = G.clone()
        d, e, tau = ops.tridiag_reduce(A)
        lam, V = sym_eig_topk(ops, G, k)
        cur = (d, e, tau, lam, V)
        if first is None:
            first = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, cur))
    d, e = first[0].cpu().numpy(), first[1].cpu().numpy()
    T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    w = np.linalg.eigvalsh(G.cpu().numpy())
    assert np.abs(np.linalg.eigvalsh(T) - w).max() < 1e-13 * w.max()
    lam, V = first[3], first[4]
    assert float(((G @ V - V * lam).norm(dim=0).max() / lam[0])) < 1e-13
