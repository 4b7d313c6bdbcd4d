"""Optimized DMD / BOP-DMD.  The reference has no code for this step (README only): PARITY IS UNPINNED by the reference.
CPU: the NumPy restatement (oracle/bopdmd_np.py) is pinned by known-answer tests - planted eigenvalues are recovered and
the Hadamard-structured normal equations equal the explicit Kaufman Jacobian.  GPU: the batched kernels against it."""
import numpy as np
import pytest

from oracle.bopdmd_np import bopdmd, dense_jacobian, initial_eigenvalues, optdmd, subsets, varpro_quantities


def planted(r, N, M, noise, seed=0, tmax=20.0):
    rng = np.random.RandomState(seed)
    om = np.sort(rng.uniform(0.2, 3.0, r // 2)) + 0.15 * np.arange(r // 2)
    gr = -rng.uniform(0.0, 0.05, r // 2)
    alpha = np.concatenate([gr + 1j * om, gr - 1j * om])
    Bh = rng.standard_normal((r // 2, N)) + 1j * rng.standard_normal((r // 2, N))
    B = np.concatenate([Bh, Bh.conj()])
    t = np.linspace(0, tmax, M)
    H = (np.exp(np.outer(t, alpha)) @ B).real + noise * rng.standard_normal((M, N))
    return H, t, alpha, B


def match(a, b):
    """max distance after matching every eigenvalue of a to its nearest in b"""
    return max(np.min(np.abs(x - b)) for x in a)


def test_hadamard_normal_equations_equal_dense_kaufman_jacobian():
    H, t, alpha, _ = planted(6, 8, 150, 0.01, seed=2)
    a = alpha * (1 + 0.02 * np.random.RandomState(1).standard_normal(6))
    rho, B, JhJ, rhs = varpro_quantities(a, t, H)
    J, res = dense_jacobian(a, t, H)
    assert abs(rho - np.sum(np.abs(res) ** 2)) < 1e-9 * rho
    assert np.abs(J.conj().T @ J - JhJ).max() < 1e-11 * np.abs(JhJ).max()
    assert np.abs(-J.conj().T @ res - rhs).max() < 1e-10 * np.abs(rhs).max()


def test_optdmd_recovers_planted_eigenvalues():
    H, t, alpha, B = planted(8, 8, 300, 0.0, seed=3)
    a0 = initial_eigenvalues(H, t, 8)
    a, Bf, rho = optdmd(H, t, a0, max_iter=40)
    assert match(a, alpha) < 1e-8 and rho < 1e-16 * np.sum(H * H)


def test_bopdmd_statistics_cover_the_truth():
    H, t, alpha, _ = planted(8, 10, 400, 0.05, seed=1)
    out = bopdmd(H, t, 8, n_trials=60, trial_size=320, seed=3)
    assert match(out["alpha_full"], alpha) < 2e-3
    # the ensemble mean is at least as close as a few ensemble standard deviations
    for j in range(8):
        d = np.min(np.abs(out["alpha_mean"][j] - alpha))
        assert d < 6 * out["alpha_std"][j] + 1e-4
    assert out["subsets"].shape == (60, 320) and np.all(np.diff(out["subsets"], axis=1) > 0)


@pytest.mark.gpu
def test_device_optdmd_matches_oracle(ops):
    import torch

    from dmd_era5_b200.bopdmd import optdmd_device

    H, t, alpha, _ = planted(10, 12, 257, 0.02, seed=5)
    a0 = initial_eigenvalues(H, t, 10)
    a_ref, B_ref, rho_ref = optdmd(H, t, a0, max_iter=30)
    idx = torch.arange(len(t), dtype=torch.int32, device="cuda").unsqueeze(0)
    a, B, rho, done, it = optdmd_device(ops, torch.from_numpy(H).cuda(), torch.from_numpy(t).cuda(), idx,
                                        torch.from_numpy(a0).cuda(), max_iter=30)
    assert np.abs(a[0].cpu().numpy() - a_ref).max() < 1e-8
    assert abs(float(rho[0]) - rho_ref) < 1e-8 * rho_ref
    assert np.abs(B[0].cpu().numpy() - B_ref).max() < 1e-7 * np.abs(B_ref).max()


@pytest.mark.gpu
def test_device_bopdmd_matches_oracle_trial_by_trial(ops):
    from dmd_era5_b200.bopdmd import bopdmd_device

    H, t, alpha, _ = planted(8, 10, 400, 0.05, seed=1)
    ref = bopdmd(H, t, 8, n_trials=40, trial_size=320, seed=3)
    out = bopdmd_device(ops, H, t, n_trials=40, trial_size=320, r=8, seed=3)
    assert np.array_equal(out["subsets"].cpu().numpy(), ref["subsets"])
    assert np.abs(out["alpha_full"].cpu().numpy() - ref["alpha_full"]).max() < 1e-8
    assert np.abs(out["alphas"].cpu().numpy() - ref["alphas"]).max() < 1e-7
    assert np.abs(out["amps"].cpu().numpy() - ref["amps"]).max() < 1e-6 * ref["amps"].max()
    assert np.abs(out["alpha_std"].cpu().numpy() - ref["alpha_std"]).max() < 1e-7
    assert match(out["alpha_mean"].cpu().numpy(), alpha) < 2e-3


@pytest.mark.gpu
def test_device_bopdmd_r100_scale(ops):
    """configs[4] shape, reduced trial count: r = 100 modes on 1460 snapshots, planted spectrum, 64 trials."""
    from dmd_era5_b200.bopdmd import bopdmd_device

    H, t, alpha, _ = planted(100, 100, 1460, 1e-3, seed=7, tmax=60.0)
    out = bopdmd_device(ops, H, t, n_trials=64, trial_size=1168, seed=1, max_iter=25)
    af = out["alpha_full"].cpu().numpy()
    assert match(af, alpha) < 1e-3
    assert float(out["alpha_std"].max()) < 1e-2
    assert match(out["alpha_mean"].cpu().numpy(), alpha) < 1e-3


@pytest.mark.gpu
def test_device_bopdmd_on_svd_output_of_noise_like_data(ops):
    """Coefficients that are NOT sums of exponentials (random orthonormal temporal patterns, as in the bench's
    synthetic field): the fit is poor, but every trial must end in a finite accepted state or be flagged, never NaN."""
    import torch

    from dmd_era5_b200.bopdmd import bopdmd_on_svd

    rng = np.random.RandomState(0)
    k, n = 20, 300
    V = np.linalg.qr(rng.standard_normal((n, k)))[0].T
    s = 100.0 * 0.9 ** np.arange(k)
    out = bopdmd_on_svd(ops, s, V, np.arange(n, dtype=np.float64), n_trials=16, trial_size=240, seed=2, max_iter=15)
    assert out["alphas"].shape == (16, k)
    assert bool(torch.isfinite(torch.view_as_real(out["alphas"])).all())
    assert bool(torch.isfinite(out["amps"]).all()) and bool(torch.isfinite(out["rhos"]).all())
    # the objective never exceeds the energy of the data it fits
    H = (V * s[:, None]).T
    assert float(out["rho_full"]) <= float((H * H).sum()) * (1 + 1e-9)
