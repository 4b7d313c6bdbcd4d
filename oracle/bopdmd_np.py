"""TEST INFRASTRUCTURE - NumPy statement of optimized DMD / BOP-DMD on SVD-projected coefficients.

The reference (ClimeTrend/DMD-ERA5) has NO code for this step: its README only names it as the consumer of the SVD
stage (README.md:85, :139 cite Askham & Kutz 2018 and Sashidhar & Kutz 2022) and `pydmd` is not a dependency.
PARITY IS THEREFORE UNPINNED by the reference; this file restates the published algorithm and is pinned by
known-answer tests (planted eigenvalues) in tests/test_bopdmd.py.  Only tests/ and bench code may import it.

Model (Askham & Kutz, "Variable projection methods for an optimized dynamic mode decomposition", SIAM J. Appl. Dyn.
Syst. 2018): snapshots of the projected coefficients  h(t_i) in R^N  (N = n_components, rows of H = (diag(s) V)^T),

        H  ~=  Phi(alpha) B,      Phi[i, j] = exp(alpha_j t_i),   alpha in C^r,  B in C^{r x N}

Variable projection eliminates B = Phi^+ H and minimises  rho(alpha) = || H - Phi Phi^+ H ||_F^2  by
Levenberg-Marquardt with Kaufman's approximation of the Jacobian.  Because d Phi / d alpha_j touches only column j,

        J[:, j] = - vec( (P_perp d_j) b_j^T ),   d_j = t * Phi[:, j],   b_j = B[j, :],   P_perp = I - Phi Phi^+

so the normal equations need only r x r and r x N quantities (G = Phi^H Phi etc.):

        J^H J   = ( D^H P_perp D ) o conj( B B^H )          (Hadamard product)
        -J^H res = rowsum( conj(B) o (D^H Res) ),            D^H Res = D^H H - (D^H Phi) B
        rho     = ||H||_F^2 - Re tr( (Phi^H H)^H B )

BOP-DMD (Sashidhar & Kutz, "Bagging, optimized dynamic mode decomposition for robust, stable forecasting with
spatial and temporal uncertainty quantification", Phil. Trans. R. Soc. A 2022): fit all snapshots once, then refit
`n_trials` random subsets of `trial_size` snapshots (without replacement, time-ordered) starting from the full fit,
and report the mean and standard deviation of eigenvalues, amplitudes and modes over the trials.
"""
from __future__ import annotations

import numpy as np


def initial_eigenvalues(H: np.ndarray, t: np.ndarray, r: int) -> np.ndarray:
    """Trapezoidal-rule DMD eigenvalues of the projected snapshots (the usual optimized-DMD initial guess):
    (h_{i+1} - h_i) / dt_i  ~=  A (h_{i+1} + h_i) / 2, alpha_0 = eig(A)."""
    H = np.asarray(H, dtype=np.float64)
    dt = np.diff(t)
    dH = (H[1:] - H[:-1]) / dt[:, None]           # (M-1, N)
    Hm = 0.5 * (H[1:] + H[:-1])
    # least squares A^T: Hm A^T = dH, restricted to the leading r directions of Hm
    U, s, Vt = np.linalg.svd(Hm, full_matrices=False)
    r = min(r, int(np.sum(s > 1e-12 * s[0])))
    Ar = (U[:, :r].T @ dH @ Vt[:r].T) / s[:r, None]          # (r, r) = S^-1 U^T dH V
    return np.linalg.eigvals(Ar.T).astype(np.complex128)


def varpro_quantities(alpha: np.ndarray, t: np.ndarray, H: np.ndarray):
    """Everything one LM iteration needs, from r x r / r x N reductions only (what the device kernels compute)."""
    Phi = np.exp(np.outer(t, alpha))                   # (M, r)
    D = t[:, None] * Phi
    G = Phi.conj().T @ Phi
    F = Phi.conj().T @ D
    E2 = D.conj().T @ D
    C = Phi.conj().T @ H                               # (r, N)
    Ct = D.conj().T @ H
    L = np.linalg.cholesky(G)
    solve = lambda R: np.linalg.solve(L.conj().T, np.linalg.solve(L, R))
    B = solve(C)
    rho = float(np.sum(H * H) - np.real(np.sum(C.conj() * B)))
    P = solve(F)
    A1 = E2 - F.conj().T @ P                           # D^H P_perp D
    E = Ct - F.conj().T @ B                            # D^H Res   (D^H Phi = F^H)
    JhJ = A1 * np.conj(B @ B.conj().T)
    rhs = np.sum(np.conj(B) * E, axis=1)
    return rho, B, JhJ, rhs


def dense_jacobian(alpha: np.ndarray, t: np.ndarray, H: np.ndarray):
    """Kaufman Jacobian and residual formed explicitly (validation of the Hadamard formulas, small cases only)."""
    Phi = np.exp(np.outer(t, alpha))
    Q, _ = np.linalg.qr(Phi)
    B = np.linalg.lstsq(Phi, H.astype(np.complex128), rcond=None)[0]
    Res = H - Phi @ B
    M, r = Phi.shape
    J = np.empty((M * H.shape[1], r), dtype=np.complex128)
    for j in range(r):
        d = t * Phi[:, j]
        pd = d - Q @ (Q.conj().T @ d)
        J[:, j] = -np.outer(pd, B[j]).ravel()
    return J, Res.ravel()


def optdmd(H: np.ndarray, t: np.ndarray, alpha0: np.ndarray, max_iter: int = 30, tol: float = 1e-9,
           lam0: float = 1.0, nu: float = 3.0):
    """Levenberg-Marquardt on alpha.  Fixed control flow shared with the device driver: every iteration evaluates
    the candidate alpha_try; a decrease of rho accepts it (lambda /= nu), otherwise lambda *= nu; the next candidate
    is alpha + delta(lambda) from the normal equations of the last accepted point.  A trial stops when an accepted
    step gains less than tol * rho, or a rejected candidate is within tol * rho of the accepted objective (rho is
    formed as ||H||^2 - tr(C^H B), so its own rounding is ~1e-16 ||H||^2: tol below ~1e-10 is meaningless)."""
    H = np.asarray(H, dtype=np.float64)
    alpha = np.array(alpha0, dtype=np.complex128)
    rho = np.inf
    lam = lam0
    JhJ = rhs = B = None
    a_try = alpha.copy()
    done = False
    for _ in range(max_iter):
        if not done:
            rho_t, B_t, JhJ_t, rhs_t = varpro_quantities(a_try, t, H)
            if rho_t < rho:
                converged = np.isfinite(rho) and (rho - rho_t) <= tol * rho
                alpha, rho, B, JhJ, rhs = a_try, rho_t, B_t, JhJ_t, rhs_t
                lam = max(lam / nu, 1e-12)
                done = converged
            else:
                lam = lam * nu
                # stagnation (the candidate is no worse than tol): converged; runaway damping: give up
                if lam > 1e12 or (np.isfinite(rho) and rho_t - rho <= tol * rho):
                    done = True
            if not done:
                Areg = JhJ + lam * np.diag(np.real(np.diag(JhJ)))
                a_try = alpha + np.linalg.solve(Areg, rhs)
    return alpha, B, rho


def subsets(n_time: int, trial_size: int, n_trials: int, seed: int) -> np.ndarray:
    """(n_trials, trial_size) sorted snapshot indices, drawn without replacement per trial."""
    rs = np.random.RandomState(seed)
    return np.stack([np.sort(rs.choice(n_time, size=trial_size, replace=False)) for _ in range(n_trials)]).astype(np.int32)


def bopdmd(H: np.ndarray, t: np.ndarray, r: int, n_trials: int, trial_size: int, seed: int = 0, max_iter: int = 30,
           tol: float = 1e-9, alpha0: np.ndarray | None = None):
    """Returns dict(alpha_full, alpha_mean, alpha_std, amp_mean, amp_std, alphas (n_trials, r))."""
    H = np.asarray(H, dtype=np.float64)
    a0 = initial_eigenvalues(H, t, r) if alpha0 is None else np.asarray(alpha0, dtype=np.complex128)
    a_full, B_full, rho_full = optdmd(H, t, a0, max_iter, tol)
    idx = subsets(len(t), trial_size, n_trials, seed)
    alphas = np.empty((n_trials, len(a_full)), dtype=np.complex128)
    amps = np.empty((n_trials, len(a_full)))
    for k in range(n_trials):
        a, B, _ = optdmd(H[idx[k]], t[idx[k]], a_full, max_iter, tol)
        alphas[k] = a
        amps[k] = np.linalg.norm(B, axis=1)
    return {"alpha_full": a_full, "B_full": B_full, "rho_full": rho_full, "alphas": alphas, "amps": amps,
            "alpha_mean": alphas.mean(axis=0), "alpha_std": np.sqrt(alphas.real.var(axis=0) + alphas.imag.var(axis=0)),
            "amp_mean": amps.mean(axis=0), "amp_std": amps.std(axis=0), "subsets": idx}
