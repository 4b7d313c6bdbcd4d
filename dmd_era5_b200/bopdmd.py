"""Optimized DMD / BOP-DMD on the SVD-projected coefficients - the consumer of the SVD stage.

The reference only NAMES this step (README.md:85, :139: "optimized DMD ... BOP-DMD", Askham & Kutz 2018, Sashidhar &
Kutz 2022); BASELINE.json lists it as configs[4] ("BOP-DMD on SVD-projected coefficients r=100, 1000 bagging trials
batched").  Inputs are exactly what `svd_on_era5` returns:  H = (diag(s) V)^T, one row per snapshot.

    H[idx] ~= Phi(alpha) B,   Phi[i, j] = exp(alpha_j t_i)

Variable projection + Levenberg-Marquardt with Kaufman's Jacobian; all trials advance together, one batched C-ABI call
per LM iteration (`era5svd_bop_iterate_f64`, csrc/bopdmd.cu): Psi = [Re Phi | Im Phi], batched FP64 Gram GEMMs, then one
CTA per trial for the complex Cholesky solves, the accept / reject decision and the next candidate.

The only host arithmetic is the initial guess (eigenvalues of an r x r trapezoidal-rule DMD matrix, NumPy) and the
subset draw (NumPy RandomState, like the reference draws its random test matrix on the host).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._cabi import check


def initial_eigenvalues(H: np.ndarray, t: np.ndarray, r: int) -> np.ndarray:
    """Trapezoidal-rule DMD eigenvalues of the projected snapshots: (h_{i+1} - h_i) / dt_i ~= A (h_{i+1} + h_i) / 2."""
    H = np.asarray(H, dtype=np.float64)
    dt = np.diff(np.asarray(t, dtype=np.float64))
    dH = (H[1:] - H[:-1]) / dt[:, None]
    Hm = 0.5 * (H[1:] + H[:-1])
    U, s, Vt = np.linalg.svd(Hm, full_matrices=False)
    r = min(int(r), int(np.sum(s > 1e-12 * s[0])))
    Ar = (U[:, :r].T @ dH @ Vt[:r].T) / s[:r, None]
    return np.linalg.eigvals(Ar.T).astype(np.complex128)


def draw_subsets(n_time: int, trial_size: int, n_trials: int, seed: int | None) -> np.ndarray:
    """(n_trials, trial_size) time-ordered snapshot indices, without replacement within a trial."""
    rs = np.random.mtrand._rand if seed is None else np.random.RandomState(seed)
    return np.stack([np.sort(rs.choice(n_time, size=trial_size, replace=False)) for _ in range(n_trials)]).astype(np.int32)


def optdmd_device(ops, H: torch.Tensor, t: torch.Tensor, idx: torch.Tensor, alpha0: torch.Tensor, *, max_iter: int = 30,
                  tol: float = 1e-9, lam0: float = 1.0, nu: float = 3.0):
    """Batched optimized DMD.  H (n_time, N) float64, t (n_time,) float64, idx (K, p) int32, alpha0 (K, r) or (r,)
    complex128, all on the device.  Returns (alpha (K, r), B (K, r, N), rho (K,), done (K,), iterations)."""
    dev = ops.device
    if H.dtype != torch.float64 or t.dtype != torch.float64 or idx.dtype != torch.int32:
        raise TypeError("optdmd_device: H, t float64 and idx int32 expected")
    n_time, N = H.shape
    K, p = idx.shape
    a0 = alpha0.to(torch.complex128)
    if a0.dim() == 1:
        a0 = a0.unsqueeze(0).expand(K, -1)
    r = a0.shape[1]
    alpha = a0.contiguous().clone()
    alpha_try = alpha.clone()
    rho = torch.full((K,), float("inf"), dtype=torch.float64, device=dev)
    lam = torch.full((K,), float(lam0), dtype=torch.float64, device=dev)
    JhJ = torch.zeros((K, r, r), dtype=torch.complex128, device=dev)
    rhs = torch.zeros((K, r), dtype=torch.complex128, device=dev)
    B = torch.zeros((K, r, N), dtype=torch.complex128, device=dev)
    done = torch.zeros((K,), dtype=torch.int32, device=dev)
    nbytes = int(ops.lib.era5svd_bop_workspace_bytes(K, p, r, N))
    ws = ops._workspace("bopdmd", nbytes)
    Hc, tc, ic = H.contiguous(), t.contiguous(), idx.contiguous()
    it = 0
    for it in range(1, max_iter + 1):
        check(ops.lib.era5svd_bop_iterate_f64(Hc.data_ptr(), n_time, N, Hc.stride(0), tc.data_ptr(), ic.data_ptr(), K, p, r,
                                              alpha.data_ptr(), alpha_try.data_ptr(), rho.data_ptr(), lam.data_ptr(),
                                              JhJ.data_ptr(), rhs.data_ptr(), B.data_ptr(), done.data_ptr(), float(nu),
                                              float(tol), int(it == 1), ws.data_ptr(), ws.numel(), ops._stream()),
              "era5svd_bop_iterate_f64")
        if it % 5 == 0 and bool(done.all()):
            break
    return alpha, B, rho, done, it


def bopdmd_device(ops, H, t, *, n_trials: int, trial_size: int, r: int | None = None, seed: int | None = 0,
                  max_iter: int = 30, tol: float = 1e-9, alpha0=None, comm=None) -> dict:
    """BOP-DMD: full fit, then `n_trials` refits of random `trial_size`-snapshot subsets started from the full fit.
    H: (n_time, N) projected coefficients (NumPy or tensor), t: (n_time,) times, r: number of DMD modes (default N).
    With a multi-rank `comm` the TRIALS are partitioned over the ranks (they are independent: no data-path collective);
    the per-trial eigenvalues / amplitudes are all-gathered and the mode statistics all-reduced, so every rank returns
    the same dictionary.  Returns tensors on the device: alpha_full (r,), B_full (r, N), alphas (n_trials, r), amps
    (n_trials, r), alpha_mean, alpha_std, amp_mean, amp_std, mode_mean (r, N) / mode_std (r, N) of the unit-norm rows
    of B, and the subsets used."""
    from .dist import LocalComm

    comm = comm or LocalComm()
    if comm.world > 1 and seed is None:
        raise ValueError("bopdmd_device: a seed is required when the trials are partitioned over several ranks")
    dev = ops.device
    H_h = H.detach().cpu().numpy() if torch.is_tensor(H) else np.asarray(H)
    t_h = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
    H_h = np.ascontiguousarray(H_h, dtype=np.float64)
    t_h = np.ascontiguousarray(t_h, dtype=np.float64)
    n_time, N = H_h.shape
    a0 = initial_eigenvalues(H_h, t_h, N if r is None else r) if alpha0 is None else np.asarray(alpha0, dtype=np.complex128)
    with torch.cuda.device(dev):
        Hd = torch.from_numpy(H_h).to(dev)
        td = torch.from_numpy(t_h).to(dev)
        full_idx = torch.arange(n_time, dtype=torch.int32, device=dev).unsqueeze(0)
        a_full, B_full, rho_full, _, it_full = optdmd_device(ops, Hd, td, full_idx, torch.from_numpy(a0).to(dev),
                                                             max_iter=max_iter, tol=tol)
        idx_h = draw_subsets(n_time, trial_size, n_trials, seed)             # identical on every rank (seeded)
        per = -(-n_trials // comm.world)
        k0, k1 = min(n_trials, comm.rank * per), min(n_trials, (comm.rank + 1) * per)
        nr = a_full.shape[1]
        ref = B_full[0] / torch.linalg.vector_norm(B_full[0], dim=1, keepdim=True).clamp_min(1e-300)
        alphas = torch.zeros((per, nr), dtype=torch.complex128, device=dev)
        amps = torch.zeros((per, nr), dtype=torch.float64, device=dev)
        rhos = torch.zeros((per,), dtype=torch.float64, device=dev)
        done = torch.zeros((per,), dtype=torch.int32, device=dev)
        msum = torch.zeros((2, nr, N), dtype=torch.complex128, device=dev)   # sum of modes, sum of |mode|^2
        iters = 0
        if k1 > k0:
            idx = torch.from_numpy(idx_h[k0:k1]).to(dev)
            al, Bs, rh, dn, iters = optdmd_device(ops, Hd, td, idx, a_full[0], max_iter=max_iter, tol=tol)
            am = torch.linalg.vector_norm(Bs, dim=2)                         # (k1 - k0, r)
            modes = Bs / am.clamp_min(1e-300).unsqueeze(-1)
            # fix the phase of every trial's mode to the full fit's before averaging
            ph = (modes * ref.conj().unsqueeze(0)).sum(dim=2)
            modes = modes * (ph.conj() / ph.abs().clamp_min(1e-300)).unsqueeze(-1)
            alphas[: k1 - k0], amps[: k1 - k0], rhos[: k1 - k0], done[: k1 - k0] = al, am, rh, dn
            msum[0] = modes.sum(dim=0)
            msum[1] = modes.abs().pow(2).sum(dim=0).to(torch.complex128)
        if comm.world > 1:
            alphas = torch.view_as_complex(comm.allgather(torch.view_as_real(alphas)).reshape(-1, nr, 2))[:n_trials]
            amps = comm.allgather(amps).reshape(-1, nr)[:n_trials]
            rhos = comm.allgather(rhos).reshape(-1)[:n_trials]
            done = comm.allgather(done).reshape(-1)[:n_trials]
            ms = torch.view_as_real(msum).contiguous()
            comm.allreduce_sum_(ms)
            msum = torch.view_as_complex(ms)
        else:
            alphas, amps, rhos, done = alphas[:n_trials], amps[:n_trials], rhos[:n_trials], done[:n_trials]
        mode_mean = msum[0] / n_trials
        mode_var = (msum[1].real / n_trials - mode_mean.abs().pow(2)).clamp_min(0.0)
        return {
            "alpha_full": a_full[0], "B_full": B_full[0], "rho_full": rho_full[0], "iterations_full": it_full,
            "alphas": alphas, "amps": amps, "rhos": rhos, "done": done, "iterations": iters,
            "alpha_mean": alphas.mean(dim=0),
            "alpha_std": torch.sqrt(alphas.real.var(dim=0, unbiased=False) + alphas.imag.var(dim=0, unbiased=False)),
            "amp_mean": amps.mean(dim=0), "amp_std": amps.std(dim=0, unbiased=False),
            "mode_mean": mode_mean, "mode_std": torch.sqrt(mode_var),
            "subsets": torch.from_numpy(idx_h).to(dev),
        }


def bopdmd_on_svd(ops, s, V, times, **kwargs) -> dict:
    """BOP-DMD on the output of the SVD stage: s (k,), V (k, n) as returned by `svd_on_era5` / stored in the stage's
    NetCDF file (README.md:97-119), `times` (n,) in any unit (hours since the first snapshot is a good choice: the
    eigenvalues come out per that unit).  Physical-space DMD modes are U @ mode.T for the stage's U."""
    s_t = torch.as_tensor(np.asarray(s) if not torch.is_tensor(s) else s).to(torch.float64)
    V_t = torch.as_tensor(np.asarray(V) if not torch.is_tensor(V) else V).to(torch.float64)
    H = (V_t * s_t.unsqueeze(1)).t().contiguous()         # (n, k): projected coefficients, one row per snapshot
    return bopdmd_device(ops, H, times, **kwargs)
