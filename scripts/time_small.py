"""Time the single-CTA small-factor kernels at the bench's sizes (l = 110, n = 744)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
rng = np.random.RandomState(0)
l, n = 110, 744


def timeit(name, fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:45s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us")


B = rng.standard_normal((3 * l, l)); G = torch.from_numpy(B.T @ B).cuda()
timeit("chol_inv l=110", lambda: ops.chol_inv(G, 1e-13))
Q = np.linalg.qr(rng.standard_normal((l, l)))[0]
w = 100.0 * 0.93 ** np.arange(l)
T_dense = torch.from_numpy((Q * w**2) @ Q.T).cuda()
for sw in (1, 2, 4, 8, 0):
    timeit(f"syevj dense l=110 max_sweeps={sw or 'conv'}", lambda: ops.syevj(T_dense.clone(), sw))
E = 1e-8 * rng.standard_normal((l, l)); T_near = torch.from_numpy(np.diag(w**2) + w[:, None] * (E + E.T) * w[None, :]).cuda()
timeit("syevj near-diagonal l=110 (converged)", lambda: ops.syevj(T_near.clone(), 0))
P = torch.from_numpy(rng.standard_normal((n, l))).cuda()
timeit("gemm_f64 P^T P (110x110, K=744)", lambda: ops.gemm(P, P, transA=True))
timeit("gemm_f64 P R (744x110x110)", lambda: ops.gemm(P, G))
timeit("col_normalize", lambda: ops.col_normalize(P.clone()))
