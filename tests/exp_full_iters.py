"""Experiment (run by hand on the GPU, prints only): effect of the number of full-precision power iterations under
precision 'tf32mix' on sigma and the vectors, against the float64 oracle on the same float32 data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dmd_era5_b200.era5_svd import get_ops, host_to_device_matrix
from dmd_era5_b200.pipeline import svd_device
from oracle.compare import recon_rel_err, sigma_rel_err, vector_angles
from oracle.svd_ref import randomized_svd_ref
from oracle.synthetic_np import lowrank_field_np

K = 100
ops = get_ops()
for (m, n, r, rho, seed) in [(40000, 1460, 160, 0.93, 4), (20000, 744, 160, 0.93, 9), (30000, 744, 300, 0.97, 11),
                             (30000, 1460, 400, 0.98, 12)]:
    X = lowrank_field_np(m, n, r=r, rho=rho, seed=seed, dtype=np.float32)
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), K, 1)
    U32, s32, V32 = randomized_svd_ref(X, K, 1)
    a32 = vector_angles(U32, U0)
    print(f"shape {m}x{n} rho={rho}: reference float32 vs float64: sigma {sigma_rel_err(s32, s0):.2e} angle first50 {a32[:50].max():.2e} all {a32.max():.2e}")
    Xd = host_to_device_matrix(ops, X)
    for prec, fi in [("tf32x3", None), ("tf32mix", 2), ("tf32mix", 1), ("tf32mix", 0)]:
        U, s, V = svd_device(ops, Xd, svd_type="randomized", n_components=K, seed=1, precision=prec, full_iters=fi)
        U, s, V = U.cpu().numpy(), s.cpu().numpy(), V.cpu().numpy()
        au, av = vector_angles(U, U0), vector_angles(V.T, V0.T)
        print(f"   {prec} full_iters={fi}: sigma {sigma_rel_err(s, s0):.2e}  U first50 {au[:50].max():.2e} all {au.max():.2e}  "
              f"V first50 {av[:50].max():.2e} all {av.max():.2e}  recon {recon_rel_err(X.astype(np.float64), U, s, V):.6e} (ref {recon_rel_err(X.astype(np.float64), U0, s0, V0):.6e})")
