"""One launch each of the single-CTA small-factor kernels at the bench's sizes (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
rng = np.random.RandomState(0)
l = 110
B = rng.standard_normal((3 * l, l)); G = torch.from_numpy(B.T @ B).cuda()
Q = np.linalg.qr(rng.standard_normal((l, l)))[0]
w = 100.0 * 0.93 ** np.arange(l)
T_dense = torch.from_numpy((Q * w**2) @ Q.T).cuda()
for _ in range(2):
    ops.chol_inv(G, 1e-13)
    ops.syevj(T_dense.clone(), 2)
torch.cuda.synchronize()
