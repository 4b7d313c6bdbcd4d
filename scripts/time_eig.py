"""Time the time-sized eigensolver (eig_tridiag.cu) for growing n."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from dmd_era5_b200.device_ops import CudaOps, KernelTimer
from dmd_era5_b200.standard import sym_eig_topk

ops = CudaOps("cuda:0")
for n in [int(a) for a in sys.argv[1:]] or [744, 1460, 2920]:
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    B = torch.randn((n + 50, n), device="cuda", dtype=torch.float64, generator=g) * (0.999 ** torch.arange(n, device="cuda", dtype=torch.float64))
    G = B.t() @ B
    sym_eig_topk(ops, G, 100); torch.cuda.synchronize()
    t0 = time.perf_counter()
    lam, V = sym_eig_topk(ops, G, 100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    resid = (G @ V - V * lam).norm(dim=0).max().item() / lam[0].item()
    orth = (V.t() @ V - torch.eye(100, device="cuda", dtype=torch.float64)).abs().max().item()
    print(f"n = {n}: top-100 eigenpairs in {dt * 1e3:.1f} ms; max residual / lam_1 = {resid:.1e}; orthogonality {orth:.1e}; "
          f"tridiagonalisation traffic 8 n^3 B = {8 * n ** 3 / 1e9:.1f} GB")
