"""Time the tall tcgen05 passes alone at the bench shape (events around each launch, L2-cold inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
m, n, l = 1038240, 744, 110
X = torch.randn((m, n), device="cuda")
Om = torch.from_numpy(np.linalg.qr(np.random.RandomState(0).standard_normal((n, l)))[0]).cuda()
ldy = ops.tf32_ldy(l)
Yh = torch.zeros((m, ldy), device="cuda")[:, :l]; Yl = torch.zeros((m, ldy), device="cuda")[:, :l]

def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print(os.environ.get("ERA5SVD_SK_DBG", "0"), os.environ.get("ERA5SVD_PJ_DBG", "0"),
      "sketch %.3f ms" % timeit(lambda: ops.sketch_tf32x3(X, None, Om, None, Yh, Yl)),
      "sketch x2 (tf32 Omega) %.3f ms" % timeit(lambda: ops.sketch_tf32x3(X, None, Om, None, Yh, Yl, om_tf32=True)),
      "project %.3f ms" % timeit(lambda: ops.project_tf32x3(X, None, Yh, Yl)),
      "project plain-Y %.3f ms" % timeit(lambda: ops.project_tf32x3(X, None, Yh, None)),
      "sketch x2 one image %.3f ms" % timeit(lambda: ops.sketch_tf32x3(X, None, Om, Yh, None, None, om_tf32=True)))
