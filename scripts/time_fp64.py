"""FP64 tall passes (parity mode) at the c2 shape: achieved FP64 TFLOP/s of the SIMT kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200._cabi import PREC_NATIVE
ops = CudaOps("cuda:0")
m, n, l = 1038240, 744, 110
only64 = "--f64" in sys.argv
print("DMMA staging of the X tile:", "strided (first version)" if os.environ.get("ERA5SVD_DMMA_STAGING", "").startswith("s") else "lane-contiguous")
for dt in ((torch.float64,) if only64 else (torch.float64, torch.float32)):
    X = torch.randn((m, n), device="cuda", dtype=dt)
    Om = torch.randn((n, l), device="cuda", dtype=dt)
    Y = torch.empty((m, l), device="cuda", dtype=dt)
    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ts = timeit(lambda: ops.sketch(X, Om, Y, PREC_NATIVE))
    tp = timeit(lambda: ops.project(X, Y, None, precision=PREC_NATIVE))
    fl = 2.0 * m * n * l
    print(f"{dt}: sketch {ts:.2f} ms = {fl / ts / 1e9:.1f} TFLOP/s, project {tp:.2f} ms = {fl / tp / 1e9:.1f} TFLOP/s; "
          f"HBM floor {X.element_size() * m * n / 6.46e9:.2f} ms")
    del X, Y
