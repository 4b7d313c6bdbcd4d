"""Time every tall kernel of the randomized schedule alone at a given shape (sustained: many back-to-back launches), with the
fraction of the measured HBM copy rate.   python scripts/time_tall_shapes.py [rows] [snapshots] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2530752
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1460
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
l = 110
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6460.5
ops = CudaOps("cuda:0")
ld = (n + 7) // 8 * 8
src = torch.randn((n, m), device="cuda")                         # native layout (time, points)
X = torch.empty((m, ld), device="cuda")[:, :n]
mean = torch.empty(m, device="cuda")
Om = torch.from_numpy(np.linalg.qr(np.random.RandomState(0).standard_normal((n, l)))[0]).cuda()
ops.round_tf32_(Om)
ldy = ops.tf32_ldy(l)
Y = torch.zeros((m, ldy), device="cuda")[:, :l]
Yh = torch.zeros((m, ldy), device="cuda")[:, :l]; Yl = torch.zeros((m, ldy), device="cuda")[:, :l]
Z = torch.zeros((n, l), device="cuda", dtype=torch.float64)


def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


xb, yb = 4.0 * m * n, 4.0 * m * l
cases = [
    ("build (center)", lambda: ops.build_rows(src, X, mean, None, None, 1 | 4, None), 2 * xb),
    ("sketch x1", lambda: ops.sketch_tf32x1(X, Om, Y), xb + yb),
    ("project x1", lambda: ops.project_tf32x1(X, Y, Z), xb + yb),
    ("sketch 2xTF32 (one Y image)", lambda: ops.sketch_tf32x3(X, None, Om, Y, None, None, om_tf32=True), xb + yb),
    ("sketch 2xTF32 (hi / lo pair)", lambda: ops.sketch_tf32x3(X, None, Om, None, Yh, Yl, om_tf32=True), xb + 2 * yb),
    ("project x2 (Y truncated)", lambda: ops.project_tf32x2(X, Y, Z), xb + yb),
    ("project 3xTF32 (plain Y)", lambda: ops.project_tf32x3(X, None, Y, None, Z), xb + yb),
    ("project 3xTF32 (hi / lo pair)", lambda: ops.project_tf32x3(X, None, Yh, Yl, Z), xb + 2 * yb),
]
only = os.environ.get("ONLY")
print(f"shape {m} x {n}, l = {l}, {reps} back-to-back launches each; HBM peak {HBM} GB/s")
for name, fn, nbytes in cases:
    if only and only not in name:
        continue
    ms = timeit(fn)
    print(f"{name:32s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s  {nbytes / ms / 1e6 / HBM:5.3f} of HBM")
