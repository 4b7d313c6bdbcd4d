"""Config-4-shaped standard SVD on ONE rank's shard: 1/8 of 1 deg x 3 vars x 37 levels (904 095 rows) x 8760 hourly
snapshots, float32, Gram route (tensor-core Gram over 112-column blocks, tridiagonal eigensolver, U = X V S^-1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps, KernelTimer
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 7232760 // 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8760
precision = sys.argv[3] if len(sys.argv) > 3 else "auto"
ops = CudaOps("cuda:0")
field = synthetic_field(T, rows, device="cuda", seed=4, rank=200, rho=0.96, chunk=1 << 17)
built = build_matrix_device(ops, [field], mean_center=True, scale=False)
del field
X = built.X
torch.cuda.synchronize()
for rep in range(2):
    ops.timer = KernelTimer()
    t0 = time.perf_counter()
    stats = {}
    U, s, V = svd_device(ops, X, svd_type="standard", n_components=100, precision=precision, stats=stats)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    summ = ops.timer.summary(); ops.timer = None
    print(f"rep {rep}: {rows} x {T} f32 ({rows * T * 4 / 1e9:.1f} GB) standard SVD k=100: {dt:.3f} s = {rows * T * 4 / 1e9 / dt:.1f} GB/s; "
          f"sigma_1..3 = {s[:3].tolist()}")
    print("   kernels:", {k: (round(v["ms"], 1), v["calls"]) for k, v in summ.items()}, "precision", precision, stats)
# invariants at full size (the CPU reference cannot run this shape): orthonormal U, V and X^T U = V^T S
UtU = ops.project_tf32x3(X[:, :1], None, *ops.split_tf32(torch.ones((rows, 4), device="cuda"))) if False else None
Ud = U[:, :8].double()
print("   ||U^T U - I|| (first 8) =", (Ud.t() @ Ud - torch.eye(8, device="cuda", dtype=torch.float64)).abs().max().item(),
      " ||V V^T - I|| =", (V @ V.t() - torch.eye(V.shape[0], device="cuda", dtype=torch.float64)).abs().max().item())
