"""Does the row pitch of X matter for the tall passes?  Times sketch / project (x1 and 3x forms) on the same m x n data
stored with different leading dimensions (CUDA events, L2-cold: the matrix is far larger than L2).

    python scripts/time_pitch.py [n] [rows] [ld,ld,...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1460
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
l = 110
Om = torch.from_numpy(np.linalg.qr(np.random.RandomState(0).standard_normal((n, l)))[0]).cuda()
ops.round_tf32_(Om)
ldy = ops.tf32_ldy(l)
Y = torch.randn((m, ldy), device="cuda")[:, :l]
Yh = torch.zeros((m, ldy), device="cuda")[:, :l]; Yl = torch.zeros((m, ldy), device="cuda")[:, :l]


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = m * n * 4 / 1e9
lds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
for ld in lds or sorted({(n + 7) // 8 * 8, (n + 31) // 32 * 32, (n + 63) // 64 * 64 + 8, (n + 63) // 64 * 64 + 32, (n + 127) // 128 * 128,
                  (n + 127) // 128 * 128 + 32, (n + 511) // 512 * 512, (n + 511) // 512 * 512 + 32}):
    X = torch.randn((m, ld), device="cuda")[:, :n]
    t = {"sketch_x1": timeit(lambda: ops.sketch_tf32x1(X, Om, Y)),
         "project_x1": timeit(lambda: ops.project_tf32x1(X, Y)),
         "sketch_x2": timeit(lambda: ops.sketch_tf32x3(X, None, Om, Y, None, None, om_tf32=True)),
         "project_x3": timeit(lambda: ops.project_tf32x3(X, None, Y, None))}
    print(f"n={n} m={m} ld={ld} (pitch {ld * 4} B): " + "  ".join(f"{k} {v:.3f} ms ({gb / v * 1e3:.0f} GB/s)" for k, v in t.items()), flush=True)
    del X
