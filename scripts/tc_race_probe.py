"""Diagnostic: run the on-chip-split sketch repeatedly and locate entries that differ from a float64 product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
m, n, l, off = [int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (40000, 744, 100, 3))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 8
rng = np.random.RandomState(1)
X = torch.randn((m, n + off + (-(n + off)) % 8), device="cuda")
Om = torch.from_numpy(rng.standard_normal((n, l))).cuda()
ref = X[:, off:off + n].double() @ Om
scale = X[:, off:off + n].double().norm(dim=1)[:, None] * Om.norm(dim=0)[None, :]
ldy = ops.tf32_ldy(l)
first = None
for rep in range(reps):
    Y = torch.zeros((m, ldy), device="cuda")[:, :l]
    Yh = torch.zeros((m, ldy), device="cuda")[:, :l]
    Yl = torch.zeros((m, ldy), device="cuda")[:, :l]
    ops.sketch_tf32x3(X[:, off:off + n], None, Om, Y, Yh, Yl)
    torch.cuda.synchronize()
    err = ((Y.double() - ref).abs() / scale)
    same = None if first is None else bool(torch.equal(Y, first))
    if first is None:
        first = Y.clone()
    bad = err > 2e-6
    rows = bad.any(dim=1).nonzero().flatten()
    tiles = torch.unique(rows // 128)
    print(f"rep {rep}: max rel err {err.max().item():.2e}, bad entries {int(bad.sum())}, bad rows {rows.numel()}, "
          f"bad tiles {tiles.numel()} {tiles[:12].tolist()}, identical to rep 0: {same}")
    if rows.numel():
        print("   bad rows by (row % 128):", torch.bincount(rows % 128, minlength=128).tolist())
        print("   bad tiles by seq (tile // 148):", torch.bincount(tiles // 148).tolist())
        e_row = err.max(dim=1).values[rows]
        print("   err quantiles over bad rows:", [float(torch.quantile(e_row, q)) for q in (0.1, 0.5, 0.9)])
        r0 = int(rows[0])
        cols = bad[r0].nonzero().flatten()
        print(f"   first bad row {r0} (tile {r0 // 128}, seq {r0 // 128 // 148}): bad cols {cols[:16].tolist()} n={cols.numel()}, "
              f"err {err[r0].max().item():.2e}")
