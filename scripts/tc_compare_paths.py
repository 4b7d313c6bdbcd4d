"""Diagnostic: on-chip-split kernels (gemm_tc2.cu) vs pre-split kernels (gemm_tc.cu) vs float64 on bench-like data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

ops = CudaOps("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
T, l = 744, 110
field = synthetic_field(T, S, device="cuda", seed=1000)
built = build_matrix_device(ops, [field], mean_center=True, scale=False, split=False, keep_x=True)
X = built.X
print("X", tuple(X.shape), X.stride(), "absmax", X.abs().max().item())
hi, lo = ops.split_tf32(X)
rng = np.random.RandomState(0)
Om = torch.from_numpy(np.linalg.qr(rng.standard_normal((T, l)))[0]).cuda()
ldy = ops.tf32_ldy(l)
def bufs():
    return [torch.zeros((S, ldy), device="cuda")[:, :l] for _ in range(3)]
Y1, Yh1, Yl1 = bufs(); Y2, Yh2, Yl2 = bufs()
ops.sketch_tf32x3(hi, lo, Om, Y1, Yh1, Yl1)
ops.sketch_tf32x3(X, None, Om, Y2, Yh2, Yl2)
ref = X.double() @ Om
sc = X.double().norm(dim=1)[:, None] * Om.norm(dim=0)[None, :]
for name, Y in (("pre-split", Y1), ("on-chip", Y2)):
    e = (Y.double() - ref).abs() / sc
    print(f"sketch {name}: max scaled err {e.max().item():.2e}; rel Frobenius {((Y.double()-ref).norm()/ref.norm()).item():.2e}")
    bad = (e > 3e-6).any(dim=1).nonzero().flatten()
    if bad.numel():
        print("   bad rows:", bad.numel(), "tiles", torch.unique(bad // 128)[:10].tolist(), "cols of first", (e[bad[0]] > 3e-6).nonzero().flatten()[:10].tolist())
Z1 = ops.project_tf32x3(hi, lo, Yh1, Yl1)
Z2 = ops.project_tf32x3(X, None, Yh1, Yl1)
zref = X.double().t() @ (Yh1.double() + Yl1.double())
zs = X.double().norm(dim=0)[:, None] * (Yh1.double() + Yl1.double()).norm(dim=0)[None, :]
for name, Z in (("pre-split", Z1), ("on-chip", Z2)):
    e = (Z - zref).abs() / zs
    print(f"project {name}: max scaled err {e.max().item():.2e}; rel Frobenius {((Z-zref).norm()/zref.norm()).item():.2e}")
    bad = (e > 3e-6).nonzero()
    if bad.numel():
        print("   bad entries:", bad.shape[0], "times", torch.unique(bad[:, 0])[:20].tolist(), "cols", torch.unique(bad[:, 1])[:20].tolist())
U1, s1, V1 = svd_device(ops, X, svd_type="randomized", n_components=100, seed=1, precision="tf32x3")
U2, s2, V2 = svd_device(ops, None, svd_type="randomized", n_components=100, seed=1, precision="tf32x3", split=(hi, lo))
U3, s3, V3 = svd_device(ops, X, svd_type="randomized", n_components=100, seed=1, precision="native")
print("sigma on-chip  ", s1[:3].tolist())
print("sigma pre-split", s2[:3].tolist())
print("sigma fp32 fma ", s3[:3].tolist())
print("max rel sigma diff on-chip vs pre-split", ((s1 - s2).abs() / s2).max().item(), " vs fp32", ((s1 - s3).abs() / s3).max().item())
# determinism
for name, fn in (("sketch pre-split", lambda: (ops.sketch_tf32x3(hi, lo, Om, Y1, Yh1, Yl1), Y1.clone())[1]),
                 ("gram pre-split", lambda: ops.project_tf32x3(Yh1, Yl1, Yh1, Yl1).clone()),
                 ("sketch on-chip", lambda: (ops.sketch_tf32x3(X, None, Om, Y2, Yh2, Yl2), Y2.clone())[1]),
                 ("project on-chip", lambda: ops.project_tf32x3(X, None, Yh1, Yl1).clone()),
                 ("project pre-split", lambda: ops.project_tf32x3(hi, lo, Yh1, Yl1).clone())):
    outs = [fn() for _ in range(4)]
    torch.cuda.synchronize()
    print(name, "bitwise repeatable:", [bool(torch.equal(outs[0], o)) for o in outs[1:]],
          "max abs diff", max(float((outs[0] - o).abs().max()) for o in outs[1:]))
sig = [svd_device(ops, X, svd_type="randomized", n_components=100, seed=1, precision="tf32x3")[1] for _ in range(4)]
print("svd on-chip sigma_1 over 4 runs:", [float(s[0]) for s in sig])
sig = [svd_device(ops, None, svd_type="randomized", n_components=100, seed=1, precision="tf32x3", split=(hi, lo))[1] for _ in range(3)]
print("svd pre-split sigma_1 over 3 runs:", [float(s[0]) for s in sig])
