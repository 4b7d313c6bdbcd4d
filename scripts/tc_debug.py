"""Diagnostics for the tcgen05 kernels (run on the GPU box): compares sketch_tf32x3 / project_tf32x3
with float64 references on random and on one-hot probes, and prints where they differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dmd_era5_b200.device_ops import CudaOps

ops = CudaOps("cuda:0")
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def report(name, got, ref, scale):
    err = np.abs(got - ref)
    print(f"[{name}] max abs err {err.max():.3e} (scale {scale:.3e}, rel {err.max()/scale:.3e}); "
          f"got finite={np.isfinite(got).all()} |got|max={np.abs(got).max():.3e} |ref|max={np.abs(ref).max():.3e}")
    if err.max() > 1e-4 * scale:
        bad = np.argwhere(err > 1e-4 * scale)
        print(f"   {len(bad)} bad entries of {got.size}; first: {bad[:6].tolist()}")
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
        print(f"   bad rows: {len(rows)} [{rows[:8].tolist()}...], bad cols: {len(cols)} [{cols[:8].tolist()}...]")
        return False
    return True


def run_sketch(m, n, l, seed=0, off=0, raw=False):
    rng = np.random.RandomState(seed)
    Xfull = rng.standard_normal((m, n + off)).astype(np.float32)
    X = dev(Xfull)[:, off:]
    Om = rng.standard_normal((n, l))
    hi, lo = ops.split_tf32(dev(Xfull))
    hi, lo = hi[:, off:off + n], lo[:, off:off + n]
    ldy = ops.tf32_ldy(l)
    Y = torch.zeros((m, ldy), device="cuda")[:, :l]; Yh = torch.zeros((m, ldy), device="cuda")[:, :l]; Yl = torch.zeros((m, ldy), device="cuda")[:, :l]
    if raw:
        ops.sketch_tf32x3(X, None, dev(Om), Y, Yh, Yl)
    else:
        ops.sketch_tf32x3(hi, lo, dev(Om), Y, Yh, Yl)
    torch.cuda.synchronize()
    ref = Xfull[:, off:].astype(np.float64) @ Om
    ok = report(f"sketch{'-raw' if raw else ''} m={m} n={n} l={l} off={off}", Y.cpu().numpy().astype(np.float64), ref, np.sqrt(n))
    ok &= report("   Yhi+Ylo", (Yh.double() + Yl.double()).cpu().numpy(), Y.cpu().numpy().astype(np.float64), 1.0)
    return ok


def run_project(m, n, l, seed=0, onehot=False):
    rng = np.random.RandomState(seed)
    Xh = rng.standard_normal((m, n)).astype(np.float32)
    Yh = rng.standard_normal((m, l)).astype(np.float32)
    if onehot:
        Yh[:] = 0
        Yh[5, 3] = 1.0
    ldy = ops.tf32_ldy(l)
    Yb = torch.zeros((m, ldy), device="cuda"); Yb[:, :l] = dev(Yh)
    xhi, xlo = ops.split_tf32(dev(Xh))
    yhi, ylo = ops.split_tf32(Yb)
    Z = ops.project_tf32x3(xhi, xlo, yhi[:, :l], ylo[:, :l])
    torch.cuda.synchronize()
    ref = Xh.astype(np.float64).T @ Yh.astype(np.float64)
    return report(f"project m={m} n={n} l={l} onehot={onehot}", Z.cpu().numpy(), ref, np.sqrt(m) if not onehot else 1.0)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if which in ("all", "sketch"):
        ok &= run_sketch(128, 32, 16)
        ok &= run_sketch(128, 64, 110)
        ok &= run_sketch(1000, 744, 110)
        ok &= run_sketch(300, 100, 130)
        ok &= run_sketch(40000, 744, 110, off=3)
    if which in ("all", "project"):
        ok &= run_project(16, 32, 16, onehot=True)
        ok &= run_project(64, 32, 32)
        ok &= run_project(1000, 744, 110)
        ok &= run_project(50000, 1460, 110)
        ok &= run_project(4097, 25, 20)
    if which in ("all", "raw"):
        ok &= run_sketch(128, 32, 16, raw=True)
        ok &= run_sketch(128, 64, 110, raw=True)
        ok &= run_sketch(1000, 744, 110, raw=True)
        ok &= run_sketch(40000, 744, 110, raw=True)
        ok &= run_sketch(2000, 742, 110, off=2, raw=True)
        import time
        X = torch.randn((1038240, 744), device="cuda"); Om = torch.randn((744, 110), device="cuda", dtype=torch.float64)
        ldy = ops.tf32_ldy(110)
        Yh = torch.zeros((1038240, ldy), device="cuda")[:, :110]; Yl = torch.zeros((1038240, ldy), device="cuda")[:, :110]
        hi, lo = ops.split_tf32(X)
        for name, a, b in (("v1 (hi/lo in HBM)", hi, lo), ("v2 (on-chip split)", X, None)):
            for _ in range(2): ops.sketch_tf32x3(a, b, Om, None, Yh, Yl)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.sketch_tf32x3(a, b, Om, None, Yh, Yl)
            e1.record(); torch.cuda.synchronize()
            print(f"sketch {name}: {e0.elapsed_time(e1) / 5:.3f} ms per pass (c2 shape)")
    print("ALL OK" if ok else "FAILURES")
