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


def run_project(m, n, l, seed=0, onehot=False, raw=False, off=0):
    rng = np.random.RandomState(seed)
    Xfull = rng.standard_normal((m, n + off)).astype(np.float32)
    Xh = Xfull[:, off:]
    Yh = rng.standard_normal((m, l)).astype(np.float32)
    if onehot:
        Yh[:] = 0
        Yh[5, 3] = 1.0
    ldy = ops.tf32_ldy(l)
    Yb = torch.zeros((m, ldy), device="cuda"); Yb[:, :l] = dev(Yh)
    yhi, ylo = ops.split_tf32(Yb)
    if raw:
        ldp = -(-(n + off) // 4) * 4
        Xp = torch.zeros((m, ldp), device="cuda"); Xp[:, :n + off] = dev(Xfull)
        Z = ops.project_tf32x3(Xp[:, off:off + n], None, yhi[:, :l], ylo[:, :l])
    else:
        xhi, xlo = ops.split_tf32(dev(Xfull))
        Z = ops.project_tf32x3(xhi[:, off:], xlo[:, off:], yhi[:, :l], ylo[:, :l])
    torch.cuda.synchronize()
    ref = Xh.astype(np.float64).T @ Yh.astype(np.float64)
    return report(f"project{'-raw' if raw else ''} m={m} n={n} l={l} onehot={onehot} off={off}", Z.cpu().numpy(), ref, np.sqrt(m) if not onehot else 1.0)


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
    if which in ("all", "rawp"):
        ok &= run_project(16, 32, 16, onehot=True, raw=True)
        ok &= run_project(64, 32, 32, raw=True)
        ok &= run_project(1000, 744, 110, raw=True)
        ok &= run_project(50000, 1460, 110, raw=True)
        ok &= run_project(4097, 25, 20, raw=True)
        ok &= run_project(3000, 742, 110, raw=True, off=2)
        X = torch.randn((1038240, 744), device="cuda")
        ldy = ops.tf32_ldy(110)
        Yb = torch.randn((1038240, ldy), device="cuda"); yhi, ylo = ops.split_tf32(Yb)
        hi, lo = ops.split_tf32(X)
        for name, a, b in (("v1 (hi/lo in HBM)", hi, lo), ("v2 (on-chip split)", X, None)):
            for _ in range(2): ops.project_tf32x3(a, b, yhi[:, :110], ylo[:, :110])
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.project_tf32x3(a, b, yhi[:, :110], ylo[:, :110])
            e1.record(); torch.cuda.synchronize()
            print(f"project {name}: {e0.elapsed_time(e1) / 5:.3f} ms per pass (c2 shape)")
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
