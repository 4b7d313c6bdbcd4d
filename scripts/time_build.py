"""Time the build kernel alone (TMA tile loads vs the register-staged kernel, flag 8 = ERA5SVD_BUILD_NO_TMA) at the
c2 shape and at a c3-shard-like shape (T = 1460).  Bytes = read m*n*4 + write m*n*4.
ERA5SVD_BUILD_PB16=1 selects the round-1 variant for long series (16-point tiles, one CTA per tile) instead of the
32-point tiles split along time over a thread-block cluster."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200._cabi import BUILD_MEAN_CENTER, BUILD_SCALE

ops = CudaOps("cuda:0")


def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for T, P in [(744, 1038240), (1460, 1265356), (2920, 600000), (8760, 200000)]:
    src = torch.randn((T, P), device="cuda") * 10 + 250
    ld = T + (-T) % 8
    X = torch.empty((P, ld), device="cuda")[:, :T]
    mean = torch.empty(P, device="cuda"); std = torch.empty(P, device="cuda")
    gb = 2.0 * T * P * 4 / 1e9
    for name, fl in [("center", BUILD_MEAN_CENTER), ("center+scale", BUILD_MEAN_CENTER | BUILD_SCALE), ("plain", 0)]:
        res = []
        for no_tma in (0, 8):
            ms = timeit(lambda: ops.build_rows(src, X, mean if fl & 1 else None, std if fl & 2 else None, None, fl | no_tma, None))
            res.append(f"{'staged' if no_tma else 'tma'} {ms:.3f} ms = {gb / ms:.2f} TB/s")
        print(f"T={T} P={P} {name}: " + " | ".join(res), flush=True)
    del src, X
