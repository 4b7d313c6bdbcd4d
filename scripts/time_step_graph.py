"""c2 step (build + randomized SVD, tf32x3 on-chip split): eager launches with / without per-kernel events vs one
CUDA-graph replay per step.  Tells how much of the step is launch gaps / event overhead rather than kernel time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps, KernelTimer
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

dev = torch.device("cuda", 0)
ops = CudaOps(dev)
S, T, k = 1038240, 744, 100
field = synthetic_field(T, S, device=dev, seed=1000)


def step(timer=None):
    ops.timer = timer
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    out = svd_device(ops, built.X, svd_type="randomized", n_components=k, seed=1, precision="tf32x3")
    ops.timer = None
    return out


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("eager, no events      %.3f ms" % timeit(step), flush=True)
tm = KernelTimer()
print("eager, kernel events  %.3f ms" % timeit(lambda: step(tm)), flush=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    U, sv, V = step()
torch.cuda.synchronize()
print("graph replay          %.3f ms" % timeit(g.replay), flush=True)
U0, s0, V0 = step()
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("graph == eager:", float((sv - s0).abs().max()), float((U - U0).abs().max()))
