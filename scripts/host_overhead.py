"""How long does the HOST need to enqueue one c2 step (no synchronisation), against the GPU time of the step?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

ops = CudaOps("cuda:0")
T, S = 744, 721 * 1440
field = synthetic_field(T, S, device="cuda", seed=1000)

def step():
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    return svd_device(ops, built.X, svd_type="randomized", n_components=100, seed=1, precision="auto")

for _ in range(3): step()
torch.cuda.synchronize()
host, total = [], []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3); total.append((t2 - t0) * 1e3)
print(f"host enqueue per step: {sorted(host)[len(host)//2]:.2f} ms (min {min(host):.2f}); enqueue + GPU drain: {sorted(total)[len(total)//2]:.2f} ms")
# the same step captured once into a CUDA graph and replayed
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        out = step()
    torch.cuda.synchronize()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"CUDA graph replay: {e0.elapsed_time(e1) / 10:.2f} ms per step; sigma_1 = {float(out[1][0]):.10f}")
except Exception as ex:
    print("graph capture failed:", repr(ex)[:300])
