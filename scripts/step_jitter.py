"""Per-step device times of the c2 step (one CUDA event pair per step), with and without the NVML clock sampler thread
of bench.py, to see where run-to-run variance of the bench value comes from."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

ops = CudaOps("cuda:0")
T, S = 744, 721 * 1440
field = synthetic_field(T, S, device="cuda", seed=1000, total_points=S)

def step():
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    return svd_device(ops, built.X, svd_type="randomized", n_components=100, seed=1, precision="auto")

def run(label, n=40, sampler_period=None):
    smp = None
    if sampler_period is not None:
        smp = bench.ClockSampler(0, uuid=str(torch.cuda.get_device_properties(0).uuid))
        bench.ClockSampler.PERIOD = sampler_period
        smp.start()
    for _ in range(5): step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        step(); ev[i + 1].record()
    torch.cuda.synchronize()
    if smp: smp.stop()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
    print(f"{label:34s} mean {ms.mean():6.2f}  median {np.median(ms):6.2f}  min {ms.min():6.2f}  max {ms.max():6.2f}  "
          f"steps > 12 ms: {(ms > 12).sum()}  sorted tail {np.sort(ms)[-4:].round(2)}")

for rep in range(2):
    run("no sampler")
    run("NVML sampler every 5 ms", sampler_period=0.005)
    run("NVML sampler every 50 ms", sampler_period=0.05)
