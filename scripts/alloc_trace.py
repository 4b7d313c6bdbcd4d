"""Which allocation makes the caching allocator call cudaMalloc in the middle of a run of identical c2 steps?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field

ops = CudaOps("cuda:0")
T, S = 744, 721 * 1440
field = synthetic_field(T, S, device="cuda", seed=1000, total_points=S)

def step():
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    return svd_device(ops, built.X, svd_type="randomized", n_components=100, seed=1, precision="auto")

def segs():
    return sorted((s["address"], s["total_size"]) for s in torch.cuda.memory_snapshot())

for _ in range(5):
    step()                       # bench-style warm-up: results dropped at once
prev = segs()
for i in range(12):
    n0 = torch.cuda.memory_stats()["num_device_alloc"]; f0 = torch.cuda.memory_stats().get("num_device_free", 0)
    t0 = time.perf_counter()
    U, s, V = step()
    dt = (time.perf_counter() - t0) * 1e3
    st = torch.cuda.memory_stats()
    cur = segs()
    new = [x for x in cur if x not in prev]; gone = [x for x in prev if x not in cur]
    if st["num_device_alloc"] != n0 or new or gone or dt > 5:
        print(f"step {i}: host {dt:.1f} ms, cudaMalloc +{st['num_device_alloc'] - n0}, cudaFree +{st.get('num_device_free', 0) - f0}, "
              f"new segments {[(hex(a), sz) for a, sz in new]}, released {[(hex(a), sz) for a, sz in gone]}, "
              f"reserved {st['reserved_bytes.all.current'] / 2**30:.2f} GiB allocated {st['allocated_bytes.all.current'] / 2**30:.2f} GiB "
              f"inactive_split {st['inactive_split_bytes.all.current'] / 2**20:.1f} MiB")
    prev = cur
torch.cuda.synchronize()
print("done; segments:", [(sz) for a, sz in segs()])
