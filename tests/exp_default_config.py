"""The reference's DEFAULT era5-svd configuration (config.ini: randomized, delay_embedding = 2, mean_center, n_components = 10)
at the c2 grid: device time and parity against the oracle on a row sample.  Run by hand on a GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field
from oracle.compare import sigma_rel_err, vector_angles
from oracle.slice_tools_np import delay_embed_np
from oracle.svd_ref import randomized_svd_ref

ops = CudaOps("cuda:0")
k, d = 10, 2
for rows, T in ((1038240, 744), (1038240, 97)):
    field = synthetic_field(T, rows, device="cuda", seed=5, rank=min(160, T - 2))
    def step():
        built = build_matrix_device(ops, [field], mean_center=True, scale=False)
        return built, svd_device(ops, built.X, svd_type="randomized", n_components=k, delay=d, seed=1, precision="auto")
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): built, (U, s, V) = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{rows} x {T}, d = {d}, k = {k}: {ms:.2f} ms per build + SVD = {rows * T * 4 / 1e9 / (ms / 1e3):.0f} GB/s of stored matrix")
    # parity on a row sample (the oracle cannot take a million rows quickly)
    sub = 65536
    Xs = built.X[:sub].cpu().numpy()
    Us, ss, Vs = svd_device(ops, built.X[:sub].contiguous(), svd_type="randomized", n_components=k, delay=d, seed=1, precision="auto")
    U0, s0, V0 = randomized_svd_ref(delay_embed_np(Xs.astype(np.float64), d), k, 1)
    print(f"   sample {sub} rows: sigma {sigma_rel_err(ss.cpu().numpy(), s0):.2e}, angle U {vector_angles(Us.cpu().numpy(), U0).max():.2e} V {vector_angles(Vs.cpu().numpy().T, V0.T).max():.2e}")
