"""Wall-clock phases of the drop-in `main()` path (slice file -> Dataset -> device build + SVD -> packaging -> NetCDF)
at a c2-like size: where a user's time goes outside the kernels.  Usage: python scripts/time_main.py [T] [A] [O]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

T = int(sys.argv[1]) if len(sys.argv) > 1 else 744
A = int(sys.argv[2]) if len(sys.argv) > 2 else 721
O = int(sys.argv[3]) if len(sys.argv) > 3 else 1440
root = tempfile.mkdtemp(prefix="era5svd_main_")
os.environ["DMD_ERA5_ROOT"] = root

from dmd_era5_b200 import stage
from dmd_era5_b200.config_parser import config_parser
from dmd_era5_b200.dataset import DataArray, Dataset, write_netcdf

cfg = {"source_path": "gs://mock", "variables": "temperature", "levels": "1000", "svd_type": "randomized",
       "delay_embedding": 1, "mean_center": True, "scale": False, "start_datetime": "2019-01-01T00",
       "end_datetime": "2019-02-01T00", "delta_time": "1h", "n_components": 100, "save_data_matrix": False,
       "random_seed": 1}
parsed = config_parser(cfg, "era5-svd")
rng = np.random.default_rng(0)
t0 = time.perf_counter()
base = (250 + 30 * np.cos(np.linspace(-np.pi / 2, np.pi / 2, A))[:, None] * np.ones((1, O))).astype(np.float32)
modes = rng.standard_normal((40, A * O)).astype(np.float32)
coef = (rng.standard_normal((T, 40)) * (0.9 ** np.arange(40))).astype(np.float32)
field = (coef @ modes).reshape(T, 1, A, O) + base[None, None]
field += 0.01 * rng.standard_normal(field.shape, dtype=np.float32)
times = np.datetime64("2019-01-01T00") + np.arange(T) * np.timedelta64(1, "h")
ds = Dataset({"temperature": DataArray(field, ("time", "level", "latitude", "longitude"))},
             {"time": times.astype("datetime64[ns]"), "level": np.array([1000]), "latitude": np.linspace(90, -90, A),
              "longitude": np.arange(O) * (360.0 / O)},
             {"source_path": cfg["source_path"], "variables": ["temperature"], "levels": [1000]})
print(f"synthetic slice {field.nbytes / 1e9:.2f} GB generated in {time.perf_counter() - t0:.1f} s", flush=True)
t0 = time.perf_counter(); write_netcdf(ds, parsed["era5_slice_path"]); print(f"slice file written in {time.perf_counter() - t0:.1f} s", flush=True)
del ds, field

stage.get_ops("cuda:0")                      # library load / context creation outside the timings
torch.cuda.synchronize()
for rep in range(2):
    t = {}
    t0 = time.perf_counter(); ds, _ = stage.retrieve_era5_slice(parsed); t["read slice"] = time.perf_counter() - t0
    t0 = time.perf_counter(); res = stage._compute(ds, parsed); torch.cuda.synchronize(); t["compute (upload, build, SVD, download, packaging)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); write_netcdf(res, parsed["save_path"]); t["write result"] = time.perf_counter() - t0
    print(f"run {rep}: " + "; ".join(f"{k} {v:.2f} s" for k, v in t.items()), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); res = stage._compute(ds, parsed); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
