"""One sketch + one projection on the FP64 tensor cores (DMMA) at a reduced c2 shape, for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:dmma -s 2 -c 2 -o gpurun_out/dmma python scripts/prof_dmma.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200._cabi import PREC_NATIVE
ops = CudaOps("cuda:0")
m, n, l = 262144, 744, 110
X = torch.randn((m, n), device="cuda", dtype=torch.float64)
Om = torch.randn((n, l), device="cuda", dtype=torch.float64)
Y = torch.empty((m, l), device="cuda", dtype=torch.float64)
for _ in range(2):
    ops.sketch(X, Om, Y, PREC_NATIVE)
    ops.project(X, Y, None, precision=PREC_NATIVE)
torch.cuda.synchronize()
print("ok")
