"""CPU: the C-ABI shared library loads and exports every symbol include/era5svd.h declares
(no compute calls - there is no GPU here), and the product path refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

from dmd_era5_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "era5svd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(era5svd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in era5svd.h but not exported"
    assert sorted(_cabi.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"
    assert lib.era5svd_version() == 100


def test_argument_validation_without_gpu():
    lib = _cabi.lib()
    # null pointers are rejected before any CUDA call
    rc = lib.era5svd_sketch(None, 0, 1, 1, 1, None, 1, 1, None, 1, 0, None)
    assert rc == -1 and b"null" in lib.era5svd_last_error()
    assert lib.era5svd_project_workspace_bytes(0, 1038240, 744, 110, 0) > 0
    assert lib.era5svd_syevj_workspace_bytes(110) == 2 * 110 * 111 * 8 + 16   # odd leading dimension + status word


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from dmd_era5_b200.device_ops import CudaOps
    from dmd_era5_b200.era5_svd import svd_on_era5

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        svd_on_era5(np.zeros((8, 4)), {"svd_type": "standard", "n_components": 2})
    with pytest.raises(RuntimeError):
        CudaOps("cpu")
    with pytest.raises(ValueError, match="SVD type foo is not supported."):
        svd_on_era5(np.zeros((8, 4)), {"svd_type": "foo", "n_components": 2})
