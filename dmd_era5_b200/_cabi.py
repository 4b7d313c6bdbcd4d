"""ctypes binding of ``libera5svd.so`` (the C ABI declared in ``include/era5svd.h``).

There is NO fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C dmd_era5_b200/csrc``)
``lib()`` raises, and every compute call raises ``Era5SvdError`` when the CUDA call fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libera5svd.so")

F32, F64 = 0, 1
PREC_NATIVE, PREC_TF32X3 = 0, 1
BUILD_MEAN_CENTER, BUILD_SCALE, BUILD_CHECK_FINITE = 1, 2, 4

_i64, _int, _dbl, _vp, _sz = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/era5svd.h
PROTOTYPES = {
    "era5svd_version": (_int, []),
    "era5svd_last_error": (C.c_char_p, []),
    "era5svd_launch_count": (C.c_ulonglong, []),
    "era5svd_probe_tf32_tflops": (_int, [_int, _int, _dbl, _vp]),
    "era5svd_probe_dmma_tflops": (_int, [_dbl, _vp]),
    "era5svd_build_rows": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _vp, _vp, C.c_uint, _vp, _vp]),
    "era5svd_build_rows_split": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.c_uint, _vp, _vp]),
    "era5svd_check_finite": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp]),
    "era5svd_sketch": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _int, _vp]),
    "era5svd_project_workspace_bytes": (_sz, [_int, _i64, _i64, _i64, _int]),
    "era5svd_project": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _int, _int, _vp, _sz, _vp]),
    "era5svd_split_tf32": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp]),
    "era5svd_sketch_tf32x3_workspace_bytes": (_sz, [_i64, _i64]),
    "era5svd_sketch_tf32x3": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "era5svd_sketch_tf32x2": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "era5svd_sketch_tf32x1": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _sz, _vp]),
    "era5svd_project_tf32x1": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _int, _vp, _sz, _vp]),
    "era5svd_project_tf32x2": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _int, _vp, _sz, _vp]),
    "era5svd_round_tf32_f64": (_int, [_vp, _i64, _i64, _i64, _vp]),
    "era5svd_project_tf32x3_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "era5svd_project_tf32x3": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _int, _vp, _sz, _vp]),
    "era5svd_gemm_f64": (_int, [_int, _int, _i64, _i64, _i64, _dbl, _vp, _i64, _vp, _i64, _dbl, _vp, _i64, _vp]),
    "era5svd_syevj_workspace_bytes": (_sz, [_i64]),
    "era5svd_syevj_f64": (_int, [_vp, _i64, _i64, _vp, _vp, _i64, _int, _dbl, _vp, _sz, _vp]),
    "era5svd_tridiag_reduce_workspace_bytes": (_sz, [_i64]),
    "era5svd_tridiag_reduce_f64": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "era5svd_tridiag_eig_topk_workspace_bytes": (_sz, [_i64, _i64]),
    "era5svd_tridiag_eig_topk_f64": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _sz, _vp]),
    "era5svd_tridiag_apply_f64": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    "era5svd_tridiag_backtransform_f64": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "era5svd_bop_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "era5svd_bop_iterate_f64": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _dbl, _dbl, _int, _vp, _sz, _vp]),
    "era5svd_chol_inv_f64": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _vp]),
    "era5svd_col_normalize_f64": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "era5svd_sigma_from_eig_f64": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "era5svd_convert": (_int, [_vp, _int, _i64, _vp, _int, _i64, _i64, _i64, _vp]),
    "era5svd_comm_handle_bytes": (_sz, []),
    "era5svd_comm_create": (_int, [_int, _int, _i64, C.POINTER(_vp), _vp]),
    "era5svd_comm_connect": (_int, [_vp, _vp]),
    "era5svd_comm_destroy": (_int, [_vp]),
    "era5svd_comm_capacity": (_i64, [_vp]),
    "era5svd_comm_fused_count": (C.c_ulonglong, [_vp]),
    "era5svd_comm_allreduce_f64": (_int, [_vp, _vp, _i64, _vp]),
    "era5svd_comm_allgather_f64": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "era5svd_comm_fuse_next_project": (_int, [_vp, _i64, _i64]),
    "era5svd_col_absmax_workspace_bytes": (_sz, [_i64, _i64]),
    "era5svd_col_absmax": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "era5svd_maxloc_combine": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "era5svd_scale_cols": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp]),
    "era5svd_scale_rows_f64": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
}


class Era5SvdError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib = None


def lib() -> C.CDLL:
    """Load libera5svd.so once; fail loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(__graft_entry__.build() or `make -C dmd_era5_b200/csrc`). "
                "dmd_era5_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().era5svd_last_error().decode("utf-8", "replace")
        raise Era5SvdError(f"{what} failed with status {status}: {msg}")


def launch_count() -> int:
    return int(lib().era5svd_launch_count())
