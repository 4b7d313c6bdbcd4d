"""dmd_era5_b200 - B200-native (sm_100a) implementation of the DMD-ERA5 SVD stage.

Scope: the hot path named by BASELINE.json: matrix build, randomized /
standard SVD, packaging - behind the reference's own function names.  Everything numerical runs in
hand-written CUDA reached through the C ABI in ``include/era5svd.h``; there is no CPU fallback.
"""
from ._cabi import Era5SvdError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"

__all__ = ["Era5SvdError", "LIB_PATH", "__version__"]
