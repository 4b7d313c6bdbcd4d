"""dmd_era5_b200 - B200-native (sm_100a) implementation of the DMD-ERA5 SVD stage.

Scope: the hot path named by BASELINE.json: matrix build, randomized /
standard SVD, packaging - behind the reference's own function names.  Everything numerical runs in
hand-written CUDA reached through the C ABI in ``include/era5svd.h``; there is no CPU fallback.

The package keeps the reference's import surface for this path (src/dmd_era5/__init__.py:22-38,
era5_svd/__init__.py:10-17, slice_tools/__init__.py:11-19), so ``from dmd_era5_b200 import
slice_era5_dataset, config_parser, add_data_to_dvc`` and ``from dmd_era5_b200.era5_svd import main, svd_on_era5``
read like the reference's own imports.  Names are resolved lazily (PEP 562) so that importing the package neither
loads torch nor needs the CUDA library.  Out of scope and therefore absent: ``download_era5_data`` (the network
stage), ``create_mock_era5`` / ``create_mock_era5_svd`` (test-data generators; their seeded restatement lives in
``oracle/``).
"""
from ._cabi import Era5SvdError, LIB_PATH  # noqa: F401
# like the reference (``from dmd_era5.core import config_parser`` after the submodule import), the FUNCTION shadows the
# submodule of the same name on the package; the module stays reachable through sys.modules / ``from . import``
from .config_parser import config_parser, config_reader  # noqa: F401

__version__ = "0.2.0"

# reference name -> module of this package that defines it
_EXPORTS = {
    # slice_tools (src/dmd_era5/slice_tools/__init__.py:11-19)
    "slice_era5_dataset": "slice_tools", "resample_era5_dataset": "slice_tools", "standardize_data": "slice_tools",
    "apply_delay_embedding": "slice_tools", "flatten_era5_variables": "slice_tools",
    "_apply_delay_embedding_np": "slice_tools", "space_coord_to_level_lat_lon": "slice_tools",
    # core (src/dmd_era5/core.py)
    "log_and_print": "era5_svd", "setup_logger": "era5_svd",
    # dvc_tools (src/dmd_era5/dvc_tools.py:50-63, :119-253)
    "add_data_to_dvc": "dvc_tools", "retrieve_data_from_dvc": "dvc_tools",
    # era5_svd (src/dmd_era5/era5_svd/__init__.py:10-17)
    "svd_on_era5": "era5_svd", "combine_svd_results": "stage", "retrieve_era5_slice": "stage",
    "retrieve_svd_results": "stage", "add_config_attributes": "stage", "main": "stage",
}

__all__ = ["Era5SvdError", "LIB_PATH", "__version__", "config_parser", "config_reader", *sorted(_EXPORTS)]


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    import importlib

    return getattr(importlib.import_module(f".{mod}", __name__), name)
