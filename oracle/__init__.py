"""CPU oracle for the DMD-ERA5 SVD stage.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``dmd_era5_b200``) never imports it and fails loudly when its CUDA library is
missing.

Parity pinning status:
  * delay embedding: pinned by the reference's own 4 integer known-answer cases
    (tests/test_02_slice_tools.py:215-231) and by golden vectors generated from the
    reference's own ``_apply_delay_embedding_np`` source (tests/golden/make_golden.py).
  * standardize / flatten / coords: pinned by the properties the reference's tests
    assert (mean 0 / std 1 / ddof=0, row order, coord tiling); the reference package
    itself cannot be imported here (xarray, dvc, pyprojroot absent, no network).
  * SVD numerics: the reference's tests hold shapes only -> "parity unpinned" by
    reference fixtures; the oracle therefore *executes the very same library calls*
    the reference makes (numpy.linalg.svd, sklearn.utils.extmath.randomized_svd,
    scikit-learn 1.9.0) on identical inputs, and the golden vectors under
    tests/golden/ were produced by those calls.
"""
