"""GPU: shapes a user of the reference can hand to svd_on_era5 (era5_svd.py:230-263) beyond the tall benchmark shapes -
wide matrices (sklearn transposes them, extmath.py:591-595), n_components larger than the matrix allows (the reference's
slices silently return fewer, era5_svd.py:252-254), k + 10 > n, a single column, a handful of rows, row counts that are no
multiple of any tile.  Result shapes must equal the reference's, values must match its float64 run."""
import numpy as np
import pytest

from dmd_era5_b200.era5_svd import svd_on_era5
from oracle.compare import sigma_rel_err, vector_angles
from oracle.svd_ref import randomized_svd_ref, standard_svd_ref

pytestmark = pytest.mark.gpu


def lowrank(m, n, r, dtype, seed):
    rng = np.random.RandomState(seed)
    r = min(r, m, n)
    A = np.linalg.qr(rng.standard_normal((m, r)))[0]
    B = np.linalg.qr(rng.standard_normal((n, r)))[0]
    return ((A * (30.0 * 0.8 ** np.arange(r))) @ B.T + 1e-6 * rng.standard_normal((m, n))).astype(dtype)


CASES = [("wide randomized f32", 500, 2000, 10, "randomized", np.float32), ("wide randomized f64", 300, 1200, 8, "randomized", np.float64),
         ("wide standard f64", 200, 900, 6, "standard", np.float64), ("wide standard f32", 200, 900, 6, "standard", np.float32),
         ("k > n randomized f64", 4000, 12, 20, "randomized", np.float64), ("k > n standard f64", 4000, 12, 20, "standard", np.float64),
         ("k = n standard f32", 3000, 16, 16, "standard", np.float32), ("tiny m", 7, 5, 3, "randomized", np.float64),
         ("single column", 1000, 1, 1, "standard", np.float64), ("k + 10 > n randomized f32", 5000, 25, 20, "randomized", np.float32),
         ("ragged m f32", 12345, 333, 17, "randomized", np.float32)]


@pytest.mark.parametrize("name,m,n,k,kind,dtype", CASES, ids=[c[0] for c in CASES])
def test_edge_shapes_match_the_reference(name, m, n, k, kind, dtype):
    X = lowrank(m, n, 30, dtype, seed=m + n)
    U, s, V = svd_on_era5(X, {"svd_type": kind, "n_components": k, "random_seed": 3})
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), k, 3) if kind == "randomized" else standard_svd_ref(X.astype(np.float64), k)
    assert U.shape == U0.shape and s.shape == s0.shape and V.shape == V0.shape
    assert U.dtype == X.dtype
    assert np.isfinite(U).all() and np.isfinite(s).all() and np.isfinite(V).all()
    good = s0 > 1e-4 * s0[0]                      # components above the noise floor of the float32 runs
    f64 = dtype == np.float64
    assert sigma_rel_err(s[good], s0[good]) < (1e-6 if f64 else 1e-4)
    assert vector_angles(U[:, good], U0[:, good]).max() < (1e-5 if f64 else 1e-3)
    assert vector_angles(V[good].T, V0[good].T).max() < (1e-5 if f64 else 1e-3)
