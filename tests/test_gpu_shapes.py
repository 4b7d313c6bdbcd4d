"""GPU: parity on BASELINE.json's OWN shapes at reduced rows ("next" item 1).

  c3 shape : n = 1460 snapshots, k = 100  ->  l = 110, n_iter = 7 (16 passes over X), d in {1, 2}
             every float32 path (3xTF32, mixed 3x / 1x TF32, FP32 FMA) and the FP64 DMMA path against the reference's
             own call (sklearn randomized_svd, era5_svd.py:258) on identical inputs and the same seeded test matrix
  c4 shape : n = 8760 snapshots, standard SVD  -> against the committed golden vectors of np.linalg.svd (era5_svd.py:251)
  plus     : an exactly rank-deficient input (rank < l) end to end, float32 input against the reference's own float32 call.

Tolerances (BASELINE.json north_star): sigma 1e-6 relative in FP64 mode, 1e-4 in the FP32-split modes; vectors up to sign,
principal angle < 1e-5 rad in FP64 mode; reconstruction error within 1 % of the reference's.
"""
import functools
import os

import numpy as np
import pytest
import torch

from dmd_era5_b200.era5_svd import get_ops, host_to_device_matrix, svd_on_era5
from dmd_era5_b200.pipeline import svd_device
from oracle.compare import recon_rel_err, sigma_rel_err, signs_agree, vector_angles
from oracle.slice_tools_np import delay_embed_np
from oracle.svd_ref import randomized_svd_ref
from oracle.synthetic_np import lowrank_field_np

pytestmark = pytest.mark.gpu

N_C3, K, ROWS_C3 = 1460, 100, 40000


@functools.lru_cache(maxsize=None)
def c3_case(d: int):
    """float32 data of the c3 shape (reduced rows) + the float64 oracle on the same values."""
    X = lowrank_field_np(ROWS_C3, N_C3, r=160, rho=0.93, seed=4, dtype=np.float32)
    Xd = delay_embed_np(X.astype(np.float64), d)
    return X, Xd, randomized_svd_ref(Xd, K, 1)


def run_device(X, d, precision, **kw):
    ops = get_ops()
    stats = {}
    U, s, V = svd_device(ops, host_to_device_matrix(ops, X), svd_type="randomized", n_components=K, delay=d, seed=1,
                         precision=precision, stats=stats, **kw)
    return U.cpu().numpy(), s.cpu().numpy(), V.cpu().numpy(), stats


# measured on B200 (this file's own printout, profiles/r02_parity_shapes.txt): see the asserts below
@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("precision", ["tf32x3", "tf32mix", "native"])
def test_c3_shape_float32_paths(precision, d):
    X, Xd, (U0, s0, V0) = c3_case(d)
    U, s, V, stats = run_device(X, d, precision)
    assert stats["tall_passes"] == 16                       # n_iter = 7: k < 0.1 * min(m, n) (extmath.py:586-589)
    if precision == "tf32mix":
        assert stats["low_precision_iters"] == 6
    err = sigma_rel_err(s, s0)
    ang_u, ang_v = vector_angles(U, U0), vector_angles(V.T, V0.T)
    ref = recon_rel_err(Xd, U0, s0, V0)
    rec = recon_rel_err(Xd, U, s, V)
    print(f"\nc3-shape {precision} d={d}: sigma {err:.2e}  angle U first50 {ang_u[:50].max():.2e} all {ang_u.max():.2e}  "
          f"V first50 {ang_v[:50].max():.2e} all {ang_v.max():.2e}  recon {rec:.6e} vs {ref:.6e}")
    assert err < 1e-4
    # measured (profiles/r02_parity_shapes.txt): first 50 <= 3.0e-5, all <= 6.4e-5 rad on every float32 path
    assert ang_u[:50].max() < 1e-4 and ang_u.max() < 3e-4
    assert ang_v[:50].max() < 1e-4 and ang_v.max() < 3e-4
    assert signs_agree(U, U0)
    assert abs(rec - ref) <= 0.01 * ref


def test_c3_shape_fp64_dmma():
    """FP64 mode at n = 1460, q = 7: sigma 1e-6, angles 1e-5 rad."""
    X = lowrank_field_np(20000, N_C3, r=160, rho=0.93, seed=5)
    U0, s0, V0 = randomized_svd_ref(X, K, 2)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": K, "random_seed": 2})
    err = sigma_rel_err(s, s0)
    ang = max(vector_angles(U, U0).max(), vector_angles(V.T, V0.T).max())
    print(f"\nc3-shape fp64: sigma {err:.2e} angle {ang:.2e}")
    assert err < 1e-6 and ang < 1e-5 and signs_agree(U, U0)
    ref = recon_rel_err(X, U0, s0, V0)
    assert abs(recon_rel_err(X, U, s, V) - ref) <= 0.01 * ref


def test_c4_shape_standard_fp64_golden(golden_dir):
    """n = 8760 (hourly year), standard SVD by the Gram route + the all-SM tridiagonal eigensolver, float64, against
    np.linalg.svd's committed output (tests/golden/make_golden_c4.py)."""
    g = np.load(os.path.join(golden_dir, "svd_standard_n8760.npz"))
    m, n, r, seed, ks, kv = [int(x) for x in g["meta"]]
    X = lowrank_field_np(m, n, r=r, rho=float(g["rho"]), seed=seed)
    U, s, V = svd_on_era5(X, {"svd_type": "standard", "n_components": ks})
    assert U.shape == (m, ks) and V.shape == (ks, n)
    err = sigma_rel_err(s, g["s"])
    ang_u = vector_angles(U[:, :kv], g["U"]).max()
    ang_v = vector_angles(V[:kv].T, g["V"].T).max()
    print(f"\nc4-shape standard fp64: sigma {err:.2e} ({ks} values)  angle U {ang_u:.2e} V {ang_v:.2e} ({kv} pairs)")
    assert err < 1e-6
    assert ang_u < 1e-5 and ang_v < 1e-5
    assert np.max(np.abs(U.T @ U - np.eye(ks))) < 1e-8


def test_c4_shape_standard_float32():
    """The same shape from float32 storage (the real ERA5 dtype): the Gram matrix is accumulated in float64 from
    fp32-exact products, so sigma keeps 1e-4 against the float64 oracle values well below sigma_1."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "svd_standard_n8760.npz"))
    m, n, r, seed, ks, kv = [int(x) for x in g["meta"]]
    X = lowrank_field_np(m, n, r=r, rho=float(g["rho"]), seed=seed, dtype=np.float32)
    U, s, V = svd_on_era5(X, {"svd_type": "standard", "n_components": ks})
    err = np.abs(s.astype(np.float64) - g["s"]) / g["s"]
    ang_v = vector_angles(V[:kv].T, g["V"].T)
    print(f"\nc4-shape standard f32: sigma max {err.max():.2e} (first 12: {err[:12].max():.2e})  angle V {ang_v.max():.2e}")
    assert err.max() < 1e-4
    assert ang_v.max() < 1e-3
    assert np.max(np.abs(U.astype(np.float64).T @ U.astype(np.float64) - np.eye(ks))) < 1e-3


@pytest.mark.parametrize("dtype,precision", [(np.float64, "native"), (np.float32, "tf32x3"), (np.float32, "tf32mix"),
                                              (np.float32, "native")])
def test_exactly_rank_deficient_input(dtype, precision):
    """rank(X) = 40 < l = 60: the trailing sketch directions are exactly dependent, Cholesky drops their pivots.
    The reference returns ~eps * sigma_1 for the trailing values; sigma and vectors of the true rank must match."""
    rng = np.random.RandomState(3)
    rank, m, n, k = 40, 6000, 200, 50
    A = np.linalg.qr(rng.standard_normal((m, rank)))[0]
    B = np.linalg.qr(rng.standard_normal((n, rank)))[0]
    X = ((A * (50.0 * 0.8 ** np.arange(rank))) @ B.T).astype(dtype)
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), k, 6)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": k, "random_seed": 6, "precision": precision})
    assert np.all(np.isfinite(U)) and np.all(np.isfinite(s)) and np.all(np.isfinite(V))
    f64 = dtype == np.float64
    assert sigma_rel_err(s[:rank], s0[:rank]) < (1e-6 if f64 else 1e-4)
    assert np.all(np.abs(s[rank:]) < (1e-9 if f64 else 1e-6) * s0[0])            # reference: ~1e-15 * sigma_1
    ang = vector_angles(U[:, :rank], U0[:, :rank])
    print(f"\nrank-deficient {np.dtype(dtype).name} {precision}: sigma {sigma_rel_err(s[:rank], s0[:rank]):.2e} "
          f"trailing {np.abs(s[rank:]).max():.2e} angle {ang.max():.2e}")
    assert ang.max() < (1e-5 if f64 else 5e-3)
    ref = recon_rel_err(X, U0, s0, V0)
    assert recon_rel_err(X, U, s, V) <= max(1.01 * ref, 1e-12 if f64 else 2e-6)


@pytest.mark.parametrize("precision", ["tf32x3", "tf32mix", "native"])
def test_float32_input_against_the_references_own_float32_call(precision):
    """Real ERA5 is float32, and then the reference itself computes in float32 (extmath.py:324-334 casts Omega; BLAS
    sgemm, sgetrf, sgeqrf, sgesdd).  Our float32 paths must (a) agree with that call to the float32 tolerance and
    (b) be at least as close to the float64 truth as the reference's own float32 result is."""
    X = lowrank_field_np(20000, 744, r=160, rho=0.93, seed=9, dtype=np.float32)
    U32, s32, V32 = randomized_svd_ref(X, K, 1)                       # the reference's float32 computation
    U64, s64, V64 = randomized_svd_ref(X.astype(np.float64), K, 1)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": K, "random_seed": 1, "precision": precision})
    ours_vs_ref32 = sigma_rel_err(s, s32)
    ours_vs_64 = sigma_rel_err(s, s64)
    ref32_vs_64 = sigma_rel_err(s32, s64)
    a_ours = vector_angles(U, U64)
    a_ref = vector_angles(U32, U64)
    print(f"\nf32 {precision}: sigma ours-ref32 {ours_vs_ref32:.2e} ours-f64 {ours_vs_64:.2e} ref32-f64 {ref32_vs_64:.2e}  "
          f"angle ours-f64 {a_ours.max():.2e} ref32-f64 {a_ref.max():.2e}")
    assert ours_vs_ref32 < 1e-4 and ours_vs_64 < 1e-4
    assert ours_vs_64 <= max(10 * ref32_vs_64, 2e-6)
    assert a_ours.max() <= max(2 * a_ref.max(), 1e-3)
    # signs follow the float64 truth; the reference's own float32 run may flip a column whose two largest |entries|
    # tie within float32 noise (svd_flip picks the first maximum, extmath.py:964-972)
    assert signs_agree(U, U64)
    flips = int(np.sum(np.sum(U32.astype(np.float64) * U64, axis=0) < 0))
    assert int(np.sum(np.sum(U.astype(np.float64) * U32, axis=0) < 0)) == flips


@pytest.mark.parametrize("layout", ["packed_odd_pitch", "transposed_view"])
def test_svd_device_accepts_unaligned_device_matrices(layout):
    """A caller's own device matrix - tightly packed with an odd row pitch (T = 97: the reference's default 4-day hourly
    slice), or a transposed view - is repacked into the padded layout instead of failing in the TMA set-up; the
    reference's default configuration otherwise (randomized, delay_embedding = 2, n_components = 10, config.ini)."""
    ops = get_ops()
    m, n, k, d = 20000, 97, 10, 2
    X = lowrank_field_np(m, n, r=40, rho=0.85, seed=21, dtype=np.float32)
    Xd = torch.from_numpy(X).cuda() if layout == "packed_odd_pitch" else torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()
    assert Xd.shape == (m, n) and (Xd.stride(0) % 4 != 0 or Xd.stride(1) != 1)
    U, s, V = svd_device(ops, Xd, svd_type="randomized", n_components=k, delay=d, seed=1, precision="auto")
    U0, s0, V0 = randomized_svd_ref(delay_embed_np(X.astype(np.float64), d), k, 1)
    assert sigma_rel_err(s.cpu().numpy(), s0) < 1e-4
    assert vector_angles(U.cpu().numpy(), U0).max() < 1e-3 and vector_angles(V.cpu().numpy().T, V0.T).max() < 1e-3
    assert signs_agree(U.cpu().numpy(), U0)
