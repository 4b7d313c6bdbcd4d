"""GPU: end-to-end parity of the CUDA path (through the reference-facing API and the C ABI)
against the oracle = the reference's own library calls on identical inputs.

Tolerances (BASELINE.json north_star): singular values 1e-6 relative in FP64 mode (1e-4 in the
FP32 modes), singular vectors up to sign with principal angle < 1e-5 rad (FP64 mode),
reconstruction error within 1 % of the reference's."""
import os

import numpy as np
import pytest
import torch

from dmd_era5_b200.era5_svd import svd_on_era5
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from oracle.compare import (orthonormality, recon_rel_err, sigma_rel_err, signs_agree,
                            vector_angles)
from oracle.slice_tools_np import build_matrix_np, delay_embed_np
from oracle.svd_ref import randomized_svd_ref, standard_svd_ref
from oracle.synthetic_np import lowrank_field_np, mock_era5_np

pytestmark = pytest.mark.gpu

SIGMA_TOL_FP64 = 1e-6
ANGLE_TOL_FP64 = 1e-5
SIGMA_TOL_FP32 = 1e-4


def c1_matrix(d=2, mean_center=True, scale=False):
    ds = mock_era5_np(25, ["temperature", "u_component_of_wind"], [1000], seed=0)
    return build_matrix_np(list(ds["vars"].values()), mean_center, scale, d)[0], ds


def test_config1_standard_golden(golden_dir):
    """BASELINE config 1: mock ERA5, standard SVD, float64 - against the committed golden vectors."""
    g = np.load(os.path.join(golden_dir, "svd_standard_c1.npz"))
    X, _ = c1_matrix()
    U, s, V = svd_on_era5(X, {"svd_type": "standard", "n_components": 6})
    assert U.shape == (X.shape[0], 6) and s.shape == (6,) and V.shape == (6, X.shape[1])
    assert U.dtype == np.float64
    assert sigma_rel_err(s, g["s"]) < SIGMA_TOL_FP64
    assert vector_angles(U, g["U"]).max() < ANGLE_TOL_FP64
    assert vector_angles(V.T, g["V"].T).max() < ANGLE_TOL_FP64
    ref = recon_rel_err(X, g["U"], g["s"], g["V"])
    assert abs(recon_rel_err(X, U, s, V) - ref) <= 0.01 * ref


def test_config1_randomized_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "svd_randomized_c1.npz"))
    X, _ = c1_matrix()
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": 6, "random_seed": int(g["meta"][6])})
    assert sigma_rel_err(s, g["s"]) < SIGMA_TOL_FP64
    # mock data is white noise: the spectrum is flat, so compare the SUBSPACE-insensitive quantities
    ref = recon_rel_err(X, g["U"], g["s"], g["V"])
    assert abs(recon_rel_err(X, U, s, V) - ref) <= 0.01 * ref
    assert orthonormality(U) < 1e-10


def test_randomized_lowrank_fp64_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "svd_randomized_lowrank_f64.npz"))
    m, n, r, k, seed, rs = [int(x) for x in g["meta"]]
    X = lowrank_field_np(m, n, r=r, rho=0.8, seed=seed)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": k, "random_seed": rs})
    assert sigma_rel_err(s, g["s"]) < SIGMA_TOL_FP64
    assert vector_angles(U, g["U"]).max() < ANGLE_TOL_FP64
    assert vector_angles(V.T, g["V"].T).max() < ANGLE_TOL_FP64
    assert signs_agree(U, g["U"])


def test_randomized_fp64_era5_like_spectrum():
    """k = 100, l = 110 as in the bench configs, reduced rows; sigma_1/sigma_100 ~ 1.3e3."""
    X = lowrank_field_np(20000, 744, r=160, rho=0.93, seed=0)
    U0, s0, V0 = randomized_svd_ref(X, 100, 1)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": 100, "random_seed": 1})
    assert sigma_rel_err(s, s0) < SIGMA_TOL_FP64
    assert vector_angles(U, U0).max() < ANGLE_TOL_FP64
    assert vector_angles(V.T, V0.T).max() < ANGLE_TOL_FP64
    assert signs_agree(U, U0)
    ref = recon_rel_err(X, U0, s0, V0)
    assert abs(recon_rel_err(X, U, s, V) - ref) <= 0.01 * ref


def test_randomized_fp32_native():
    """float32 storage, FP32 FMA passes: 1e-4 on sigma against the float64 oracle on the same data."""
    X = lowrank_field_np(20000, 744, r=160, rho=0.93, seed=0, dtype=np.float32)
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), 100, 1)
    U, s, V = svd_on_era5(X, {"svd_type": "randomized", "n_components": 100, "random_seed": 1})
    assert U.dtype == np.float32 and s.dtype == np.float32
    assert sigma_rel_err(s, s0) < SIGMA_TOL_FP32
    ang = vector_angles(U, U0)
    assert ang[:50].max() < 1e-3 and ang.max() < 2e-2      # float32 vectors: eps * sigma_1 / gap
    assert signs_agree(U, U0)
    ref = recon_rel_err(X, U0, s0, V0)
    assert abs(recon_rel_err(X, U, s, V) - ref) <= 0.01 * ref


@pytest.mark.parametrize("svd_type", ["standard", "randomized"])
@pytest.mark.parametrize("mean_center,scale,d", [(True, False, 2), (True, True, 1), (False, False, 3)])
def test_device_build_plus_svd_vs_oracle(ops, svd_type, mean_center, scale, d):
    """Native (T, L, A, O) arrays -> device build (stack, centre, scale, virtual delay) -> SVD."""
    ds = mock_era5_np(25, ["temperature", "u_component_of_wind"], [1000, 500], seed=3)
    arrs = list(ds["vars"].values())
    Xref, mu_ref, sd_ref = build_matrix_np(arrs, mean_center, scale, d)
    blocks = [torch.from_numpy(a.reshape(a.shape[0], -1).copy()).cuda() for a in arrs]
    built = build_matrix_device(ops, blocks, mean_center=mean_center, scale=scale)
    Xdev = built.X.cpu().numpy()
    assert np.max(np.abs(delay_embed_np(Xdev, d) - Xref)) < 1e-10
    k = 8
    U, s, V = svd_device(ops, built.X, svd_type=svd_type, n_components=k, delay=d, seed=2)
    if svd_type == "standard":
        U0, s0, V0 = standard_svd_ref(Xref, k)
    else:
        U0, s0, V0 = randomized_svd_ref(Xref, k, 2)
    assert U.shape == (Xref.shape[0], k) and V.shape == (k, Xref.shape[1])
    assert sigma_rel_err(s.cpu().numpy(), s0) < SIGMA_TOL_FP64
    ref = recon_rel_err(Xref, U0, s0, V0)
    assert abs(recon_rel_err(Xref, U.cpu().numpy(), s.cpu().numpy(), V.cpu().numpy()) - ref) <= 0.01 * ref


@pytest.mark.parametrize("svd_type", ["standard", "randomized"])
@pytest.mark.parametrize("mean_center,scale,d", [(True, False, 2), (True, True, 1), (False, False, 3), (True, True, 3)])
def test_device_build_plus_svd_float32_slices(ops, svd_type, mean_center, scale, d):
    """The same chain on FLOAT32 slices (the real ERA5 dtype) for every route x delay combination, default precision
    ("auto": tensor-core passes; the standard route = Gram of the base matrix + refinement).  float32 + standard +
    delay > 1 once failed in the projection kernel's alignment check (a delay window as its Y operand) - found by
    tests/test_compute_phase_golden.py, fixed in standard.gram_device.  Oracle: float64 arithmetic on the same float32 data."""
    ds = mock_era5_np(25, ["temperature", "u_component_of_wind"], [1000, 500], seed=3)
    arrs = [a.astype(np.float32) for a in ds["vars"].values()]
    Xref, _, _ = build_matrix_np([a.astype(np.float64) for a in arrs], mean_center, scale, d)
    blocks = [torch.from_numpy(a.reshape(a.shape[0], -1).copy()).cuda() for a in arrs]
    built = build_matrix_device(ops, blocks, mean_center=mean_center, scale=scale)
    assert built.X.dtype == torch.float32
    Xdev = built.X.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(delay_embed_np(Xdev, d) - Xref)) < (1e-4 if mean_center else 1e-30)   # float32 storage of x - mean
    k = 8
    U, s, V = svd_device(ops, built.X, svd_type=svd_type, n_components=k, delay=d, seed=2)
    U, s, V = U.cpu().numpy().astype(np.float64), s.cpu().numpy().astype(np.float64), V.cpu().numpy().astype(np.float64)
    Xd = delay_embed_np(Xdev, d)                                       # the matrix the device factorised, exactly
    U0, s0, V0 = standard_svd_ref(Xd, k) if svd_type == "standard" else randomized_svd_ref(Xd, k, 2)
    assert U.shape == (Xref.shape[0], k) and V.shape == (k, Xref.shape[1])
    err = sigma_rel_err(s, s0)
    ref = recon_rel_err(Xd, U0, s0, V0)
    rec = recon_rel_err(Xd, U, s, V)
    print(f"float32 {svd_type} mc={mean_center} sc={scale} d={d}: sigma {err:.2e} recon {rec:.6e} (reference {ref:.6e})")
    assert err < SIGMA_TOL_FP32
    assert abs(rec - ref) <= 0.01 * ref
    assert np.max(np.abs(U.T @ U - np.eye(k))) < 1e-4


def test_invariants_at_bench_shape_reduced_rows(ops):
    """Size-independent properties on a float32 device-generated field (no CPU oracle):
    orthonormal U and V, X^T U = V^T S, descending sigma."""
    from dmd_era5_b200.synthetic import synthetic_field

    T, S, k = 744, 120000, 100
    field = synthetic_field(T, S, device="cuda", seed=0)
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    U, s, V = svd_device(ops, built.X, svd_type="randomized", n_components=k, seed=1)
    torch.cuda.synchronize()
    U64, V64, s64 = U.double(), V.double(), s.double()
    assert torch.all(s64[:-1] >= s64[1:])
    assert float((U64.t() @ U64 - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max()) < 5e-5
    assert float((V64 @ V64.t() - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max()) < 5e-5
    resid = built.X.double().t() @ U64 - V64.t() * s64
    assert float(resid.norm() / s64.norm()) < 1e-4
    # spectrum of the generator: sigma_i ~ 100 * 0.93**i
    expect = 100.0 * 0.93 ** torch.arange(k, device="cuda", dtype=torch.float64)
    assert float(((s64 - expect).abs() / expect).max()) < 0.15


def test_invariants_at_full_c2_size_tf32x3(ops):
    """BASELINE configs[1] at FULL size (1 038 240 x 744 float32, k = 100) on the tensor-core path; the CPU oracle
    cannot run this in test time, so size-independent properties are checked: descending sigma, orthonormal U / V,
    X^T U = V^T S, linearity (scaling X by 4 scales sigma by 4 exactly and leaves U, V unchanged bit for bit),
    run-to-run bitwise reproducibility, and agreement of sigma with the FP32-FMA path (different kernels, same data)."""
    from dmd_era5_b200.synthetic import synthetic_field

    T, S, k = 744, 721 * 1440, 100
    field = synthetic_field(T, S, device="cuda", seed=1000)
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    del field
    X = built.X
    U, s, V = svd_device(ops, X, svd_type="randomized", n_components=k, seed=1, precision="tf32x3")
    U2, s2, V2 = svd_device(ops, X, svd_type="randomized", n_components=k, seed=1, precision="tf32x3")
    assert torch.equal(s, s2) and torch.equal(V, V2) and torch.equal(U, U2)
    eye = torch.eye(k, device="cuda", dtype=torch.float64)
    assert torch.all(s[:-1] >= s[1:])
    UtU = torch.zeros((k, k), device="cuda", dtype=torch.float64)
    XtU = torch.zeros((T, k), device="cuda", dtype=torch.float64)
    for r0 in range(0, S, 1 << 17):                      # float64 reductions in row chunks (bounded temporaries)
        Uc = U[r0 : r0 + (1 << 17)].double()
        UtU += Uc.t() @ Uc
        XtU += X[r0 : r0 + (1 << 17)].double().t() @ Uc
    assert float((UtU - eye).abs().max()) < 5e-5
    assert float((V @ V.t() - eye).abs().max()) < 5e-5
    assert float((XtU - V.t() * s).norm() / s.norm()) < 1e-4
    # scaling by a power of two is exact in every kernel of the path
    X.mul_(4.0)
    U4, s4, V4 = svd_device(ops, X, svd_type="randomized", n_components=k, seed=1, precision="tf32x3")
    assert torch.equal(s4, 4.0 * s) and torch.equal(V4, V) and torch.equal(U4, U)
    X.mul_(0.25)
    Un, sn, Vn = svd_device(ops, X, svd_type="randomized", n_components=k, seed=1, precision="native")
    assert float(((s - sn).abs() / sn).max()) < 1e-4
