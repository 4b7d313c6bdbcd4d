"""CPU: host-side drivers (randomized / standard schedule, delay blocks, row sharding, MAXLOC flip)
exercised with the test-only FakeOps stand-in, single process and 2-rank gloo."""
import os
import sys

import numpy as np
import pytest
import torch

from fake_ops import FakeOps
from dmd_era5_b200.dist import LocalComm, shard_rows
from dmd_era5_b200.rsvd import draw_omega, n_iter_auto, randomized_svd_device
from dmd_era5_b200.standard import standard_svd_device
from oracle.compare import sigma_rel_err, signs_agree, vector_angles
from oracle.slice_tools_np import delay_embed_np
from oracle.svd_ref import omega_ref, randomized_svd_ref, standard_svd_ref
from oracle.synthetic_np import lowrank_field_np


def test_draw_omega_matches_reference_rng():
    for dt, nd in ((torch.float64, np.float64), (torch.float32, np.float32)):
        assert np.array_equal(draw_omega(50, 7, 3, dt), omega_ref(50, 7, 3, nd).astype(np.float64))
    np.random.seed(5)
    a = draw_omega(9, 2, None, torch.float64)          # unseeded: global RandomState, like the reference
    np.random.seed(5)
    assert np.array_equal(a, np.random.normal(size=(9, 12)))
    assert n_iter_auto(1038240, 744, 100) == 4 and n_iter_auto(40491360, 1460, 100) == 7


@pytest.mark.parametrize("d", [1, 2])
def test_randomized_driver_matches_oracle(d):
    X = lowrank_field_np(3000, 160, r=60, rho=0.88, seed=3)
    k = 12
    Xd = delay_embed_np(X, d)
    U0, s0, V0 = randomized_svd_ref(Xd, k, 7)
    U, s, Vt = randomized_svd_device(FakeOps(), torch.from_numpy(X), k, draw_omega(Xd.shape[1], k, 7, torch.float64), delay=d)
    assert sigma_rel_err(s, s0) < 1e-9
    assert vector_angles(U.numpy(), U0).max() < 1e-6 and vector_angles(Vt.numpy().T, V0.T).max() < 1e-6
    assert signs_agree(U.numpy(), U0)


@pytest.mark.parametrize("d", [1, 2])
def test_randomized_driver_tensor_core_schedule(d):
    """The tensor-core schedule of the driver on CPU stand-ins (exact products of the float32 operands): Omega kept in
    tf32-representable values after every orthonormalisation, ONE plain Y image through the power iterations, the
    hi / lo pair only in the final pass.  Rounding Omega is a choice of basis, not an approximation: sigma agrees with
    the float64 oracle on the same float32 data to 1e-6 (the float32 storage of Y), vectors to 1e-4 rad."""
    from dmd_era5_b200._cabi import PREC_TF32X3

    X = lowrank_field_np(3000, 160, r=60, rho=0.88, seed=3).astype(np.float32)
    k = 12
    Xd = delay_embed_np(X.astype(np.float64), d)
    U0, s0, V0 = randomized_svd_ref(Xd, k, 7)
    ops = FakeOps()
    U, s, Vt = randomized_svd_device(ops, torch.from_numpy(X), k, draw_omega(Xd.shape[1], k, 7, torch.float32), delay=d,
                                     precision=PREC_TF32X3)
    assert U.dtype == torch.float32
    assert sigma_rel_err(s, s0) < 1e-6
    assert vector_angles(U.double().numpy(), U0).max() < 1e-4 and vector_angles(Vt.numpy().T, V0.T).max() < 1e-4
    assert signs_agree(U.double().numpy(), U0)
    q = n_iter_auto(Xd.shape[0], Xd.shape[1], k)
    assert ops.calls["round_tf32"] == q + 1                       # the start matrix and every orthonormalised basis
    assert ops.calls["sketch_tc"] == d * (q + 1) + 1              # q + 1 tall sketches per delay block + U = Y M
    assert ops.calls["project_tc"] == d * (q + 1) + 1             # q + 1 projections per block + the Gram of Y


def test_standard_driver_matches_oracle():
    X = lowrank_field_np(2000, 64, r=64, rho=0.9, seed=5)
    U0, s0, V0 = standard_svd_ref(delay_embed_np(X, 2), 10)
    U, s, Vt = standard_svd_device(FakeOps(), torch.from_numpy(X), 10, delay=2)
    assert sigma_rel_err(s, s0) < 1e-9 and vector_angles(U.numpy(), U0).max() < 1e-6


@pytest.mark.parametrize("precision", ["native", "tf32x3", "tf32mix"])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_gram_of_the_delay_embedded_matrix_from_the_base_gram(precision, d):
    """standard.gram_device: G = sum_j X_j^T X_j over the overlapping delay windows is the sum of d shifted diagonal
    blocks of ONE Gram matrix of the base matrix - same values, one pass set over X instead of d, and no tensor-core
    operand starts at an unaligned column (float32 + standard + delay > 1 failed in the kernel's alignment check on the
    device before; the stand-in enforces the same rule)."""
    from dmd_era5_b200.rsvd import PRECISIONS
    from dmd_era5_b200.standard import gram_device

    dtype = np.float64 if precision == "native" else np.float32
    X = lowrank_field_np(700, 301, r=40, rho=0.9, seed=8).astype(dtype)        # 301 columns: two 256-column blocks
    n = X.shape[1] - d + 1
    ops = FakeOps()
    G = gram_device(ops, torch.from_numpy(X), n, d, PRECISIONS[precision]).numpy()
    Xe = X.astype(np.float64)
    if precision == "tf32mix":                                                  # what the single-product kernel computes
        Xe = FakeOps._tf32_trunc(torch.from_numpy(X)).double().numpy()
    want = sum(Xe[:, j : j + n].T @ Xe[:, j : j + n] for j in range(d))
    assert G.shape == (n, n) and np.allclose(G, G.T, rtol=0, atol=0)
    assert np.max(np.abs(G - want)) <= 1e-12 * np.abs(want).max()
    if precision != "native":
        # column blocks of the BASE matrix (256 / 112 columns), whatever d is
        assert (ops.calls["project_x1"] == 2) if precision == "tf32mix" else (ops.calls["project_tc"] == 3)


def test_standard_driver_float32_with_delay():
    """float32 data, standard SVD, delay 2 (the configuration that failed on the device): Gram route + refinement."""
    from dmd_era5_b200.rsvd import PREC_TF32MIX

    X = lowrank_field_np(1500, 40, r=40, rho=0.8, seed=6).astype(np.float32)
    U0, s0, V0 = standard_svd_ref(delay_embed_np(X.astype(np.float64), 2), 6)
    U, s, Vt = standard_svd_device(FakeOps(), torch.from_numpy(X), 6, delay=2, precision=PREC_TF32MIX)
    assert sigma_rel_err(s, s0) < 1e-6 and vector_angles(U.double().numpy(), U0).max() < 1e-4


def test_sketch_wider_than_time_axis():
    X = lowrank_field_np(500, 12, r=12, rho=0.7, seed=1)
    U0, s0, V0 = randomized_svd_ref(X, 8, 2)           # l = 18 > n = 12
    U, s, Vt = randomized_svd_device(FakeOps(), torch.from_numpy(X), 8, draw_omega(12, 8, 2, torch.float64))
    assert sigma_rel_err(s, s0) < 1e-9 and vector_angles(U.numpy(), U0).max() < 1e-6


def test_shard_rows():
    assert shard_rows(1000, 1, 0) == (0, 1000)
    parts = [shard_rows(1038240, 8, r) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 1038240
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert all(p[0] % 128 == 0 for p in parts)
    # balanced: no rank is left without rows while there are at least as many 128-row tiles as ranks (2592 rows = the
    # reference's mock grid with one level: 21 tiles on 8 ranks), and shard sizes differ by at most one tile
    for m0, world in ((2592, 8), (1024, 8), (129, 2), (40491360, 8), (7232760, 8), (1000, 3)):
        parts = [shard_rows(m0, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == m0
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(a % 128 == 0 for a, _ in parts)
        sizes = [b - a for a, b in parts]
        assert min(sizes) > 0 and max(sizes) - min(sizes) <= 128 + 127
    # fewer tiles than ranks: the surplus ranks are empty, at the END (callers cap the rank count: stage._compute)
    assert [shard_rows(200, 4, r) for r in range(4)] == [(0, 128), (128, 200), (200, 200), (200, 200)]


def _gloo_worker(rank, world, port, q, tensor_core=False, route="randomized"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dmd_era5_b200._cabi import PREC_NATIVE, PREC_TF32X3
        from dmd_era5_b200.dist import TorchDistComm

        comm = TorchDistComm()
        X = lowrank_field_np(1030, 90, r=40, rho=0.85, seed=9)
        if tensor_core:
            X = X.astype(np.float32)
        k, d = 9, 2
        r0, r1 = shard_rows(X.shape[0], world, rank, align=128)
        n = X.shape[1] - d + 1
        if route == "standard":
            from dmd_era5_b200.rsvd import PREC_TF32MIX

            U, s, Vt = standard_svd_device(FakeOps(), torch.from_numpy(X[r0:r1].copy()), k, delay=d, comm=comm,
                                           precision=PREC_TF32MIX if tensor_core else PREC_NATIVE)
            q.put((rank, r0, r1, U.double().numpy(), s.numpy(), Vt.numpy()))
            return
        U, s, Vt = randomized_svd_device(FakeOps(), torch.from_numpy(X[r0:r1].copy()), k,
                                         draw_omega(n, k, 4, torch.float32 if tensor_core else torch.float64), delay=d,
                                         comm=comm, row_offset=r0, m0_global=X.shape[0],
                                         precision=PREC_TF32X3 if tensor_core else PREC_NATIVE)
        q.put((rank, r0, r1, U.double().numpy(), s.numpy(), Vt.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tensor_core", [False, True])
def test_two_rank_gloo_row_sharding(tensor_core):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if tensor_core else 0)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q, tensor_core)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = lowrank_field_np(1030, 90, r=40, rho=0.85, seed=9)
    if tensor_core:                                      # float32 data, float64 oracle on the same values
        X = X.astype(np.float32).astype(np.float64)
    k, d, m0 = 9, 2, 1030
    U0, s0, V0 = randomized_svd_ref(delay_embed_np(X, d), k, 4)
    s_tol, a_tol = (1e-6, 1e-4) if tensor_core else (1e-9, 1e-6)
    U = np.zeros((m0 * d, k))
    for rank, r0, r1, Ul, s, Vt in res:
        ml = r1 - r0
        for j in range(d):                               # local block-major -> global rows
            U[j * m0 + r0 : j * m0 + r1] = Ul[j * ml : (j + 1) * ml]
        assert sigma_rel_err(s, s0) < s_tol and vector_angles(Vt.T, V0.T).max() < a_tol
    assert np.array_equal(res[0][4], res[1][4])          # replicated small factors agree bitwise
    assert vector_angles(U, U0).max() < a_tol and signs_agree(U, U0)


@pytest.mark.parametrize("tensor_core", [False, True])
def test_two_rank_gloo_standard_route_with_delay(tensor_core):
    """The standard route row-sharded over 2 ranks (gloo), delay 2: each rank forms the Gram matrix of ITS rows of the base
    matrix once, the shifted diagonal blocks are summed, the n x n result is all-reduced; float32 data additionally runs
    the sharded refinement passes.  Against np.linalg.svd of the embedded matrix (signs are not part of the contract)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (1 if tensor_core else 0)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q, tensor_core, "standard")) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = lowrank_field_np(1030, 90, r=40, rho=0.85, seed=9)
    if tensor_core:
        X = X.astype(np.float32).astype(np.float64)
    k, d, m0 = 9, 2, 1030
    U0, s0, V0 = standard_svd_ref(delay_embed_np(X, d), k)
    s_tol, a_tol = (1e-6, 1e-4) if tensor_core else (1e-9, 1e-6)
    U = np.zeros((m0 * d, k))
    for rank, r0, r1, Ul, s, Vt in res:
        ml = r1 - r0
        for j in range(d):
            U[j * m0 + r0 : j * m0 + r1] = Ul[j * ml : (j + 1) * ml]
        assert sigma_rel_err(s, s0) < s_tol and vector_angles(Vt.T, V0.T).max() < a_tol
    assert np.array_equal(res[0][4], res[1][4]) and np.array_equal(res[0][5], res[1][5])      # replicas agree bitwise
    assert vector_angles(U, U0).max() < a_tol


@pytest.mark.parametrize("d", [1, 2])
def test_randomized_driver_mixed_precision_schedule(d):
    """precision "tf32mix": the early power iterations run single-product TF32 passes on truncated operands (emulated on
    the bit patterns here), only the last MIX_FULL_ITERS iteration(s) and the final range / projection passes keep
    fp32-level products.  Subspace iteration contracts the early error, so sigma stays at the float32-storage level."""
    from dmd_era5_b200.rsvd import MIX_FULL_ITERS, PREC_TF32MIX

    X = lowrank_field_np(3000, 160, r=60, rho=0.88, seed=3).astype(np.float32)
    k = 12
    Xd = delay_embed_np(X.astype(np.float64), d)
    U0, s0, V0 = randomized_svd_ref(Xd, k, 7)
    ops, stats = FakeOps(), {}
    U, s, Vt = randomized_svd_device(ops, torch.from_numpy(X), k, draw_omega(Xd.shape[1], k, 7, torch.float32), delay=d,
                                     precision=PREC_TF32MIX, stats=stats)
    q = n_iter_auto(Xd.shape[0], Xd.shape[1], k)
    low = q - MIX_FULL_ITERS
    assert stats["low_precision_iters"] == low and stats["tall_passes"] == 2 * q + 2
    assert ops.calls["sketch_x1"] == d * low and ops.calls["project_x1"] == d * low
    # full-precision iteration(s): 2-product sketch + projection with Y truncated (project_x2); final passes: 3xTF32
    assert ops.calls["sketch_tc"] == d * (q + 1 - low) + 1 and ops.calls["project_x2"] == d * (q - low)
    assert ops.calls["project_tc"] == d + 1
    assert sigma_rel_err(s, s0) < 1e-6
    assert vector_angles(U.double().numpy(), U0).max() < 2e-4 and vector_angles(Vt.numpy().T, V0.T).max() < 2e-4
    assert signs_agree(U.double().numpy(), U0)
    # full_iters = q switches the low-precision passes off entirely
    ops2 = FakeOps()
    randomized_svd_device(ops2, torch.from_numpy(X), k, draw_omega(Xd.shape[1], k, 7, torch.float32), delay=d,
                          precision=PREC_TF32MIX, full_iters=q)
    assert "sketch_x1" not in ops2.calls


def test_sketch_wider_than_the_rank_drops_noise_directions():
    """rank(X) = 20 < l = 40, float32 mixed precision: the truncated X of the early iterations has full rank, so the
    surplus directions survive until the Rayleigh-Ritz rotation of the last iteration, where the 1e-8 pivot threshold
    removes them (rsvd.py): trailing singular values come out exactly zero, the true ones at float32 accuracy."""
    from dmd_era5_b200.rsvd import PREC_TF32MIX

    rng = np.random.RandomState(0)
    rank, m, n, k = 20, 1500, 90, 30
    A = np.linalg.qr(rng.standard_normal((m, rank)))[0]
    B = np.linalg.qr(rng.standard_normal((n, rank)))[0]
    X = ((A * (10.0 * 0.8 ** np.arange(rank))) @ B.T).astype(np.float32)
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), k, 2)
    U, s, Vt = randomized_svd_device(FakeOps(), torch.from_numpy(X), k, draw_omega(n, k, 2, torch.float32),
                                     precision=PREC_TF32MIX)
    assert torch.isfinite(U).all() and torch.isfinite(s).all()
    assert sigma_rel_err(s[:rank], s0[:rank]) < 1e-5
    assert float(s[rank:].abs().max()) < 1e-6 * s0[0]


def test_subspace_eigensolver_topk_and_fallback(monkeypatch):
    """standard.sym_eig_topk_subspace (the top-k route of the Gram-route standard SVD for time-sized matrices): block
    subspace iteration with Rayleigh-Ritz on the CPU stand-ins.  A decaying spectrum converges to eigh's leading pairs to
    the requested residual; a flat spectrum does not converge and the function says so (the caller then takes the
    tridiagonal route)."""
    from dmd_era5_b200 import standard

    rng = np.random.RandomState(0)
    n, k = 1100, 20
    Qm = np.linalg.qr(rng.standard_normal((n, n)))[0]
    lam = 50.0 * 0.85 ** np.arange(n)
    G = (Qm * lam) @ Qm.T
    stats = {}
    out = standard.sym_eig_topk_subspace(FakeOps(), torch.from_numpy(G), k, 1e-13, stats)
    assert out is not None and "subspace iteration" in stats["eig_route"]
    w, V = out
    assert np.max(np.abs(w.numpy() - lam[:k]) / lam[:k]) < 1e-12
    assert vector_angles(V.numpy(), Qm[:, :k]).max() < 1e-9
    assert np.max(np.abs(V.numpy().T @ V.numpy() - np.eye(k))) < 1e-12
    resid = np.linalg.norm(G @ V.numpy() - V.numpy() * w.numpy(), axis=0).max() / lam[0]
    assert resid < 2e-13
    # dispatch: sym_eig_topk takes this route for n >= SUBSPACE_MIN_N, 8 k <= n
    w2, V2 = standard.sym_eig_topk(FakeOps(), torch.from_numpy(G), k)
    assert np.array_equal(w2.numpy(), w.numpy())
    # flat spectrum: no convergence within the iteration budget
    monkeypatch.setattr(standard, "SUBSPACE_MAX_ITERS", 8)
    Gf = (Qm * (1.0 + 1e-3 * rng.rand(n))) @ Qm.T
    assert standard.sym_eig_topk_subspace(FakeOps(), torch.from_numpy(Gf), k, 1e-13) is None


def test_shard_rows_weighted_partitions_all_rows():
    """Row shards proportional to per-rank speed (bench.py north-star leg): contiguous, aligned, complete; equal weights
    reproduce the equal split up to one tile."""
    from dmd_era5_b200.dist import shard_rows_weighted

    m0 = 40491360
    for w in ([1.0] * 8, [1.05, 1, 0.95, 1, 1.1, 0.9, 1, 1], [1, 2], [3.0]):
        sh = [shard_rows_weighted(m0, w, r) for r in range(len(w))]
        assert sh[0][0] == 0 and sh[-1][1] == m0
        assert all(sh[i][1] == sh[i + 1][0] for i in range(len(w) - 1))
        assert all(a % 128 == 0 for a, _ in sh)
        sizes = [b - a for a, b in sh]
        tot = sum(w)
        assert all(abs(sz - m0 * wi / tot) <= 256 for sz, wi in zip(sizes, w))
    eq = [shard_rows_weighted(m0, [1.0] * 8, r) for r in range(8)]
    ref = [shard_rows(m0, 8, r) for r in range(8)]
    assert all(abs((b - a) - (d - c)) <= 1024 for (a, b), (c, d) in zip(eq, ref))


def test_auto_precision_keeps_three_products_for_uncentred_data():
    """pipeline.svd_device, precision "auto" on float32 data: the mixed schedule (single-product early power iterations)
    for centred data, 3xTF32 in every iteration when the stage says the time mean was NOT removed; float64 data stays on
    the native path either way.  The emulation shows why: with a 250 K mean the truncation of X in the single-product
    passes costs a factor ~10 in sigma."""
    from dmd_era5_b200.pipeline import svd_device

    rng = np.random.RandomState(3)
    X = (5.0 * rng.standard_normal((4000, 120)) + 250.0).astype(np.float32)
    k = 12
    U0, s0, V0 = randomized_svd_ref(X.astype(np.float64), k, 1)
    err = {}
    for centred in (True, False):
        stats = {}
        U, s, Vt = svd_device(FakeOps(), torch.from_numpy(X), svd_type="randomized", n_components=k, seed=1, stats=stats,
                              centred=centred)
        q = n_iter_auto(X.shape[0], X.shape[1], k)
        assert stats["low_precision_iters"] == (q - 1 if centred else 0)
        err[centred] = sigma_rel_err(s.numpy(), s0)
    print(err)
    assert err[False] < 3e-5 and err[False] < 0.5 * err[True]      # a flat (noise) spectrum amplifies every rounding difference
    stats = {}
    svd_device(FakeOps(), torch.from_numpy(X.astype(np.float64)), svd_type="randomized", n_components=k, seed=1, stats=stats,
               centred=False)
    assert stats["low_precision_iters"] == 0 and "sketch_tc" not in stats


def test_looks_centred_separates_anomaly_fields_from_fields_with_their_mean():
    from dmd_era5_b200.era5_svd import looks_centred

    rng = np.random.RandomState(1)
    assert looks_centred(lowrank_field_np(5000, 200, r=40, rho=0.9, seed=2).astype(np.float32))
    assert looks_centred(rng.standard_normal((3000, 50)).astype(np.float32))
    assert not looks_centred((5.0 * rng.standard_normal((3000, 50)) + 250.0).astype(np.float32))
    raw = (30.0 * rng.rand(25, 2000) + 250.0).astype(np.float32).T                  # the reference's mock temperature
    assert not looks_centred(raw) and looks_centred(raw - raw.mean(axis=1, keepdims=True))
    assert looks_centred(np.ones((10, 1), np.float32)) and looks_centred(np.full((10, 5), 3.0, np.float32))   # degenerate
    bad = rng.standard_normal((100, 8)).astype(np.float32)
    bad[3, 2] = np.nan
    assert looks_centred(bad)                                                        # NaN rows are skipped, not decisive
