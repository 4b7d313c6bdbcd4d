"""CPU: the oracle against the reference's golden vectors / known-answer tests and against the
reference's own library calls (pins the oracle)."""
import os

import numpy as np
import pytest

from oracle import slice_tools_np as st
from oracle.compare import sigma_rel_err, vector_angles
from oracle.svd_ref import (n_iter_auto, omega_ref, randomized_svd_ref, randomized_svd_restated,
                            standard_svd_ref, svd_flip_u_np)
from oracle.synthetic_np import lowrank_field_np, mock_era5_np


def test_delay_embedding_reference_kats():
    # tests/test_02_slice_tools.py:215-231 of the reference, verbatim cases
    cases = [
        (np.array([[0, 1, 2, 3, 4]]), 1, np.array([[0, 1, 2, 3, 4]])),
        (np.array([[0, 1, 2, 3, 4]]), 2, np.array([[0, 1, 2, 3], [1, 2, 3, 4]])),
        (np.array([[0, 1, 2, 3, 4]]), 3, np.array([[0, 1, 2], [1, 2, 3], [2, 3, 4]])),
        (np.array([[0, 1, 2], [3, 4, 5]]), 2, np.array([[0, 1], [3, 4], [1, 2], [4, 5]])),
    ]
    for X, d, expected in cases:
        assert np.array_equal(st.delay_embed_np(X, d), expected)


def test_delay_embedding_golden_from_reference_source(golden_dir):
    g = np.load(os.path.join(golden_dir, "delay_embedding.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in g.files})
    assert len(names) == 8
    for nm in names:
        out = st.delay_embed_np(g[nm + "_X"], int(g[nm + "_d"]))
        assert np.array_equal(out, g[nm + "_out"]), nm


@pytest.mark.parametrize("X", [np.zeros(3), np.zeros((3, 3, 3))])
def test_delay_embedding_invalid_matrix(X):
    with pytest.raises(ValueError, match="Input array must be 2D."):
        st.delay_embed_np(X, 1)


@pytest.mark.parametrize("d", [0, 0.5, -1])
def test_delay_embedding_invalid_delay(d):
    with pytest.raises(ValueError, match="Delay must be an integer greater than 0."):
        st.delay_embed_np(np.zeros((3, 3)), d)


def test_standardize_properties():
    # reference tests/test_02_slice_tools.py:108-174: mean 0 / std 1 (ddof=0) to 1e-6
    ds = mock_era5_np(25, ["temperature", "u_component_of_wind"], [1000, 850], seed=1)
    for a in ds["vars"].values():
        out, mu, sd = st.standardize_np(a, scale=True)
        assert np.allclose(out.mean(axis=0), 0, atol=1e-6)
        assert np.allclose(out.std(axis=0), 1, atol=1e-6)
        assert np.allclose(mu, a.mean(axis=0)) and np.allclose(sd, a.std(axis=0))
        out2, mu2, sd2 = st.standardize_np(a, scale=False)
        assert sd2 is None and np.allclose(out2.std(axis=0), a.std(axis=0))


def test_flatten_row_order():
    # row r = v*S + (l*A + a)*O + o, X[r, t] = var_v[t, l, a, o]  (slice_tools.py:323-336)
    rng = np.random.RandomState(0)
    a0, a1 = rng.rand(5, 2, 3, 4), rng.rand(5, 2, 3, 4)
    X = st.flatten_np([a0, a1])
    assert X.shape == (2 * 2 * 3 * 4, 5)
    S = 24
    for (v, arr) in enumerate((a0, a1)):
        for (l, a, o) in [(0, 0, 0), (1, 2, 3), (0, 1, 2)]:
            assert np.array_equal(X[v * S + (l * 3 + a) * 4 + o], arr[:, l, a, o])
    lev, lat, lon = st.space_coords_np([1000, 850], [10., 5., 0.], [0., 1., 2., 3.], 2, d=2)
    assert lev.shape == (96,) and lev[13] == 850 and lat[13] == 10. and lon[13] == 1.
    assert np.array_equal(st.delay_coord_np(3, 2), [1, 1, 1, 0, 0, 0])


def test_build_matrix_quirks():
    ds = mock_era5_np(9, ["temperature"], [1000], seed=2)
    arrs = list(ds["vars"].values())
    X, mu, sd = st.build_matrix_np(arrs, True, True, 1)
    assert mu is None and sd is None                      # Q3: dropped when d == 1
    X2, mu2, sd2 = st.build_matrix_np(arrs, False, True, 2)
    assert mu2 is None and np.array_equal(X2, st.delay_embed_np(st.flatten_np(arrs), 2))   # Q4
    X3, mu3, sd3 = st.build_matrix_np(arrs, True, False, 2)
    assert mu3.shape == (X3.shape[0],) and sd3 is None


def test_resample_25_hourly_to_6h():
    # reference tests/test_02_slice_tools.py:85-101: 25 hourly samples -> 5 six-hourly
    t = (np.datetime64("2019-01-01T00", "ns") + np.arange(25) * np.timedelta64(1, "h")).astype(np.int64)
    labels, idx = st.resample_nearest_index(t, 6 * 3600 * 10**9)
    assert len(labels) == 5 and list(idx) == [0, 6, 12, 18, 24]


def test_restatement_matches_sklearn_bitwise():
    X = lowrank_field_np(1500, 120, r=50, rho=0.85, seed=4)
    for Xc in (X, X.astype(np.float32)):
        a = randomized_svd_ref(Xc, 11, 3)
        b = randomized_svd_restated(Xc, 11, 3)
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    assert n_iter_auto(1038240, 744, 100) == 4 and n_iter_auto(40491360, 1460, 100) == 7
    assert omega_ref(10, 3, 1, np.float32).dtype == np.float32


def test_svd_flip_first_max():
    u = np.array([[1.0, -2.0], [-1.0, 2.0], [0.5, 0.0]])
    v = np.eye(2)
    uf, vf = svd_flip_u_np(u.copy(), v.copy())
    assert uf[0, 0] > 0 and uf[0, 1] > 0 and vf[1, 1] == -1.0   # ties -> first row decides


def test_golden_svd_vectors(golden_dir):
    from oracle.slice_tools_np import build_matrix_np

    g = np.load(os.path.join(golden_dir, "svd_standard_c1.npz"))
    T, nv, nl, d, k, seed = [int(x) for x in g["meta"]]
    ds = mock_era5_np(T, ["temperature", "u_component_of_wind"][:nv], [1000][:nl], seed=seed)
    X, _, _ = build_matrix_np(list(ds["vars"].values()), True, False, d)
    U, s, V = standard_svd_ref(X, k)
    assert sigma_rel_err(s, g["s"]) < 1e-12 and vector_angles(U, g["U"]).max() < 1e-9
    g = np.load(os.path.join(golden_dir, "svd_randomized_c1.npz"))
    U, s, V = randomized_svd_ref(X, k, int(g["meta"][6]))
    assert sigma_rel_err(s, g["s"]) < 1e-12 and vector_angles(U, g["U"]).max() < 1e-9
    g = np.load(os.path.join(golden_dir, "svd_randomized_lowrank_f64.npz"))
    m, n, r, k, seed, rs = [int(x) for x in g["meta"]]
    X = lowrank_field_np(m, n, r=r, rho=0.8, seed=seed)
    U, s, V = randomized_svd_ref(X, k, rs)
    assert sigma_rel_err(s, g["s"]) < 1e-12 and vector_angles(V.T, g["V"].T).max() < 1e-9


H_NS = 3600 * 10**9
RESAMPLE_CASES = {
    "25 hourly -> 6h": (np.datetime64("2019-01-01T00", "ns").astype(np.int64) + H_NS * np.arange(25), 6 * H_NS),
    "hourly -> 90 min (ties)": (np.datetime64("2019-01-01T00", "ns").astype(np.int64) + H_NS * np.arange(13), 90 * 60 * 10**9),
    "start 05:00, 6h": (np.datetime64("2019-01-01T05", "ns").astype(np.int64) + H_NS * np.arange(30), 6 * H_NS),
    "3-hourly -> 2h (ties)": (np.datetime64("2019-03-01T00", "ns").astype(np.int64) + 3 * H_NS * np.arange(17), 2 * H_NS),
    "irregular": (np.datetime64("2019-01-01T00", "ns").astype(np.int64) + np.array([0, 1, 2, 5, 6, 7, 11, 12, 20, 21, 30]) * H_NS, 4 * H_NS),
    "identity": (np.datetime64("2019-01-02T03", "ns").astype(np.int64) + H_NS * np.arange(10), H_NS),
    "6-hourly -> 1h (upsampling, ties)": (np.datetime64("2019-01-01T00", "ns").astype(np.int64) + 6 * H_NS * np.arange(5), H_NS),
}


@pytest.mark.parametrize("name", list(RESAMPLE_CASES))
def test_resample_nearest_pinned_against_pandas(name):
    """``ds.resample(time=delta).nearest()`` (slice_tools.py:139) is xarray's reindex onto the resample bins' labels with
    pandas' method='nearest' - both the oracle restatement and the product's index form must reproduce pandas itself,
    including its tie rule (a label half way between two samples takes the later one)."""
    pd = pytest.importorskip("pandas")
    from dmd_era5_b200.slice_tools import resample_nearest_index as product_index
    from oracle.slice_tools_np import resample_nearest_index as oracle_index

    t, delta = RESAMPLE_CASES[name]
    idx = pd.DatetimeIndex(t)
    labels = pd.Series(np.arange(len(idx)), index=idx).resample(pd.Timedelta(delta, "ns")).first().index
    src = idx.get_indexer(labels, method="nearest")
    want_labels = labels.values.astype("datetime64[ns]").astype(np.int64)
    for fn in (oracle_index, product_index):
        got_labels, got_src = fn(t, delta)
        assert np.array_equal(got_labels, want_labels), fn.__module__
        assert np.array_equal(got_src, src), fn.__module__


def test_resample_nearest_against_pandas_on_random_time_axes():
    """300 seeded random cases - irregular, sorted, minute-resolution time axes (gaps, exact half-way ties, starts off
    midnight) and bin widths from 30 minutes to 2 days: the oracle and the product index form give pandas' labels and
    pandas' nearest sample for every label."""
    pd = pytest.importorskip("pandas")
    from dmd_era5_b200.slice_tools import resample_nearest_index as product_index
    from oracle.slice_tools_np import resample_nearest_index as oracle_index

    rng = np.random.RandomState(12)
    minute = 60 * 10**9
    t0 = np.datetime64("2019-01-01T00", "ns").astype(np.int64)
    for case in range(300):
        n = int(rng.randint(2, 60))
        steps = rng.choice([15, 30, 60, 60, 60, 120, 180, 360], size=n)          # minutes between samples
        t = t0 + int(rng.randint(0, 3 * 24 * 60)) * minute + np.concatenate([[0], np.cumsum(steps[:-1])]) * minute
        delta = int(rng.choice([30, 60, 90, 120, 180, 240, 360, 720, 1440, 2880])) * minute
        idx = pd.DatetimeIndex(t)
        labels = pd.Series(np.arange(n), index=idx).resample(pd.Timedelta(delta, "ns")).first().index
        src = idx.get_indexer(labels, method="nearest")
        want_labels = labels.values.astype("datetime64[ns]").astype(np.int64)
        for fn in (oracle_index, product_index):
            got_labels, got_src = fn(t, delta)
            assert np.array_equal(got_labels, want_labels), (case, fn.__module__)
            assert np.array_equal(got_src, src), (case, fn.__module__, t.tolist(), delta)


def _svd_on_era5_cases():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "svd_on_era5_reference.npz"), allow_pickle=True)
    return g, [tuple(row) for row in g["meta"]]


@pytest.mark.parametrize("row", _svd_on_era5_cases()[1], ids=[r[0] for r in _svd_on_era5_cases()[1]])
def test_oracle_reproduces_the_references_own_svd_on_era5_bit_for_bit(row):
    """tests/golden/svd_on_era5_reference.npz holds what the reference's OWN svd_on_era5 (extracted from
    src/dmd_era5/era5_svd/era5_svd.py:230-263 with ast and executed unchanged, tests/golden/make_golden_svd_on_era5.py)
    returned for seeded inputs.  The oracle's calls must give exactly those arrays (same libraries in this image: NumPy /
    OpenBLAS gesdd, scikit-learn 1.9 randomized_svd), for float64 and float32 input, tall and wide, k > n - bit for bit on
    the CPU type that generated them, to rounding on any other (OpenBLAS dispatches per CPU)."""
    g, _ = _svd_on_era5_cases()
    name, m, n, k, kind, dt, seed, data_seed = row
    X = lowrank_field_np(int(m), int(n), r=min(30, int(m), int(n)), rho=0.85, seed=int(data_seed), dtype=np.dtype(dt))
    if kind == "standard":
        U, s, V = standard_svd_ref(X, int(k))
    else:
        U, s, V = randomized_svd_ref(X, int(k), int(seed))
    f64 = np.dtype(dt) == np.float64
    kk = min(int(k), int(m), int(n))
    lead = min(kk, 8)                              # well-separated leading part (the generator's spectrum is 0.85^i)
    for got, key in ((U, "U"), (s, "s"), (V, "V")):
        want = g[f"{name}_{key}"]
        assert got.shape == want.shape and got.dtype == want.dtype
        if np.array_equal(got, want):
            continue
        # another CPU type dispatches other BLAS kernels: then the last bits may differ - but nothing more than that
        a, b = (got[:, :lead], want[:, :lead]) if key == "U" else ((got[:lead], want[:lead]) if key == "V" else (got[:lead], want[:lead]))
        assert np.allclose(a, b, rtol=1e-9 if f64 else 2e-4, atol=(1e-11 if f64 else 2e-5) * float(np.max(np.abs(b)))), (name, key)


def test_mock_data_generator_reproduces_the_references_own_bit_for_bit():
    """BASELINE config 1's input: the reference's OWN _generate_variable_data (create_mock_data.py:112-155, executed from
    source by tests/golden/make_golden_mock.py with NumPy's global RNG seeded) against oracle.synthetic_np.mock_era5_np -
    same draws in the same order, so the arrays agree bit for bit (sha256 of the bytes)."""
    import hashlib
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "mock_era5.json")) as f:
        g = json.load(f)
    for rec in g["cases"]:
        ds = mock_era5_np(rec["n_times"], rec["variables"], rec["levels"], seed=rec["seed"])
        for var, want in rec["arrays"].items():
            a = ds["vars"][var]
            assert list(a.shape) == want["shape"] and str(a.dtype) == want["dtype"]
            assert np.allclose(a.ravel()[:5], want["first"], rtol=0, atol=0)
            assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == want["sha256"], var
