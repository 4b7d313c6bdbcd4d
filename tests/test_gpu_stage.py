"""GPU: the drop-in `main()` path end to end (slice file -> device build -> SVD -> Dataset -> NetCDF ->
cache hit), against the oracle's restatement of the reference flow (era5_svd.py:384-425), with the
presence rules of tests/test_05_dvc_era5_svd.py:326-373 (X iff save_data_matrix, X_mean iff
mean_center and d > 1, X_std iff scale)."""
import os

import numpy as np
import pytest

from dmd_era5_b200.dataset import DataArray, Dataset, read_netcdf, write_netcdf
from oracle.compare import recon_rel_err, sigma_rel_err, vector_angles
from oracle.slice_tools_np import build_matrix_np, standardize_np
from oracle.svd_ref import randomized_svd_ref, standard_svd_ref
from oracle.synthetic_np import mock_era5_np

pytestmark = pytest.mark.gpu

VARS = ["temperature", "u_component_of_wind"]
LEVELS = [1000, 850, 500]


def make_slice(root, config):
    """What era5_download writes (era5_download.py:36-42, 114): per-variable (time, level, lat, lon)."""
    from dmd_era5_b200.config_parser import config_parser

    parsed = config_parser(config, "era5-svd")
    m = mock_era5_np(25, VARS, LEVELS, seed=4)
    dv = {k: DataArray(v, ("time", "level", "latitude", "longitude")) for k, v in m["vars"].items()}
    ds = Dataset(dv, {"time": m["time"], "level": m["level"], "latitude": m["latitude"], "longitude": m["longitude"]},
                 {"source_path": config["source_path"], "variables": VARS, "levels": LEVELS})
    write_netcdf(ds, parsed["era5_slice_path"])
    return m, parsed


def base_config(**kw):
    cfg = {"source_path": "gs://mock", "variables": "u_component_of_wind,temperature", "levels": "850,1000",
           "svd_type": "randomized", "delay_embedding": 2, "mean_center": True, "scale": False,
           "start_datetime": "2019-01-01T00", "end_datetime": "2019-01-02T00", "delta_time": "1h",
           "n_components": 6, "save_data_matrix": True, "random_seed": 3}
    cfg.update(kw)
    return cfg


def oracle_matrix(m, cfg):
    # variables and levels in CONFIG order (era5_svd.py:384-385)
    vs = [v.strip() for v in cfg["variables"].split(",")]
    lv = [LEVELS.index(int(x)) for x in cfg["levels"].split(",")]
    arrs = [m["vars"][v][:, lv] for v in vs]
    return build_matrix_np(arrs, cfg["mean_center"], cfg["scale"], cfg["delay_embedding"])


@pytest.mark.parametrize("svd_type,d,mc,sc", [("randomized", 2, True, False), ("standard", 1, True, True),
                                               ("randomized", 3, False, False)])
def test_main_end_to_end(tmp_path, monkeypatch, svd_type, d, mc, sc):
    from dmd_era5_b200.era5_svd import main

    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    cfg = base_config(svd_type=svd_type, delay_embedding=d, mean_center=mc, scale=sc)
    m, parsed = make_slice(tmp_path, cfg)
    res, added, retrieved = main(cfg, write_to_netcdf=True)
    assert added is False and retrieved is False
    X, Xm, Xs = oracle_matrix(m, cfg)
    k = 6
    U0, s0, V0 = (randomized_svd_ref(X, k, 3) if svd_type == "randomized" else standard_svd_ref(X, k))
    assert res["U"].dims == ("space", "components") and res["V"].dims == ("components", "time")
    assert res["U"].shape == (X.shape[0], k) and res["V"].shape == (k, X.shape[1]) and res["s"].shape == (k,)
    assert sigma_rel_err(res["s"].values, s0) < 1e-6
    ref = recon_rel_err(X, U0, s0, V0)
    assert abs(recon_rel_err(X, res["U"].values, res["s"].values, res["V"].values) - ref) <= 0.01 * ref
    assert np.max(np.abs(res["X"].values - X)) < 1e-10                      # save_data_matrix
    assert ("X_mean" in res) == (mc and d > 1) and ("X_std" in res) == (sc and mc and d > 1)      # quirk Q3
    if "X_mean" in res:
        assert np.allclose(res["X_mean"].values, Xm)
    S = 2 * 36 * 72
    assert sorted(res.coords) == sorted(["space", "time", "components", "original_variable", "delay", "level",
                                         "latitude", "longitude"])
    assert np.array_equal(res.coord("space"), np.arange(2 * S * d))
    assert res.coord("original_variable")[0] == "u_component_of_wind" and res.coord("original_variable")[S] == "temperature"
    assert res.coord("level")[0] == 850 and res.coord("level")[36 * 72] == 1000
    assert np.array_equal(res.coord("delay"), np.repeat(np.flip(np.arange(d)), 2 * S))
    assert np.array_equal(res.coord("time"), m["time"][d - 1:])
    a = res.attrs
    assert a["variables"] == ["u_component_of_wind", "temperature"] and a["levels"] == [850, 1000]
    assert a["mean_center"] == int(mc) and a["scale"] == int(sc) and a["delay_embedding"] == d
    assert a["svd_type"] == svd_type and a["n_components"] == k and a["save_data_matrix"] == 1
    # written file + result cache (era5_svd.py:157-227): second call returns the stored result
    assert os.path.exists(parsed["save_path"])
    res2, _, _ = main(cfg, write_to_netcdf=False)
    assert np.allclose(res2["s"].values, res["s"].values) and sorted(res2.data_vars) == sorted(res.data_vars)
    back = read_netcdf(parsed["save_path"])
    assert np.array_equal(back["U"].values, res["U"].values)


def test_main_matrix_dtype_cast(tmp_path, monkeypatch):
    """Opt-in float cast (north_star (a); absent in the reference): float64 mock slice, float32 snapshot matrix written
    by the build kernel, tensor-core passes; sigma within the FP32-split tolerance (1e-4) of the float64 oracle."""
    from dmd_era5_b200.era5_svd import main

    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    cfg = base_config(delay_embedding=1, matrix_dtype="float32")
    m, parsed = make_slice(tmp_path, cfg)
    res, _, _ = main(cfg)
    X, _, _ = oracle_matrix(m, cfg)
    U0, s0, V0 = randomized_svd_ref(X, 6, 3)
    assert res["U"].values.dtype == np.float32 and res["X"].values.dtype == np.float32
    # the time mean is held in the matrix dtype (it is what X_mean stores): one float32 rounding of a ~300 K mean
    raw_max = max(float(np.max(np.abs(v))) for v in m["vars"].values())
    assert np.max(np.abs(res["X"].values - X)) <= 1.2e-7 * raw_max
    assert sigma_rel_err(res["s"].values, s0) < 1e-4
    ref = recon_rel_err(X, U0, s0, V0)
    assert abs(recon_rel_err(X, res["U"].values.astype(np.float64), res["s"].values.astype(np.float64),
                             res["V"].values.astype(np.float64)) - ref) <= 0.01 * ref
    with pytest.raises(Exception, match="matrix_dtype"):
        main(base_config(delay_embedding=1, matrix_dtype="float16"))


def test_main_phase_errors(tmp_path, monkeypatch):
    from dmd_era5_b200.era5_svd import main

    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    with pytest.raises(Exception, match="Error retrieving ERA5 slice"):
        main(base_config())
    cfg = base_config(levels="850,300")                  # level missing from the slice file attrs -> no match
    make_slice(tmp_path, base_config())
    with pytest.raises(Exception, match="Error retrieving ERA5 slice"):
        main(cfg)


def test_standardize_data_gpu():
    from dmd_era5_b200 import slice_tools as st

    m = mock_era5_np(25, VARS, [1000, 850], seed=1)
    ds = Dataset({k: DataArray(v, ("time", "level", "latitude", "longitude")) for k, v in m["vars"].items()},
                 {"time": m["time"], "level": m["level"], "latitude": m["latitude"], "longitude": m["longitude"]})
    out, mean, std = st.standardize_data(ds)              # reference tests/test_02_slice_tools.py:108-152
    for v in VARS:
        ref, mu, sd = standardize_np(m["vars"][v], scale=True)
        assert np.allclose(out[v].values.mean(axis=0), 0, atol=1e-6) and np.allclose(out[v].values.std(axis=0), 1, atol=1e-6)
        assert np.allclose(out[v].values, ref, atol=1e-10) and np.allclose(mean[v].values, mu) and np.allclose(std[v].values, sd)
    out2, mean2, std2 = st.standardize_data(ds, scale=False)
    assert std2 is None and np.allclose(out2["temperature"].values.std(axis=0), m["vars"]["temperature"].std(axis=0))
    out3, _, _ = st.standardize_data(ds, dim="level")     # test_02:180-212
    assert np.allclose(out3["u_component_of_wind"].values.mean(axis=1), 0, atol=1e-6)
    assert np.allclose(out3["u_component_of_wind"].values.std(axis=1), 1, atol=1e-6)


def test_stage_stream_matches_one_shot():
    """SvdStageStream (H2D | compute | D2H overlapped, buffers reused) returns, slice by slice, exactly what the
    one-shot path returns for the same slice."""
    import torch

    from dmd_era5_b200.era5_svd import get_ops
    from dmd_era5_b200.pipeline import build_matrix_device, svd_device
    from dmd_era5_b200.stage_stream import SvdStageStream
    from dmd_era5_b200.synthetic import synthetic_field

    ops = get_ops()
    T, S, k = 96, 20000, 12
    hosts = []
    for i in range(5):
        f = synthetic_field(T, S, device="cuda", seed=40 + i, rank=30, rho=0.8)
        h = torch.empty((T, S), dtype=torch.float32, pin_memory=True)
        h.copy_(f)
        hosts.append(h)
    torch.cuda.synchronize()
    got = {}
    st = SvdStageStream(ops, T, S, n_components=k, precision="tf32x3", seed=5)
    n = st.run(hosts, consume=lambda i, U, s, V: got.__setitem__(i, (U.clone(), s.clone(), V.clone())))
    assert n == 5 and sorted(got) == [0, 1, 2, 3, 4]
    for i, h in enumerate(hosts):
        built = build_matrix_device(ops, [h.cuda()], mean_center=True, scale=False)
        U, s, V = svd_device(ops, built.X, svd_type="randomized", n_components=k, seed=5, precision="tf32x3")
        assert torch.equal(got[i][1], s.cpu()) and torch.equal(got[i][2], V.cpu()) and torch.equal(got[i][0], U.cpu())


@pytest.mark.parametrize("dtype", [np.float32, np.float64, ">f4"])
def test_stage_blocks_chunked_upload(dtype):
    """Pinned-ring staging: many small chunks, pending time (resample) and level selections fused into the copy,
    native and big-endian sources, several variables: the device blocks equal the materialised host selection."""
    from datetime import timedelta

    from dmd_era5_b200 import slice_tools as st
    from dmd_era5_b200.era5_svd import get_ops
    from dmd_era5_b200.stage import stage_blocks

    m = mock_era5_np(49, VARS, LEVELS, seed=2)
    dv = {k: DataArray(v.astype(dtype), ("time", "level", "latitude", "longitude")) for k, v in m["vars"].items()}
    ds = Dataset(dv, {"time": m["time"], "level": m["level"], "latitude": m["latitude"], "longitude": m["longitude"]})
    ops = get_ops("cuda:0")
    for sel in (ds, st.slice_era5_dataset(ds, levels=[500, 1000]),
                st.resample_era5_dataset(st.slice_era5_dataset(ds, levels=[850]), timedelta(hours=6))):
        want = {v: np.asarray(sel[v].lazy().materialise()).astype(np.dtype(dtype).newbyteorder("=")) for v in VARS}
        for chunk in (50_000, 1 << 20, None):
            blocks, S = stage_blocks(sel, VARS, ops, chunk_bytes=chunk)
            for v, b in zip(VARS, blocks):
                T = want[v].shape[0]
                assert S == want[v][0].size and tuple(b.shape) == (T, S)
                assert np.array_equal(b.cpu().numpy(), want[v].reshape(T, -1))


def test_main_use_dvc_without_dvc(tmp_path, monkeypatch):
    """use_dvc=True where dvc / GitPython are not installed: retrieval attempts end in the reference's warnings and the
    stage computes (era5_svd.py:116-154, 190-227); adding the written result to DVC fails with the reference's error
    wrapper (:442-451).  With dvc installed this path is the reference's docker-marked test (test_05_dvc_era5_svd.py)."""
    from dmd_era5_b200 import dvc_tools
    from dmd_era5_b200.era5_svd import main

    if dvc_tools.dvc_available():
        pytest.skip("dvc is installed: covered by the reference's own DVC tests")
    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    cfg = base_config(delay_embedding=1)
    make_slice(tmp_path, cfg)
    res, added, retrieved = main(cfg, write_to_netcdf=False, use_dvc=True)        # nothing to add: no error
    assert added is False and retrieved is False and res["s"].shape == (6,)
    with pytest.raises(Exception, match="Error adding SVD results to DVC"):
        main(base_config(delay_embedding=1, n_components=5), write_to_netcdf=True, use_dvc=True)


def test_stage_blocks_pieces_are_the_row_shards(tmp_path, monkeypatch, ops):
    """Row-sharded staging (n_gpus > 1): the pieces a rank stages are exactly the columns [p0, p1) of the full per-variable
    blocks, with level / time selections still pending on the lazy slice."""
    import torch
    from dmd_era5_b200.stage import _prepare, retrieve_era5_slice, shard_pieces, stage_blocks

    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    cfg = base_config(delta_time="6h")
    m, parsed = make_slice(tmp_path, cfg)
    ds, _ = retrieve_era5_slice(parsed)
    ds = _prepare(ds, parsed)
    variables = parsed["variables"]
    full, S = stage_blocks(ds, variables, ops)
    assert S == 2 * 36 * 72 and len(full) == 2
    m0 = 2 * S
    for r0, r1 in ((0, 1000), (1000, S + 77), (S - 5, S + 5), (S + 77, m0), (0, m0), (2592 * 2 - 1, 2592 * 2 + 3)):
        pieces = shard_pieces(S, 2, r0, r1)
        blocks, S2 = stage_blocks(ds, variables, ops, pieces=pieces)
        assert S2 == S and sum(b.shape[1] for b in blocks) == r1 - r0
        for b, (v, p0, p1) in zip(blocks, pieces):
            assert torch.equal(b, full[v][:, p0:p1])
    torch.cuda.synchronize()


def test_main_n_gpus_matches_single_device(tmp_path, monkeypatch):
    """The opt-in ``n_gpus`` key: main() row-shards the stage over 2 worker processes / GPUs (stage_multi.py) and returns
    the same Dataset as the single-device path (sigma, V to float64 round-off; U rows in the reference's order)."""
    import torch
    from dmd_era5_b200.era5_svd import main

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    for svd_type, d in (("randomized", 2), ("standard", 1)):
        cfg = base_config(svd_type=svd_type, delay_embedding=d, mean_center=True, scale=True)
        make_slice(tmp_path, cfg)
        one, _, _ = main(cfg)
        two, _, _ = main(dict(cfg, n_gpus=2))
        assert sorted(one.data_vars) == sorted(two.data_vars)
        assert sigma_rel_err(two["s"].values, one["s"].values) < 1e-10
        assert np.max(np.abs(two["X"].values - one["X"].values)) == 0.0
        assert vector_angles(two["U"].values, one["U"].values).max() < 1e-7
        assert vector_angles(two["V"].values.T, one["V"].values.T).max() < 1e-7
        if d > 1:
            assert np.array_equal(two["X_mean"].values, one["X_mean"].values)
        for c in one.coords:
            assert np.array_equal(one.coord(c), two.coord(c)), c


def _golden_compute():
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "main_flow.json")) as f:
        return json.load(f)["compute"]


@pytest.mark.parametrize("name", list(_golden_compute()))
def test_compute_phase_follows_the_references_own_main(name, tmp_path, monkeypatch):
    """Which pre-processing a configuration switches on and which variables the result carries, as recorded from the
    reference's OWN ``main`` (era5_svd.py:383-425, executed unchanged with recording stand-ins,
    tests/golden/make_golden_main.py): X iff save_data_matrix; X_mean / X_std only when they exist AND delay_embedding > 1
    (quirk Q3); `scale` without `mean_center` does nothing (quirk Q4) - for all 16 combinations."""
    from dmd_era5_b200.era5_svd import main

    rec = _golden_compute()[name]
    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    cfg = base_config(**rec["config"])
    m, parsed = make_slice(tmp_path, cfg)
    res, added, retrieved = main(cfg, write_to_netcdf=False)
    assert ("X" in res) == rec["has_X"]
    assert ("X_mean" in res) == rec["has_X_mean"]
    assert ("X_std" in res) == rec["has_X_std"]
    d = rec["config"]["delay_embedding"]
    if rec["has_X"]:
        X = np.asarray(res["X"].values, dtype=np.float64)
        rows = X[: X.shape[0] // d]                                  # first delay block = the base rows, n columns each
        if rec["standardize"] is None:
            vs = [v.strip() for v in cfg["variables"].split(",")]
            lv = [LEVELS.index(int(x)) for x in cfg["levels"].split(",")]
            raw = np.concatenate([m["vars"][v][:, lv].reshape(25, -1).T for v in vs], axis=0)
            assert np.allclose(rows, raw[:, : rows.shape[1]], rtol=1e-12, atol=0)       # untouched data
    if rec["has_X_mean"]:
        assert res["X_mean"].values.shape[0] == res["U"].values.shape[0]                 # replicated d times along space
