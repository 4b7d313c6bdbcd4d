"""CPU: the compute phase of the reference's OWN ``main`` with all of its own collaborators (slice_era5_dataset,
resample_era5_dataset, standardize_data, flatten_era5_variables, apply_delay_embedding, svd_on_era5, combine_svd_results,
add_config_attributes, space_coord_to_level_lat_lon - executed unchanged on the xarray stand-in of
tests/golden/xr_contract.py, see tests/golden/make_golden_compute_phase.py) against

* the ORACLE's build (oracle/slice_tools_np.build_matrix_np - what every GPU parity test compares the kernels with):
  X, X_mean, X_std bit for bit, so reference -> oracle -> device is a closed chain;
* the product's host functions flatten_era5_variables / apply_delay_embedding on the stage's own containers;
* the product's ``stage._compute`` packaging - complete result Dataset: variables, dims, dtypes, values, coordinates and
  their values, the attributes of X, the global attributes - with the device arrays supplied by a NumPy stand-in (no GPU
  here; the device arrays themselves are compared with the oracle in tests/test_gpu_stage.py).
"""
import json
import os
import time as _time

import numpy as np
import pytest

from dmd_era5_b200.config_parser import config_parser
from dmd_era5_b200.dataset import DataArray, Dataset
from oracle.slice_tools_np import build_matrix_np, delay_embed_np, flatten_np, resample_nearest_index, standardize_np

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "compute_phase.json")) as f:
    GOLD = json.load(f)
CASES = list(GOLD["cases"])


def golden_array(rec):
    if rec["dtype"] == "datetime64[ns]":
        return np.asarray(rec["values"], dtype=np.int64).astype("datetime64[ns]")
    if rec["dtype"] == "str":
        return np.asarray(rec["values"])
    return np.asarray(rec["values"], dtype=rec["dtype"])


def make_slice(case):
    """The generator's mock slice (tests/golden/make_golden_compute_phase.py::make_slice), same seed."""
    s = GOLD["slice"]
    dtype, n_times, kind = case["slice_dtype"], case["n_times"], case["kind"]
    rng = np.random.RandomState(s["seed"])
    shape = (n_times, len(s["levels"]), len(s["latitude"]), len(s["longitude"]))
    dv = {}
    for i, v in enumerate(s["variables"]):
        if kind == "noise":
            a = rng.standard_normal(shape) * (3.0 + i) + (250.0 if i == 0 else 0.0)
        else:
            sp = rng.standard_normal((10,) + shape[1:])
            tm = rng.standard_normal((10, n_times))
            a = np.einsum("i,it,ilao->tlao", 10.0 * 0.6 ** np.arange(10), tm, sp) + 1e-3 * rng.standard_normal(shape)
            a = a + (250.0 if i == 0 else 0.0)
        dv[v] = np.asarray(a, dtype=dtype)
    times = np.datetime64("2019-01-01T00", "ns") + np.arange(n_times) * np.timedelta64(1, "h")
    return dv, times


def reference_svd(X, cfg):
    """The reference's own two calls (era5_svd.py:251, :258; the randomized one draws from NumPy's global RNG, Q6)."""
    k = cfg["n_components"]
    if cfg["svd_type"] == "standard":
        U, s, Vt = np.linalg.svd(X, full_matrices=False)
        return U[:, :k], s[:k], Vt[:k]
    from sklearn.utils.extmath import randomized_svd

    np.random.seed(GOLD["np_random_seed"])
    return randomized_svd(X, n_components=k)


def selected_arrays(case):
    """Variable / level / time selection of the case in plain NumPy (request order, nearest resampling)."""
    cfg, s = case["config"], GOLD["slice"]
    dv, times = make_slice(case)
    variables = cfg["variables"].split(",")
    levels = [int(x) for x in cfg["levels"].split(",")]
    li = [s["levels"].index(lv) for lv in levels]
    hours = int(cfg["delta_time"][:-1])
    labels, ti = resample_nearest_index(times.astype(np.int64), hours * 3600 * 10**9)
    return variables, levels, labels.astype("datetime64[ns]"), [dv[v][ti][:, li] for v in variables]


@pytest.mark.parametrize("name", CASES)
def test_oracle_build_reproduces_the_references_own_chain_bit_for_bit(name):
    case = GOLD["cases"][name]
    cfg, res = case["config"], case["result"]
    _, _, _, arrays = selected_arrays(case)
    X, X_mean, X_std = build_matrix_np(arrays, cfg["mean_center"], cfg["scale"], cfg["delay_embedding"])
    if "X" in res["data_vars"]:
        want = golden_array(res["data_vars"]["X"])
        assert X.dtype == want.dtype and np.array_equal(X, want)
    for key, got in (("X_mean", X_mean), ("X_std", X_std)):
        if key in res["data_vars"]:
            want = golden_array(res["data_vars"][key])
            assert got is not None and got.dtype == want.dtype and np.array_equal(got, want), key
        else:
            assert got is None, key                      # Q3 (d == 1) / Q4 (scale without centring)
    # the reference's U, s, V are its own library call on that X
    U, sv, Vt = reference_svd(X, cfg)
    assert np.array_equal(sv, golden_array(res["data_vars"]["s"]))
    assert np.array_equal(U, golden_array(res["data_vars"]["U"]))
    assert np.array_equal(Vt, golden_array(res["data_vars"]["V"]))


def product_slice(case):
    s = GOLD["slice"]
    dv, times = make_slice(case)
    dims = ("time", "level", "latitude", "longitude")
    return Dataset({k: DataArray(v, dims) for k, v in dv.items()},
                   {"time": times, "level": np.asarray(s["levels"]), "latitude": np.asarray(s["latitude"]),
                    "longitude": np.asarray(s["longitude"])}, dict(s["attrs"]))


@pytest.mark.parametrize("name", CASES)
def test_host_flatten_and_delay_embedding_follow_the_references_own(name, monkeypatch):
    """flatten_era5_variables + apply_delay_embedding of the product (host index work on the stage's containers) on the
    selected / standardised variables: data, dims, time / original_variable / delay coordinates, the three attributes the
    reference sets on the matrix, and level / latitude / longitude = the components of the reference's space tuples."""
    from dmd_era5_b200.slice_tools import apply_delay_embedding, flatten_era5_variables, resample_era5_dataset, slice_era5_dataset

    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    case = GOLD["cases"][name]
    cfg, res = case["config"], case["result"]
    if "X" not in res["data_vars"]:
        pytest.skip("no data matrix recorded for this case")
    parsed = config_parser(dict(cfg), section="era5-svd")
    ds = product_slice(case)[parsed["variables"]]
    ds = resample_era5_dataset(slice_era5_dataset(ds, levels=parsed["levels"]), parsed["delta_time"])
    dv = {}
    for v in parsed["variables"]:
        a = np.asarray(ds[v].values)
        if parsed["mean_center"]:
            a = standardize_np(a, scale=bool(parsed["scale"]))[0]
        dv[v] = DataArray(a, ds[v].dims)
    flat = flatten_era5_variables(Dataset(dv, ds.coords, ds.attrs if not parsed["mean_center"] else {}))
    da = apply_delay_embedding(flat, parsed["delay_embedding"])
    want = res["data_vars"]["X"]
    assert list(da.dims) == want["dims"] and np.array_equal(np.asarray(da.values), golden_array(want))
    for key in ("time", "original_variable", "delay", "level", "latitude", "longitude"):
        g = golden_array(res["coords"][key])
        got = np.asarray(da.coord(key))
        assert np.array_equal(got.astype(g.dtype) if g.dtype.kind == "M" else got, g), key
    for key in ("original_variables", "space_coords", "delay_embedding"):
        assert da.attrs[key] == want["attrs"][key], key
    assert list(da.attrs) == list(want["attrs"])          # incl. the slice's own attributes when nothing was centred


def fake_device_arrays(ds, parsed_config, ops, comm=None, rank=0, world=1):
    """NumPy stand-in for stage._device_arrays (the device build + SVD): same contract, reference numerics."""
    variables, d = parsed_config["variables"], parsed_config["delay_embedding"]
    mc = bool(parsed_config["mean_center"])
    sc = bool(parsed_config["scale"]) and mc
    arrays = [np.asarray(ds[v].values) for v in variables]
    X0, _, _ = build_matrix_np(arrays, mc, sc, 1)
    stats = [standardize_np(a, scale=sc) for a in arrays] if mc else None
    mean = flatten_np([st[1] for st in stats]) if mc else None
    std = flatten_np([st[2] for st in stats]) if sc else None
    U, s, Vt = reference_svd(delay_embed_np(X0, d), parsed_config)
    S = X0.shape[0] // len(variables)
    return {"U": U, "s": s, "V": Vt, "X": X0 if parsed_config["save_data_matrix"] else None, "mean": mean,
            "std": std, "rows": (0, X0.shape[0]), "m0": X0.shape[0], "S": S}


@pytest.mark.parametrize("name", CASES)
def test_compute_phase_packaging_matches_the_references_own_result(name, monkeypatch):
    from dmd_era5_b200 import stage

    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", "/ROOT")
    monkeypatch.setattr(stage, "_device_arrays", fake_device_arrays)
    monkeypatch.setattr(stage, "get_ops", lambda *a, **k: None)
    from dmd_era5_b200 import slice_tools

    log = []
    for mod in (stage, slice_tools):
        monkeypatch.setattr(mod, "log_and_print", lambda lg, msg, level="info": log.append([level, " ".join(str(msg).split())]))
    case = GOLD["cases"][name]
    res = case["result"]
    parsed = config_parser(dict(case["config"]), section="era5-svd")
    out = stage._compute(product_slice(case), parsed)
    # progress lines of the phase, in order: slicing, resampling, standardisation (the fused device build replaces the
    # call but keeps its lines); the last two of the reference's are the SVD pair, emitted by the device leg
    kind = case["config"]["svd_type"]
    assert case["log"][-2:] == [["info", f"Performing {kind} SVD..."], ["info", f"{kind.capitalize()} SVD complete."]]
    assert log == case["log"][:-2]
    # variables: names in the reference's order, dims, dtype, values
    assert list(out.data_vars) == res["var_order"]
    for k, rec in res["data_vars"].items():
        got = out[k]
        want = golden_array(rec)
        assert list(got.dims) == rec["dims"], k
        assert np.asarray(got.values).dtype == want.dtype, k
        assert np.array_equal(np.asarray(got.values), want), k
    # coordinates: the reference's set (ORDER is not part of a NetCDF file's contract), dims, values, dtype kind
    assert sorted(out.coords) == sorted(res["coord_order"])
    for k, rec in res["coords"].items():
        want = golden_array(rec)
        dims, got = out.coords[k]
        got = np.asarray(got)
        assert list(dims) == rec["dims"], k
        assert got.dtype.kind == want.dtype.kind and np.array_equal(got.astype(want.dtype), want), k
    # attributes of the data matrix (original_variables / space_coords / delay_embedding, + the slice's own attributes
    # when nothing was centred: xarray arithmetic drops them otherwise)
    if "X" in res["data_vars"]:
        assert out["X"].attrs == res["data_vars"]["X"]["attrs"]
        assert list(out["X"].attrs) == list(res["data_vars"]["X"]["attrs"])
    for k in ("U", "s", "V", "X_mean", "X_std"):
        if k in res["data_vars"]:
            want_attrs = res["data_vars"][k]["attrs"]
            assert dict(out[k].attrs) == want_attrs, k
    # global attributes: names in order, values
    assert list(out.attrs) == list(res["attrs"])
    for k, v in res["attrs"].items():
        if k == "date_processed":
            continue
        got = out.attrs[k]
        assert (got.replace("/ROOT", "<ROOT>") if isinstance(got, str) else got) == v, k


@pytest.mark.parametrize("version", ["netcdf3", "cdf5"])
def test_result_with_the_references_variable_attributes_survives_the_file_round_trip(version, tmp_path, monkeypatch):
    """X / X_mean / X_std carry list-valued attributes (original_variables, space_coords): both classic writers hold them
    (comma-joined strings, like the global ``variables`` attribute) and every value comes back."""
    from dmd_era5_b200 import dataset, stage

    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", "/ROOT")
    monkeypatch.setattr(stage, "_device_arrays", fake_device_arrays)
    monkeypatch.setattr(stage, "get_ops", lambda *a, **k: None)
    if version == "cdf5":
        monkeypatch.setattr(dataset, "NETCDF3_VAR_LIMIT", 64)          # force the CDF-5 writer
    case = GOLD["cases"][CASES[1]]                                      # X, X_mean and X_std present
    out = stage._compute(product_slice(case), config_parser(dict(case["config"]), section="era5-svd"))
    path = str(tmp_path / "svd.nc")
    fmt = dataset.write_netcdf(out, path)
    if not dataset._have_xarray():
        assert fmt == ("NETCDF3_64BIT_DATA" if version == "cdf5" else "NETCDF3_64BIT")
    back = dataset.read_netcdf(path)
    assert sorted(back.data_vars) == sorted(out.data_vars)
    for k in out.data_vars:
        assert np.array_equal(np.asarray(back[k].values), np.asarray(out[k].values)), k
    for k in ("X", "X_mean", "X_std"):
        assert back[k].attrs["original_variables"] == ",".join(out[k].attrs["original_variables"])
        assert back[k].attrs["space_coords"] == "level,latitude,longitude"
    assert int(back["X"].attrs["delay_embedding"]) == case["config"]["delay_embedding"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_main_on_the_device_reproduces_the_references_own_main(name, tmp_path, monkeypatch):
    """The drop-in ``main`` (slice file -> device build -> SVD -> Dataset) against the result of the reference's OWN main on
    the same slice and configuration: X / X_mean / X_std, singular values, singular vectors up to sign (LAPACK's signs
    are not part of the contract: no svd_flip on the standard route, era5_svd.py:251), coordinates, attributes.
    Tolerances: float64 slices sigma 1e-9 relative, vectors 1e-6 rad; the float32 slice 1e-4 / 1e-3."""
    from dmd_era5_b200.dataset import write_netcdf
    from dmd_era5_b200.era5_svd import main
    from oracle.compare import sigma_rel_err, vector_angles

    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    case = GOLD["cases"][name]
    res, cfg = case["result"], dict(case["config"])
    parsed = config_parser(dict(cfg), section="era5-svd")
    os.makedirs(os.path.dirname(parsed["era5_slice_path"]), exist_ok=True)
    write_netcdf(product_slice(case), parsed["era5_slice_path"])
    np.random.seed(GOLD["np_random_seed"])          # the randomized route draws Omega from NumPy's global RNG, like the reference
    out, added, retrieved = main(cfg, write_to_netcdf=False)
    assert not added and not retrieved
    f32 = case["slice_dtype"] == "float32"
    assert list(out.data_vars) == res["var_order"]
    for k in ("X", "X_mean", "X_std"):
        if k in res["data_vars"]:
            want, got = golden_array(res["data_vars"][k]), np.asarray(out[k].values)
            assert got.dtype == want.dtype and list(out[k].dims) == res["data_vars"][k]["dims"], k
            assert np.max(np.abs(got.astype(np.float64) - want)) <= (1e-4 if f32 else 1e-10), k
    U0, s0, V0 = (golden_array(res["data_vars"][k]).astype(np.float64) for k in ("U", "s", "V"))
    U, s, V = (np.asarray(out[k].values) for k in ("U", "s", "V"))
    assert U.dtype == golden_array(res["data_vars"]["U"]).dtype and U.shape == U0.shape and V.shape == V0.shape
    assert sigma_rel_err(s.astype(np.float64), s0) < (1e-4 if f32 else 1e-9)
    tol = 1e-3 if f32 else 1e-6
    assert vector_angles(U.astype(np.float64), U0).max() < tol and vector_angles(V.astype(np.float64).T, V0.T).max() < tol
    sign = np.sign(np.sum(U.astype(np.float64) * U0, axis=0))             # the same sign on U_j and V_j
    assert np.allclose(V.astype(np.float64) * sign[:, None], V0, atol=10 * tol)
    if cfg["svd_type"] == "randomized":
        assert np.all(sign == 1.0)                                       # svd_flip (extmath.py:964-972): the SAME signs
    for k, rec in res["coords"].items():
        want = golden_array(rec)
        dims, got = out.coords[k]
        got = np.asarray(got)
        assert list(dims) == rec["dims"] and got.dtype.kind == want.dtype.kind, k
        assert np.array_equal(got.astype(want.dtype), want), k
    if "X" in res["data_vars"]:
        for k in ("original_variables", "space_coords", "delay_embedding"):
            assert out["X"].attrs[k] == res["data_vars"]["X"]["attrs"][k], k
        assert list(out["X"].attrs) == list(res["data_vars"]["X"]["attrs"])
    assert list(out.attrs) == list(res["attrs"])
    for k in ("n_components", "variables", "levels", "mean_center", "scale", "delay_embedding", "svd_type", "save_data_matrix"):
        assert out.attrs[k] == res["attrs"][k], k


def test_the_xarray_stand_in_passed_the_references_own_tests():
    """The stand-in the golden vectors were generated on is itself checked by the reference's OWN tests (collected
    unchanged, ``xarray`` -> stand-in, ``dmd_era5`` -> the reference's functions from source):
    tests/golden/run_reference_tests_on_standin.py writes the outcome next to the vectors."""
    with open(os.path.join(HERE, "golden", "reference_tests_on_standin.log")) as f:
        lines = [ln.split() for ln in f if ln.strip() and not ln.startswith("#")]
    assert len(lines) == 33 and all(ln[0] == "PASSED" for ln in lines)
    names = {ln[1].split("[")[0] for ln in lines}
    for must in ("test_02_slice_tools.py::test_flatten_era5_variables_array_dims",
                 "test_02_slice_tools.py::test_apply_delay_embedding_dataarray_values",
                 "test_02_slice_tools.py::test_standardize_data", "test_02_slice_tools.py::test_resample_era5_dataset",
                 "test_03_era5_svd.py::test_svd_on_era5", "test_03_era5_svd.py::test_combine_svd_results"):
        assert must in names, must


def test_n_gpus_is_capped_by_the_number_of_row_tiles(monkeypatch):
    """``n_gpus`` larger than the number of 128-row tiles of the base matrix (a small slice): the stage uses as many GPUs
    as there are tiles - here one, i.e. the single-device path - instead of handing a rank an empty shard."""
    from dmd_era5_b200 import slice_tools, stage, stage_multi

    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", "/ROOT")
    monkeypatch.setattr(stage, "_device_arrays", fake_device_arrays)
    monkeypatch.setattr(stage, "get_ops", lambda *a, **k: None)
    monkeypatch.setattr(stage_multi, "compute_multi", lambda *a, **k: pytest.fail("no multi-GPU run for a one-tile matrix"))
    log = []
    for mod in (stage, slice_tools):
        monkeypatch.setattr(mod, "log_and_print", lambda lg, msg, level="info": log.append([level, str(msg)]))
    case = GOLD["cases"][CASES[0]]                                  # 48 base rows: one tile
    parsed = config_parser(dict(case["config"], n_gpus=4), section="era5-svd")
    out = stage._compute(product_slice(case), parsed)
    assert ["warning", "n_gpus = 4, but the matrix has only 1 row tile(s) of 128: using 1 GPU(s)."] in log
    assert np.array_equal(np.asarray(out["U"].values), golden_array(case["result"]["data_vars"]["U"]))


def test_xarray_branches_of_the_file_io_run_on_the_contract_stand_in(tmp_path, monkeypatch):
    """dataset.write_netcdf / read_netcdf take xarray + netCDF4 when they are importable (the reference's
    ``to_netcdf(path, format="NETCDF4")``, era5_svd.py:434) - never the case in this image.  With ``xarray`` resolving to
    the contract stand-in (pickle as the container) those branches at least RUN: Dataset.to_xarray (constructor
    signatures, coords as (dims, values) pairs, attributes), the NETCDF4 format argument, and the conversion back."""
    import sys
    import types

    from dmd_era5_b200 import dataset, stage

    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import xr_contract
    finally:
        sys.path.pop(0)
    monkeypatch.setitem(sys.modules, "xarray", xr_contract)
    monkeypatch.setitem(sys.modules, "netCDF4", types.ModuleType("netCDF4"))
    assert dataset._have_xarray()
    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", "/ROOT")
    monkeypatch.setattr(stage, "_device_arrays", fake_device_arrays)
    monkeypatch.setattr(stage, "get_ops", lambda *a, **k: None)
    case = GOLD["cases"][CASES[1]]
    out = stage._compute(product_slice(case), config_parser(dict(case["config"]), section="era5-svd"))
    path = str(tmp_path / "svd.nc")
    assert dataset.write_netcdf(out, path) == "NETCDF4"
    assert xr_contract.open_dataset(path).file_format == "NETCDF4"

    def norm(attrs):            # attributes come back the netCDF4 way (one-element lists as scalars, NumPy types)
        return {k: np.atleast_1d(np.asarray(v)).tolist() for k, v in attrs.items()}

    back = dataset.read_netcdf(path)
    assert list(back.data_vars) == list(out.data_vars) and sorted(back.coords) == sorted(out.coords)
    for k in out.data_vars:
        assert back[k].dims == out[k].dims and np.array_equal(np.asarray(back[k].values), np.asarray(out[k].values)), k
        assert norm(back[k].attrs) == norm(out[k].attrs), k
    for k, (dims, v) in out.coords.items():
        assert back.coords[k][0] == tuple(dims) and np.array_equal(np.asarray(back.coords[k][1]), np.asarray(v)), k
    assert norm(back.attrs) == norm(out.attrs)


def test_main_write_and_cache_hit_through_the_xarray_branches(tmp_path, monkeypatch):
    """main(write_to_netcdf=True) followed by a second main() with ``xarray`` resolving to the contract stand-in: the slice
    file is read, the result written, and the second call finds and MATCHES the stored result (list-valued attributes come
    back as lists here, not as the comma-joined strings of the classic formats) - the cache-hit path of
    era5_svd.py:157-227 on xarray-read attributes."""
    import sys
    import types

    from dmd_era5_b200 import dataset, slice_tools, stage

    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import xr_contract
    finally:
        sys.path.pop(0)
    monkeypatch.setitem(sys.modules, "xarray", xr_contract)
    monkeypatch.setitem(sys.modules, "netCDF4", types.ModuleType("netCDF4"))
    monkeypatch.setenv("TZ", "UTC")
    _time.tzset()
    monkeypatch.setenv("DMD_ERA5_ROOT", str(tmp_path))
    monkeypatch.setattr(stage, "_device_arrays", fake_device_arrays)
    monkeypatch.setattr(stage, "get_ops", lambda *a, **k: None)
    log = []
    for mod in (stage, slice_tools):
        monkeypatch.setattr(mod, "log_and_print", lambda lg, msg, level="info": log.append(str(msg)))
    case = GOLD["cases"][CASES[0]]
    cfg = dict(case["config"])
    parsed = config_parser(dict(cfg), section="era5-svd")
    dataset.write_netcdf(product_slice(case), parsed["era5_slice_path"])
    first, added, retrieved = stage.main(cfg, write_to_netcdf=True)
    assert os.path.exists(parsed["save_path"]) and not added and not retrieved
    log.clear()
    second, added, retrieved = stage.main(cfg, write_to_netcdf=False)
    assert "SVD results match configuration." in log and "Performing standard SVD..." not in log
    for k in first.data_vars:
        assert np.array_equal(np.asarray(second[k].values), np.asarray(first[k].values)), k
    # a request the stored result does not satisfy is recomputed
    log.clear()
    stage.main(dict(cfg, n_components=2), write_to_netcdf=False)
    assert "SVD results do not match configuration." in log or "SVD results in working directory do not match configuration." in log
