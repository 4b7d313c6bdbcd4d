"""The compute phase of the reference's OWN ``main`` (src/dmd_era5/era5_svd/era5_svd.py:383-425) with ALL of its own
collaborators: ``slice_era5_dataset``, ``resample_era5_dataset``, ``standardize_data``, ``flatten_era5_variables``,
``apply_delay_embedding`` (slice_tools.py), ``svd_on_era5``, ``combine_svd_results``, ``add_config_attributes``
(era5_svd.py) and ``space_coord_to_level_lat_lon`` - every function extracted from the reference's source with ``ast`` and
executed UNCHANGED, on a small slice held in the xarray stand-in of xr_contract.py (xarray itself is not installable here;
the stand-in implements the documented behaviour of the dozen xarray calls the chain makes, resampling and label selection
through pandas).  Only the file / DVC look-ups of ``main`` are stubbed.  Recorded: the complete result Dataset - variables
(dims, dtype, values, attributes), coordinates (dims, dtype, values, ORDER), global attributes.  Run with TZ=UTC.

    TZ=UTC python tests/golden/make_golden_compute_phase.py
"""
import ast
import json
import os
import sys
from datetime import datetime, timedelta
from typing import cast

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from sklearn.utils.extmath import randomized_svd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import xr_contract as xr  # noqa: E402
from make_golden_config import BASE, load_reference_config_parser  # noqa: E402

REF_SVD = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"
REF_SLICE = "/root/reference/src/dmd_era5/slice_tools/slice_tools.py"

FILE_VARIABLES = ["temperature", "u_component_of_wind", "v_component_of_wind"]
FILE_LEVELS = [1000, 925, 850]
LAT = [10.0, 7.5, 5.0]
LON = [0.0, 2.5, 5.0, 7.5]
SLICE_ATTRS = {"source_path": BASE["source_path"], "variables": FILE_VARIABLES, "levels": FILE_LEVELS, "note": "mock slice"}
NP_RANDOM_SEED = 123               # np.random.seed(...) right before main: the randomized route draws from the global RNG (Q6)


def make_slice(dtype, n_times, kind):
    """Deterministic mock slice, dims (time, level, latitude, longitude) like the file era5_download writes; hourly from
    2019-01-01T00.  kind "noise": independent normal values (+ 250 on the first variable); kind "modes": ten space x time
    modes with amplitudes 10 * 0.6^i plus 1e-3 noise, i.e. a decaying spectrum (a well-posed randomized SVD)."""
    rng = np.random.RandomState(11)
    shape = (n_times, len(FILE_LEVELS), len(LAT), len(LON))
    dv = {}
    for i, v in enumerate(FILE_VARIABLES):
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


def functions(path, names=None):
    src = open(path).read()
    return {n.name: ast.get_source_segment(src, n) for n in ast.parse(src).body
            if isinstance(n, ast.FunctionDef) and (names is None or n.name in names)}


def run(config_delta, dtype, n_times, kind):
    dv, times = make_slice(dtype, n_times, kind)
    dims = ("time", "level", "latitude", "longitude")
    ds = xr.Dataset({k: xr.DataArray(v, dims) for k, v in dv.items()},
                    {"time": times, "level": np.asarray(FILE_LEVELS), "latitude": np.asarray(LAT), "longitude": np.asarray(LON)},
                    SLICE_ATTRS)
    log = []
    ns = {"np": np, "xr": xr, "datetime": datetime, "timedelta": timedelta, "cast": cast, "logger": None, "config": None,
          "sliding_window_view": sliding_window_view, "randomized_svd": randomized_svd,
          "log_and_print": lambda lg, msg, level="info": log.append([level, " ".join(str(msg).split())]),
          "config_parser": load_reference_config_parser(),
          "retrieve_svd_results": lambda parsed, use_dvc: (None, False),
          "retrieve_era5_slice": lambda parsed, use_dvc: (ds, False),
          "add_data_to_dvc": None}
    for path, names in ((REF_SLICE, None), (REF_SVD, ("svd_on_era5", "combine_svd_results", "add_config_attributes", "main"))):
        for name, code in functions(path, names).items():
            exec(compile(code, path, "exec"), ns)
    cfg = dict(BASE, **config_delta)
    np.random.seed(NP_RANDOM_SEED)
    res, added, retrieved = ns["main"](cfg, write_to_netcdf=False, use_dvc=False)
    assert not added and not retrieved
    return cfg, res, log


def enc_values(a):
    a = np.asarray(a)
    if a.dtype.kind == "M":
        return {"dtype": "datetime64[ns]", "values": a.astype("datetime64[ns]").astype(np.int64).tolist()}
    if a.dtype.kind == "O":                     # labels that are Python objects (must not survive to the result)
        return {"dtype": "object", "values": [list(x) if isinstance(x, tuple) else x for x in a.tolist()]}
    if a.dtype.kind in "US":
        return {"dtype": "str", "values": a.tolist()}
    return {"dtype": str(a.dtype), "values": a.astype(np.float64).tolist() if a.dtype.kind == "f" else a.tolist()}


def enc_attr(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.integer, np.floating)):
        return v.item()
    return v


def record(res):
    out = {"data_vars": {}, "coords": {}, "coord_order": list(res.coords.keys()), "var_order": list(res.data_vars.keys())}
    for k, v in res.data_vars.items():
        out["data_vars"][k] = {"dims": list(v.dims), "attrs": {a: enc_attr(b) for a, b in v.attrs.items()}, **enc_values(v.values)}
    for k, c in res.coords.items():
        out["coords"][k] = {"dims": list(c.dims), **enc_values(c.values)}
    attrs = {a: enc_attr(b) for a, b in res.attrs.items()}
    datetime.fromisoformat(attrs["date_processed"])
    attrs["date_processed"] = "<now>"
    out["attrs"] = {k: (v.replace("/ROOT", "<ROOT>") if isinstance(v, str) else v) for k, v in attrs.items()}
    return out


CASES = {
    "centre, d=2, two variables, levels in request order, 2h resampling": (
        {"variables": "temperature,u_component_of_wind", "levels": "850,1000", "delta_time": "2h", "svd_type": "standard",
         "n_components": 3, "delay_embedding": 2, "mean_center": True, "scale": False, "save_data_matrix": True}, np.float64, 13, "noise"),
    "centre + scale, d=3, one variable, all file levels": (
        {"variables": "v_component_of_wind", "levels": "1000,925,850", "delta_time": "1h", "svd_type": "standard",
         "n_components": 4, "delay_embedding": 3, "mean_center": True, "scale": True, "save_data_matrix": True}, np.float64, 13, "noise"),
    "centre + scale, d=1 (Q3: no X_mean / X_std)": (
        {"variables": "u_component_of_wind,temperature", "levels": "925", "delta_time": "1h", "svd_type": "standard",
         "n_components": 2, "delay_embedding": 1, "mean_center": True, "scale": True, "save_data_matrix": True}, np.float64, 13, "noise"),
    "scale without centring (Q4), d=2, no data matrix": (
        {"variables": "temperature", "levels": "1000,850", "delta_time": "3h", "svd_type": "standard",
         "n_components": 2, "delay_embedding": 2, "mean_center": False, "scale": True, "save_data_matrix": False}, np.float64, 13, "noise"),
    "no centring, d=1, data matrix kept (slice attributes reach X)": (
        {"variables": "temperature,v_component_of_wind", "levels": "850", "delta_time": "1h", "svd_type": "standard",
         "n_components": 3, "delay_embedding": 1, "mean_center": False, "scale": False, "save_data_matrix": True}, np.float64, 13, "noise"),
    "float32 slice, centre, d=2": (
        {"variables": "temperature,u_component_of_wind,v_component_of_wind", "levels": "1000", "delta_time": "1h",
         "svd_type": "standard", "n_components": 3, "delay_embedding": 2, "mean_center": True, "scale": False,
         "save_data_matrix": True}, np.float32, 13, "noise"),
    "randomized, float64, centre, d=2, two variables, 49 hourly snapshots": (
        {"variables": "u_component_of_wind,temperature", "levels": "1000,850", "delta_time": "1h", "svd_type": "randomized",
         "n_components": 5, "delay_embedding": 2, "mean_center": True, "scale": False, "save_data_matrix": True}, np.float64, 49, "modes"),
    "randomized, float32, centre, d=1, three variables, 2h resampling of 97 hourly snapshots": (
        {"variables": "temperature,u_component_of_wind,v_component_of_wind", "levels": "925,1000,850", "delta_time": "2h",
         "svd_type": "randomized", "n_components": 4, "delay_embedding": 1, "mean_center": True, "scale": False,
         "save_data_matrix": False}, np.float32, 97, "modes"),
}


def main():
    assert datetime.fromtimestamp(0) == datetime(1970, 1, 1), "run with TZ=UTC"
    out = {"_generated_by": "tests/golden/make_golden_compute_phase.py from " + REF_SVD + " and " + REF_SLICE,
           "slice": {"variables": FILE_VARIABLES, "levels": FILE_LEVELS, "latitude": LAT, "longitude": LON, "seed": 11, "attrs": SLICE_ATTRS},
           "np_random_seed": NP_RANDOM_SEED,
           "cases": {}}
    for name, (delta, dtype, n_times, kind) in CASES.items():
        cfg, res, log = run(delta, dtype, n_times, kind)
        rec = record(res)
        out["cases"][name] = {"config": cfg, "slice_dtype": np.dtype(dtype).name, "n_times": n_times, "kind": kind, "result": rec,
                              "log": [[lv, m.replace("/ROOT", "<ROOT>")] for lv, m in log]}
        print(f"{name}\n    vars {rec['var_order']}  coords {rec['coord_order']}\n    X attrs {list(rec['data_vars'].get('X', {}).get('attrs', {}))}"
              f"  U {res.data_vars['U'].shape} {res.data_vars['U'].dtype}")
    with open(os.path.join(HERE, "compute_phase.json"), "w") as f:
        json.dump(out, f, indent=None, separators=(",", ":"))
    print("wrote compute_phase.json", os.path.getsize(os.path.join(HERE, "compute_phase.json")), "bytes")


if __name__ == "__main__":
    main()
