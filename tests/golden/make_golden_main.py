"""Control flow of the reference's OWN ``main`` (src/dmd_era5/era5_svd/era5_svd.py:336-453) over a table of scenarios:
return flags, exception types and messages, log lines, and - for the compute phase - which pre-processing the configuration
switches on and which variables the result carries (quirks Q3: no X_mean / X_std when delay_embedding == 1, Q4: scale
without mean_center does nothing).  Run in the build container.

``main`` is extracted from the reference's source with ``ast`` and executed unchanged; every collaborator it calls is a
recorder / stand-in driven by the scenario (the REAL config_parser is used, loaded like in make_golden_config.py).

    python tests/golden/make_golden_main.py
"""
import ast
import json
import os
import sys
import types
from typing import cast

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_config import BASE, load_reference_config_parser  # noqa: E402

REF = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"


class Obj:
    """Stand-in for Dataset / DataArray values flowing through main: truthy, indexable, remembers what it is."""

    def __init__(self, what, **kw):
        self.what, self.kw, self.attrs, self.coords = what, kw, {"attr": 1}, "coords"

    def __getitem__(self, key):
        return Obj("ds[variables]", variables=key)

    def __bool__(self):
        return True


def run(config_delta, *, cached=None, cached_from_dvc=False, cache_raises=None, slice_found=True, compute_raises=None,
        write=False, write_raises=None, use_dvc=False, dvc_raises=None):
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "main")
    log, calls = [], []

    def rec(name, ret=None, raises=None):
        def f(*a, **k):
            calls.append([name, {kk: (vv if isinstance(vv, (int, float, str, bool, type(None))) else type(vv).__name__) for kk, vv in k.items()}])
            if raises is not None:
                raise raises
            return ret(*a, **k) if callable(ret) else ret
        return f

    results = Obj("svd_results")
    results.to_netcdf = rec("to_netcdf", raises=write_raises)

    def concat(objs, dim):
        calls.append(["xr.concat", {"copies": len(objs), "dim": dim}])
        return Obj("concat")

    ns = {"config": None, "cast": cast, "logger": None,
          "xr": types.SimpleNamespace(Dataset=object, DataArray=object, concat=concat),
          "config_parser": load_reference_config_parser(),
          "log_and_print": lambda lg, msg, level="info": log.append([level, " ".join(str(msg).split())]),
          "retrieve_svd_results": rec("retrieve_svd_results", ret=(cached, cached_from_dvc), raises=cache_raises),
          "retrieve_era5_slice": rec("retrieve_era5_slice", ret=(Obj("slice") if slice_found else None, False)),
          "slice_era5_dataset": rec("slice_era5_dataset", ret=Obj("sliced")),
          "resample_era5_dataset": rec("resample_era5_dataset", ret=Obj("resampled")),
          "standardize_data": rec("standardize_data", ret=lambda ds, scale=True: (Obj("std"), Obj("mean"), Obj("stddev") if scale else None)),
          "flatten_era5_variables": rec("flatten_era5_variables", ret=Obj("flat")),
          "apply_delay_embedding": rec("apply_delay_embedding", ret=Obj("embedded")),
          "svd_on_era5": rec("svd_on_era5", ret=("U", "s", "V"), raises=compute_raises),
          "combine_svd_results": rec("combine_svd_results", ret=results),
          "add_config_attributes": rec("add_config_attributes", ret=results),
          "space_coord_to_level_lat_lon": rec("space_coord_to_level_lat_lon", ret=results),
          "add_data_to_dvc": rec("add_data_to_dvc", raises=dvc_raises)}
    exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    cfg = dict(BASE, **config_delta)
    out = {"config": cfg, "write_to_netcdf": write, "use_dvc": use_dvc}
    try:
        res, added, retrieved = ns["main"](cfg, write_to_netcdf=write, use_dvc=use_dvc)
        out["returned"] = {"results": getattr(res, "what", None), "added_to_dvc": bool(added), "retrieved_from_dvc": bool(retrieved)}
    except Exception as e:  # noqa: BLE001
        out["raised"] = {"type": type(e).__name__, "message": " ".join(str(e).split()),
                         "cause": type(e.__cause__).__name__ if e.__cause__ is not None else None}
    out["log"] = [[lv, m.replace("/ROOT", "<ROOT>")] for lv, m in log]
    out["calls"] = calls
    return out


def main():
    out = {"_generated_by": "tests/golden/make_golden_main.py from " + REF, "flow": {}, "compute": {}}
    flow = {
        "cache hit": dict(cached=Obj("cached")),
        "cache hit from DVC": dict(cached=Obj("cached"), cached_from_dvc=True, use_dvc=True),
        "cache lookup raises": dict(cache_raises=OSError("disk on fire")),
        "slice missing": dict(slice_found=False),
        "slice missing, DVC on": dict(slice_found=False, use_dvc=True),
        "computed, not written": dict(),
        "compute raises": dict(compute_raises=ValueError("Input contains NaN.")),
        "written": dict(write=True),
        "write raises": dict(write=True, write_raises=PermissionError("read-only file system")),
        "written and added to DVC": dict(write=True, use_dvc=True),
        "DVC add raises": dict(write=True, use_dvc=True, dvc_raises=RuntimeError("not a DVC repository")),
        "DVC on but not written": dict(use_dvc=True),
    }
    for name, kw in flow.items():
        r = run({}, **kw)
        r.pop("calls")
        out["flow"][name] = {"scenario": {k: (repr(v) if isinstance(v, BaseException) else (v.what if isinstance(v, Obj) else v)) for k, v in kw.items()}, **r}
    for mc in (True, False):
        for sc in (True, False):
            for d in (1, 2):
                for save in (True, False):
                    r = run({"mean_center": mc, "scale": sc, "delay_embedding": d, "save_data_matrix": save})
                    calls = r["calls"]
                    std = [c for c in calls if c[0] == "standardize_data"]
                    comb = next(c for c in calls if c[0] == "combine_svd_results")[1]
                    out["compute"][f"mean_center={mc} scale={sc} d={d} save={save}"] = {
                        "config": {"mean_center": mc, "scale": sc, "delay_embedding": d, "save_data_matrix": save},
                        "standardize": None if not std else ("center" if std[0][1].get("scale") is False else "center+scale"),
                        "has_X": "X" in comb, "has_X_mean": comb.get("X_mean") is not None, "has_X_std": comb.get("X_std") is not None,
                        "copies_of_mean": [c[1]["copies"] for c in calls if c[0] == "xr.concat"],
                        "call_order": [c[0] for c in calls]}
    with open(os.path.join(HERE, "main_flow.json"), "w") as f:
        json.dump(out, f, indent=1)
    for name, r in out["flow"].items():
        print(f"{name:28s} {r.get('returned') or r.get('raised')}")
    for name, r in out["compute"].items():
        print(f"{name:48s} std={r['standardize']!s:13s} X={r['has_X']!s:5s} mean={r['has_X_mean']!s:5s} std={r['has_X_std']!s:5s} copies={r['copies_of_mean']}")


if __name__ == "__main__":
    main()
