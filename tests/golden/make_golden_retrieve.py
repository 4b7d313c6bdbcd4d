"""Decisions and log lines of the reference's OWN retrieve_era5_slice / retrieve_svd_results
(src/dmd_era5/era5_svd/era5_svd.py:69-227: which file in the working directory counts as a match, what is logged, when DVC
is consulted) for a table of scenarios.  Run in the build container.

The two functions are extracted from the reference's source with ``ast`` and executed unchanged.  Their free names are
bound to recorders / stand-ins: ``log_and_print`` records (level, message); ``xr.open_dataset`` returns an object whose
``.attrs`` are what xarray hands back for a NETCDF4 file written by the reference (length-1 attribute arrays come back as
scalars: a single variable as ``str``, a single level as ``numpy.int64``; longer ones as ``list[str]`` / ``numpy.ndarray``;
flags and counts as ``numpy.int64`` - README.md:97-119, era5_svd.py:87-99, 175-176); ``retrieve_data_from_dvc`` raises the
reference's own "nothing in DVC" errors; ``os.path.exists`` is answered from the scenario.

    python tests/golden/make_golden_retrieve.py
"""
import ast
import json
import os
import types
from typing import cast

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"
SRC = "gs://gcp-public-data-arco-era5/ar/1959-2022-full_37-1h-0p25deg-chunk-1.zarr-v2"


def as_file_attrs(attrs: dict) -> dict:
    """What xarray returns for these attributes after a NETCDF4 round trip."""
    out = {}
    for k, v in attrs.items():
        if isinstance(v, list) and v and isinstance(v[0], str):
            out[k] = v[0] if len(v) == 1 else list(v)
        elif isinstance(v, list):
            out[k] = np.int64(v[0]) if len(v) == 1 else np.asarray(v, dtype=np.int64)
        elif isinstance(v, bool) or isinstance(v, int):
            out[k] = np.int64(int(v))
        else:
            out[k] = v
    return out


def load(names, log, files, dvc_error):
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np, "os": types.SimpleNamespace(path=types.SimpleNamespace(exists=lambda p: p in files)), "cast": cast,
          "logger": None, "log_and_print": lambda lg, msg, level="info": log.append([level, " ".join(str(msg).split())]),
          "xr": types.SimpleNamespace(Dataset=object, DataArray=object,
                                      open_dataset=lambda p: types.SimpleNamespace(attrs=as_file_attrs(files[p]), path=p))}

    def retrieve_data_from_dvc(parsed_config, data_type="era5_slice"):
        raise dvc_error

    ns["retrieve_data_from_dvc"] = retrieve_data_from_dvc
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    return [ns[n] for n in names]


SLICE_FILE = {"variables": ["temperature", "u_component_of_wind"], "levels": [1000, 850], "source_path": SRC}
SVD_FILE = {"source_path": SRC, "n_components": 10, "variables": ["temperature"], "levels": [1000], "mean_center": True,
            "scale": False, "delay_embedding": 2, "svd_type": "randomized", "save_data_matrix": True}
SVD_CFG = {"source_path": SRC, "n_components": 10, "variables": ["temperature"], "levels": [1000], "mean_center": True,
           "scale": False, "delay_embedding": 2, "svd_type": "randomized", "save_data_matrix": True}

SLICE_SCENARIOS = {
    "subset of the file": ({"variables": ["temperature"], "levels": [850]}, SLICE_FILE, False),
    "all of the file, other order": ({"variables": ["u_component_of_wind", "temperature"], "levels": [850, 1000]}, SLICE_FILE, False),
    "variable missing": ({"variables": ["v_component_of_wind"], "levels": [850]}, SLICE_FILE, False),
    "level missing": ({"variables": ["temperature"], "levels": [500]}, SLICE_FILE, False),
    "other source": ({"variables": ["temperature"], "levels": [850], "source_path": "other"}, SLICE_FILE, False),
    "single-variable single-level file": ({"variables": ["temperature"], "levels": [1000]},
                                          {"variables": ["temperature"], "levels": [1000], "source_path": SRC}, False),
    "mismatch, DVC consulted": ({"variables": ["temperature"], "levels": [500]}, SLICE_FILE, True),
    "no file": ({"variables": ["temperature"], "levels": [850]}, None, False),
    "no file, DVC consulted": ({"variables": ["temperature"], "levels": [850]}, None, True),
}
SVD_SCENARIOS = {
    "exact": ({}, SVD_FILE, False),
    "svd_type differs (not compared)": ({"svd_type": "standard"}, SVD_FILE, False),
    "save_data_matrix differs (not compared)": ({"save_data_matrix": False}, SVD_FILE, False),
    "n_components differs": ({"n_components": 20}, SVD_FILE, False),
    "delay differs": ({"delay_embedding": 1}, SVD_FILE, False),
    "scale differs": ({"scale": True}, SVD_FILE, False),
    "multi-variable file, same order": ({"variables": ["temperature", "u_component_of_wind"], "levels": [1000, 850]},
                                        dict(SVD_FILE, variables=["temperature", "u_component_of_wind"], levels=[1000, 850]), False),
    "multi-variable file, other order": ({"variables": ["u_component_of_wind", "temperature"], "levels": [1000, 850]},
                                         dict(SVD_FILE, variables=["temperature", "u_component_of_wind"], levels=[1000, 850]), False),
    "levels in other order": ({"variables": ["temperature", "u_component_of_wind"], "levels": [850, 1000]},
                              dict(SVD_FILE, variables=["temperature", "u_component_of_wind"], levels=[1000, 850]), False),
    "mismatch, DVC consulted": ({"n_components": 20}, SVD_FILE, True),
    "no file": ({}, None, False),
    "no file, DVC consulted": ({}, None, True),
}


def main():
    out = {"_generated_by": "tests/golden/make_golden_retrieve.py from " + REF, "era5_slice": {}, "era5_svd": {}}
    for kind, fn_name, scenarios, key in (("era5_slice", "retrieve_era5_slice", SLICE_SCENARIOS, "era5_slice_path"),
                                          ("era5_svd", "retrieve_svd_results", SVD_SCENARIOS, "save_path")):
        for name, (cfg_delta, file_attrs, use_dvc) in scenarios.items():
            path = f"/WORK/{kind}.nc"
            if kind == "era5_slice":
                cfg = {"source_path": SRC, **cfg_delta}
            else:
                cfg = dict(SVD_CFG, **cfg_delta)
            cfg[key] = path
            log = []
            files = {path: file_attrs} if file_attrs is not None else {}
            (fn,) = load([fn_name], log, files, FileNotFoundError("DVC file or log file does not exist."))
            ds, from_dvc = fn(cfg, use_dvc=use_dvc)
            cfg.pop(key)
            out[kind][name] = {"config": cfg, "file_attrs": file_attrs, "use_dvc": use_dvc, "found": ds is not None,
                               "retrieved_from_dvc": bool(from_dvc), "log": [[lv, m.replace(path, "<PATH>")] for lv, m in log]}
    with open(os.path.join(HERE, "retrieve_cache.json"), "w") as f:
        json.dump(out, f, indent=1)
    for kind in ("era5_slice", "era5_svd"):
        for name, r in out[kind].items():
            print(f"{kind:10s} {name:42s} found={r['found']!s:5s} dvc={r['retrieved_from_dvc']!s:5s} {[m for _, m in r['log']][-1][:60]}")


if __name__ == "__main__":
    main()
