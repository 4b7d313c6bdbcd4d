"""Golden vectors of the reference's OWN dvc_tools.py (src/dmd_era5/dvc_tools.py) for the parts that need no live DVC
repository: the YAML side-log written by add_config_to_dvc_log (:11-47) and the version-matching rules of
retrieve_data_from_dvc (:119-253).

Run in the build container (where /root/reference exists).  The reference module is loaded from its source file; `dvc`,
`git` and `pyprojroot` (absent from this image) are stubbed - none of the exercised code calls into them:
find_first_commit_with_md5_hash is replaced by a recorder that notes WHICH md5 the reference's matching loop selected and
returns None (so the function ends in its own "could not retrieve" ValueError without touching git).

    python tests/golden/make_golden_dvc.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types
from datetime import datetime

REF = "/root/reference/src/dmd_era5/dvc_tools.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dvc_tools.json")


def load_reference():
    for name, attrs in (("pyprojroot", {"here": lambda *a: "/ROOT"}), ("dvc", {}), ("dvc.repo", {"Repo": object}),
                        ("git", {"Repo": object})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_dvc_tools", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


SRC = "gs://gcp-public-data-arco-era5/ar/1959-2022-full_37-1h-0p25deg-chunk-1.zarr-v2"
SLICE_ATTRS = [
    {"source_path": SRC, "start_datetime": "2019-01-01T00", "end_datetime": "2019-01-05T00", "hours_delta_time": 1,
     "variables": ["temperature"], "levels": [1000], "date_downloaded": "2024-05-01T10:00:00"},
    {"source_path": SRC, "start_datetime": "2019-01-01T00", "end_datetime": "2019-01-05T00", "hours_delta_time": 1,
     "variables": ["temperature", "u_component_of_wind"], "levels": [1000, 850], "date_downloaded": "2024-06-01T10:00:00"},
    {"source_path": "other", "start_datetime": "2019-01-01T00", "end_datetime": "2019-01-05T00", "hours_delta_time": 1,
     "variables": ["temperature", "u_component_of_wind", "v_component_of_wind"], "levels": [1000, 850, 500],
     "date_downloaded": "2024-07-01T10:00:00"},
]
SVD_ATTRS = [
    {"source_path": SRC, "n_components": 10, "variables": ["temperature"], "levels": [1000], "mean_center": 1, "scale": 0,
     "delay_embedding": 2, "svd_type": "randomized", "era5_slice_path": "/ROOT/data/era5_download/a.nc",
     "date_processed": "2024-05-02T10:00:00", "save_data_matrix": 1},
    {"source_path": SRC, "n_components": 10, "variables": ["temperature"], "levels": [1000], "mean_center": 1, "scale": 0,
     "delay_embedding": 2, "svd_type": "standard", "era5_slice_path": "/ROOT/data/era5_download/a.nc",
     "date_processed": "2024-05-03T10:00:00", "save_data_matrix": 0},
    {"source_path": SRC, "n_components": 10, "variables": ["u_component_of_wind", "temperature"], "levels": [1000, 850],
     "mean_center": 1, "scale": 1, "delay_embedding": 1, "svd_type": "randomized",
     "era5_slice_path": "/ROOT/data/era5_download/a.nc", "date_processed": "2024-05-04T10:00:00", "save_data_matrix": 0},
]
SLICE_REQUESTS = {
    "exact": {"variables": ["temperature"], "levels": [1000], "source_path": SRC},
    "subset of a richer slice": {"variables": ["u_component_of_wind"], "levels": [850], "source_path": SRC},
    "both variables": {"variables": ["u_component_of_wind", "temperature"], "levels": [1000], "source_path": SRC},
    "level not downloaded": {"variables": ["temperature"], "levels": [500], "source_path": SRC},
    "other source": {"variables": ["v_component_of_wind"], "levels": [500], "source_path": "other"},
    "unknown source": {"variables": ["temperature"], "levels": [1000], "source_path": "nowhere"},
}
SVD_BASE = {"source_path": SRC, "variables": ["temperature"], "levels": [1000], "delay_embedding": 2, "mean_center": True,
            "scale": False, "n_components": 10, "svd_type": "randomized", "save_data_matrix": True}
SVD_REQUESTS = {
    "exact (svd_type is not compared: the later entry wins)": {},
    "other svd_type still matches": {"svd_type": "standard"},
    "other n_components": {"n_components": 20},
    "other delay": {"delay_embedding": 1},
    "variable ORDER matters": {"variables": ["temperature", "u_component_of_wind"], "levels": [1000, 850], "scale": True,
                               "delay_embedding": 1},
    "same order as logged": {"variables": ["u_component_of_wind", "temperature"], "levels": [1000, 850], "scale": True,
                             "delay_embedding": 1},
    "scale differs": {"scale": True},
}


def main():
    ref = load_reference()
    picked = []
    ref.find_first_commit_with_md5_hash = lambda md5, path: picked.append(md5) or None
    out = {"_generated_by": "tests/golden/make_golden_dvc.py from /root/reference/src/dmd_era5/dvc_tools.py", "logs": {}, "matching": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for kind, attrs_list, requests, key in (("era5_slice", SLICE_ATTRS, SLICE_REQUESTS, "era5_slice_path"),
                                                ("era5_svd", SVD_ATTRS, SVD_REQUESTS, "era5_svd_path")):
            data_path = os.path.join(tmp, f"{kind}.nc")
            dvc_file = data_path + ".dvc"
            for i, attrs in enumerate(attrs_list):
                with open(dvc_file, "w") as f:                       # what `dvc add` leaves behind (only outs[0].md5 is read)
                    f.write(f"outs:\n- md5: {"abcdef0123456789" * 2}"[:-1] + f"{i}\n  size: 1\n  path: {kind}.nc\n")
                ref.add_config_to_dvc_log(dvc_file, data_path, attrs, git_add=False)
            with open(data_path + ".yaml") as f:
                out["logs"][kind] = {"attrs": attrs_list, "text": f.read()}
            out["matching"][kind] = {}
            for name, req in requests.items():
                cfg = dict(SVD_BASE, **req) if kind == "era5_svd" else dict(req)
                cfg[key] = data_path
                picked.clear()
                try:
                    ref.retrieve_data_from_dvc(cfg, kind)
                    res = {"selected": picked[-1] if picked else None, "error": None}
                except Exception as e:  # noqa: BLE001
                    res = {"selected": picked[-1] if picked else None,
                           "error": {"type": type(e).__name__, "message": " ".join(str(e).split())}}
                cfg.pop(key)
                out["matching"][kind][name] = {"request": cfg, **res}
            # error contract on missing files / keys / bad type
        errs = {}
        for label, cfg, kind in (("missing files", {"era5_slice_path": os.path.join(tmp, "absent.nc")}, "era5_slice"),
                                 ("missing key", {}, "era5_svd"), ("bad type", {}, "era5_other")):
            try:
                ref.retrieve_data_from_dvc(cfg, kind)
            except Exception as e:  # noqa: BLE001
                errs[label] = {"type": type(e).__name__, "message": " ".join(str(e).replace("'", "").split())}
        out["errors"] = errs
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)          # insertion order matters: the side-log lists the attributes in dict order
    print("wrote", OUT)
    for kind in out["matching"]:
        for name, r in out["matching"][kind].items():
            print(f"  {kind:10s} {name:55s} -> {r['selected']} {r['error']['message'][:50] if r['error'] else ''}")
    print(out["errors"])


if __name__ == "__main__":
    main()
