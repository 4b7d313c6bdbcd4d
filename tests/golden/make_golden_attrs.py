"""The attributes the reference's OWN ``add_config_attributes`` (src/dmd_era5/era5_svd/era5_svd.py:42-66) puts on the
result Dataset - names, ORDER, values, Python types - for parsed configurations produced by its own config_parser.
Extracted with ``ast`` and executed unchanged on an object with an ``attrs`` dict.  Run in the build container.

    python tests/golden/make_golden_attrs.py
"""
import ast
import json
import os
import sys
import types
from datetime import datetime

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_config import BASE, load_reference_config_parser  # noqa: E402

REF = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"


def main():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "add_config_attributes")
    ns = {"xr": types.SimpleNamespace(Dataset=object), "datetime": datetime}
    exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    parser = load_reference_config_parser()
    out = {"_generated_by": "tests/golden/make_golden_attrs.py from " + REF, "cases": []}
    for delta in ({}, {"variables": "temperature,u_component_of_wind", "levels": "1000,850", "scale": True, "mean_center": True,
                       "svd_type": "standard", "delay_embedding": 1, "save_data_matrix": False, "n_components": 3}):
        cfg = dict(BASE, **delta)
        parsed = parser(cfg, "era5-svd")
        ds = ns["add_config_attributes"](types.SimpleNamespace(attrs={"pre_existing": "kept"}), parsed)
        attrs = {k: (v.replace("/ROOT", "<ROOT>") if isinstance(v, str) else v) for k, v in ds.attrs.items()}
        datetime.fromisoformat(attrs["date_processed"])
        attrs["date_processed"] = "<ISO DATETIME>"
        out["cases"].append({"config": cfg, "attrs": attrs, "types": {k: type(v).__name__ for k, v in ds.attrs.items()}})
    with open(os.path.join(HERE, "config_attributes.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["cases"][1]["attrs"]))


if __name__ == "__main__":
    main()
