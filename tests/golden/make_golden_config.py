"""Golden vectors of the reference's OWN config_parser (src/dmd_era5/config_parser.py:51-...) for the era5-svd section.

Run in the build container (where /root/reference exists); the tests only read the committed JSON.  The reference module
is loaded from its source file; its two non-stdlib imports are satisfied without touching its code: ``pyprojroot.here`` is
stubbed (absent from this image; it only supplies the project root that the output paths are joined to - recorded as
<ROOT>) and ``dmd_era5.constants`` is loaded from the reference's own constants.py.

    python tests/golden/make_golden_config.py
"""
import importlib.util
import json
import os
import sys
import types
from datetime import datetime, timedelta

REF = "/root/reference/src/dmd_era5"
ROOT = "/ROOT"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_parser_era5_svd.json")


def load_reference_config_parser():
    stub = types.ModuleType("pyprojroot")
    stub.here = lambda *a: ROOT
    sys.modules["pyprojroot"] = stub
    pkg = types.ModuleType("dmd_era5")
    pkg.__path__ = [REF]
    sys.modules["dmd_era5"] = pkg
    for name in ("constants", "config_parser"):
        spec = importlib.util.spec_from_file_location(f"dmd_era5.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"dmd_era5.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["dmd_era5.config_parser"].config_parser


def encode(v):
    if isinstance(v, datetime):
        return {"__datetime__": v.isoformat()}
    if isinstance(v, timedelta):
        return {"__timedelta_s__": v.total_seconds()}
    if isinstance(v, str):
        return v.replace(ROOT, "<ROOT>")
    if isinstance(v, (list, tuple)):
        return [encode(x) for x in v]
    return v


BASE = {"source_path": "gs://gcp-public-data-arco-era5/ar/1959-2022-full_37-1h-0p25deg-chunk-1.zarr-v2",
        "start_datetime": "2019-01-01T00", "end_datetime": "2019-01-05T00", "delta_time": "1h",
        "variables": "temperature", "levels": "1000", "svd_type": "randomized", "delay_embedding": 2,
        "mean_center": True, "scale": False, "n_components": 10, "save_data_matrix": True}

CASES = {
    "default (config.ini)": {},
    "standard, d = 1": {"svd_type": "standard", "delay_embedding": 1},
    "several variables and levels": {"variables": "temperature,u_component_of_wind,v_component_of_wind", "levels": "1000,850,500"},
    "levels with spaces": {"levels": "1000, 850"},
    "delta 6h": {"delta_time": "6h"},
    "delta 1d": {"delta_time": "1d"},
    "delta 2w": {"delta_time": "2w", "end_datetime": "2019-03-01T00"},
    "delta 1m": {"delta_time": "1m", "end_datetime": "2019-06-01T00"},
    "delta 1y": {"delta_time": "1y", "end_datetime": "2021-01-01T00"},
    "delta upper case": {"delta_time": "6H"},
    "all levels": {"levels": "all"},
    "scale without mean_center": {"mean_center": False, "scale": True},
    "bad svd_type": {"svd_type": "truncated"},
    "bad delay 0": {"delay_embedding": 0},
    "bad delay float": {"delay_embedding": 1.5},
    "bad n_components 0": {"n_components": 0},
    "bad n_components str": {"n_components": "10"},
    "bad mean_center": {"mean_center": "yes"},
    "bad scale": {"scale": 1},
    "bad save_data_matrix": {"save_data_matrix": "True"},
    "end before start": {"end_datetime": "2018-12-31T00"},
    "range shorter than delta": {"delta_time": "2w"},
    "bad datetime": {"start_datetime": "2019-13-01T00"},
    "bad delta unit": {"delta_time": "5x"},
    "bad delta number": {"delta_time": "xh"},
    "unknown variable": {"variables": "temperature,not_a_variable"},
    "unknown level": {"levels": "1000,999"},
    "single level vars": {"variables": "all_single_level_vars"},
    "missing field svd_type": {"__drop__": "svd_type"},
    "missing field n_components": {"__drop__": "n_components"},
    "missing field variables": {"__drop__": "variables"},
}


def main():
    parser = load_reference_config_parser()
    out = {"_generated_by": "tests/golden/make_golden_config.py from /root/reference/src/dmd_era5/config_parser.py", "cases": {}}
    for name, delta in CASES.items():
        cfg = dict(BASE)
        drop = delta.get("__drop__")
        cfg.update({k: v for k, v in delta.items() if k != "__drop__"})
        if drop:
            cfg.pop(drop)
        rec = {"config": cfg}
        try:
            parsed = parser(cfg, "era5-svd")
            rec["parsed"] = {k: encode(v) for k, v in parsed.items()}
        except Exception as e:  # noqa: BLE001
            rec["error"] = {"type": type(e).__name__, "message": " ".join(str(e).split())}
        out["cases"][name] = rec
    try:
        parser(dict(BASE), "nope")
    except Exception as e:  # noqa: BLE001
        out["bad_section"] = {"type": type(e).__name__, "message": str(e)}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    n_ok = sum("parsed" in r for r in out["cases"].values())
    print(f"wrote {OUT}: {n_ok} parsed, {len(out['cases']) - n_ok} errors")


if __name__ == "__main__":
    main()
