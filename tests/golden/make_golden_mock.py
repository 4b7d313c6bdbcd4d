"""The reference's OWN mock-data generator (src/dmd_era5/create_mock_data/create_mock_data.py:112-155,
``_generate_variable_data``: BASELINE config 1's input distribution) executed from source with NumPy's global RNG seeded,
against which ``oracle/synthetic_np.mock_era5_np`` is pinned (tests/test_oracle.py).  Run in the build container.

    python tests/golden/make_golden_mock.py
"""
import ast
import hashlib
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/create_mock_data/create_mock_data.py"


def load_reference_generator():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "_generate_variable_data")
    ns = {"np": np, "pd": pd}
    exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    return ns["_generate_variable_data"]


def main():
    gen = load_reference_generator()
    lats, lons = np.arange(90, -90, -5.0), np.arange(-180, 180, 5.0)       # create_mock_era5 :66-70
    out = {"_generated_by": "tests/golden/make_golden_mock.py from " + REF, "cases": []}
    for seed, n_times, variables, levels in ((0, 25, ["temperature", "u_component_of_wind"], [1000, 850, 500]),
                                             (3, 7, ["v_component_of_wind", "temperature", "specific_humidity"], [500])):
        times = pd.date_range(start="2019-01-01", periods=n_times, freq="h")
        np.random.seed(seed)
        rec = {"seed": seed, "n_times": n_times, "variables": variables, "levels": levels, "arrays": {}}
        for var in variables:                                             # same order of RNG draws as create_mock_era5 :74-77
            a = gen(var, times, levels, lats, lons)
            rec["arrays"][var] = {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest(),
                                  "first": a.ravel()[:5].tolist(), "sum": float(a.sum())}
        out["cases"].append(rec)
    with open(os.path.join(HERE, "mock_era5.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote mock_era5.json")


if __name__ == "__main__":
    main()
