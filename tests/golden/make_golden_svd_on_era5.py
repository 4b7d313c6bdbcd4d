"""Outputs of the reference's OWN ``svd_on_era5`` (src/dmd_era5/era5_svd/era5_svd.py:230-263) on seeded inputs.

Run in the build container (where /root/reference exists).  The function is extracted from the reference's source file with
``ast`` and executed unchanged; its free names are bound to what the reference binds them to (numpy, scikit-learn's
``randomized_svd``) or to inert stand-ins (``xr`` - annotations only -, ``log_and_print``, ``logger``).  The unseeded
randomized call (quirk Q6) is made reproducible the way the oracle does it: ``np.random.seed(seed)`` right before the call.
The committed file pins ``oracle/svd_ref.py`` (tests/test_oracle.py): the oracle must reproduce these arrays BIT FOR BIT.

    python tests/golden/make_golden_svd_on_era5.py
"""
import ast
import os
import sys
import types

import numpy as np
from sklearn.utils.extmath import randomized_svd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"


def load_reference_svd_on_era5():
    src = open(REF).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "svd_on_era5")
    code = ast.get_source_segment(src, node)
    ns = {"np": np, "randomized_svd": randomized_svd, "xr": types.SimpleNamespace(DataArray=object),
          "log_and_print": lambda *a, **k: None, "logger": None}
    exec(compile(code, REF, "exec"), ns)
    return ns["svd_on_era5"]


class DA:                      # the only thing svd_on_era5 reads from its DataArray argument
    def __init__(self, values):
        self.values = values


def main():
    from oracle.synthetic_np import lowrank_field_np

    ref = load_reference_svd_on_era5()
    out = {}
    cases = [("std_f64", 600, 48, 7, "standard", np.float64, None), ("std_f32", 600, 48, 7, "standard", np.float32, None),
             ("rnd_f64_q7", 3000, 200, 12, "randomized", np.float64, 4), ("rnd_f32_q7", 3000, 200, 12, "randomized", np.float32, 4),
             ("rnd_f64_q4", 900, 60, 12, "randomized", np.float64, 8), ("rnd_wide", 80, 400, 6, "randomized", np.float64, 2),
             ("std_k_gt_n", 300, 10, 20, "standard", np.float64, None)]
    meta = []
    for name, m, n, k, kind, dt, seed in cases:
        X = lowrank_field_np(m, n, r=min(30, m, n), rho=0.85, seed=len(meta) + 1, dtype=dt)
        if seed is not None:
            np.random.seed(seed)
        U, s, V = ref(DA(X), {"svd_type": kind, "n_components": k})
        out[f"{name}_U"], out[f"{name}_s"], out[f"{name}_V"] = U, s, V
        meta.append((name, m, n, k, kind, np.dtype(dt).name, -1 if seed is None else seed, len(meta) + 1))
    try:
        ref(DA(np.zeros((3, 3))), {"svd_type": "truncated", "n_components": 1})
    except ValueError as e:
        out["bad_type_message"] = np.array(str(e))
    out["meta"] = np.array(meta, dtype=object)
    np.savez_compressed(os.path.join(HERE, "svd_on_era5_reference.npz"), **out)
    print("wrote svd_on_era5_reference.npz", os.path.getsize(os.path.join(HERE, "svd_on_era5_reference.npz")), "bytes;", str(out["bad_type_message"]))


if __name__ == "__main__":
    main()
