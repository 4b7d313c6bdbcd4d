"""The reference's OWN ``apply_delay_embedding`` (src/dmd_era5/slice_tools/slice_tools.py:214-274: the DataArray-level
wrapper - data, ``time`` / ``original_variable`` / ``delay`` coordinates, attrs, error messages), extracted with ``ast``
together with ``_apply_delay_embedding_np`` and executed unchanged; ``xr.DataArray`` is a plain record class (the function
only constructs one and reads ``.dims / .coords / .values / .attrs`` of its argument).  Run in the build container.

    python tests/golden/make_golden_delay_da.py
"""
import ast
import os
import types

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/slice_tools/slice_tools.py"


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None):
        self.values, self.dims, self.coords, self.attrs = np.asarray(data), tuple(dims), dict(coords), attrs


def main():
    src = open(REF).read()
    ns = {"np": np, "xr": types.SimpleNamespace(DataArray=DataArray), "sliding_window_view": sliding_window_view}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("_apply_delay_embedding_np", "apply_delay_embedding"):
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    fn = ns["apply_delay_embedding"]
    out = {}
    rng = np.random.RandomState(5)
    t0 = np.datetime64("2019-01-01T00", "ns")
    for i, (m0, T, d, variables) in enumerate([(6, 9, 1, ["temperature"]), (6, 9, 3, ["temperature", "u_component_of_wind"]),
                                               (4, 5, 5, ["t", "u"]), (10, 25, 2, ["a", "b"])]):
        X = rng.standard_normal((m0, T))
        time = t0 + np.arange(T) * np.timedelta64(1, "h")
        ov = np.repeat(variables, m0 // len(variables))
        da = DataArray(X, ("space", "time"), {"space": np.arange(m0), "time": time, "original_variable": ov}, {"source": "mock"})
        r = fn(da, d)
        out[f"c{i}_X"], out[f"c{i}_d"], out[f"c{i}_variables"] = X, d, np.array(variables)
        out[f"c{i}_values"], out[f"c{i}_dims"] = r.values, np.array(r.dims)
        out[f"c{i}_time"] = np.asarray(r.coords["time"]).astype("datetime64[ns]").astype(np.int64)
        out[f"c{i}_original_variable"] = np.asarray(r.coords["original_variable"][1])
        out[f"c{i}_delay"] = np.asarray(r.coords["delay"][1])
        out[f"c{i}_space"] = np.asarray(r.coords["space"])
        out[f"c{i}_delay_attr"] = r.attrs["delay_embedding"]
    out["n_cases"] = 4
    errs = {}
    good = DataArray(np.zeros((2, 3)), ("space", "time"), {"space": np.arange(2), "time": np.arange(3), "original_variable": np.array(["a", "a"])}, {})
    for label, arg, d in (("not a DataArray", np.zeros((2, 3)), 1),
                          ("bad dims", DataArray(np.zeros((2, 3)), ("x", "time"), good.coords, {}), 1),
                          ("bad coords", DataArray(np.zeros((2, 3)), ("space", "time"), {"space": np.arange(2), "time": np.arange(3)}, {}), 1),
                          ("d = 0", good, 0), ("d float", good, 1.5), ("d too large", good, 4)):
        try:
            fn(arg, d)
            errs[label] = "no error"
        except Exception as e:  # noqa: BLE001
            errs[label] = f"{type(e).__name__}: {e}"
    out["errors"] = np.array([f"{k} -> {v}" for k, v in errs.items()])
    np.savez_compressed(os.path.join(HERE, "delay_embedding_dataarray.npz"), **out)
    print("wrote delay_embedding_dataarray.npz"); print("\n".join(out["errors"]))


if __name__ == "__main__":
    main()
