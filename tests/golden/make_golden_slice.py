"""The reference's OWN ``slice_era5_dataset`` / ``_get_dataset_time_bounds`` (src/dmd_era5/slice_tools/slice_tools.py:20-123),
extracted with ``ast`` and executed unchanged on a stand-in dataset that exposes ``.time.values``, ``.level.values`` and a
label-based ``.sel`` (time by inclusive slice, levels by list, KeyError on an unknown level - xarray's contract): which
time labels and which levels IN WHICH ORDER are selected, and every error message.  Run with TZ=UTC (quirk Q1: the bounds
go through datetime.fromtimestamp).

    TZ=UTC python tests/golden/make_golden_slice.py
"""
import ast
import json
import os
import types
from datetime import datetime

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/slice_tools/slice_tools.py"


class DS:
    def __init__(self, times, levels):
        self.time = types.SimpleNamespace(values=times)
        self.level = types.SimpleNamespace(values=np.asarray(levels))

    def sel(self, time, level):
        t = self.time.values
        keep = [i for i, x in enumerate(t) if np.datetime64(time.start, "ns") <= x <= np.datetime64(time.stop, "ns")]
        have = list(self.level.values)
        idx = []
        for lv in level:
            if lv not in have:
                raise KeyError(lv)
            idx.append(have.index(lv))
        return {"time_index": keep, "level_index": idx}


def main():
    assert datetime.fromtimestamp(0) == datetime(1970, 1, 1), "run with TZ=UTC"
    src = open(REF).read()
    log = []
    ns = {"np": np, "xr": types.SimpleNamespace(Dataset=object), "datetime": datetime, "logger": None,
          "log_and_print": lambda lg, msg, level="info": log.append([level, str(msg)])}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("slice_era5_dataset", "_get_dataset_time_bounds"):
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    fn = ns["slice_era5_dataset"]
    times = np.datetime64("2019-01-01T00", "ns") + np.arange(49) * np.timedelta64(1, "h")
    levels = [1000, 925, 850, 500]
    cases = {"defaults": {}, "levels in request order": {"levels": [500, 1000]}, "single level": {"levels": [925]},
             "time window (strings)": {"start_datetime": "2019-01-01T06", "end_datetime": "2019-01-02T00"},
             "time window (datetimes)": {"start_datetime": datetime(2019, 1, 1, 12), "end_datetime": datetime(2019, 1, 3)},
             "window + levels": {"start_datetime": "2019-01-02T00", "end_datetime": "2019-01-02T12", "levels": [850, 925]},
             "start before data": {"start_datetime": "2018-12-31T00"}, "end after data": {"end_datetime": "2019-02-01T00"},
             "start == end": {"start_datetime": "2019-01-01T06", "end_datetime": "2019-01-01T06"},
             "start after end": {"start_datetime": "2019-01-02T00", "end_datetime": "2019-01-01T00"},
             "unknown level": {"levels": [1000, 700]}, "empty level list = all": {"levels": []}}
    out = {"_generated_by": "tests/golden/make_golden_slice.py from " + REF, "n_times": 49, "levels": levels, "cases": {}}
    for name, kw in cases.items():
        log.clear()
        enc = {k: (v.isoformat() if isinstance(v, datetime) else v) for k, v in kw.items()}
        rec = {"kwargs": enc, "datetime_args": [k for k, v in kw.items() if isinstance(v, datetime)]}
        try:
            rec["selected"] = fn(DS(times, levels), **kw)
        except Exception as e:  # noqa: BLE001
            rec["error"] = {"type": type(e).__name__, "message": str(e), "cause": type(e.__cause__).__name__ if e.__cause__ else None}
        rec["log"] = [list(x) for x in log]
        out["cases"][name] = rec
    with open(os.path.join(HERE, "slice_era5_dataset.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out["cases"].items():
        print(f"{k:28s}", (f"times {v['selected']['time_index'][0]}..{v['selected']['time_index'][-1]} levels {v['selected']['level_index']}") if "selected" in v else v["error"]["message"][:90])


if __name__ == "__main__":
    main()
