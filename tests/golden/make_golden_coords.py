"""The per-row level / latitude / longitude coordinates the reference writes to its NetCDF file:
the reference's OWN ``space_coord_to_level_lat_lon`` (src/dmd_era5/slice_tools/slice_tools.py:368-414), extracted with
``ast`` and executed unchanged on the ``space`` labels the preceding steps produce - (level, latitude, longitude) tuples
in the order of ``stack(space=["level", "latitude", "longitude"])`` (:323; last name fastest), tiled once per variable
(:346) and once per delay block (:259).  Pins ``slice_tools.space_coords`` (closed form).  Run in the build container.

    python tests/golden/make_golden_coords.py
"""
import ast
import itertools
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/slice_tools/slice_tools.py"


class Coords(dict):
    pass


class DS:
    def __init__(self, space):
        self.coords = Coords(space=types.SimpleNamespace(values=space))
        self.assigned = None

    def assign_coords(self, **kw):
        self.assigned = kw
        return self


def main():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "space_coord_to_level_lat_lon")
    ns = {"np": np, "xr": types.SimpleNamespace(Dataset=object), "logger": None, "log_and_print": lambda *a, **k: None}
    exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    fn = ns["space_coord_to_level_lat_lon"]
    out = {"_generated_by": "tests/golden/make_golden_coords.py from " + REF, "cases": []}
    for levels, lats, lons, n_vars, d in (([1000, 850, 500], [90.0, 85.0], [-180.0, -175.0, -170.0], 2, 1),
                                         ([1000], list(np.arange(90, 70, -5.0)), list(np.arange(-180, -160, 5.0)), 1, 2),
                                         ([850, 1000], [10.0, 5.0, 0.0], [0.0, 0.25], 3, 3)):
        tuples = list(itertools.product(levels, lats, lons))                       # stack(space=[level, latitude, longitude])
        space = np.empty(len(tuples), dtype=object)
        space[:] = tuples
        space = np.tile(np.tile(space, n_vars), d)                                 # :346 per variable, :259 per delay block
        ds = fn(DS(space))
        a = ds.assigned
        out["cases"].append({"levels": levels, "latitudes": lats, "longitudes": lons, "n_vars": n_vars, "d": d,
                             "space": ds.coords["space"].tolist(), "space_dtype": str(ds.coords["space"].dtype),
                             "level": a["level"][1].tolist(), "level_dtype": str(a["level"][1].dtype), "level_dim": a["level"][0],
                             "latitude": a["latitude"][1].tolist(), "latitude_dtype": str(a["latitude"][1].dtype),
                             "longitude": a["longitude"][1].tolist(), "longitude_dtype": str(a["longitude"][1].dtype)})
    with open(os.path.join(HERE, "space_coords.json"), "w") as f:
        json.dump(out, f)
    print("wrote space_coords.json;", [len(c["level"]) for c in out["cases"]], out["cases"][0]["level_dtype"], out["cases"][0]["latitude_dtype"], out["cases"][0]["space_dtype"])


if __name__ == "__main__":
    main()
