"""The calls the reference's OWN dvc_tools.py (src/dmd_era5/dvc_tools.py:50-63, :95-116, :209-250) makes INTO dvc and git -
the part that needs a live repository - recorded with stand-ins for ``dvc.repo.Repo`` / ``git.Repo`` that note every
call (constructor root, ``add``, ``index.add``, ``git.checkout``, ``config["remote"]``, ``fetch``, ``checkout``) and
answer according to the scenario: adding a result, retrieval from the local cache, from the default remote, and the three
ways a retrieval fails after a version matched.  Also recorded: what is printed, and the exception.

    python tests/golden/make_golden_dvc_calls.py
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/dvc_tools.py"
MD5 = "0123456789abcdef0123456789abcdef"
SRC = "gs://mock"
ATTRS = {"source_path": SRC, "n_components": 10, "variables": ["temperature"], "levels": [1000], "mean_center": 1, "scale": 0,
         "delay_embedding": 2, "svd_type": "randomized", "era5_slice_path": "<ROOT>/data/era5_download/a.nc",
         "date_processed": "2024-05-02T10:00:00", "save_data_matrix": 1}
REQUEST = {"source_path": SRC, "variables": ["temperature"], "levels": [1000], "delay_embedding": 2, "mean_center": True,
           "scale": False, "n_components": 10}


def make_standins(calls, root, scenario):
    """dvc.repo.Repo / git.Repo stand-ins; paths are recorded relative to the project root."""
    rel = lambda p: os.path.relpath(p, root) if isinstance(p, str) and p.startswith(root) else p   # noqa: E731

    class DvcRepo:
        def __init__(self, r):
            calls.append(["DvcRepo", rel(r) if r != root else "<ROOT>"])
            self.config = {"remote": {"origin": {"url": "s3://bucket"}} if scenario.get("remote") else {}}

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def add(self, path):
            calls.append(["dvc.add", rel(path)])
            with open(path + ".dvc", "w") as f:                      # what ``dvc add`` leaves behind
                f.write(f"outs:\n- md5: {MD5}\n  size: 1\n  path: {os.path.basename(path)}\n")

        def fetch(self, targets):
            calls.append(["dvc.fetch", [rel(t) for t in targets]])
            return scenario.get("fetched", 0)

        def checkout(self, targets):
            calls.append(["dvc.checkout", [rel(t) for t in targets]])
            return {"modified": [rel(t)[:-4] for t in targets]}

    class GitRepo:
        def __init__(self, r):
            calls.append(["GitRepo", "<ROOT>" if r == root else rel(r)])
            outer = self

            class Index:
                def add(self, files):
                    calls.append(["git.index.add", [rel(f) for f in files]])

            class Git:
                def checkout(self, commit, path):
                    calls.append(["git.checkout", commit, rel(path)])

            self.index, self.git = Index(), Git()

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    return DvcRepo, GitRepo


def load_reference(root, DvcRepo, GitRepo):
    for name, attrs in (("pyprojroot", {"here": lambda *a: os.path.join(root, *a)}), ("dvc", {}),
                        ("dvc.repo", {"Repo": DvcRepo}), ("git", {"Repo": GitRepo})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_dvc_tools_calls", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


SCENARIOS = {
    "add a result": {"op": "add"},
    "retrieve: in the local cache": {"op": "retrieve", "commit": "abc1234", "cached": True},
    "retrieve: fetched from the default remote": {"op": "retrieve", "commit": "abc1234", "remote": True, "fetched": 1},
    "retrieve: no default remote": {"op": "retrieve", "commit": "abc1234"},
    "retrieve: remote has nothing": {"op": "retrieve", "commit": "abc1234", "remote": True, "fetched": 0},
    "retrieve: no commit mentions the md5": {"op": "retrieve", "commit": None},
}


def run(scenario):
    calls = []
    with tempfile.TemporaryDirectory() as root:
        DvcRepo, GitRepo = make_standins(calls, root, scenario)
        mod = load_reference(root, DvcRepo, GitRepo)
        os.makedirs(os.path.join(root, "data/era5_svd"))
        data_path = os.path.join(root, "data/era5_svd/result.nc")
        open(data_path, "w").write("x")
        out = io.StringIO()
        rec = {"scenario": scenario}
        try:
            with contextlib.redirect_stdout(out):
                if scenario["op"] == "add":
                    mod.add_data_to_dvc(data_path, dict(ATTRS))
                    rec["log_file"] = open(data_path + ".yaml").read()
                else:
                    open(data_path + ".dvc", "w").write(f"outs:\n- md5: {MD5}\n")
                    open(data_path + ".yaml", "w").write(f"{MD5}:\n" + "".join(f"  {k}: {v}\n" for k, v in ATTRS.items()))
                    if scenario.get("cached"):
                        os.makedirs(os.path.join(root, ".dvc/cache/files/md5", MD5[:2]))
                        open(os.path.join(root, ".dvc/cache/files/md5", MD5[:2], MD5[2:]), "w").write("x")
                    mod.find_first_commit_with_md5_hash = lambda md5, path: (calls.append(["git log -S", md5, os.path.relpath(path, root)]),
                                                                               scenario["commit"])[1]
                    mod.retrieve_data_from_dvc(dict(REQUEST, era5_svd_path=data_path), "era5_svd")
        except Exception as e:  # noqa: BLE001
            rec["raised"] = {"type": type(e).__name__, "message": str(e)}
        rec["calls"] = calls
        rec["printed"] = out.getvalue().replace(root, "<ROOT>")
    return rec


if __name__ == "__main__":
    res = {"_generated_by": "tests/golden/make_golden_dvc_calls.py from " + REF, "md5": MD5, "attrs": ATTRS, "request": REQUEST,
           "scenarios": {name: run(sc) for name, sc in SCENARIOS.items()}}
    with open(os.path.join(HERE, "dvc_calls.json"), "w") as f:
        json.dump(res, f, indent=1)
    for name, r in res["scenarios"].items():
        print(f"{name:45s} {[c[0] for c in r['calls']]} {r.get('raised', {}).get('type', '')}")
