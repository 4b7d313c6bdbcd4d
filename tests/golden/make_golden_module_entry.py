"""The ``if __name__ == "__main__":`` block of the reference's era5_svd.py (:455-478), extracted with ``ast`` and executed
unchanged with ``DvcRepo`` / ``main`` / ``log_and_print`` bound to recorders: which warnings are logged and how ``main`` is
called when the project is / is not a DVC repository.

    python tests/golden/make_golden_module_entry.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/era5_svd/era5_svd.py"


def run(is_repo: bool):
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.If) and "__main__" in ast.get_source_segment(src, n.test))
    block = compile(ast.Module(body=node.body, type_ignores=[]), REF, "exec")       # the block's own statements
    log, calls = [], []

    class Repo:
        def __init__(self, root):
            if not is_repo:
                raise RuntimeError("not a dvc repository")

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    ns = {"DvcRepo": Repo, "here": lambda: "/ROOT", "logger": None,
          "log_and_print": lambda lg, msg, level="info": log.append([level, str(msg)]),
          "main": lambda *a, **k: calls.append({"args": list(a), "kwargs": k})}
    exec(block, ns)
    return {"log": log, "main_calls": calls}


if __name__ == "__main__":
    out = {"_generated_by": "tests/golden/make_golden_module_entry.py from " + REF,
           "not a DVC repository": run(False), "DVC repository": run(True)}
    with open(os.path.join(HERE, "module_entry.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
