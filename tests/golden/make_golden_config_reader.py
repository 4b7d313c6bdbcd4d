"""Golden vectors of the reference's OWN config_reader (src/dmd_era5/config_reader.py:16-62), loaded from its source file
(``pyprojroot.here`` and ``dmd_era5.logger.setup_logger`` stubbed: they only supply the default path and a logger): the
values it returns for the reference's own config files and for a set of small INI texts, and the exception TYPE and TEXT of
a missing file, a missing section and an unparsable value.

    python tests/golden/make_golden_config_reader.py
"""
import importlib.util
import json
import logging
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/dmd_era5/config_reader.py"

TEXTS = {
    "literals": '[s]\na = "text"\nb = 3\nc = 2.5\nd = True\ne = None\nf = [1, 2]\ng = {"k": 1}\nUPPER = "keys are lower-cased"\n',
    "utf-8 byte order mark": '﻿[s]\na = "bom"\n',
    "comments and blank lines": '# leading comment\n[s]\n\n; another\na = "x"  \n   b = 1\n',
    "unquoted string": '[s]\na = temperature\n',
    "syntax error": '[s]\na = "unterminated\n',
    "empty value": '[s]\na =\n',
    "two sections": '[s]\na = 1\n[t]\na = 2\n',
}
REQUESTS = [("literals", "s"), ("utf-8 byte order mark", "s"), ("comments and blank lines", "s"), ("unquoted string", "s"),
            ("syntax error", "s"), ("empty value", "s"), ("two sections", "t"), ("two sections", "missing")]


def load():
    stub = types.ModuleType("pyprojroot")
    stub.here = lambda *a: "/ROOT"
    sys.modules["pyprojroot"] = stub
    pkg = types.ModuleType("dmd_era5")
    pkg.__path__ = []
    lg = types.ModuleType("dmd_era5.logger")
    lg.setup_logger = lambda *a, **k: logging.getLogger("ConfigReader")
    sys.modules["dmd_era5"], sys.modules["dmd_era5.logger"] = pkg, lg
    spec = importlib.util.spec_from_file_location("dmd_era5.config_reader", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.config_reader


def record(fn, section, path):
    try:
        return {"values": fn(section, path)}
    except BaseException as e:  # noqa: BLE001
        return {"error": {"type": type(e).__name__, "message": str(e).replace(path, "<PATH>")}}


def main():
    logging.getLogger("ConfigReader").addHandler(logging.NullHandler())
    logging.getLogger("ConfigReader").propagate = False
    fn = load()
    out = {"_generated_by": "tests/golden/make_golden_config_reader.py from " + REF, "texts": TEXTS, "requests": [], "files": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for name, section in REQUESTS:
            p = os.path.join(tmp, "config.ini")
            with open(p, "w", encoding="utf-8") as f:
                f.write(TEXTS[name])
            out["requests"].append({"text": name, "section": section, **record(fn, section, p)})
        out["requests"].append({"text": None, "section": "s", **record(fn, "s", os.path.join(tmp, "does-not-exist.ini"))})
    for label, path, sections in (("reference config.ini", "/root/reference/config.ini", ["era5-download", "era5-svd"]),
                                  ("reference tests/config.ini", "/root/reference/tests/config.ini", ["test-section-0", "test-section-1"])):
        out["files"][label] = {"text": open(path, encoding="utf-8").read(), "sections": {s: record(fn, s, path) for s in sections}}
    with open(os.path.join(HERE, "config_reader.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in out["requests"]:
        print(f"{str(r['text']):28s} [{r['section']}]", r.get("values", r.get("error")))


if __name__ == "__main__":
    main()
