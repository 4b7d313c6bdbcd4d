"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<.*", "", name).replace("void ", "").strip()
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = {"nsecond": v / 1e6, "ns": v / 1e6, "usecond": v / 1e3, "us": v / 1e3, "msecond": v, "ms": v,
             "second": v * 1e3, "s": v * 1e3}[unit]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot:.3f} ms total (ncu serialised, cold cache: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% | {ms / c * 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
