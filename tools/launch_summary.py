"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + compact sequence."""
import collections
import csv
import re
import sys


def load(path):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for row in csv.DictReader(lines[start:]):
        try:
            rows.append((row["Kernel Name"], row["Grid Size"], float(row["Metric Value"].replace(",", ""))))
        except Exception:
            pass
    return rows


def short(name):
    s = re.sub(r"\(.*", "", name).replace("_kernel", "").replace("void ", "")
    return re.sub(r"<.*", "", s).split("::")[-1]


if __name__ == "__main__":
    rows = load(sys.argv[1])
    tot = sum(r[2] for r in rows)
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms (serialised, cold-cache)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, _, t in rows:
        agg[short(name)][0] += 1
        agg[short(name)][1] += t
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:28s} n={n:4d} {t / 1e6:8.3f} ms {100 * t / tot:5.1f}%")
    if "--seq" in sys.argv:
        print(" ".join(f"{i}:{short(n)[:9]}:{t / 1e3:.0f}" for i, (n, _, t) in enumerate(rows)))
