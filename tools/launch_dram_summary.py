"""Per-launch time / DRAM bytes / tensor-pipe utilisation from an ncu multi-metric launch list."""
import collections
import csv
import re
import sys


def load(path):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        d = rows.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1].replace("_kernel", "")})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    return list(rows.values())


if __name__ == "__main__":
    rows = load(sys.argv[1])
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    tot = sum(r["gpu__time_duration.sum"] for r in rows) / 1e6
    gb = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in rows) / 1e9
    print(f"{len(rows)} launches, {tot:.3f} ms, {gb:.1f} GB DRAM traffic, {gb / tot:.2f} TB/s average")
    for i, r in enumerate(rows):
        t = r["gpu__time_duration.sum"] / 1e3
        g = (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]) / 1e9
        if t > thr:
            tp = r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
            print(f"  {i:3d} {r['name'][:14]:14s} {t:7.0f} us {g:6.2f} GB {g / t * 1e3:5.1f} TB/s  tensor {tp:4.0f}%")
