"""profiles/traffic.json from an ncu launch list of ONE plan replay of THIS build.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \\
        python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-torch-gpu-baseline [--workload cfg5]
    python tools/update_traffic.py cfg3 gpurun_out/launches.csv profiles/r2_final_launches_dram_b64.csv 1024

The fL stage of a replay = the launches from the focus network's crop + space-to-depth prepass (stem_s2d_kernel<1>)
through its global average pool (the second avgpool launch of the replay).  bench.py reads the entry back and reports it
as roofline.traffic only while the hash of the CUDA sources still matches (adafocus_b200.build.source_hash)."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from launch_dram_summary import load  # noqa: E402


def main():
    workload, src, dst, patches = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    from adafocus_b200 import build
    rows = load(src)
    names = [r["name"] for r in rows]
    start = next(i for i, n in enumerate(names) if n.startswith("stem_s2d") and "<1>" in n or n.startswith("stem_s2d<1"))
    pools = [i for i, n in enumerate(names) if n.startswith("avgpool") and i > start]
    end = pools[0]
    fl = rows[start:end + 1]
    dram = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in fl)
    t_us = sum(r["gpu__time_duration.sum"] for r in fl) / 1e3
    # the whole replay: from the glance network's prepass (first stem_s2d<2>) to the launch before the next one
    g0 = next(i for i, n in enumerate(names) if n.startswith("stem_s2d") and i < start)
    nxt = [i for i, n in enumerate(names) if n.startswith("stem_s2d") and i > start]
    replay = rows[g0:(nxt[0] if nxt else len(rows))]
    if os.path.abspath(src) != os.path.abspath(dst):
        shutil.copyfile(src, dst)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    rec = {}
    if os.path.exists(path):
        with open(path) as f:
            rec = json.load(f)
    rec[workload] = {
        "src_hash": build.source_hash(), "launch_list": os.path.relpath(dst, ROOT), "patches": patches,
        "fl_launches": [start - g0, end - g0], "fl_dram_bytes": dram, "fl_time_us_serialised": t_us,
        "replay_launches": len(replay),
        "replay_dram_bytes": sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in replay),
        "replay_time_us_serialised": sum(r["gpu__time_duration.sum"] for r in replay) / 1e3,
    }
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec[workload], indent=1))


if __name__ == "__main__":
    main()
