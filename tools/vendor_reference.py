#!/usr/bin/env python
"""Make the UNMODIFIED reference available on the GPU box: copies the Python sources and YAML configs of
/root/reference (read-only, present only in the build container) into baseline/_ref/{ACT,STH}/.

baseline/_ref/ is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so the copy travels with the snapshot.  It is used only as test / measurement infrastructure:

  * bench.py `torch_gpu_baseline` and `--impl reference` time the reference's own classes,
  * tests/test_gpu_reference_band.py measures the reference's fp32 / TF32 / fp16 error band on the B200,
  * tests/test_gpu_validate_dropin.py runs the reference's own validate() loops over adafocus_b200/dropin/.

Nothing under adafocus_b200/ imports it.  Run by __graft_entry__.build() whenever /root/reference exists.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = {"ACT": "Experiments on ActivityNet, FCVID and Mini-Kinetics",
         "STH": "Experiments on Something-Something V1&V2"}
KEEP_EXT = (".py", ".yaml", ".sh", ".md")


def vendor(force=False):
    """Returns the destination directory, or None when the reference is not mounted (the GPU box)."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    stamp = os.path.join(DST, ".vendored")
    if os.path.exists(stamp) and not force:
        return DST
    for alias, name in TREES.items():
        src_tree = os.path.join(SRC, name)
        dst_tree = os.path.join(DST, alias)
        if os.path.isdir(dst_tree):
            shutil.rmtree(dst_tree)
        for dirpath, dirnames, filenames in os.walk(src_tree):
            dirnames[:] = [d for d in dirnames if d != "__pycache__" and not d.startswith(".")]
            rel = os.path.relpath(dirpath, src_tree)
            out_dir = os.path.join(dst_tree, rel) if rel != "." else dst_tree
            os.makedirs(out_dir, exist_ok=True)
            for fn in filenames:
                if fn.endswith(KEEP_EXT):
                    shutil.copyfile(os.path.join(dirpath, fn), os.path.join(out_dir, fn))
    with open(stamp, "w") as f:
        f.write("verbatim copy of /root/reference (*.py, *.yaml, *.sh, *.md); see tools/vendor_reference.py\n")
    return DST


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv))
