"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference classes.

Finds the reference either at /root/reference (build container) or at baseline/_ref/ (the verbatim, git-ignored copy
tools/vendor_reference.py makes so that it travels to the GPU box) and imports one of its two trees with the shims of
SURVEY.md section 8(c), all applied from OUTSIDE the reference files:

  1. `mobilenet_v2` / `resnet50` constructed with pretrained=False (the constructors otherwise download weights,
     ACT/models/gfv_net.py:237,261; no network);
  2. on a CUDA-less host only: Tensor.cuda / Module.cuda -> identity (hard-coded .cuda() at ACT/models/gfv_net.py:279,
     429 and ACT/models/ppo.py:70,135);
  3. `hydra` / `omegaconf` stand-ins (oracle/stubs) so that main_dist.py / evaluate.py can be imported and their
     validate() called directly.

Both trees name their packages `models`, `ops`, `basic_tools`: only one tree can be imported at a time; `unload()`
purges them from sys.modules / sys.path.  Only tests/, bench.py (reference arms) and tools/ may import this module.
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_NAMES = {"ACT": "Experiments on ActivityNet, FCVID and Mini-Kinetics",
          "STH": "Experiments on Something-Something V1&V2"}
_PKGS = ("models", "ops", "basic_tools", "main_dist", "evaluate", "stage3", "stage2", "stage1")
_STUBS = os.path.join(ROOT, "oracle", "stubs")


def tree_path(tree):
    """Directory of the reference tree 'ACT' / 'STH', or None when the reference is nowhere on this machine."""
    for cand in (os.path.join(ROOT, "baseline", "_ref", tree), os.path.join("/root/reference", _NAMES[tree])):
        if os.path.isfile(os.path.join(cand, "models", "gfv_net.py")):
            return cand
    return None


def available(tree="ACT"):
    return tree_path(tree) is not None


def unload():
    for k in [k for k in sys.modules if k.split(".")[0] in _PKGS]:
        del sys.modules[k]
    keep = []
    for p in sys.path:
        if any(p == tree_path(t) for t in _NAMES) or p.rstrip("/").endswith(("/dropin/act", "/dropin/sth")):
            continue
        keep.append(p)
    sys.path[:] = keep


def _cpu_shims():
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


class force_cpu:
    """Context manager for running the reference on the HOST cores of a machine that also has a GPU (bench.py's
    cpu_baseline / --impl reference): inside it `.cuda()` is the identity, exactly like shim 2 above, so the
    reference's hard-coded `.cuda()` calls (ACT/models/ppo.py:70,135; ACT/models/gfv_net.py:279,429) leave everything on
    the CPU.  The original methods are restored on exit."""

    def __enter__(self):
        self._saved = (torch.Tensor.cuda, torch.nn.Module.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda, torch.nn.Module.cuda = self._saved
        return False


def _stub_hydra():
    try:
        importlib.import_module("hydra")
        importlib.import_module("omegaconf")
    except ImportError:
        if _STUBS not in sys.path:
            sys.path.append(_STUBS)


def import_tree(tree, dropin=None):
    """Puts the reference tree first on sys.path (optionally with one of OUR drop-in `models` packages in front of it)
    and returns its path.  `dropin` = path of adafocus_b200/dropin/{act,sth} shadows the reference's `models`."""
    path = tree_path(tree)
    if path is None:
        raise FileNotFoundError(f"reference tree {tree} not found (neither baseline/_ref nor /root/reference)")
    unload()
    _cpu_shims()
    _stub_hydra()
    sys.path.insert(0, path)
    if dropin is not None:
        sys.path.insert(0, dropin)
    return path


def import_act():
    """-> the reference's ACT `models.gfv_net` module (GFV, get_patch, ...), weights not downloaded."""
    import_tree("ACT")
    ref = importlib.import_module("models.gfv_net")
    ref_mb = importlib.import_module("models.mobilenet")
    ref_rn = importlib.import_module("models.resnet")
    ref.mobilenet_v2 = lambda pretrained=True, **kw: ref_mb.mobilenet_v2(pretrained=False, **kw)
    ref.resnet50 = lambda pretrained=True, **kw: ref_rn.resnet50(pretrained=False, **kw)
    return ref


def import_sth():
    """-> the reference's STH `models.gfv_net` module (already constructs its backbones with pretrained=False)."""
    import_tree("STH")
    return importlib.import_module("models.gfv_net")


def import_entry(tree, dropin=None):
    """-> the reference's entry script module (ACT main_dist / STH evaluate) with `models` resolved either to the
    reference's own package or to the drop-in given.  validate(val_loader, model, criterion, args) is then callable."""
    import_tree(tree, dropin=dropin)
    mod = importlib.import_module("main_dist" if tree == "ACT" else "evaluate")
    if dropin is None and tree == "ACT":
        ref = sys.modules["models.gfv_net"]
        ref_mb = importlib.import_module("models.mobilenet")
        ref_rn = importlib.import_module("models.resnet")
        ref.mobilenet_v2 = lambda pretrained=True, **kw: ref_mb.mobilenet_v2(pretrained=False, **kw)
        ref.resnet50 = lambda pretrained=True, **kw: ref_rn.resnet50(pretrained=False, **kw)
    return mod
