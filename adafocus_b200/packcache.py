"""Persisted packed-weight cache (reference call sites that precede every evaluation:
checkpoint loads at ACT/main_dist.py:92-110 and STH/evaluate.py:40-81,136-146).

The kernels consume derived tensors -- BatchNorm folded, OIHW -> K-major fp16, stems re-indexed for the space-to-depth
view, fused-block bias columns -- built on the host by adafocus_b200.engine.pack_*.  With a cache directory set
(`AF_PACK_CACHE=/path` or `set_cache_dir()`), a runner is stored under a key derived from the CONTENT of the module's
state_dict, so a later process that loads the same checkpoint maps the packed blob straight to the device instead of
repacking.  Off by default: nothing is written outside the directory the user names."""
import hashlib
import os
import tempfile

import torch

PACK_FORMAT = "r2.2"
_cache_dir = os.environ.get("AF_PACK_CACHE") or None
stats = {"hits": 0, "misses": 0, "stores": 0}


def set_cache_dir(path):
    global _cache_dir
    _cache_dir = path or None


def cache_dir():
    return _cache_dir


def _code_tag():
    h = hashlib.blake2b(digest_size=8)
    here = os.path.dirname(os.path.abspath(__file__))
    for rel in ("engine.py", os.path.join("models", "mobilenet.py"), os.path.join("models", "resnet.py"),
                os.path.join("models", "ppo.py"), os.path.join("models", "gfv_net.py"),
                os.path.join("models_sth", "mobilenetv2.py")):
        with open(os.path.join(here, rel), "rb") as f:
            h.update(f.read())
    for knob in ("AF_NO_MBCONV_FUSED", "AF_MB_NO_BIAS_MMA", "AF_NO_BOTTLENECK_FUSED"):
        h.update(f"{knob}={os.environ.get(knob)}".encode())
    return h.hexdigest()


def content_hash(module, extra=""):
    """blake2b over names, shapes, dtypes and raw bytes of module.state_dict() (one D2H copy per tensor)."""
    h = hashlib.blake2b(digest_size=16)
    h.update(f"{PACK_FORMAT}|{_code_tag()}|{extra}".encode())
    for name, t in module.state_dict().items():
        c = t.detach().to("cpu").contiguous()
        h.update(f"{name}|{tuple(c.shape)}|{c.dtype}".encode())
        h.update(c.reshape(-1).view(torch.uint8).numpy().tobytes() if c.numel() else b"")
    return h.hexdigest()


def cached_runner(module, tag, build, key, device=None, extra=""):
    """build() -> runner, or the stored runner for identical weights.  `key` is the in-process identity (_param_key)
    stamped onto the returned object."""
    if _cache_dir is None:
        return build()
    # structure that is not in the state_dict: TemporalShift wrappers (segment count / fold divisor)
    extra += ";".join(f"{n}:{m.n_segment}:{m.fold_div}" for n, m in module.named_modules() if hasattr(m, "n_segment"))
    digest = content_hash(module, extra=tag + "|" + extra)
    path = os.path.join(_cache_dir, f"{tag}-{digest}.pt")
    if device is None:
        device = next(module.parameters()).device
    if os.path.exists(path):
        try:
            runner = torch.load(path, map_location=device, weights_only=False)
            runner.key = key
            stats["hits"] += 1
            return runner
        except Exception:
            pass                     # unreadable / truncated blob: rebuild and overwrite
    stats["misses"] += 1
    runner = build()
    os.makedirs(_cache_dir, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=_cache_dir, suffix=".tmp")
    os.close(fd)
    try:
        torch.save(runner, tmp)
        os.replace(tmp, path)        # atomic: concurrent ranks race benignly
        stats["stores"] += 1
    except Exception:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return runner
