"""Correctness / timing driver of af_mbconv_rows: python tools/mbrows_debug.py [time | b2 b3 ...]"""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import (get_engine, mbconv_rows_spr, mbconv_rows_supported, pack_mbconv, pack_mbconv_rows)

dev = torch.device("cuda", 0)
eng = get_engine(dev)
CASES_ALL = {"b2": (32, 96, 24, 2, 112, False), "b3": (24, 144, 24, 1, 56, True), "b4": (24, 144, 32, 2, 56, False),
         "b5": (32, 192, 32, 1, 28, True), "b7": (32, 192, 64, 2, 28, False), "b8": (64, 384, 64, 1, 14, True)}
CASES = {k: v for k, v in CASES_ALL.items() if mbconv_rows_supported(1024, v[4], v[4], v[0], v[1], v[2], v[3])}


def make(cin, cexp, cout, stride, hw, res, n):
    torch.manual_seed(cexp + hw)
    x = torch.randn(n, hw, hw, cin, device=dev).half()
    w1 = torch.randn(cexp, cin, device=dev) / math.sqrt(cin)
    wd = torch.randn(cexp, 1, 3, 3, device=dev) / 3
    w2 = torch.randn(cout, cexp, device=dev) / math.sqrt(cexp)
    s1, b1 = torch.rand(cexp, device=dev) + 0.5, torch.randn(cexp, device=dev) * 0.2
    s2, b2 = torch.rand(cexp, device=dev) + 0.5, torch.randn(cexp, device=dev) * 0.2
    s3, b3 = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.2
    return x, (w1, s1, b1, wd, s2, b2, w2, s3, b3)


def ref_fn(x, ws, stride, res):
    w1, s1, b1, wd, s2, b2, w2, s3, b3 = ws
    cexp = w1.shape[0]
    xf = x.float().permute(0, 3, 1, 2)
    e = (F.conv2d(xf, (w1 * s1[:, None])[:, :, None, None]) + b1.view(1, -1, 1, 1)).clamp(0, 6)
    d = (F.conv2d(e, wd * s2.view(-1, 1, 1, 1), None, stride, 1, 1, cexp) + b2.view(1, -1, 1, 1)).clamp(0, 6)
    r = F.conv2d(d, (w2 * s3[:, None])[:, :, None, None]) + b3.view(1, -1, 1, 1)
    if res:
        r = r + xf
    return r.permute(0, 2, 3, 1)


if len(sys.argv) > 1 and sys.argv[1] == "time":
    n = 1024
    for name, (cin, cexp, cout, stride, hw, res) in CASES.items():
        x, ws = make(cin, cexp, cout, stride, hw, res, n)
        pr = pack_mbconv_rows(*ws, stride, mbconv_rows_spr(hw, stride), device=dev)
        pm = pack_mbconv(*ws, stride, device=dev)
        for label, fn in (("rows", lambda: eng.mbconv_rows(x, pr, residual=x if res else None)),
                          ("tile", lambda: eng.mbconv(x, pm, residual=x if res else None))):
            for _ in range(3):
                eng.release(fn())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.release(fn())
            e1.record()
            torch.cuda.synchronize()
            print(f"{name} {label}: {e0.elapsed_time(e1) * 100:.1f} us", flush=True)
else:
    names = sys.argv[1:] or list(CASES)
    for name in names:
        cin, cexp, cout, stride, hw, res = CASES[name]
        for n in (1, 3, 200):
            x, ws = make(cin, cexp, cout, stride, hw, res, n)
            pr = pack_mbconv_rows(*ws, stride, mbconv_rows_spr(hw, stride), device=dev)
            out = eng.mbconv_rows(x, pr, residual=x if res else None)
            torch.cuda.synchronize()
            ref = ref_fn(x, ws, stride, res)
            err = (out.float() - ref).abs()
            bad = err > 2e-2 * max(1.0, ref.abs().max().item())
            print(f"{name} n={n}: max err {err.max().item():.4f} (ref max {ref.abs().max().item():.2f}), bad {bad.sum().item()} / {bad.numel()}", flush=True)
            if bad.any():
                idx = bad.nonzero()
                print("  first bad (n,y,x,c):", idx[:5].tolist(), " rows:", sorted(set(idx[:, 1].tolist()))[:20], " cols:", sorted(set(idx[:, 2].tolist()))[:20],
                      " frames:", sorted(set(idx[:, 0].tolist()))[:10], " ch:", sorted(set(idx[:, 3].tolist()))[:20])

