"""CUDA-event timings of the stem paths at cfg3 sizes (1024 frames / patches)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_stem


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / iters


def main():
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    n = int(os.environ.get("N", 1024))
    frames = torch.randn(n, 3, 224, 224, device=dev)
    yx = torch.randint(0, 97, (n, 2), dtype=torch.int32, device=dev)
    wt = torch.randn(64, 3, 7, 7, device=dev) / math.sqrt(147)
    pc = pack_stem(wt, torch.ones(64, device=dev), torch.zeros(64, device=dev), stride=2, pad=3, act=1, device=dev)
    for mode in ("s2d", "fused"):
        eng.s2d_stem = mode == "s2d"
        print(f"resnet stem {mode}: {timeit(lambda: eng.stem(frames, pc, yx=yx, patch=128)):.1f} us", flush=True)
    wt = torch.randn(32, 3, 3, 3, device=dev) / math.sqrt(27)
    sc, bi = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    pm = pack_stem(wt, sc, bi, stride=2, pad=1, act=2, device=dev)
    w27 = wt.permute(2, 3, 1, 0).reshape(27, 32).contiguous()
    eng.s2d_stem = True
    print(f"mobilenet stem s2d: {timeit(lambda: eng.stem(frames, pm)):.1f} us", flush=True)
    print(f"mobilenet stem direct: {timeit(lambda: eng.stem_conv3x3s2_c32(frames, w27, sc, bi)):.1f} us", flush=True)


if __name__ == "__main__" and "--pieces" not in sys.argv:
    main()


def pieces():
    """prepass and windowed conv of the two s2d stems timed separately"""
    from adafocus_b200.engine import _ptr, check
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    n = int(os.environ.get("N", 1024))
    frames = torch.randn(n, 3, 224, 224, device=dev)
    yx = torch.randint(0, 97, (n, 2), dtype=torch.int32, device=dev)
    for name, k, pad, cout, patch in (("resnet", 7, 3, 64, 128), ("mobilenet", 3, 1, 32, 224)):
        wt = torch.randn(cout, 3, k, k, device=dev) / math.sqrt(3 * k * k)
        pc = pack_stem(wt, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), stride=2, pad=pad, act=1, device=dev)
        q = pc.s2d
        pe = 16 * q.vt
        ho = wo = patch // 2
        hs, ws = ho + q.kh - 1, wo + 64 // pe - 1
        buf = torch.zeros(n * hs * ws + ws, pe, device=dev, dtype=torch.float16)
        out = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.float16)
        yxp = yx if patch != 224 else None
        t1 = timeit(lambda: check(eng.lib.af_stem_s2d(eng.h, _ptr(frames), _ptr(yxp), 1, _ptr(buf), n, 224, 224, patch, pad,
                                                      hs, ws, q.vt, eng._stream()), "s2d"))
        t2 = timeit(lambda: eng.conv(buf, q, out=out, out_stride=cout, shape=(n, hs, wo, 64, pe), row_stride=ws * pe,
                                     img_stride=hs * ws * pe))
        print(f"{name}: prepass {t1:.1f} us ({(frames.numel() * 4 * (patch / 224) ** 2 + buf.numel() * 2) / t1 / 1e6:.2f} TB/s), "
              f"conv {t2:.1f} us (out {out.numel() * 2 / t2 / 1e6:.2f} TB/s)", flush=True)


if __name__ == "__main__" and "--pieces" in sys.argv:
    pieces()
