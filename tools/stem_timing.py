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


if __name__ == "__main__":
    main()
