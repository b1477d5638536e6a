"""Single launches of the HBM-bound helper kernels at cfg3 sizes (for ncu)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_stem

def timed(fn, byts, name, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    print(f"{name}: {us:.1f} us  {byts / us / 1e6:.2f} TB/s", flush=True)

def main():
    which = sys.argv[1:] or ["dw32", "dw144", "dw96s2", "stem", "im2col", "maxpool"]
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    n = 1024
    for w in which:
        if w.startswith("dw"):
            c, hw, s = {"dw32": (32, 112, 1), "dw144": (144, 56, 1), "dw96s2": (96, 112, 2), "dw384": (384, 14, 1),
                        "dw576": (576, 14, 1), "dw576s2": (576, 14, 2), "dw960": (960, 7, 1)}[w]
            x = torch.randn(n, hw, hw, c, device=dev).half()
            w9 = torch.randn(9, c, device=dev)
            sc, bi = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            out_elems = n * (hw // s) ** 2 * c
            timed(lambda: eng.dwconv3x3(x, w9, sc, bi, s), (x.numel() + out_elems) * 2, w)
        elif w == "stem":
            fr = torch.randn(n, 3, 224, 224, device=dev)
            w27 = torch.randn(27, 32, device=dev)
            sc, bi = torch.ones(32, device=dev), torch.zeros(32, device=dev)
            timed(lambda: eng.stem_conv3x3s2_c32(fr, w27, sc, bi), fr.numel() * 4 + n * 112 * 112 * 32 * 2, w)
        elif w == "im2col":
            fr = torch.randn(n, 3, 224, 224, device=dev)
            yx = torch.randint(0, 97, (n, 2), device=dev, dtype=torch.int32)
            out = torch.empty(n * 64 * 64, 192, device=dev, dtype=torch.float16)
            def f():
                from adafocus_b200._lib import check
                from ctypes import c_void_p
                check(eng.lib.af_stem_im2col(eng.h, c_void_p(fr.data_ptr()), c_void_p(yx.data_ptr()), 1,
                                             c_void_p(out.data_ptr()), n, 224, 224, 128, 7, 7, 2, 3, 192,
                                             eng._stream()), "im2col")
            timed(f, n * 3 * 128 * 128 * 4 + out.numel() * 2, w)
        elif w == "maxpool":
            x = torch.randn(n, 64, 64, 64, device=dev).half()
            timed(lambda: eng.maxpool3x3s2(x), (x.numel() + x.numel() // 4) * 2, w)

if __name__ == "__main__":
    main()
