"""Single launches of the fused inverted-residual block kernel at cfg3 sizes (for ncu / quick timing)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_mbconv

CASES = {
    # name: (n, hw, cin, cexp, cout, stride, res)
    "b2": (1024, 112, 32, 96, 24, 2, False),
    "b3": (1024, 56, 24, 144, 24, 1, True),
    "b4": (1024, 56, 24, 144, 32, 2, False),
    "b5": (1024, 28, 32, 192, 32, 1, True),
    "b7": (1024, 28, 32, 192, 64, 2, False),
    "b8": (1024, 14, 64, 384, 64, 1, True),
}

def main():
    names = sys.argv[1:] or list(CASES)
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    for name in names:
        n, hw, cin, cexp, cout, s, res = CASES[name]
        x = torch.randn(n, hw, hw, cin, device=dev).half()
        g = lambda *sh: torch.randn(*sh, device=dev)
        pm = pack_mbconv(g(cexp, cin) / math.sqrt(cin), torch.ones(cexp), g(cexp) * 0.1, g(cexp, 1, 3, 3) / 3,
                         torch.ones(cexp), g(cexp) * 0.1, g(cout, cexp) / math.sqrt(cexp), torch.ones(cout),
                         g(cout) * 0.1, s, device=dev)
        for _ in range(2):
            eng.mbconv(x, pm, residual=x if res else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.mbconv(x, pm, residual=x if res else None)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) * 200:.1f} us", flush=True)

if __name__ == "__main__":
    main()
