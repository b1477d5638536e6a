"""One MobileNet-V2 block through af_mbconv_rows at 1024 frames, three launches, for
`ncu --set full -k regex:mbconv_rows --launch-skip 2 -c 1`:  python tools/ncu_mbrows_case.py b2|b3|b4|b5|b7|b8"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, mbconv_rows_spr, pack_mbconv_rows

CASES = {"b2": (32, 96, 24, 2, 112, False), "b3": (24, 144, 24, 1, 56, True), "b4": (24, 144, 32, 2, 56, False),
         "b5": (32, 192, 32, 1, 28, True), "b7": (32, 192, 64, 2, 28, False), "b8": (64, 384, 64, 1, 14, True)}
dev = torch.device("cuda", 0)
eng = get_engine(dev)
cin, cexp, cout, stride, hw, res = CASES[sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
x = torch.randn(n, hw, hw, cin, device=dev).half()
ws = (torch.randn(cexp, cin) / math.sqrt(cin), torch.ones(cexp), torch.zeros(cexp), torch.randn(cexp, 1, 3, 3) / 3,
      torch.ones(cexp), torch.zeros(cexp), torch.randn(cout, cexp) / math.sqrt(cexp), torch.ones(cout), torch.zeros(cout))
pr = pack_mbconv_rows(*ws, stride, mbconv_rows_spr(hw, stride), device=dev)
for _ in range(3):
    eng.release(eng.mbconv_rows(x, pr, residual=x if res else None))
torch.cuda.synchronize()
