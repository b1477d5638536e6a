"""ResNet stem at cfg3 size (1024 patches of 128^2): prepass + conv + separate max-pool vs conv with the pool fused."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_stem
from stem_timing import timeit

dev = torch.device("cuda", 0)
eng = get_engine(dev)
n = int(os.environ.get("N", 1024))
frames = torch.randn(n, 3, 224, 224, device=dev)
yx = torch.randint(0, 97, (n, 2), dtype=torch.int32, device=dev)
wt = torch.randn(64, 3, 7, 7, device=dev) / math.sqrt(147)
pc = pack_stem(wt, torch.ones(64, device=dev), torch.zeros(64, device=dev), stride=2, pad=3, act=1, device=dev)
x = eng.stem(frames, pc, yx=yx, patch=128)
print(f"prepass + conv          : {timeit(lambda: eng.stem(frames, pc, yx=yx, patch=128)):.1f} us")
print(f"maxpool                 : {timeit(lambda: eng.maxpool3x3s2(x)):.1f} us")
print(f"prepass + conv(+pool)   : {timeit(lambda: eng.stem(frames, pc, yx=yx, patch=128, pool=True)):.1f} us")
