"""Round-2 kernels at cfg3 sizes, three launches each, for `ncu --set full -k regex:<kernel> --launch-skip 2 -c 1`:
  stem_pool    ResNet stem conv with the max-pool in the epilogue (conv_gemm_kernel<1,0,0,1>)
  shortcut_l1  layer1.0 conv3 + projection shortcut as a second GEMM (conv_gemm_kernel<0,0,0,0>)
  shortcut_l3  layer3.0 conv3 + stride-2 shortcut (CTA pair)
  gru_policy   persistent tensor-core GRU, policy (gru_tc_kernel<false>)
  gru_head     persistent tensor-core GRU, split-precision classifier (gru_tc_kernel<true>)"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import AF_ACT_NONE, AF_ACT_RELU, get_engine, pack_conv, pack_conv_split, pack_stem

dev = torch.device("cuda", 0)
eng = get_engine(dev)
case = sys.argv[1]
n = 1024
if case == "stem_pool":
    frames = torch.randn(n, 3, 224, 224, device=dev)
    yx = torch.randint(0, 97, (n, 2), dtype=torch.int32, device=dev)
    pc = pack_stem(torch.randn(64, 3, 7, 7, device=dev) / math.sqrt(147), torch.ones(64), torch.zeros(64), stride=2, pad=3,
                   act=AF_ACT_RELU, device=dev)
    fn = lambda: eng.stem(frames, pc, yx=yx, patch=128, pool=True)
elif case in ("shortcut_l1", "shortcut_l3"):
    hw, cmid, cin2, cout, s2 = (32, 64, 64, 256, 1) if case == "shortcut_l1" else (8, 256, 512, 1024, 2)
    h = torch.randn(n, hw, hw, cmid, device=dev).half()
    x = torch.randn(n, hw * s2, hw * s2, cin2, device=dev).half()
    c3 = pack_conv(torch.randn(cout, cmid, device=dev) / math.sqrt(cmid), torch.ones(cout), torch.zeros(cout),
                   act=AF_ACT_RELU, device=dev, fold_scale=True)
    ds = pack_conv(torch.randn(cout, cin2, device=dev) / math.sqrt(cin2), torch.ones(cout), None, stride=s2,
                   act=AF_ACT_NONE, device=dev, fold_scale=True, block_n=c3.block_n)
    out = torch.empty(n, hw, hw, cout, device=dev, dtype=torch.float16)
    fn = lambda: eng.conv(h, c3, out=out, shortcut=(x, ds))
else:
    b, t, hd = 64, 16, 1024
    split = case == "gru_head"
    w = torch.randn(3 * hd, hd) / math.sqrt(hd)
    pc = (pack_conv_split(w, torch.zeros(3 * hd), device=dev, block_n=32) if split
          else pack_conv(w, None, torch.zeros(3 * hd), device=dev, block_n=32))
    xg = torch.randn(b * t, 3 * hd, device=dev)
    hseq = torch.zeros(b * t, (3 if split else 1) * hd, device=dev, dtype=torch.float16)
    fn = lambda: eng.gru_sequence_tc(xg, pc, b, t, hseq)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{case}: {e0.elapsed_time(e1) * 200:.1f} us per call")
