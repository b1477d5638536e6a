"""Runs single conv layers of the cfg3 workload a few times (for `ncu --set full -k regex:conv_gemm`)."""
import sys, os, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_conv

CASES = {
    # name: (n, h, w, cin, cout, k, stride, pad, act, residual)
    "expand112": (256, 112, 112, 32, 96, 1, 1, 0, 2, False),     # fG merged project->expand, store-heavy
    "conv3_l1": (1024, 32, 32, 64, 256, 1, 1, 0, 1, True),       # fL layer1 conv3 + residual
    "conv2_l3": (1024, 8, 8, 256, 256, 3, 1, 1, 1, False),       # fL layer3 3x3 (MMA-bound)
    "conv1_l1": (1024, 32, 32, 256, 64, 1, 1, 0, 1, False),      # fL layer1 conv1 (read-heavy)
    "conv2_l1": (1024, 32, 32, 64, 64, 3, 1, 1, 1, False),       # fL layer1 3x3 (L2 operand traffic)
    "conv3_l2": (1024, 16, 16, 128, 512, 1, 1, 0, 1, True),      # fL layer2 conv3 + residual
    "conv3_l3": (1024, 8, 8, 256, 1024, 1, 1, 0, 1, True),       # fL layer3 conv3 + residual
    "expand56": (1024, 56, 56, 24, 144, 1, 1, 0, 2, False),      # fG block 3 expand
    "conv1_l2": (1024, 16, 16, 512, 128, 1, 1, 0, 1, False),     # fL layer2 conv1
    "conv2_l2": (1024, 16, 16, 128, 128, 3, 1, 1, 1, False),     # fL layer2 3x3
    "conv1_l3": (1024, 8, 8, 1024, 256, 1, 1, 0, 1, False),      # fL layer3 conv1
    "conv1_l4": (1024, 4, 4, 2048, 512, 1, 1, 0, 1, False),      # fL layer4 conv1
    "conv2_l4": (1024, 4, 4, 512, 512, 3, 1, 1, 1, False),       # fL layer4 3x3
    "conv3_l4": (1024, 4, 4, 512, 2048, 1, 1, 0, 1, True),       # fL layer4 conv3 + residual
    "ds_l3": (1024, 16, 16, 512, 1024, 1, 2, 0, 0, False),       # fL layer3 downsample (stride 2)
    "conv2s2_l3": (1024, 16, 16, 256, 256, 3, 2, 1, 1, False),   # fL layer3.0 3x3 stride 2
    "last_fg": (1024, 7, 7, 320, 1280, 1, 1, 0, 2, False),       # fG last 1x1
    "exp_fg14": (1024, 14, 14, 96, 576, 1, 1, 0, 2, False),      # fG 14x14 expand
    "proj_fg14": (1024, 14, 14, 576, 96, 1, 1, 0, 0, True),      # fG 14x14 project + residual
}

def main():
    names = sys.argv[1:] or list(CASES)
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    for name in names:
        n, h, w, cin, cout, k, stride, pad, act, res = CASES[name]
        x = torch.randn(n, h, w, cin, device=dev).half()
        wt = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)
        pc = pack_conv(wt, torch.ones(cout), torch.zeros(cout), stride, pad, act, device=dev, fold_scale=res)
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        r = torch.randn(n, ho, wo, cout, device=dev).half() if res else None
        out = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.float16)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            eng.conv(x, pc, out=out, residual=r)
        ev0.record()
        for _ in range(10):
            eng.conv(x, pc, out=out, residual=r)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 100
        byts = (x.numel() + out.numel() * (2 if res else 1)) * 2
        fl = 2.0 * n * ho * wo * cout * cin * k * k
        print(f"{name}: {us:.1f} us  {byts / us / 1e6:.2f} TB/s  {fl / us / 1e6:.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    main()
