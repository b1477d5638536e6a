"""GPU bring-up script for the tcgen05 conv kernel: each case runs in its own subprocess with a timeout so a
hang in one case cannot take the others (or the box) down.  Usage: python tools/gpu_conv_debug.py [case_idx]"""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, n,h,w,cin,cout,k,stride,pad,act,residual,out_f32,block_n
    ("gemm_small", 1, 1, 128, 64, 64, 1, 1, 0, 0, False, True, 64),
    ("gemm_mask", 1, 1, 300, 192, 96, 1, 1, 0, 0, False, False, None),
    ("c1x1_relu", 2, 16, 16, 64, 128, 1, 1, 0, 1, False, False, None),
    ("c3x3_res", 2, 16, 16, 64, 64, 3, 1, 1, 1, True, False, None),
    ("c3x3_s2", 3, 16, 16, 128, 128, 3, 2, 1, 1, False, False, None),
    ("c1x1_s2", 2, 16, 16, 256, 512, 1, 2, 0, 0, False, False, None),
    ("odd_mbv2", 2, 9, 9, 24, 144, 3, 1, 1, 2, False, False, None),
    ("odd_s2", 2, 9, 9, 32, 48, 3, 2, 1, 0, False, False, None),
    ("tiny_sp", 20, 4, 4, 512, 512, 3, 1, 1, 1, True, False, None),
    ("actor49", 1, 1, 77, 1024, 49, 1, 1, 0, 0, False, True, None),
    ("persist", 64, 32, 32, 64, 256, 1, 1, 0, 1, False, False, None),
    ("persist3", 32, 16, 16, 128, 128, 3, 1, 1, 1, True, False, 64),
]


def run_case(i):
    import torch
    import torch.nn.functional as F
    from adafocus_b200.engine import get_engine, pack_conv
    name, n, h, w, cin, cout, k, stride, pad, act, res, out_f32, bn = CASES[i]
    torch.manual_seed(100 + i)
    dev = torch.device("cuda", 0)
    eng = get_engine(dev)
    x = torch.randn(n, h, w, cin, device=dev).half()
    wt = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).half().float()
    scale = torch.rand(cout, device=dev) + 0.5
    bias = torch.randn(cout, device=dev) * 0.1
    pc = pack_conv(wt, scale, bias, stride, pad, act, block_n=bn, device=dev)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    r = torch.randn(n, ho, wo, cout, device=dev).half() if res else None
    out = eng.conv(x, pc, residual=r, out_f32=out_f32)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, stride, pad)
    ref = ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    if act == 1:
        ref = ref.clamp(min=0)
    if act == 2:
        ref = ref.clamp(0, 6)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    err = (out.float() - ref).abs()
    tol = 2e-2 if not out_f32 else 2e-3
    bad = (err > tol + 1e-2 * ref.abs())
    print(f"[{name}] out {tuple(out.shape)} max_err {err.max().item():.4e} ref_absmax {ref.abs().max().item():.3f} "
          f"bad {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        for ix in idx:
            print("   bad at", ix, "got", out[tuple(ix)].item(), "ref", ref[tuple(ix)].item())
        # pattern summary: which channels / pixels are bad
        bc = bad.reshape(-1, cout).any(0).nonzero().flatten().tolist()
        bp = bad.reshape(-1, cout).any(1).nonzero().flatten().tolist()
        print("   bad channels:", bc[:40], "... n=", len(bc))
        print("   bad pixels:", bp[:40], "... n=", len(bp))
        return 1
    return 0


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(run_case(int(sys.argv[1])))
    import torch  # warm the import cache
    print(torch.__version__, torch.cuda.get_device_name(0), flush=True)
    fails = 0
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], timeout=120, capture_output=True, text=True)
            print(r.stdout.strip())
            if r.returncode != 0:
                fails += 1
                print(f"  case {c[0]} rc={r.returncode}\n  stderr tail: {r.stderr[-1500:]}")
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"[{c[0]}] TIMEOUT (hang)")
    print("FAILS:", fails)
