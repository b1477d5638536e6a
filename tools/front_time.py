import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200.engine import get_engine, pack_mbconv_rows, pack_stem, stem_s2d_weights, mbconv_rows_spr, pack_mbconv
dev = torch.device("cuda", 0); eng = get_engine(dev)
n = 1024
frames = torch.randn(n, 3, 224, 224, device=dev)
w0 = torch.randn(32, 3, 3, 3) / math.sqrt(27); wd = torch.randn(32, 1, 3, 3) / 3; wp = torch.randn(16, 32) / math.sqrt(32)
one32, z32, one16, z16 = torch.ones(32), torch.zeros(32), torch.ones(16), torch.zeros(16)
stem = pack_stem(w0, one32, z32, stride=2, pad=1, act=2, device=dev)
w64, vt = stem_s2d_weights(w0)
front = pack_mbconv_rows(w64.reshape(32, 64), one32, z32, wd, one32, z32, wp, one16, z16, 1, 4, device=dev)
def t(fn, label):
    for _ in range(3): eng.release(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.release(fn())
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) * 100:.1f} us", flush=True)
t(lambda: eng.stem_front(frames, stem, front), "stem_front (s2d prepass + rows kernel)")
t(lambda: eng.stem(frames, stem), "stem (s2d prepass + conv)")
x16 = torch.randn(n, 112, 112, 16, device=dev).half()
x32 = torch.randn(n, 112, 112, 32, device=dev).half()
for cin, x in ((16, x16), (32, x32)):
    ws = (torch.randn(96, cin) / math.sqrt(cin), torch.ones(96), torch.zeros(96), torch.randn(96, 1, 3, 3) / 3, torch.ones(96), torch.zeros(96), torch.randn(24, 96) / 10, torch.ones(24), torch.zeros(24))
    pr = pack_mbconv_rows(*ws, 2, 4, device=dev)
    t(lambda: eng.mbconv_rows(x, pr), f"block 2 rows cin={cin}")
