"""Torch-facing wrappers over the C ABI: weight packing (one-off, torch ops) and kernel launches on NHWC fp16
tensors.  torch is used for device memory and streams only; every compute launch goes through
libadafocus_b200.so.  Nothing here falls back to torch ops for the hot path."""
import ctypes
import math
import os
from ctypes import byref, c_void_p

import torch

from . import _lib
from ._lib import AF_ACT_NONE, AF_ACT_RELU, AF_ACT_RELU6, ConvDesc, MbconvDesc, MbconvRowsDesc, check

BLOCK_K = 64


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def round_up(a, b):
    return (a + b - 1) // b * b


def default_block_n(cout):
    """Split Cout into the fewest <=256-wide tiles; a single tile is a multiple of 16, several tiles are multiples
    of 64 so that the 64-channel TMA-store slices of the epilogue never straddle two tiles."""
    nb = (cout + 255) // 256
    if nb == 1:
        return round_up(cout, 16)
    return min(256, round_up((cout + nb - 1) // nb, 64))


class PackedConv:
    """A convolution / linear layer in kernel layout.

    w     fp16 [cout_pad, kh*kw*cblk*64]   K index = (r*kw+s)*cblk*64 + ci, zero padded
    scale fp32 [cout_pad], bias fp32 [cout_pad]   (folded BatchNorm or linear bias)
    """

    def __init__(self, w, scale, bias, cin, cout, kh, kw, stride, pad, block_n, act):
        self.w, self.scale, self.bias = w, scale, bias
        self.cin, self.cout, self.kh, self.kw = cin, cout, kh, kw
        self.stride, self.pad, self.block_n, self.act = stride, pad, block_n, act


def host(t):
    """fp32 HOST copy of a parameter / buffer (one D2H memcpy).  All weight packing below is host arithmetic followed
    by one H2D copy per packed tensor: repacking a model issues no device kernels at all (round 1 spent ~1000 small
    torch launches per model here, which also hid the product kernels from launch-capped profilers)."""
    return None if t is None else t.detach().to(dtype=torch.float32, device="cpu")


def fold_bn(bn_weight, bn_bias, running_mean, running_var, eps):
    """BatchNorm2d in eval mode as y = x*scale + bias (fp32, host tensors)."""
    scale = host(bn_weight) / torch.sqrt(host(running_var) + eps)
    bias = host(bn_bias) - host(running_mean) * scale
    return scale, bias


def pack_conv(weight, scale=None, bias=None, stride=1, pad=0, act=AF_ACT_NONE, block_n=None, device=None,
              cin_perm=None, fold_scale=False):
    """weight: (cout, cin, kh, kw) or (cout, cin) fp32 in torch layout -> PackedConv on `device`.

    fold_scale: multiply the per-channel scale into the fp32 weights before the fp16 rounding and pass no scale to
    the kernel.  Layers that add a residual use it: the kernel then adds the residual on the tensor core.

    cin_perm: optional LongTensor; packed input channel j reads torch input channel cin_perm[j] (used to absorb the
    NCHW-flatten vs NHWC-flatten difference of ACT/models/ppo.py:36-37)."""
    device = device or weight.device
    w = host(weight)
    if w.dim() == 2:
        w = w[:, :, None, None]
    cout, cin, kh, kw = w.shape
    if cin_perm is not None:
        w = w[:, cin_perm.cpu()]
    if fold_scale and scale is not None:
        w = w * host(scale)[:, None, None, None]
        scale = None
    block_n = block_n or default_block_n(cout)
    cout_pad = round_up(cout, block_n)
    cblk = (cin + BLOCK_K - 1) // BLOCK_K
    wp = torch.zeros(cout_pad, kh * kw, cblk * BLOCK_K, dtype=torch.float16)
    wp[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).half()
    wp = wp.reshape(cout_pad, kh * kw * cblk * BLOCK_K).contiguous()
    sc = None
    bi = torch.zeros(cout_pad, dtype=torch.float32)
    if scale is not None or not fold_scale:
        sc = torch.ones(cout_pad, dtype=torch.float32)
    if scale is not None:
        sc[:cout] = host(scale)
    if bias is not None:
        bi[:cout] = host(bias)
    return PackedConv(wp.to(device), sc.to(device) if sc is not None else None, bi.to(device), cin, cout, kh, kw,
                      stride, pad, block_n, act)


def pack_conv_split(weight, bias=None, device=None, block_n=None):
    """Linear layer (cout, cin) for split-precision operands: K-concatenation [W_hi | W_hi | W_lo] with
    W_hi = fp16(W), W_lo = fp16(W - W_hi), to be multiplied with activation rows [x_hi | x_lo | x_hi]
    (af_split3_f16 / af_gru_gates split=1).  One ordinary tcgen05 GEMM with K = 3*cin then accumulates
    x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32, i.e. ~22-bit operands.  cin must be a multiple of 64."""
    w = host(weight)
    assert w.dim() == 2 and w.shape[1] % BLOCK_K == 0, w.shape
    hi = w.half().float()
    lo = (w - hi).half().float()
    pc = pack_conv(torch.cat([hi, hi, lo], 1), None, bias, device=device or weight.device, block_n=block_n)
    pc.split = True
    return pc


def pack_stem(weight, scale, bias, stride, pad, act, device=None, kpad=None):
    """3-input-channel stem conv as a GEMM over the im2col matrix produced by af_stem_im2col:
    k = (r*kw+s)*3 + c."""
    device = device or weight.device
    w = host(weight)
    cout, cin, kh, kw = w.shape
    assert cin == 3
    kreal = kh * kw * 3
    # the im2col rows only need 16-byte granularity: the GEMM's TMA box zero-fills k beyond the row (147 -> 152, not 192)
    kpad = kpad or round_up(kreal, 8)
    flat = torch.zeros(cout, kpad, dtype=torch.float32)
    flat[:, :kreal] = w.permute(0, 2, 3, 1).reshape(cout, kreal)
    pc = pack_conv(flat, scale, bias, 1, 0, act, device=device)
    pc.stem = dict(kh=kh, kw=kw, stride=stride, pad=pad, kpad=kpad)
    pc.s2d = None
    if stride == 2 and kh == kw and kh % 2 == 1 and pad == kh // 2 and (kh + 1) // 2 <= 4:
        pc.s2d = pack_stem_s2d(w, scale, bias, act, device)
    return pc


def pack_stem_s2d(w, scale, bias, act, device):
    """Stride-2 k x k conv over 3 channels as a stride-1 R x 1 conv (R = (k+1)//2) over the sliding-window view of the
    2x2 space-to-depth input written by af_stem_s2d: view channel sx*16 + (dy*2+dx)*3 + c at view pixel (Y, X) is
    padded[c][2Y+dy][2(X+sx)+dx], so tap (r, s) of the original filter lands on vertical tap r//2 and view channel
    (s//2)*16 + ((r%2)*2 + s%2)*3 + c."""
    w64, vt = stem_s2d_weights(w)
    pc = pack_conv(w64, scale, bias, 1, 0, act, device=device)
    pc.vt = vt
    return pc


def stem_s2d_weights(w):
    """-> (w64 fp32 (cout, 64, R // vt, 1), vt): the stem filter re-indexed for the window view (pack_stem_s2d)."""
    w = host(w)
    cout, _, k, _ = w.shape
    R = (k + 1) // 2
    vt = 2 if R == 2 else 1   # 3x3: fold the two vertical taps into the pixel too -> one k-block, a 1x1 conv
    w64 = torch.zeros(cout, 64, R // vt, 1, dtype=torch.float32)
    for r in range(k):
        for s_ in range(k):
            ch = (s_ // 2) * 16 * vt + ((r // 2) % vt) * 16 + ((r % 2) * 2 + (s_ % 2)) * 3
            w64[:, ch:ch + 3, (r // 2) // vt, 0] = w[:, :, r, s_]
    return w64, vt


class PackedMbconv:
    """An inverted-residual block (expand 1x1 -> depthwise 3x3 -> project 1x1, BatchNorms folded) in the layout of
    af_mbconv_fused: every BN scale is folded into the weights, the kernel only adds biases."""

    def __init__(self, w1, b1, dw, b2, w2, b3, cin, cexp, cout, stride, bias1_in_w1=False):
        self.w1, self.b1, self.dw, self.b2, self.w2, self.b3 = w1, b1, dw, b2, w2, b3
        self.cin, self.cexp, self.cout, self.stride = cin, cexp, cout, stride
        self.bias1_in_w1 = bias1_in_w1


def mbconv_supported(n, h, w, cin, cexp, cout, stride):
    return bool(_lib.load().af_mbconv_fused_supported(int(n), int(h), int(w), int(cin), int(cexp), int(cout),
                                                      int(stride)))


def pack_mbconv(w_exp, s1, b1, w_dw, s2, b2, w_proj, s3, b3, stride, device=None):
    """w_exp (cexp, cin[,1,1]), w_dw (cexp, 1, 3, 3), w_proj (cout, cexp[,1,1]) fp32 in torch layout; (s, b) = folded
    BatchNorm of each conv.  Returns a PackedMbconv on `device`."""
    device = device or w_exp.device
    w_exp = host(w_exp).flatten(1)
    w_proj = host(w_proj).flatten(1)
    cexp, cin = w_exp.shape
    cout = w_proj.shape[0]
    ce = round_up(cexp, 64)
    e = pack_conv(w_exp, s1, b1, act=AF_ACT_RELU6, block_n=64, device="cpu", fold_scale=True)
    pj = pack_conv(w_proj, s3, b3, act=AF_ACT_NONE, block_n=round_up(cout, 16), device="cpu", fold_scale=True)
    assert e.w.shape == (ce, 64) and pj.w.shape == (round_up(cout, 16), ce), (e.w.shape, pj.w.shape)
    dw = torch.zeros(9, ce, dtype=torch.float32)
    dw[:, :cexp] = (host(w_dw).reshape(cexp, 9) * host(s2)[:, None]).t()
    bias2 = torch.zeros(ce, dtype=torch.float32)
    bias2[:cexp] = host(b2)
    # expand bias as two extra K columns (fp16 hi + lo parts) multiplied by a constant-1 channel pair the kernel adds
    # to the input tile: the bias add and the zeroing outside the image leave the epilogue (include/adafocus_b200.h)
    bias_in_w1 = cin + 2 <= 64 and os.environ.get("AF_MB_NO_BIAS_MMA") is None
    if bias_in_w1:
        hi = e.bias.half()
        e.w[:, cin] = hi
        e.w[:, cin + 1] = (e.bias - hi.float()).half()
    d = device
    return PackedMbconv(e.w.to(d), e.bias.to(d), dw.contiguous().to(d), bias2.to(d), pj.w.to(d), pj.bias.to(d), cin,
                        cexp, cout, stride, bias_in_w1)


class PackedMbconvRows:
    """The same block in the layout of af_mbconv_rows (row-streaming kernel) for ONE strips-per-row value `spr`."""

    def __init__(self, w1, dwp, w2, b3, cin, cexp, cout, stride, spr):
        self.w1, self.dwp, self.w2, self.b3 = w1, dwp, w2, b3
        self.cin, self.cexp, self.cout, self.stride, self.spr = cin, cexp, cout, stride, spr


def mbconv_rows_spr(w, stride):
    """Strips per row segment of af_mbconv_rows for an input of width w (None: width not handled)."""
    if stride == 1:
        return {14: 1, 28: 2, 56: 4, 112: 4}.get(int(w))
    return {28: 2, 56: 4, 112: 4}.get(int(w))


def mbconv_rows_supported(n, h, w, cin, cexp, cout, stride):
    return bool(_lib.load().af_mbconv_rows_supported(int(n), int(h), int(w), int(cin), int(cexp), int(cout),
                                                     int(stride)))


def mbconv_rows_layout(cexp, spr):
    """-> (nchunks, lane_ch int16 [3,128], lane_kpos int16 [3,128]) from af_mbconv_rows_layout, or None."""
    lib = _lib.load()
    nch = ctypes.c_int32(0)
    lane_ch = torch.zeros(3, 128, dtype=torch.int16)
    lane_kpos = torch.zeros(3, 128, dtype=torch.int16)
    rc = lib.af_mbconv_rows_layout(int(cexp), int(spr), byref(nch), lane_ch.data_ptr(), lane_kpos.data_ptr())
    if rc != 0:
        return None
    return nch.value, lane_ch, lane_kpos


def pack_mbconv_rows(w_exp, s1, b1, w_dw, s2, b2, w_proj, s3, b3, stride, spr, device=None):
    """Same arguments as pack_mbconv plus `spr`; returns a PackedMbconvRows on `device`, or None when the kernel has no
    lane placement for (cexp, spr).  ReLU6(x) = 6 * sat(x / 6): 1/6 goes into the expand weights and both biases, 6 into
    the project weights (include/adafocus_b200.h, af_mbconv_rows_desc)."""
    device = device or w_exp.device
    w_exp = host(w_exp).flatten(1).float()
    w_proj = host(w_proj).flatten(1).float()
    cexp, cin = w_exp.shape
    cout = w_proj.shape[0]
    if cin > 64 or cout > 64:
        return None
    lay = mbconv_rows_layout(cexp, spr)
    if lay is None:
        return None
    nch, lane_ch, lane_kpos = lay
    s1, b1, s2, b2, s3, b3 = (host(t).float() for t in (s1, b1, s2, b2, s3, b3))
    we = w_exp * s1[:, None] / 6.0
    wd = host(w_dw).reshape(cexp, 9).float() * s2[:, None]
    wp = w_proj * s3[:, None] * 6.0
    cp = round_up(cout, 16)
    w1 = torch.zeros(nch * 128, 64, dtype=torch.float32)
    dwp = torch.zeros(nch, 11, 128, dtype=torch.float32)
    w2 = torch.zeros(cp, nch * 128, dtype=torch.float32)
    for c in range(nch):
        ch = lane_ch[c].long()
        live = ch >= 0
        idx = ch[live]
        w1[c * 128:(c + 1) * 128][live, :cin] = we[idx]
        dwp[c, :9][:, live] = wd[idx].t()
        dwp[c, 9][live] = b1[idx] / 6.0
        dwp[c, 10][live] = b2[idx] / 6.0
        w2[:cout, c * 128 + lane_kpos[c].long()[live]] = wp[:, idx]
    bias3 = torch.zeros(cp, dtype=torch.float32)
    bias3[:cout] = b3
    d = device
    return PackedMbconvRows(w1.half().to(d), dwp.contiguous().to(d), w2.half().to(d), bias3.to(d), cin, cexp, cout,
                            stride, spr)


class Workspace:
    """Size-bucketed reuse of device buffers while a plan is being laid out.  Replay is in stream order, so a
    buffer may be handed out again as soon as its last reader has been recorded."""

    def __init__(self, device):
        self.device = device
        self.free_lists = {}
        self.all = []
        self.total_bytes = 0

    def alloc(self, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = round_up(max(n, 1) * torch.empty((), dtype=dtype).element_size(), 512)
        lst = self.free_lists.get(nbytes)
        if lst:
            raw = lst.pop()
        else:
            raw = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.all.append(raw)
            self.total_bytes += nbytes
        t = raw[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        t._af_raw = raw
        return t

    def release(self, t):
        raw = getattr(t, "_af_raw", None)
        if raw is not None:
            self.free_lists.setdefault(raw.numel(), []).append(raw)
            t._af_raw = None


class Engine:
    """One per device: wraps af_ctx and exposes the kernels on torch tensors."""

    def __init__(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.AfError("adafocus_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = _lib.Context(self.index)
        self.lib = self.ctx.lib
        self.h = self.ctx.handle
        self.fused_stem = os.environ.get("AF_NO_FUSED_STEM") is None   # crop + stem conv as one implicit-GEMM kernel
        self.s2d_stem = os.environ.get("AF_NO_S2D_STEM") is None       # crop + space-to-depth, then a windowed R x 1 conv
        self.ws = None          # Workspace while recording
        self._keep = None
        self.launch_count = 0   # launches issued eagerly (not recorded)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, shape, dtype):
        if self.ws is not None:
            return self.ws.alloc(shape, dtype)
        return torch.empty(shape, dtype=dtype, device=self.device)

    def release(self, t):
        if self.ws is not None:
            self.ws.release(t)

    def begin_plan(self):
        check(self.lib.af_plan_begin(self.h), "af_plan_begin")
        self.ws = Workspace(self.device)
        self._keep = []

    def keep(self, *tensors):
        if self._keep is not None:
            self._keep.extend(tensors)

    def end_plan(self):
        p = c_void_p()
        check(self.lib.af_plan_end(self.h, byref(p)), "af_plan_end")
        plan = _lib.Plan(self.lib, p, (self.ws.all, self._keep))
        plan.workspace_bytes = self.ws.total_bytes
        self.ws, self._keep = None, None
        return plan

    def mark(self, name=None):
        """Insert a timing mark into the plan being recorded; returns its index (no-op -> None when eager)."""
        if self.ws is None:
            return None
        i = ctypes.c_int()
        check(self.lib.af_plan_mark(self.h, byref(i)), "af_plan_mark")
        return i.value

    def _count(self):
        if self.ws is None:
            self.launch_count += 1

    # ------------------------------------------------------------------ kernels
    def conv_tsm_ok(self, x, t, fold):
        """Can conv() fold TemporalShift(n_segment=t, fold channels) into this input's loads?  (af_conv_tsm_supported)"""
        if os.environ.get("AF_NO_TSM_FOLD") is not None or not x.is_contiguous():
            return False
        n, h, w, c = x.shape
        return bool(self.lib.af_conv_tsm_supported(int(n), int(h), int(w), int(c), int(x.stride(2)), int(fold), int(t)))

    def conv(self, x, pc, out=None, residual=None, act=None, out_f32=False, out_stride=None, shape=None,
             row_stride=0, img_stride=0, tsm=None, pool=False, shortcut=None):
        """x: NHWC fp16 (n,h,w,cin) [or any tensor when `shape`=(n,h,w,cin,in_stride) is given; row_stride /
        img_stride (elements) describe a sliding-window view, see af_conv_desc].  tsm=(T, fold): the convolution reads
        TemporalShift.shift(x) (STH/ops/temporal_shift.py:29-46) without materialising it (1x1 convs, conv_tsm_ok)."""
        if shape is None:
            n, h, w, cin = x.shape
            in_stride = x.stride(2)
        else:
            n, h, w, cin, in_stride = shape
        assert cin == pc.cin, (cin, pc.cin)
        ho = (h + 2 * pc.pad - pc.kh) // pc.stride + 1
        wo = (w + 2 * pc.pad - pc.kw) // pc.stride + 1
        if pool:
            assert out is not None and tuple(out.shape) == (n, ho // 2, wo // 2, pc.cout)
        if out is None:
            out = self.empty((n, ho, wo, pc.cout), torch.float32 if out_f32 else torch.float16)
            out_stride = pc.cout
        elif out_stride is None:
            out_stride = out.stride(-2)
        d = ConvDesc()
        d.in_, d.w, d.bias = x.data_ptr(), pc.w.data_ptr(), pc.bias.data_ptr()
        d.scale = pc.scale.data_ptr() if pc.scale is not None else None
        d.residual = residual.data_ptr() if residual is not None else None
        d.out = out.data_ptr()
        d.n, d.h, d.w_, d.cin, d.cout = n, h, w, cin, pc.cout
        d.kh, d.kw, d.stride, d.pad = pc.kh, pc.kw, pc.stride, pc.pad
        d.block_n = pc.block_n
        d.act = pc.act if act is None else act
        d.out_f32 = 1 if out_f32 else 0
        d.in_stride, d.out_stride = in_stride, out_stride
        d.res_stride = residual.stride(-2) if residual is not None else 0
        d.in_row_stride, d.in_img_stride = row_stride, img_stride
        d.tsm_t, d.tsm_fold = (int(tsm[0]), int(tsm[1])) if tsm is not None else (0, 0)
        d.pool = 1 if pool else 0
        if shortcut is not None:
            # (x2, pc2): out = act(conv(x) + conv1x1_stride(x2) + bias) -- the projection shortcut accumulated in the
            # same TMEM tile (af_conv_desc.in2); pc.bias must already hold the sum of both folded BN biases
            x2, pc2 = shortcut
            assert pc.scale is None and pc2.scale is None and residual is None and pc2.kh == 1 and pc2.cout == pc.cout
            assert pc2.block_n == pc.block_n and x2.shape[-1] == pc2.cin
            d.in2, d.w2 = x2.data_ptr(), pc2.w.data_ptr()
            d.cin2, d.stride2, d.h2, d.w2_ = pc2.cin, pc2.stride, x2.shape[1], x2.shape[2]
            d.in2_stride = x2.stride(2)
            self.keep(x2, pc2.w)
        check(self.lib.af_conv2d_nhwc_f16(self.h, byref(d), self._stream()), "af_conv2d_nhwc_f16")
        self._count()
        self.keep(x, pc.w, pc.scale, pc.bias, out, residual)
        return out

    def linear(self, x2d, pc, out=None, out_f32=False, act=None, out_stride=None):
        """x2d: fp16 [M, K] (row stride >= K); C = x W^T as the 1x1 'conv' n=1,h=1,w=M."""
        m, k = x2d.shape
        if out is None:
            out = self.empty((m, pc.cout), torch.float32 if out_f32 else torch.float16)
            out_stride = pc.cout
        self.conv(x2d, pc, out=out, act=act, out_f32=out_f32, out_stride=out_stride,
                  shape=(1, 1, m, k, x2d.stride(0)))
        return out

    def stem_pool_ok(self, pc, patch):
        """Can stem(..., pool=True) fuse MaxPool2d(3, 2, 1) into the stem convolution?  (64 x 64 conv output = 128^2
        patches, s2d form, cout <= 64, ReLU; af_conv_desc.pool)"""
        s = pc.stem
        ho = (patch + 2 * s["pad"] - s["kh"]) // s["stride"] + 1
        return (os.environ.get("AF_NO_STEM_POOL") is None and self.s2d_stem and getattr(pc, "s2d", None) is not None
                and patch % 2 == 0 and ho == 64 and pc.cout <= 64 and pc.s2d.kh >= 2 and pc.act != AF_ACT_NONE)

    def stem(self, frames, pc, yx=None, patch=None, yx_div=1, pool=False):
        """frames (N,3,H,W) fp32 NCHW -> NHWC fp16 stem output; crop at yx (N,2 int32) of size `patch` fused in.
        pool=True (stem_pool_ok): returns MaxPool2d(3, 2, 1) of the stem output, pooled inside the conv kernel."""
        n, c, h, w = frames.shape
        assert c == 3 and frames.dtype == torch.float32 and frames.is_contiguous()
        s = pc.stem
        p = patch if patch is not None else h
        ho = (p + 2 * s["pad"] - s["kh"]) // s["stride"] + 1
        wo = (p + 2 * s["pad"] - s["kw"]) // s["stride"] + 1
        if self.s2d_stem and getattr(pc, "s2d", None) is not None and p % 2 == 0:
            q = pc.s2d
            pe = 16 * q.vt                      # elements per s2d pixel
            hs, ws = ho + q.kh - 1, wo + 64 // pe - 1
            # + one pixel row of slack: the window of the last view pixel extends past its row
            buf = self.empty((n * hs * ws + ws, pe), torch.float16)
            check(self.lib.af_stem_s2d(self.h, _ptr(frames), _ptr(yx), int(yx_div), _ptr(buf), n, h, w, p, s["pad"],
                                       hs, ws, q.vt, self._stream()), "af_stem_s2d")
            self._count()
            self.keep(frames, yx, buf)
            if pool:
                out = self.empty((n, ho // 2, wo // 2, pc.cout), torch.float16)
            else:
                out = self.empty((n, ho, wo, pc.cout), torch.float16)
            self.conv(buf, q, out=out, out_stride=pc.cout, shape=(n, hs, wo, 64, pe), row_stride=ws * pe,
                      img_stride=hs * ws * pe, pool=pool)
            self.release(buf)
            return out
        assert not pool, "stem(pool=True) needs the space-to-depth form (stem_pool_ok)"
        fused_ok = (self.fused_stem and pc.cout % 16 == 0 and pc.cout <= 64 and s["kh"] * s["kw"] * 3 <= 256
                    and ho * wo >= 128 and s["stride"] <= 2 and s["kh"] <= 7 and s["kw"] <= 7)
        if fused_ok:
            out = self.empty((n, ho, wo, pc.cout), torch.float16)
            check(self.lib.af_stem_conv_fused(self.h, _ptr(frames), _ptr(yx), int(yx_div), _ptr(pc.w), _ptr(pc.scale),
                                              _ptr(pc.bias), _ptr(out), n, h, w, p, pc.cout, s["kh"], s["kw"],
                                              s["stride"], s["pad"], pc.act, self._stream()), "af_stem_conv_fused")
            self._count()
            self.keep(frames, yx, pc.w, pc.scale, pc.bias, out)
            return out
        col = self.empty((n * ho * wo, s["kpad"]), torch.float16)
        check(self.lib.af_stem_im2col(self.h, _ptr(frames), _ptr(yx), int(yx_div), _ptr(col), n, h, w, p, s["kh"],
                                      s["kw"], s["stride"], s["pad"], s["kpad"], self._stream()), "af_stem_im2col")
        self._count()
        self.keep(frames, yx, col)
        out = self.empty((n, ho, wo, pc.cout), torch.float16)
        self.conv(col, pc, out=out, out_stride=pc.cout, shape=(1, 1, n * ho * wo, s["kpad"], s["kpad"]))
        self.release(col)
        return out

    def stem_front_ok(self, pc, front, frames):
        """Can stem_front() run features[0..1] of MobileNet-V2 as one row-streaming launch on these frames?"""
        n, c, h, w = frames.shape
        q = getattr(pc, "s2d", None)
        return (front is not None and self.s2d_stem and q is not None and q.vt == 2 and q.kh == 1 and h == w and h % 2 == 0
                and mbconv_rows_spr(h // 2, 1) == front.spr
                and mbconv_rows_supported(n, h // 2, w // 2, 64, front.cexp, front.cout, 1))

    def stem_front(self, frames, pc, front):
        """frames (N,3,H,W) fp32 -> features[1] output (N,H/2,W/2,cout) NHWC fp16: space-to-depth prepass, then the 3x3/2
        stem conv (a 1x1 conv over the 64-channel window view, pack_stem_s2d) as the EXPAND GEMM of af_mbconv_rows,
        followed by block 1's depthwise 3x3 and 1x1 project -- the 32-channel stem output never reaches HBM."""
        n, c, h, w = frames.shape
        s, q = pc.stem, pc.s2d
        ho, wo = h // 2, w // 2
        pe = 16 * q.vt
        hs, ws = ho + q.kh - 1, wo + 64 // pe - 1
        buf = self.empty((n * hs * ws + ws, pe), torch.float16)
        check(self.lib.af_stem_s2d(self.h, _ptr(frames), None, 1, _ptr(buf), n, h, w, h, s["pad"], hs, ws, q.vt,
                                   self._stream()), "af_stem_s2d")
        self._count()
        self.keep(frames, buf)
        out = self.mbconv_rows(buf, front, view=(n, ho, wo, 64, pe, ws * pe, hs * ws * pe))
        self.release(buf)
        return out

    def stem_conv3x3s2_c32(self, frames, w27, scale, bias, act=AF_ACT_RELU6):
        """frames (N,3,H,W) fp32 NCHW -> (N,H/2,W/2,32) NHWC fp16 (MobileNet-V2 features[0])."""
        n, c, h, w = frames.shape
        assert c == 3 and frames.dtype == torch.float32 and frames.is_contiguous()
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        out = self.empty((n, ho, wo, 32), torch.float16)
        check(self.lib.af_stem_conv3x3s2_c32(self.h, _ptr(frames), _ptr(w27), _ptr(scale), _ptr(bias), _ptr(out), n, h,
                                             w, act, self._stream()), "af_stem_conv3x3s2_c32")
        self._count()
        self.keep(frames, w27, scale, bias, out)
        return out

    def mbconv(self, x, pm, residual=None):
        """Fused inverted-residual block: x NHWC fp16 (n,h,w,cin) contiguous -> (n,ho,wo,cout)."""
        n, h, w, cin = x.shape
        assert cin == pm.cin and x.is_contiguous()
        s = pm.stride
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        out = self.empty((n, ho, wo, pm.cout), torch.float16)
        d = MbconvDesc()
        d.in_, d.w1, d.bias1, d.dw_w, d.bias2 = x.data_ptr(), pm.w1.data_ptr(), pm.b1.data_ptr(), pm.dw.data_ptr(), pm.b2.data_ptr()
        d.w2, d.bias3, d.out = pm.w2.data_ptr(), pm.b3.data_ptr(), out.data_ptr()
        d.residual = residual.data_ptr() if residual is not None else None
        d.n, d.h, d.w_, d.cin, d.cexp, d.cout, d.stride = n, h, w, cin, pm.cexp, pm.cout, s
        d.res_stride = residual.stride(-2) if residual is not None else 0
        d.bias1_in_w1 = 1 if pm.bias1_in_w1 else 0
        check(self.lib.af_mbconv_fused(self.h, byref(d), self._stream()), "af_mbconv_fused")
        self._count()
        self.keep(x, pm.w1, pm.b1, pm.dw, pm.b2, pm.w2, pm.b3, residual, out)
        return out

    def mbconv_rows(self, x, pr, residual=None, view=None):
        """Row-streaming fused inverted-residual block: x NHWC fp16 (n,h,w,cin) contiguous -> (n,ho,wo,cout).
        view = (n, h, w, cin, pix_stride, row_stride, img_stride): x is a buffer read through that (possibly
        overlapping-pixel) view instead (af_mbconv_rows_desc.in_*_stride)."""
        if view is not None:
            n, h, w, cin = view[:4]
        else:
            n, h, w, cin = x.shape
            assert x.is_contiguous()
        assert cin == pr.cin and mbconv_rows_spr(w, pr.stride) == pr.spr
        s = pr.stride
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        out = self.empty((n, ho, wo, pr.cout), torch.float16)
        d = MbconvRowsDesc()
        d.in_, d.w1, d.dwp, d.w2, d.bias3 = x.data_ptr(), pr.w1.data_ptr(), pr.dwp.data_ptr(), pr.w2.data_ptr(), pr.b3.data_ptr()
        d.out = out.data_ptr()
        d.residual = residual.data_ptr() if residual is not None else None
        d.n, d.h, d.w_, d.cin, d.cexp, d.cout, d.stride = n, h, w, cin, pr.cexp, pr.cout, s
        d.res_stride = residual.stride(-2) if residual is not None else 0
        if view is not None:
            d.in_pix_stride, d.in_row_stride, d.in_img_stride = int(view[4]), int(view[5]), int(view[6])
        check(self.lib.af_mbconv_rows(self.h, byref(d), self._stream()), "af_mbconv_rows")
        self._count()
        self.keep(x, pr.w1, pr.dwp, pr.w2, pr.b3, residual, out)
        return out

    def dwconv3x3(self, x, w9c, scale, bias, stride, act=AF_ACT_RELU6):
        n, h, w, c = x.shape
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = self.empty((n, ho, wo, c), torch.float16)
        check(self.lib.af_dwconv3x3_nhwc_f16(self.h, _ptr(x), _ptr(w9c), _ptr(scale), _ptr(bias), _ptr(out), n, h, w,
                                             c, stride, act, self._stream()), "af_dwconv3x3_nhwc_f16")
        self._count()
        self.keep(x, w9c, scale, bias, out)
        return out

    def maxpool3x3s2(self, x):
        n, h, w, c = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        out = self.empty((n, ho, wo, c), torch.float16)
        check(self.lib.af_maxpool3x3s2_nhwc_f16(self.h, _ptr(x), _ptr(out), n, h, w, c, self._stream()),
              "af_maxpool3x3s2_nhwc_f16")
        self._count()
        self.keep(x, out)
        return out

    def avgpool(self, x, out_f32=None, out_f32_stride=0, out_f16=None, out_f16_stride=0):
        n, h, w, c = x.shape
        check(self.lib.af_avgpool_nhwc_f16(self.h, _ptr(x), _ptr(out_f32), out_f32_stride, _ptr(out_f16),
                                           out_f16_stride, n, h * w, c, self._stream()), "af_avgpool_nhwc_f16")
        self._count()
        self.keep(x, out_f32, out_f16)

    def nhwc_to_nchw_f32(self, x, out=None):
        n, h, w, c = x.shape
        if out is None:
            out = self.empty((n, c, h, w), torch.float32)
        check(self.lib.af_nhwc_f16_to_nchw_f32(self.h, _ptr(x), _ptr(out), n, h * w, c, self._stream()),
              "af_nhwc_f16_to_nchw_f32")
        self._count()
        self.keep(x, out)
        return out

    def nchw_to_nhwc_f16(self, x, cpad=None):
        n, c, h, w = x.shape
        cpad = cpad or round_up(c, 8)
        out = self.empty((n, h, w, cpad), torch.float16)
        check(self.lib.af_nchw_f32_to_nhwc_f16(self.h, _ptr(x), _ptr(out), n, c, h * w, cpad, self._stream()),
              "af_nchw_f32_to_nhwc_f16")
        self._count()
        self.keep(x, out)
        return out

    def crop(self, img, action=None, yx=None, patch=None, out=None, yx_out=None):
        n, c, h, w = img.shape
        assert img.dtype == torch.float32 and img.is_contiguous()
        if out is None:
            out = self.empty((n, c, patch, patch), torch.float32)
        if n == 0:
            return out
        check(self.lib.af_crop_nchw_f32(self.h, _ptr(img), _ptr(action), _ptr(yx), _ptr(out), _ptr(yx_out), n, c, h, w,
                                        patch, self._stream()), "af_crop_nchw_f32")
        self._count()
        self.keep(img, action, yx, out, yx_out)
        return out

    def action_to_yx(self, action, h, patch):
        n = action.shape[0]
        yx = self.empty((n, 2), torch.int32)
        check(self.lib.af_action_to_yx(self.h, _ptr(action), _ptr(yx), n, h, patch, self._stream()),
              "af_action_to_yx")
        self._count()
        self.keep(action, yx)
        return yx

    def gru_gates(self, xg, xg_stride, hg, h_prev, h_new, h_new_f16=None, hseq_f16=None, hseq_stride=0,
                  hseq_f32=None, hseq_f32_stride=0, split=False):
        """split=True: the fp16 outputs are split-precision rows [hi | lo | hi] (h_new_f16 (B,3H); hseq rows 3H wide)."""
        b, hd = h_prev.shape
        check(self.lib.af_gru_gates(self.h, _ptr(xg), xg_stride, _ptr(hg), _ptr(h_prev), _ptr(h_new), _ptr(h_new_f16),
                                    _ptr(hseq_f16), hseq_stride, _ptr(hseq_f32), hseq_f32_stride, b, hd,
                                    1 if split else 0, self._stream()), "af_gru_gates")
        self._count()
        self.keep(xg, hg, h_prev, h_new, h_new_f16, hseq_f16, hseq_f32)

    GRU_SEQ_MAX_BATCH = 8    # measured crossover: from 16 clips on the per-step tcgen05 GEMM path is as fast or faster

    def can_gru_sequence(self, b, hd):
        return b <= self.GRU_SEQ_MAX_BATCH and hd % 256 == 0 and hd <= 1024 and hd // 8 <= self.ctx.sm_count

    def gru_sequence(self, xg, pc_hh, b, t, hseq16, h0=None, h_out=None):
        """All T steps of a GRU in one persistent launch. xg (B*T,3H) fp32, pc_hh = PackedConv of W_hh (+ b_hh);
        a split-precision packing (pack_conv_split) makes hseq16 rows [hi | lo | hi]."""
        split = bool(getattr(pc_hh, "split", False))
        hd = pc_hh.cin // 3 if split else pc_hh.cin
        assert pc_hh.w.shape == (3 * hd, pc_hh.cin) and xg.shape == (b * t, 3 * hd)
        hbuf = self.empty((2, b, hd), torch.float32)
        counter = self.empty((1,), torch.int32)
        check(self.lib.af_gru_sequence(self.h, _ptr(xg), _ptr(pc_hh.w), _ptr(pc_hh.bias), _ptr(h0), _ptr(hbuf),
                                       _ptr(hseq16), hseq16.stride(0), _ptr(h_out), _ptr(counter), b, t, hd,
                                       1 if split else 0, self._stream()), "af_gru_sequence")
        self._count()
        self.keep(xg, pc_hh.w, pc_hh.bias, h0, hbuf, hseq16, h_out, counter)
        self.release(hbuf)
        self.release(counter)

    def can_gru_sequence_tc(self, b, hd, split=False):
        """Whole recurrence in one persistent tensor-core launch (af_gru_sequence_tc): up to 64 sequences."""
        return (os.environ.get("AF_NO_GRU_TC") is None and
                bool(self.lib.af_gru_sequence_tc_supported(self.h, int(b), int(hd), 1 if split else 0)))

    def gru_sequence_tc(self, xg, pc_hh, b, t, hseq16, h0=None, h_out=None):
        """All T steps of a GRU for up to 64 sequences in one persistent tcgen05 launch.  xg (B*T,3H) fp32, pc_hh =
        PackedConv of W_hh (+ b_hh), plain or split-precision (then hseq16 rows are [hi | lo | hi])."""
        split = bool(getattr(pc_hh, "split", False))
        hd = pc_hh.cin // 3 if split else pc_hh.cin
        assert pc_hh.w.shape[0] == 3 * hd and xg.shape == (b * t, 3 * hd)
        hbuf = self.empty((2, b, (2 if split else 1) * hd), torch.float16)
        counter = self.empty((1,), torch.int32)
        check(self.lib.af_gru_sequence_tc(self.h, _ptr(xg), _ptr(pc_hh.w), pc_hh.w.shape[1], _ptr(pc_hh.bias), _ptr(h0),
                                          _ptr(hbuf), _ptr(hseq16), hseq16.stride(0), _ptr(h_out), _ptr(counter), b, t,
                                          hd, 1 if split else 0, self._stream()), "af_gru_sequence_tc")
        self._count()
        self.keep(xg, pc_hh.w, pc_hh.bias, h0, hbuf, hseq16, h_out, counter)
        self.release(hbuf)
        self.release(counter)

    def policy_head(self, logits, action_dim, h, patch, action_idx=None, action_yx=None, yx=None):
        rows = logits.shape[0]
        grid_n = int(round(math.sqrt(action_dim)))
        check(self.lib.af_policy_head(self.h, _ptr(logits), logits.stride(0), action_dim, grid_n, rows, h, patch,
                                      _ptr(action_idx), _ptr(action_yx), _ptr(yx), self._stream()), "af_policy_head")
        self._count()
        self.keep(logits, action_idx, action_yx, yx)

    def policy_head_continuous(self, logits, h, patch, action_yx=None, yx=None):
        rows = logits.shape[0]
        check(self.lib.af_policy_head_continuous(self.h, _ptr(logits), logits.stride(0), rows, h, patch,
                                                 _ptr(action_yx), _ptr(yx), self._stream()),
              "af_policy_head_continuous")
        self._count()
        self.keep(logits, action_yx, yx)

    def tsm_shift(self, x, t, fold):
        nt, h, w, c = x.shape
        out = self.empty((nt, h, w, c), torch.float16)
        check(self.lib.af_tsm_shift_nhwc_f16(self.h, _ptr(x), _ptr(out), nt, t, h * w, c, fold, self._stream()),
              "af_tsm_shift_nhwc_f16")
        self._count()
        self.keep(x, out)
        return out

    def tsm_shift_nchw_f32(self, x, t, fold):
        nt, c, h, w = x.shape
        out = self.empty((nt, c, h, w), torch.float32)
        check(self.lib.af_tsm_shift_nchw_f32(self.h, _ptr(x), _ptr(out), nt, t, c, h * w, fold, self._stream()),
              "af_tsm_shift_nchw_f32")
        self._count()
        self.keep(x, out)
        return out

    def consensus_avg(self, x, b, t, add=None, out=None):
        c = x.shape[-1]
        if out is None:
            out = self.empty((b, c), torch.float32)
        check(self.lib.af_consensus_avg(self.h, _ptr(x), _ptr(add), _ptr(out), b, t, c, self._stream()),
              "af_consensus_avg")
        self._count()
        self.keep(x, add, out)
        return out

    def fill(self, t, v):
        assert t.dtype == torch.float32
        check(self.lib.af_fill_f32(self.h, _ptr(t), float(v), t.numel(), self._stream()), "af_fill_f32")
        self._count()
        self.keep(t)

    def frames_u8_to_f32(self, frames_u8, mean, std, out=None):
        """Stack -> ToTorchFormatTensor(div=True) -> GroupNormalize (ACT/ops/transforms.py:303-336, 64-77) on the
        device: frames_u8 (B, H, W, 3T) uint8 -> (B, 3T, H, W) fp32, bit-identical to the torch ops."""
        b, h, w, c = frames_u8.shape
        assert frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.is_cuda
        if out is None:
            out = self.empty((b, c, h, w), torch.float32)
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == frames_u8.numel()
        m = (ctypes.c_float * 3)(*[float(v) for v in mean])
        sd = (ctypes.c_float * 3)(*[float(v) for v in std])
        check(self.lib.af_frames_u8_to_f32(self.h, _ptr(frames_u8), _ptr(out), b, h * w, c, m, sd, self._stream()),
              "af_frames_u8_to_f32")
        self._count()
        self.keep(frames_u8, out)
        return out

    def split3(self, x, out=None):
        """fp32 (rows, cols) [row stride >= cols] -> split-precision fp16 rows [hi | lo | hi] (rows, 3*cols)."""
        rows, cols = x.shape
        if out is None:
            out = self.empty((rows, 3 * cols), torch.float16)
        check(self.lib.af_split3_f16(self.h, _ptr(x), x.stride(0), _ptr(out), rows, cols, self._stream()),
              "af_split3_f16")
        self._count()
        self.keep(x, out)
        return out

    def f32_to_f16(self, x, out=None):
        if out is None:
            out = self.empty(tuple(x.shape), torch.float16)
        check(self.lib.af_f32_to_f16(self.h, _ptr(x), _ptr(out), x.numel(), self._stream()), "af_f32_to_f16")
        self._count()
        self.keep(x, out)
        return out


_engines = {}


def get_engine(device):
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(torch.device("cuda", idx))
    return _engines[idx]
