"""MobileNet-V2 glance network (fG): parameter tree + engine runner.

Mirror of ACT/models/mobilenet.py (MobileNetV2 :71-152, get_featmap :146-148): the nn.Module tree only carries
parameters under the reference's names (`features.N...`, `classifier.1.*`) so reference checkpoints load unchanged;
the arithmetic runs in adafocus_b200's CUDA kernels (NHWC fp16, tcgen05 1x1 convs + depthwise 3x3 kernels).
"""
import os

import torch
from torch import nn

from ..packcache import cached_runner
from ..engine import (AF_ACT_NONE, AF_ACT_RELU6, fold_bn, get_engine, host, mbconv_rows_spr, mbconv_rows_supported,
                      stem_s2d_weights,
                      mbconv_supported, pack_conv, pack_mbconv, pack_mbconv_rows,
                      pack_stem)

# Experiment knob: blocks whose input is smaller than this run unfused (1x1 conv -> depthwise -> 1x1 conv).  Measured
# at 1024 frames: unfusing the 14x14 blocks changes the glance stage by < 1 %, unfusing the 28x28 ones costs 5 %.
_FUSE_MIN_HW = int(os.environ.get("AF_MB_MIN_HW", "0"))

# (expand t, channels c, repeats n, first stride s) -- the MobileNet-V2 paper's table, as at ACT/models/mobilenet.py:89-98
_MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                 (6, 320, 1, 1))


def _round_channels(v, divisor=8):
    r = max(divisor, int(v + divisor / 2) // divisor * divisor)
    return r + divisor if r < 0.9 * v else r


def _cbr(cin, cout, k=3, stride=1, groups=1):
    """conv -> BN -> ReLU6 triple; children are named 0/1/2 like the reference's ConvBNReLU."""
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    """Parameter container for one inverted-residual block (`conv` Sequential as in the reference)."""

    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        hidden = int(round(cin * expand))
        self.stride, self.expand, self.cin, self.cout, self.hidden = stride, expand, cin, cout, hidden
        self.use_res_connect = stride == 1 and cin == cout
        layers = [] if expand == 1 else [_cbr(cin, hidden, k=1)]
        layers += [_cbr(hidden, hidden, stride=stride, groups=hidden), nn.Conv2d(hidden, cout, 1, bias=False),
                   nn.BatchNorm2d(cout)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        raise NotImplementedError("InvertedResidual runs inside the fused engine plan (MobileNetV2.get_featmap)")


class MobileNetV2(nn.Module):
    def __init__(self, num_classes=1000, width_mult=1.0, round_nearest=8):
        super().__init__()
        cin = _round_channels(32 * width_mult, round_nearest)
        self.last_channel = _round_channels(1280 * max(1.0, width_mult), round_nearest)
        feats = [_cbr(3, cin, stride=2)]
        for t, c, n, s in _MBV2_SETTING:
            cout = _round_channels(c * width_mult, round_nearest)
            for i in range(n):
                feats.append(InvertedResidual(cin, cout, s if i == 0 else 1, t))
                cin = cout
        feats.append(_cbr(cin, self.last_channel, k=1))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(self.last_channel, num_classes))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)
        self._runner = None

    @property
    def feature_dim(self):
        return self.last_channel

    def runner(self):
        """Packed-weight runner, rebuilt when parameters were reloaded / moved."""
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "MobileNetV2Runner", lambda: MobileNetV2Runner(self, key), key)
        return self._runner

    def get_featmap(self, x):
        """(N,3,H,W) fp32 -> (feature map (N,1280,h,w) fp32, spatial mean (N,1280) fp32), ACT/models/mobilenet.py:146-148."""
        eng = get_engine(x.device)
        r = self.runner()
        fmap = r.run(eng, x.contiguous())
        n, h, w, c = fmap.shape
        vec = torch.empty(n, c, dtype=torch.float32, device=x.device)
        eng.avgpool(fmap, out_f32=vec, out_f32_stride=c)
        return eng.nhwc_to_nchw_f32(fmap), vec

    def forward(self, x):
        raise NotImplementedError("classification head of fG is outside the inference hot path (stage-0 training)")


class _PackState:
    __slots__ = ("epoch", "probes")

    def __init__(self):
        self.epoch = 0
        self.probes = ()


_ROWS_MODE = os.environ.get("AF_MBROWS", "auto")      # "0": never, "force": whenever supported, "auto": heuristic below


def _rows_choice(eng, e, y):
    """The row-streaming packing of block e for input y (N,H,W,C), or None to use the tiled kernel.  The row kernel gives
    every CTA whole frames: it needs at least one frame (segment) per SM; with five depthwise warps per lane quarter
    (cexp = 144 at stride 1) its 72-register budget spills and the tiled kernel stays faster (profiles/README.md)."""
    if _ROWS_MODE == "0" or not e["rows"]:
        return None
    n, h, w, cin = y.shape
    if h != w:
        return None
    pr = e["rows"].get(mbconv_rows_spr(w, e["stride"]))
    if pr is None or not mbconv_rows_supported(n, h, w, cin, pr.cexp, pr.cout, pr.stride):
        return None
    if _ROWS_MODE == "force":
        return pr
    units = n * (2 if (pr.stride == 2 and w == 112) else 1)
    # CTAs own whole frames: the launch lasts ceil(units / SMs) frame times, and a frame costs ~0.75 of what it costs
    # the tiled kernel -- so skip the counts just above a multiple of the SM count (cfg5: 256 frames on 148 SMs gains 2 %)
    sms = eng.ctx.sm_count
    if units < sms or units < 0.75 * sms * ((units + sms - 1) // sms):
        return None
    if pr.stride == 1 and pr.cexp % 128 == 16:
        return None
    return pr


def _pack_state(module):
    st = module.__dict__.get("_af_pack_state")
    if st is None:
        st = _PackState()
        tensors = list(module.parameters()) + list(module.buffers())
        st.probes = tuple(tensors[i] for i in sorted({0, len(tensors) // 2, len(tensors) - 1})) if tensors else ()

        def bump(_m, _incompatible):
            st.epoch += 1
        # load_state_dict() runs the post hooks of every module it visits: registering on each sub-module catches a
        # load on this module, on any ancestor and on any descendant
        for sub in module.modules():
            sub.register_load_state_dict_post_hook(bump)
        module.__dict__["_af_pack_state"] = st
    return st


def invalidate_packed(module):
    """Force the packed (kernel-layout) copy of `module`'s weights to be rebuilt on next use.  Needed only after
    writes the hooks cannot see: in-place edits through `.data` / `torch.no_grad()` of individual tensors."""
    _pack_state(module).epoch += 1


def _param_key(module):
    """O(1) identity of a module's weights for the packed-copy caches: an epoch bumped by load_state_dict() post
    hooks (on the module, its ancestors or descendants) + (data_ptr, _version) of three probe tensors, which change on
    .to()/.cuda()/.half() and on in-place optimiser-style updates.  (Round 1 walked the whole state_dict per call.)"""
    st = _pack_state(module)
    return (id(module), st.epoch) + tuple((p.data_ptr(), p._version) for p in st.probes)


class MobileNetV2Runner:
    """Weights of a MobileNetV2 in kernel layout + the layer schedule."""

    def __init__(self, net, key):
        self.key = key
        dev = next(net.parameters()).device
        f = net.features
        c0, b0 = f[0][0], f[0][1]
        s, b = fold_bn(b0.weight, b0.bias, b0.running_mean, b0.running_var, b0.eps)
        self.stem_direct = tuple(c0.weight.shape) == (32, 3, 3, 3) and c0.stride == (2, 2)
        if self.stem_direct:
            # fp32 [27][32], k = (r*3+s)*3 + c
            self.stem_w = host(c0.weight).permute(2, 3, 1, 0).reshape(27, 32).contiguous().to(dev)
            self.stem_s, self.stem_b = s.contiguous().to(dev), b.contiguous().to(dev)
        # tensor-core form (space-to-depth + windowed 2x1 conv) -- preferred for even frame sizes
        self.stem = pack_stem(c0.weight, s, b, stride=2, pad=1, act=AF_ACT_RELU6, device=dev)
        self.blocks = []
        for blk in list(f)[1:-1]:
            seq = list(blk.conv)
            entry = {"res": blk.use_res_connect, "stride": blk.stride}
            if blk.expand != 1:
                cv, bn = seq[0][0], seq[0][1]
                s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
                entry["expand"] = pack_conv(cv.weight, s, b, act=AF_ACT_RELU6, device=dev)
                seq = seq[1:]
            else:
                entry["expand"] = None
            dw, bn = seq[0][0], seq[0][1]
            s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            entry["dw_w"] = host(dw.weight).reshape(dw.weight.shape[0], 9).t().contiguous().to(dev)
            entry["dw_s"], entry["dw_b"] = s.contiguous().to(dev), b.contiguous().to(dev)
            entry["_dw_raw"], entry["_dw_sb"] = host(dw.weight), (s, b)
            pw, bn = seq[1], seq[2]
            s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            entry["project"] = pack_conv(pw.weight, s, b, act=AF_ACT_NONE, device=dev, fold_scale=entry["res"])
            entry["_proj"] = (host(pw.weight).flatten(1), s, b)
            entry["_exp"] = None
            if blk.expand != 1:
                cv, bne = list(blk.conv)[0][0], list(blk.conv)[0][1]
                se, be = fold_bn(bne.weight, bne.bias, bne.running_mean, bne.running_var, bne.eps)
                entry["_exp"] = (host(cv.weight).flatten(1), se, be)
            self.blocks.append(entry)
        # Front end as ONE row-streaming launch for large batches (engine.stem_front): the stem conv is the expand GEMM of
        # block 1 (t = 1: depthwise + project), whose 16-channel output then feeds block 2 un-merged (cin = 16).
        self.front, self.block2_rows = None, None
        b0e, b1e = self.blocks[0], (self.blocks[1] if len(self.blocks) > 1 else None)
        if (_ROWS_MODE != "0" and self.stem.s2d is not None and self.stem.s2d.vt == 2 and b0e["_exp"] is None
                and b0e["stride"] == 1 and not b0e["res"] and b1e is not None and b1e["_exp"] is not None):
            w64, _ = stem_s2d_weights(c0.weight)
            s0, bb0 = fold_bn(b0.weight, b0.bias, b0.running_mean, b0.running_var, b0.eps)
            wp, sp, bp = b0e["_proj"]
            self.front = pack_mbconv_rows(w64.reshape(w64.shape[0], 64), s0, bb0, b0e["_dw_raw"], b0e["_dw_sb"][0],
                                          b0e["_dw_sb"][1], wp, sp, bp, 1, 4, device=dev)
            we, se, be = b1e["_exp"]
            wp1, sp1, bp1 = b1e["_proj"]
            self.block2_rows = {spr: pack_mbconv_rows(we, se, be, b1e["_dw_raw"], b1e["_dw_sb"][0], b1e["_dw_sb"][1],
                                                      wp1, sp1, bp1, b1e["stride"], spr, device=dev) for spr in (2, 4)}
        # A linear project conv (+BN) whose only consumer is the next block's expand conv (+BN+ReLU6) -- i.e. neither
        # block has a residual connection -- composes into ONE 1x1 conv: W = W_e diag(s_p) W_p, bias = s_e (W_e b_p) + b_e.
        # Exact algebra (no activation in between, ACT/models/mobilenet.py:55-62); the narrow tensor is never stored.
        for k in range(len(self.blocks) - 1):
            a, nxt = self.blocks[k], self.blocks[k + 1]
            if not a["res"] and not nxt["res"] and nxt["_exp"] is not None:
                wp, sp, bp = a["_proj"]
                we, se, be = nxt["_exp"]
                wm = we @ (sp[:, None] * wp)
                bm = se * (we @ bp) + be
                a["project"] = None
                nxt["expand"] = pack_conv(wm, se, bm, act=AF_ACT_RELU6, device=dev)
                nxt["_exp"] = (wm, se, bm)
        # Blocks whose channel counts fit the fused inverted-residual kernel (expand -> depthwise -> project in one
        # launch, the expanded tensor never reaches HBM) also get that packing; run() uses it wherever the spatial
        # size is supported.
        self.fuse_blocks = os.environ.get("AF_NO_MBCONV_FUSED") is None
        for e in self.blocks:
            e["fused"] = None
            e["rows"] = {}
            if self.fuse_blocks and e["_exp"] is not None and e["project"] is not None:
                we, se, be = e["_exp"]
                wp, sp, bp = e["_proj"]
                if mbconv_supported(1, 32, 32, we.shape[1], we.shape[0], wp.shape[0], e["stride"]):
                    e["fused"] = pack_mbconv(we, se, be, e["_dw_raw"], e["_dw_sb"][0], e["_dw_sb"][1], wp, sp, bp,
                                             e["stride"], device=dev)
                    # row-streaming form (large batches), one packing per strips-per-row value the kernel has a lane
                    # placement for
                    e["rows"] = {spr: pack_mbconv_rows(we, se, be, e["_dw_raw"], e["_dw_sb"][0], e["_dw_sb"][1], wp, sp,
                                                       bp, e["stride"], spr, device=dev) for spr in (1, 2, 4)}
        for e in self.blocks:
            e.pop("_proj"), e.pop("_exp"), e.pop("_dw_raw"), e.pop("_dw_sb")
        cl, bl = f[-1][0], f[-1][1]
        s, b = fold_bn(bl.weight, bl.bias, bl.running_mean, bl.running_var, bl.eps)
        self.last = pack_conv(cl.weight, s, b, act=AF_ACT_RELU6, device=dev)

    def _front_choice(self, eng, frames):
        if self.front is None or self.block2_rows is None or _ROWS_MODE == "0":
            return False
        n, _, h, w = frames.shape
        if not eng.stem_front_ok(self.stem, self.front, frames):
            return False
        pr = self.block2_rows.get(mbconv_rows_spr(w // 2, self.blocks[1]["stride"]))
        if pr is None or not mbconv_rows_supported(n, h // 2, w // 2, pr.cin, pr.cexp, pr.cout, pr.stride):
            return False
        # Measured at 1024 frames (tools/front_time.py): prepass + fused launch 1334 us vs prepass + stem conv 570 us +
        # depthwise 343 us -- 32 expanded channels give the row kernel ONE depthwise warp per scheduler, which is
        # latency-bound (2.5 k cycles per two-row step).  Opt-in (AF_STEM_FRONT=1) and exercised by the forced-rows
        # parity tests only.
        return _ROWS_MODE == "force" or (os.environ.get("AF_STEM_FRONT") == "1" and n >= eng.ctx.sm_count)

    def run_chunked(self, eng, frames, chunk, tsm=None):
        """run() over sub-batches of `chunk` frames so that every intermediate tensor of a sub-batch stays resident in
        the 126 MB L2 between the layer that writes it and the layer that reads it (the workspace arena hands the same
        hot buffers to every sub-batch).  Returns the full (N,h,w,1280) map."""
        n = frames.shape[0]
        if chunk is not None and tsm is not None and chunk % tsm[0]:
            chunk = max(tsm[0], chunk // tsm[0] * tsm[0])      # the temporal shift needs whole clips in a sub-batch
        if chunk is None or chunk >= n:
            return self.run(eng, frames, tsm=tsm)
        out = None
        for s0 in range(0, n, chunk):
            s1 = min(n, s0 + chunk)
            part = self.run(eng, frames[s0:s1], tsm=tsm, out_full=out, out_slice=(s0, s1), n_total=n)
            out = part
        return out

    def run(self, eng, frames, tsm=None, out_full=None, out_slice=None, n_total=None):
        """frames (N,3,H,W) fp32 contiguous -> (N,h,w,1280) NHWC fp16. tsm=(T, shift_div) applies the temporal shift
        to the input of every residual block's first 1x1 conv (STH/models/gfv_net.py:238-241)."""
        blocks = self.blocks
        if tsm is None and self._front_choice(eng, frames):
            # stem conv + block 1 in one launch, block 2 from the 16-channel tensor (its own expand conv)
            x = eng.stem_front(frames, self.stem, self.front)
            pr = self.block2_rows.get(mbconv_rows_spr(x.shape[2], self.blocks[1]["stride"]))
            y = eng.mbconv_rows(x, pr)
            eng.release(x)
            x = y
            blocks = self.blocks[2:]
        elif self.stem_direct and not (eng.s2d_stem and self.stem.s2d is not None and frames.shape[-1] % 2 == 0
                                       and frames.shape[-2] == frames.shape[-1]):
            x = eng.stem_conv3x3s2_c32(frames, self.stem_w, self.stem_s, self.stem_b)
        else:
            x = eng.stem(frames, self.stem)
        for e in blocks:
            inp = x
            y = x
            if tsm is not None and e["res"]:
                y = eng.tsm_shift(y, tsm[0], y.shape[-1] // tsm[1])
            pr = _rows_choice(eng, e, y)
            if pr is not None:
                x = eng.mbconv_rows(y, pr, residual=inp if e["res"] else None)
                if y is not inp:
                    eng.release(y)
                eng.release(inp)
                continue
            if (e["fused"] is not None and y.shape[1] >= _FUSE_MIN_HW
                    and mbconv_supported(*y.shape, e["fused"].cexp, e["fused"].cout, e["stride"])):
                x = eng.mbconv(y, e["fused"], residual=inp if e["res"] else None)
                if y is not inp:
                    eng.release(y)
                eng.release(inp)
                continue
            if e["expand"] is not None:
                h = eng.conv(y, e["expand"])
                if y is not inp:
                    eng.release(y)
            else:
                h = y
            d = eng.dwconv3x3(h, e["dw_w"], e["dw_s"], e["dw_b"], e["stride"])
            if h is not inp:
                eng.release(h)
            if e["project"] is None:         # merged into the next block's expand conv
                x = d
            else:
                x = eng.conv(d, e["project"], residual=inp if e["res"] else None)
                eng.release(d)
            eng.release(inp)
        if out_slice is None:
            out = eng.conv(x, self.last)
            eng.release(x)
            return out
        if out_full is None:
            out_full = eng.empty((n_total, x.shape[1], x.shape[2], self.last.cout), torch.float16)
        eng.conv(x, self.last, out=out_full[out_slice[0]:out_slice[1]], out_stride=self.last.cout)
        eng.release(x)
        return out_full


def mobilenet_v2(pretrained=False, progress=True, **kwargs):
    """Constructor with the reference's signature (ACT/models/mobilenet.py:155-169). ImageNet weights are not
    downloadable here; `pretrained` is accepted and ignored (checkpoints are loaded by the caller)."""
    return MobileNetV2(**kwargs)
