"""TSM-MobileNet-V2 glancer of the Something-Something tree -- mirror of STH/models/mobilenetv2.py (tonylins layout:
flat `features.N.conv.K` Sequentials, `classifier` is a bare Linear; get_featmap :116-121 returns (map, logits))."""
import os

import torch
from torch import nn

from ..engine import (AF_ACT_NONE, AF_ACT_RELU6, fold_bn, get_engine, host, mbconv_supported, pack_conv,
                      pack_conv_split, pack_mbconv, pack_mbconv_rows, pack_stem)
from ..packcache import cached_runner
from ..models.mobilenet import _FUSE_MIN_HW, _MBV2_SETTING, MobileNetV2Runner, _param_key, _rows_choice


def conv_bn(inp, oup, stride):
    return nn.Sequential(nn.Conv2d(inp, oup, 3, stride, 1, bias=False), nn.BatchNorm2d(oup), nn.ReLU6(inplace=True))


def conv_1x1_bn(inp, oup):
    return nn.Sequential(nn.Conv2d(inp, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    """Parameter container; `conv` is the reference's flat Sequential (5 entries for t=1, 8 otherwise)."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        hidden = int(inp * expand_ratio)
        self.stride, self.expand = stride, expand_ratio
        self.use_res_connect = stride == 1 and inp == oup
        dw = [nn.Conv2d(hidden, hidden, 3, stride, 1, groups=hidden, bias=False), nn.BatchNorm2d(hidden),
              nn.ReLU6(inplace=True), nn.Conv2d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup)]
        pw = [] if expand_ratio == 1 else [nn.Conv2d(inp, hidden, 1, 1, 0, bias=False), nn.BatchNorm2d(hidden),
                                           nn.ReLU6(inplace=True)]
        self.conv = nn.Sequential(*(pw + dw))

    def forward(self, x):
        raise NotImplementedError("InvertedResidual runs inside the fused engine plan")


class MobileNetV2(nn.Module):
    def __init__(self, n_class=1000, input_size=224, width_mult=1.0):
        super().__init__()
        assert input_size % 32 == 0 and width_mult == 1.0
        cin, self.last_channel = 32, 1280
        feats = [conv_bn(3, cin, 2)]
        for t, c, n, s in _MBV2_SETTING:
            for i in range(n):
                feats.append(InvertedResidual(cin, c, s if i == 0 else 1, t))
                cin = c
        feats.append(conv_1x1_bn(cin, self.last_channel))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Linear(self.last_channel, n_class)
        self._runner = None

    @property
    def feature_dim(self):
        return self.last_channel

    def runner(self):
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "SthGlancerRunner", lambda: SthGlancerRunner(self, key), key)
        return self._runner

    def get_featmap(self, x):
        """(N,3,H,W) fp32 -> (feature map (N,1280,h,w) fp32, logits (N,n_class) fp32)."""
        eng = get_engine(x.device)
        r = self.runner()
        fmap = r.run(eng, x.contiguous())
        logits = r.logits(eng, fmap)
        return eng.nhwc_to_nchw_f32(fmap), logits

    def forward(self, x):
        return self.get_featmap(x)[1]


class SthGlancerRunner(MobileNetV2Runner):
    """Same layer schedule as the ACT runner, fed from the flat tonylins module layout; TSM-wrapped first convs
    (STH/models/gfv_net.py:238-241) turn on the temporal shift in front of the residual blocks' expand conv."""

    def __init__(self, net, key):
        self.key = key
        dev = next(net.parameters()).device
        f = list(net.features)
        c0, b0 = f[0][0], f[0][1]
        s, b = fold_bn(b0.weight, b0.bias, b0.running_mean, b0.running_var, b0.eps)
        self.stem_direct = True
        self.stem_w = host(c0.weight).permute(2, 3, 1, 0).reshape(27, 32).contiguous().to(dev)
        self.stem_s, self.stem_b = s.contiguous().to(dev), b.contiguous().to(dev)
        self.stem = pack_stem(c0.weight, s, b, stride=2, pad=1, act=AF_ACT_RELU6, device=dev)
        self.blocks = []
        self.tsm = None
        fuse = os.environ.get("AF_NO_MBCONV_FUSED") is None
        for blk in f[1:-1]:
            seq = list(blk.conv)
            e = {"res": blk.use_res_connect, "stride": blk.stride, "expand": None, "shift": False}
            if blk.expand != 1:
                cv = seq[0]
                if hasattr(cv, "net"):                       # TemporalShift wrapper
                    self.tsm = (cv.n_segment, cv.fold_div)
                    e["shift"] = True
                    cv = cv.net
                s1, b1 = fold_bn(seq[1].weight, seq[1].bias, seq[1].running_mean, seq[1].running_var, seq[1].eps)
                e["expand"] = pack_conv(cv.weight, s1, b1, act=AF_ACT_RELU6, device=dev)
                seq = seq[3:]
            dw, bn = seq[0], seq[1]
            s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            e["dw_w"] = host(dw.weight).reshape(dw.weight.shape[0], 9).t().contiguous().to(dev)
            e["dw_s"], e["dw_b"] = s.contiguous().to(dev), b.contiguous().to(dev)
            dw_s, dw_b = s, b
            pw, bn = seq[3], seq[4]
            s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            e["project"] = pack_conv(pw.weight, s, b, act=AF_ACT_NONE, device=dev, fold_scale=e["res"])
            e["fused"] = None
            e["rows"] = {}
            if (fuse and blk.expand != 1 and
                    mbconv_supported(1, 32, 32, cv.weight.shape[1], cv.weight.shape[0], pw.weight.shape[0], blk.stride)):
                # expand -> depthwise -> project as one launch (adafocus_b200/csrc/mbconv_fused.cu)
                e["fused"] = pack_mbconv(cv.weight, s1, b1, dw.weight, dw_s, dw_b, pw.weight, s, b, blk.stride,
                                         device=dev)
                # row-streaming form for large batches (adafocus_b200/csrc/mbconv_rows.cu)
                e["rows"] = {spr: pack_mbconv_rows(cv.weight, s1, b1, dw.weight, dw_s, dw_b, pw.weight, s, b, blk.stride,
                                                   spr, device=dev) for spr in (1, 2, 4)}
            self.blocks.append(e)
        cl, bl = f[-1][0], f[-1][1]
        s, b = fold_bn(bl.weight, bl.bias, bl.running_mean, bl.running_var, bl.eps)
        self.last = pack_conv(cl.weight, s, b, act=AF_ACT_RELU6, device=dev)
        self.fc = pack_conv_split(net.classifier.weight, net.classifier.bias, device=dev)
        self.num_classes = net.classifier.weight.shape[0]
        self.logit_stride = (self.num_classes + 7) // 8 * 8

    def run(self, eng, frames, tsm=None):
        if eng.s2d_stem and self.stem.s2d is not None and frames.shape[-1] % 2 == 0 and frames.shape[-2] == frames.shape[-1]:
            x = eng.stem(frames, self.stem)
        else:
            x = eng.stem_conv3x3s2_c32(frames, self.stem_w, self.stem_s, self.stem_b)
        for e in self.blocks:
            inp, y = x, x
            if e["shift"]:
                y = eng.tsm_shift(x, self.tsm[0], x.shape[-1] // self.tsm[1])
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
            x = eng.conv(d, e["project"], residual=inp if e["res"] else None)
            eng.release(d)
            eng.release(inp)
        out = eng.conv(x, self.last)
        eng.release(x)
        return out

    def logits(self, eng, fmap, vec16=None, padded=False):
        """classifier(mean over H,W) -> fp32 (N, n_class) [or the row-padded (N, logit_stride) buffer]."""
        n, h, w, c = fmap.shape
        assert vec16 is None, "the glancer head consumes fp32 pooled features (split-precision GEMM)"
        vec = eng.empty((n, c), torch.float32)
        eng.avgpool(fmap, out_f32=vec, out_f32_stride=c)
        vec3 = eng.split3(vec)
        eng.release(vec)
        out = eng.empty((n, self.logit_stride), torch.float32)
        eng.linear(vec3, self.fc, out=out, out_f32=True, out_stride=self.logit_stride)
        eng.release(vec3)
        if padded or self.logit_stride == self.num_classes:
            return out
        return out[:, : self.num_classes].contiguous()


def mobilenet_v2(n_class, pretrained=True):
    """STH/models/mobilenetv2.py:148-160; the Dropbox checkpoint is unreachable offline, `pretrained` is ignored."""
    return MobileNetV2(n_class=n_class, width_mult=1)
