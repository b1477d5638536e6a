"""ResNet-50/101 focus network (fL): parameter tree + engine runner.

Mirror of ACT/models/resnet.py (Bottleneck :74-114, ResNet :117-240, get_featmap :211-225).  The module tree only
carries parameters under the reference's names (`conv1`, `bn1`, `layerL.B.convK`, `downsample.0/1`, `fc`); the
trunk runs as tcgen05 implicit-GEMM convolutions on NHWC fp16 with BatchNorm folded into the epilogue.
"""
import os

import torch
from torch import nn

from ..engine import AF_ACT_NONE, AF_ACT_RELU, fold_bn, get_engine, pack_conv, pack_stem


class Bottleneck(nn.Module):
    """Parameter container for a 1x1 -> 3x3 -> 1x1 bottleneck (stride on the 3x3, as in torchvision >= 0.4)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        raise NotImplementedError("Bottleneck runs inside the fused engine plan (ResNet.get_featmap)")


class ResNet(nn.Module):
    def __init__(self, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._stage(64, layers[0], 1)
        self.layer2 = self._stage(128, layers[1], 2)
        self.layer3 = self._stage(256, layers[2], 2)
        self.layer4 = self._stage(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2048, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._runner = None

    def _stage(self, planes, blocks, stride):
        ds = None
        if stride != 1 or self.inplanes != planes * 4:
            ds = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
        mods = [Bottleneck(self.inplanes, planes, stride, ds)]
        self.inplanes = planes * 4
        mods += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    @property
    def feature_dim(self):
        return self.fc.weight.shape[-1]

    def runner(self):
        from ..packcache import cached_runner
        from .mobilenet import _param_key
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "ResNetRunner", lambda: ResNetRunner(self, key), key)
        return self._runner

    def get_featmap(self, x, pooled=True):
        """(N,3,P,P) fp32 patches -> (N,2048,1,1) fp32 (pooled) or (N,2048,h,w); ACT/models/resnet.py:211-225."""
        eng = get_engine(x.device)
        fmap = self.runner().run(eng, x.contiguous())
        n, h, w, c = fmap.shape
        if not pooled:
            return eng.nhwc_to_nchw_f32(fmap)
        out = torch.empty(n, c, dtype=torch.float32, device=x.device)
        eng.avgpool(fmap, out_f32=out, out_f32_stride=c)
        return out.view(n, c, 1, 1)

    def get_featvec(self, x):
        return self.get_featmap(x, pooled=True).flatten(1)

    def forward(self, x):
        raise NotImplementedError("the ImageNet/fc head of fL is outside the inference hot path (stage-0 training)")


def _fold(bn):
    return fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


class ResNetRunner:
    """Bottleneck ResNet trunk in kernel layout.  Works on the reference's ResNet, on this mirror and on
    torchvision.models.resnet50/101 (STH/models/tsn.py:114), including a trunk wrapped in nn.Sequential with the fc
    stripped (STH/evaluate.py:83) -- `from_children` handles that form."""

    def __init__(self, net, key=None):
        self.key = key
        mods = _trunk_modules(net)
        dev = mods["conv1"].weight.device
        s, b = _fold(mods["bn1"])
        self.stem = pack_stem(mods["conv1"].weight, s, b, stride=2, pad=3, act=AF_ACT_RELU, device=dev)
        self.blocks = []
        for stage in mods["stages"]:
            for blk in stage:
                e = {}
                conv1 = blk.conv1
                shift = None
                if hasattr(conv1, "net"):          # STH TemporalShift wrapper around conv1 (temporal_shift.py:13-27)
                    shift = (conv1.n_segment, conv1.fold_div)
                    conv1 = conv1.net
                e["shift"] = shift
                s, b = _fold(blk.bn1)
                e["c1"] = pack_conv(conv1.weight, s, b, act=AF_ACT_RELU, device=dev)
                s, b = _fold(blk.bn2)
                e["c2"] = pack_conv(blk.conv2.weight, s, b, stride=blk.conv2.stride[0], pad=1, act=AF_ACT_RELU,
                                    device=dev)
                s, b = _fold(blk.bn3)
                e["c3"] = pack_conv(blk.conv3.weight, s, b, act=AF_ACT_RELU, device=dev, fold_scale=True)   # relu after the add
                e["c3ds"] = None
                if blk.downsample is not None:
                    sd, bd = _fold(blk.downsample[1])
                    e["ds"] = pack_conv(blk.downsample[0].weight, sd, bd, stride=blk.downsample[0].stride[0],
                                        act=AF_ACT_NONE, device=dev)
                    if os.environ.get("AF_NO_SHORTCUT_FUSION") is None:
                        # projection shortcut accumulated inside conv3's tile: both BN scales folded into the fp16
                        # weights, one bias; the downsample tensor is never written or re-read
                        c3f = pack_conv(blk.conv3.weight, s, b + bd, act=AF_ACT_RELU, device=dev, fold_scale=True)
                        dsf = pack_conv(blk.downsample[0].weight, sd, None, stride=blk.downsample[0].stride[0],
                                        act=AF_ACT_NONE, device=dev, fold_scale=True, block_n=c3f.block_n)
                        e["c3ds"] = (c3f, dsf)
                else:
                    e["ds"] = None
                self.blocks.append(e)

    def run_pooled_chunked(self, eng, frames, out, out_stride, chunk, yx=None, patch=None, yx_div=1, out_f32=False):
        """Trunk + global average pool over sub-batches of `chunk` patches (L2-resident intermediates, see
        MobileNetV2Runner.run_chunked); pooled features go to `out` rows (fp16, or fp32 with out_f32; row stride
        out_stride).  A chunk never splits a group of yx_div frames that share one crop origin."""
        n = frames.shape[0]
        chunk = n if chunk is None else max(1, min(chunk, n))
        if chunk < n and chunk % yx_div:
            chunk = max(yx_div, chunk // yx_div * yx_div)
        for s0 in range(0, n, chunk):
            s1 = min(n, s0 + chunk)
            sub_yx = None if yx is None else yx[s0 // yx_div:(s1 + yx_div - 1) // yx_div]
            fmap = self.run(eng, frames[s0:s1], yx=sub_yx, patch=patch, yx_div=yx_div)
            if out_f32:
                eng.avgpool(fmap, out_f32=out[s0:s1], out_f32_stride=out_stride)
            else:
                eng.avgpool(fmap, out_f16=out[s0:s1], out_f16_stride=out_stride)
            eng.release(fmap)

    def run(self, eng, frames, yx=None, patch=None, yx_div=1):
        """frames (N,3,H,W) fp32; with yx (N,2 int32) + patch the crop of ACT/models/utils.py:37-51 is fused into the
        stem staging.  Returns the layer4 output (N,h,w,2048) NHWC fp16."""
        p_eff = patch if patch is not None else frames.shape[-1]
        if frames.shape[-1] == frames.shape[-2] and eng.stem_pool_ok(self.stem, p_eff):
            x = eng.stem(frames, self.stem, yx=yx, patch=patch, yx_div=yx_div, pool=True)   # conv + BN + ReLU + maxpool
        else:
            x = eng.stem(frames, self.stem, yx=yx, patch=patch, yx_div=yx_div)
            y = eng.maxpool3x3s2(x)
            eng.release(x)
            x = y
        for e in self.blocks:
            inp = x
            a = x
            tsm = None
            if e["shift"] is not None:
                t_seg, fold = e["shift"][0], x.shape[-1] // e["shift"][1]
                if eng.conv_tsm_ok(x, t_seg, fold):
                    tsm = (t_seg, fold)         # the shift rides in conv1's TMA loads (no shifted copy in HBM)
                else:
                    a = eng.tsm_shift(x, t_seg, fold)
            h1 = eng.conv(a, e["c1"], tsm=tsm)
            if a is not inp:
                eng.release(a)
            h2 = eng.conv(h1, e["c2"])
            eng.release(h1)
            if e["c3ds"] is not None:
                x = eng.conv(h2, e["c3ds"][0], shortcut=(inp, e["c3ds"][1]))
                eng.release(h2)
                eng.release(inp)
                continue
            if e["ds"] is not None:
                idn = eng.conv(inp, e["ds"])
            else:
                idn = inp
            x = eng.conv(h2, e["c3"], residual=idn)
            eng.release(h2)
            if idn is not inp:
                eng.release(idn)
            eng.release(inp)
        return x


def _trunk_modules(net):
    if hasattr(net, "conv1") and hasattr(net, "layer1"):
        return {"conv1": net.conv1, "bn1": net.bn1, "stages": [net.layer1, net.layer2, net.layer3, net.layer4]}
    # nn.Sequential(*children[:-1]) form: conv1, bn1, relu, maxpool, layer1..4, avgpool
    ch = list(net.children())
    return {"conv1": ch[0], "bn1": ch[1], "stages": [m for m in ch if isinstance(m, nn.Sequential)]}


def resnet50(pretrained=False, progress=True, **kwargs):
    """ACT/models/resnet.py:280-289. `pretrained` is accepted and ignored (no network; callers load checkpoints)."""
    return ResNet([3, 4, 6, 3], **kwargs)


def resnet101(pretrained=False, progress=True, **kwargs):
    """ACT/models/resnet.py:292-301."""
    return ResNet([3, 4, 23, 3], **kwargs)
