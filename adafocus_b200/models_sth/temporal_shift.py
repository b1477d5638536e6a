"""Temporal Shift Module -- mirror of STH/ops/temporal_shift.py (TemporalShift :13-46, make_temporal_shift :99-142).

`TemporalShift` keeps the reference's wrapper form (`.net`, `.n_segment`, `.fold_div`) so checkpoint keys read
`...conv1.net.weight`; inside the engine the shift is the NHWC fp16 kernel af_tsm_shift_nhwc_f16 in front of the
wrapped 1x1 convolution.  The static `shift()` works on reference-layout fp32 tensors (bit-exact copy kernel)."""
import torch
from torch import nn

from ..engine import get_engine


class TemporalShift(nn.Module):
    def __init__(self, net, n_segment=3, n_div=8, inplace=False):
        super().__init__()
        if inplace:
            raise NotImplementedError("in-place shift raises in the reference too (STH/ops/temporal_shift.py:38)")
        self.net = net
        self.n_segment = n_segment
        self.fold_div = n_div
        self.inplace = inplace

    def forward(self, x):
        raise NotImplementedError("TemporalShift runs inside the fused engine plan (shift + wrapped conv)")

    @staticmethod
    def shift(x, n_segment, fold_div=3, inplace=False):
        """(nt, c, h, w) fp32 CUDA -> shifted copy: channels [0, c/fold_div) come from frame t+1, the next c/fold_div
        from frame t-1, zeros at the clip ends (STH/ops/temporal_shift.py:29-46)."""
        if inplace:
            raise NotImplementedError
        if not x.is_cuda:
            raise RuntimeError("adafocus_b200 has no CPU path")
        nt, c, h, w = x.shape
        eng = get_engine(x.device)
        return eng.tsm_shift_nchw_f32(x.contiguous().float(), n_segment, c // fold_div)


def make_temporal_shift(net, n_segment, n_div=8, place="blockres", temporal_pool=False):
    """Wrap conv1 of the residual blocks of a bottleneck ResNet (every block; every 2nd block when layer3 has >= 23
    blocks, i.e. ResNet-101) -- STH/ops/temporal_shift.py:122-142."""
    if temporal_pool:
        raise NotImplementedError("temporal_pool is not used by any shipped configuration")
    if "blockres" not in place:
        raise NotImplementedError(place)
    n_round = 2 if len(list(net.layer3.children())) >= 23 else 1
    for stage in (net.layer1, net.layer2, net.layer3, net.layer4):
        for i, blk in enumerate(stage.children()):
            if i % n_round == 0:
                blk.conv1 = TemporalShift(blk.conv1, n_segment=n_segment, n_div=n_div)
