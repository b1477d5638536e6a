"""Segment consensus -- mirror of STH/ops/basic_ops.py (ConsensusModule :30-37, 'avg' = mean over dim 1, keepdim)."""
import torch

from ..engine import get_engine


class Identity(torch.nn.Module):
    def forward(self, input):
        return input


class ConsensusModule(torch.nn.Module):
    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = consensus_type if consensus_type != "rnn" else "identity"
        self.dim = dim

    def forward(self, input):
        if self.consensus_type == "identity":
            return input
        if self.consensus_type != "avg":
            return None
        if self.dim != 1 or input.dim() != 3 or not input.is_cuda:
            raise NotImplementedError("consensus is implemented for (B, T, C) CUDA tensors, dim=1")
        b, t, c = input.shape
        eng = get_engine(input.device)
        return eng.consensus_avg(input.contiguous().float().view(b * t, c), b, t).view(b, 1, c)
