"""Evaluation metrics on the device: the consumers of the path's logits in the reference's
validate() loops -- top-k `accuracy` (ACT/ops/utils.py:35-49) and `cal_map` (ACT/ops/utils.py:68-88) -- without the
`.cpu()` round trip of all logits and the per-class CPU sort loop (ACT/main_dist.py:381,392-408)."""
from ctypes import c_void_p

import torch

from ._lib import check
from .engine import get_engine


def accuracy(output, target, topk=(1,)):
    """output (B, C) fp32 CUDA logits, target (B,) int64 -> [tensor([acc_k in %]) for k in topk], like the reference."""
    if len(topk) > 2:
        raise NotImplementedError("at most two k values per call")
    eng = get_engine(output.device)
    out = output.float()
    if out.stride(1) != 1:
        out = out.contiguous()
    tgt = target.to(device=output.device, dtype=torch.int64).contiguous()
    b, c = out.shape
    hits = torch.zeros(2, dtype=torch.float32, device=output.device)
    k0, k1 = topk[0], topk[-1]
    check(eng.lib.af_topk_hits(eng.h, c_void_p(out.data_ptr()), out.stride(0), c_void_p(tgt.data_ptr()), b, c, k0, k1,
                               c_void_p(hits.data_ptr()), eng._stream()), "af_topk_hits")
    res = hits * (100.0 / b)
    return [res[0:1], res[1:2]][: len(topk)] if len(topk) == 2 else [res[0:1]]


def cal_map(output, old_test_y):
    """output (N, C) fp32 CUDA logits, old_test_y (N, L) int64 labels (-1 = none) -> (mAP in %, per-class AP in %).

    Reproduces the reference including its label re-ranking (`get_multi_hot(..., assumes_starts_zero=False)`,
    ACT/ops/utils.py:51-66: labels are replaced by their rank among the distinct non-negative labels present)."""
    eng = get_engine(output.device)
    out = output.float().contiguous()
    n, c = out.shape
    y = old_test_y.to(device=output.device, dtype=torch.int64).reshape(n, -1).contiguous()
    uniq = torch.unique(y[y >= 0])
    remapped = torch.where(y >= 0, torch.searchsorted(uniq, y.clamp(min=0)), y)
    # the reference writes labels into a (classes + 1)-wide table and drops the last column: labels == C vanish
    remapped = torch.where(remapped >= c, torch.full_like(remapped, -1), remapped).contiguous()
    probs = torch.empty(n, c, dtype=torch.float32, device=output.device)
    ap = torch.empty(c, dtype=torch.float32, device=output.device)
    check(eng.lib.af_softmax_rows(eng.h, c_void_p(out.data_ptr()), out.stride(0), c_void_p(probs.data_ptr()), n, c,
                                  eng._stream()), "af_softmax_rows")
    check(eng.lib.af_class_ap(eng.h, c_void_p(probs.data_ptr()), c_void_p(remapped.data_ptr()), n, c, remapped.shape[1],
                              c_void_p(ap.data_ptr()), eng._stream()), "af_class_ap")
    return ap.mean() * 100, ap * 100
