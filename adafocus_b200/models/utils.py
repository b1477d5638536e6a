"""Patch cropping -- mirror of ACT/models/utils.py:37-51 (= STH/models/utils.py:44-58)."""
import torch

from ..engine import get_engine


def get_patch(images, action_sequence, patch_size):
    """images (N,C,H,W) fp32 CUDA, action_sequence (N,2) fp32 in [0,1] -> (N,C,P,P) fp32.

    Same result as the reference's per-sample slicing loop -- coordinates are floor(action * (H - P)) evaluated in
    fp32, rows first -- but as one gather kernel with no device->host synchronisation (the reference does 4
    `.item()` syncs per sample)."""
    if not images.is_cuda:
        raise RuntimeError("adafocus_b200.get_patch needs CUDA tensors (no CPU path)")
    eng = get_engine(images.device)
    images = images.contiguous()
    if images.dtype != torch.float32:
        raise TypeError("get_patch expects fp32 images like the reference pipeline")
    action = action_sequence.to(device=images.device, dtype=torch.float32).contiguous()
    return eng.crop(images, action=action, patch=int(patch_size))


def random_crop(im, size, pad_size=0):
    raise NotImplementedError("random_crop (numpy host RNG) belongs to stage-1/2 training, outside the hot path")
