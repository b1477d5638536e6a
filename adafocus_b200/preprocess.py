"""Device-side validation transform for decoded frames: GroupScale -> GroupCenterCrop -> Stack
-> ToTorchFormatTensor -> GroupNormalize of ACT/main_dist.py:213-220 (ACT/ops/transforms.py:78-93, 37-43, 303-336,
64-77), bit-identical to the reference's PIL / torchvision path.

Decoding (JPEG -> uint8 frames, ACT/ops/dataset.py:82-136) stays on the host; the decoded frames of a batch cross PCIe
as bytes at their stored size and the rest runs on the GPU:

    af_resize_crop_u8   Pillow-exact bilinear resize (22-bit fixed point, two passes) restricted to the centre crop
    af_frames_u8_to_f32 /255, mean / std, frame by frame -> (B, 3T, crop, crop) fp32, the model's input contract

The resampling windows and fixed-point weights depend only on (source size, scale size, crop size); they are computed
on the host exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc (double precision, round half away from
zero) and cached per geometry."""
import functools
import math

import numpy as np
import torch

from .engine import _ptr, check, get_engine

PRECISION_BITS = 32 - 8 - 2


def resized_output_size(h, w, size):
    """torchvision Resize(size:int): smaller edge -> size, the other int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)     # (new_h, new_w)


def center_crop_origin(h, w, th, tw):
    """torchvision CenterCrop: int(round((h - th) / 2.0)) with round-half-to-even."""
    return int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))


def _coeffs(in_size, out_size):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * inv)
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            p = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + p * (1 << PRECISION_BITS)) if p < 0 else int(0.5 + p * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


@functools.lru_cache(maxsize=64)
def resize_tables(h, w, scale_size, crop_size):
    """Host tables for one geometry: dict(hb, hk, vb, vk (numpy int32), row0, rows, oh, ow, y0, x0)."""
    oh, ow = resized_output_size(h, w, scale_size)
    if crop_size > oh or crop_size > ow:
        raise ValueError(f"crop {crop_size} larger than the scaled frame {oh}x{ow}")
    y0, x0 = center_crop_origin(oh, ow, crop_size, crop_size)
    hb, hk = _coeffs(w, ow)
    vb, vk = _coeffs(h, oh)
    hb, hk = hb[x0:x0 + crop_size].copy(), hk[x0:x0 + crop_size].copy()
    vb, vk = vb[y0:y0 + crop_size].copy(), vk[y0:y0 + crop_size].copy()
    row0 = int(vb[:, 0].min())
    rows = int((vb[:, 0] + vb[:, 1]).max()) - row0
    return dict(hb=hb, hk=hk, vb=vb, vk=vk, row0=row0, rows=rows, oh=oh, ow=ow, y0=y0, x0=x0)


class FramePreprocessor:
    """GroupScale(scale_size) + GroupCenterCrop(crop_size) + Stack + ToTorchFormatTensor + GroupNormalize on the device.

    frames: uint8 CUDA tensor (B*T, H, W, 3), the decoded frames of B clips (T consecutive frames per clip).
    __call__ -> fp32 (B, 3T, crop, crop); `cropped_u8` -> the uint8 (B*T, crop, crop, 3) frames after scale + crop."""

    def __init__(self, scale_size, crop_size, mean, std, device):
        self.scale_size, self.crop_size = int(scale_size), int(crop_size)
        self.mean, self.std = list(mean), list(std)
        self.device = torch.device(device)
        self.eng = get_engine(self.device)
        self._dev_tables = {}
        self._tmp = None

    def _tables(self, h, w):
        key = (h, w)
        if key not in self._dev_tables:
            t = resize_tables(h, w, self.scale_size, self.crop_size)
            dev = {k: torch.from_numpy(t[k]).to(self.device) for k in ("hb", "hk", "vb", "vk")}
            self._dev_tables[key] = (t, dev)
        return self._dev_tables[key]

    def cropped_u8(self, frames, out=None):
        if frames.dtype != torch.uint8 or not frames.is_cuda or not frames.is_contiguous() or frames.dim() != 4:
            raise ValueError("frames must be a contiguous uint8 CUDA tensor (N, H, W, C)")
        n, h, w, c = frames.shape
        t, dev = self._tables(h, w)
        cs = self.crop_size
        if out is None:
            out = torch.empty(n, cs, cs, c, dtype=torch.uint8, device=self.device)
        need = n * t["rows"] * cs * c
        if self._tmp is None or self._tmp.numel() < need:
            self._tmp = torch.empty(need, dtype=torch.uint8, device=self.device)     # horizontal-pass scratch, reused
        eng = self.eng
        check(eng.lib.af_resize_crop_u8(eng.h, _ptr(frames), _ptr(self._tmp), _ptr(out), n, h, w, c, _ptr(dev["hb"]),
                                        _ptr(dev["hk"]), t["hk"].shape[1], cs, _ptr(dev["vb"]), _ptr(dev["vk"]),
                                        t["vk"].shape[1], cs, t["row0"], t["rows"], eng._stream()),
              "af_resize_crop_u8")
        eng._count()
        return out

    def __call__(self, frames, frames_per_clip, out=None, u8_out=None):
        n = frames.shape[0]
        if n % frames_per_clip:
            raise ValueError("the number of frames is not a multiple of frames_per_clip")
        u8 = self.cropped_u8(frames, out=u8_out)                              # (B*T, crop, crop, 3)
        cs = self.crop_size
        if out is None:
            out = torch.empty(n // frames_per_clip, frames_per_clip * frames.shape[3], cs, cs, dtype=torch.float32,
                              device=self.device)
        # per-frame ingest: (N, 3, HW) fp32 is exactly the (B, 3T, H, W) layout of Stack() + ToTorchFormatTensor
        self.eng.frames_u8_to_f32(u8, self.mean, self.std, out=out)
        return out
