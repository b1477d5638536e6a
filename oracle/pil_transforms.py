"""ORACLE -- test infrastructure only.  Integer restatement of the decoded-frame transforms the reference applies in front
of the model at validation time (ACT/main_dist.py:213-220): GroupScale -> GroupCenterCrop -> Stack ->
ToTorchFormatTensor -> GroupNormalize (ACT/ops/transforms.py:37-93, 303-336).

The arithmetic of GroupScale lives in a third-party dependency that is not under /root/reference: torchvision's
`Resize` on PIL images calls `PIL.Image.resize(..., BILINEAR)`, i.e. Pillow's libImaging/Resample.c (Pillow 12.2.0 and
torchvision 0.26.0 in this image; the reference pins neither).  Its published algorithm, restated here:

  * per axis, per output index xx: center = (xx + 0.5) * scale, support = max(scale, 1), window
    [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image, triangle weights
    1 - |x + 0.5 - center| / max(scale, 1), normalised to sum 1 in double precision (precompute_coeffs);
  * weights -> fixed point with 22 fractional bits, round half away from zero (normalize_coeffs_8bpc);
  * horizontal pass into a uint8 image, then vertical pass: acc = 2^21 + sum(pixel * k) >> 22, clipped to [0, 255]
    (ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc).

Pinned by tests/test_oracle.py against tests/golden/transforms.npz, which tests/golden/make_golden_transforms.py produced
by running the REFERENCE's own GroupScale / GroupCenterCrop classes (real Pillow + torchvision), and -- where Pillow is
importable -- directly against Pillow."""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_output_size(h, w, size):
    """torchvision Resize(size:int) on an (h, w) image: the smaller edge becomes `size`
    (torchvision.transforms.functional._compute_resized_output_size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def center_crop_origin(h, w, th, tw):
    """torchvision CenterCrop: int(round(.)) with Python's round-half-to-even."""
    return int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))


def bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the triangle filter.  -> (bounds int32 [out,2] = (first input
    index, tap count), kk int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0 else v
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            p = w[x]
            kk[xx, x] = int(-0.5 + p * (1 << PRECISION_BITS)) if p < 0 else int(0.5 + p * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img, bounds, kk, axis):
    """One pass of ImagingResample*_8bpc along `axis` of a uint8 (H, W, C) image."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.int64)
    for i in range(bounds.shape[0]):
        x0, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[x0 + t] * int(kk[i, t])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """PIL.Image.resize((out_w, out_h), BILINEAR) of a uint8 (H, W, C) array: horizontal pass, then vertical pass (each
    skipped by Pillow when the size does not change -- the identity tables give the same bytes)."""
    h, w = img.shape[:2]
    if out_w != w:
        img = _resample_axis(img, *bilinear_coeffs(w, out_w), axis=1)
    if out_h != h:
        img = _resample_axis(img, *bilinear_coeffs(h, out_h), axis=0)
    return img


def group_scale_center_crop(frames, scale_size, crop_size):
    """GroupScale(scale_size) + GroupCenterCrop(crop_size) on a list / array of uint8 (H, W, 3) frames
    (ACT/ops/transforms.py:37-43, 78-93) -> uint8 (N, crop, crop, 3)."""
    out = []
    for f in frames:
        h, w = f.shape[:2]
        oh, ow = resized_output_size(h, w, scale_size)
        r = resize_bilinear_u8(f, oh, ow)
        y0, x0 = center_crop_origin(oh, ow, crop_size, crop_size)
        out.append(r[y0:y0 + crop_size, x0:x0 + crop_size])
    return np.stack(out, 0)


def stack_to_tensor_normalize(frames_u8, mean, std):
    """Stack(roll=False) + ToTorchFormatTensor(div=True) + GroupNormalize (ACT/ops/transforms.py:303-336, 64-77) on the
    (T, H, W, 3) uint8 frames of ONE clip -> float32 (3T, H, W)."""
    t, h, w, _ = frames_u8.shape
    x = np.concatenate(list(frames_u8), axis=2)                    # (H, W, 3T), np.concatenate(img_group, axis=2)
    x = np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
    for c in range(3 * t):
        x[c] = (x[c] - np.float32(mean[c % 3])) / np.float32(std[c % 3])
    return x
