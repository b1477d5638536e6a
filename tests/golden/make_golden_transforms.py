"""Golden vectors for the decoded-frame transforms: runs the REFERENCE's own GroupScale / GroupCenterCrop / Stack /
ToTorchFormatTensor / GroupNormalize classes (ACT/ops/transforms.py, real Pillow + torchvision) on seeded synthetic
frames.  Build container only:  python tests/golden/make_golden_transforms.py  -> tests/golden/transforms.npz
The inputs are regenerated from the seed by `synthetic_frames` below (imported by the tests), only outputs are stored."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# (tag, frames, H, W, scale_size, crop_size)
CASES = [
    ("actnet_340x256", 4, 256, 340, 256, 224),      # the usual extracted-frame size (ACT/ops/video_jpg.py: height 256)
    ("portrait_240x320", 2, 320, 240, 256, 224),
    ("upscale_224", 2, 224, 224, 256, 224),         # smaller than the scale size: upsampling branch (support 1)
    ("small_131x97", 3, 97, 131, 64, 56),           # odd sizes, strong down-scaling (5-tap windows)
    ("big_480x360", 2, 360, 480, 128, 112),
]


def synthetic_frames(tag, n, h, w):
    seed = int(hashlib.sha256(tag.encode()).hexdigest()[:8], 16)
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (n, (h + 15) // 16, (w + 15) // 16, 3)).astype(np.float32)
    up = np.repeat(np.repeat(coarse, 16, axis=1), 16, axis=2)[:, :h, :w]
    noise = rng.integers(-40, 41, (n, h, w, 3))
    return np.clip(up + noise, 0, 255).astype(np.uint8)


def main():
    from PIL import Image
    import torch
    from oracle import reference_loader as rl
    rl.import_tree("ACT")
    from ops.transforms import GroupCenterCrop, GroupNormalize, GroupScale, Stack, ToTorchFormatTensor
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = {}
    for tag, n, h, w, scale, crop in CASES:
        frames = synthetic_frames(tag, n, h, w)
        imgs = [Image.fromarray(f) for f in frames]
        cropped = GroupCenterCrop(crop)(GroupScale(scale)(imgs))
        u8 = np.stack([np.asarray(im) for im in cropped], 0)                         # (n, crop, crop, 3)
        tensor = GroupNormalize(mean, std)(ToTorchFormatTensor(div=True)(Stack(roll=False)(cropped)))
        out[f"{tag}_sha256"] = np.frombuffer(hashlib.sha256(u8.tobytes()).digest(), dtype=np.uint8)
        out[f"{tag}_tensor_sha256"] = np.frombuffer(hashlib.sha256(tensor.numpy().tobytes()).digest(), dtype=np.uint8)
        out[f"{tag}_shape"] = np.array(u8.shape)
        if tag in ("actnet_340x256", "small_131x97"):
            out[f"{tag}_u8"] = u8[:1]                                                 # one full frame for debugging
        print(tag, u8.shape, tuple(tensor.shape))
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **out)
    print("wrote transforms.npz")


if __name__ == "__main__":
    main()
