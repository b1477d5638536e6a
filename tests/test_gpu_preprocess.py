"""f-2: the reference's validation transform chain on the device -- GroupScale -> GroupCenterCrop -> Stack ->
ToTorchFormatTensor -> GroupNormalize (ACT/main_dist.py:213-220) -- bit-exact against the oracle and against the outputs
of the reference's own classes (tests/golden/transforms.npz, real Pillow + torchvision)."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _cases():
    spec = importlib.util.spec_from_file_location("mgt", os.path.join(GOLDEN, "make_golden_transforms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_device_transform_chain_bit_exact_vs_reference_golden():
    from adafocus_b200.preprocess import FramePreprocessor
    from oracle import pil_transforms as pt
    mod = _cases()
    gold = np.load(os.path.join(GOLDEN, "transforms.npz"))
    for tag, n, h, w, scale, crop in mod.CASES:
        frames = mod.synthetic_frames(tag, n, h, w)
        pre = FramePreprocessor(scale, crop, MEAN, STD, DEV)
        fd = torch.from_numpy(frames).to(DEV)
        # (a) per-frame resize + crop == the reference's GroupScale + GroupCenterCrop
        u8 = pre.cropped_u8(fd).cpu().numpy()
        assert tuple(u8.shape) == tuple(gold[f"{tag}_shape"])
        assert np.array_equal(u8, pt.group_scale_center_crop(frames, scale, crop))
        assert hashlib.sha256(u8.tobytes()).digest() == gold[f"{tag}_sha256"].tobytes(), tag
        # (b) the whole chain for one clip of n frames == Stack + ToTorchFormatTensor + GroupNormalize
        x = pre(fd, n)
        assert x.shape == (1, 3 * n, crop, crop) and x.dtype == torch.float32
        assert hashlib.sha256(x[0].cpu().numpy().tobytes()).digest() == gold[f"{tag}_tensor_sha256"].tobytes(), tag


def test_device_transform_batched_clips_and_full_size():
    """64 clips x 16 frames at the usual 340x256 frame size (the bench's e2e shape): every clip equals the oracle on a
    sample, Stack() order is (clip, y, x, frame*3 + rgb), and the result is deterministic."""
    from adafocus_b200.preprocess import FramePreprocessor
    from oracle import pil_transforms as pt
    g = torch.Generator(device=DEV).manual_seed(9)
    b, t, h, w = 64, 16, 256, 340
    frames = torch.randint(0, 256, (b * t, h, w, 3), dtype=torch.uint8, device=DEV, generator=g)
    pre = FramePreprocessor(256, 224, MEAN, STD, DEV)
    u8 = pre.cropped_u8(frames)
    assert u8.shape == (b * t, 224, 224, 3)
    for clip, fr in ((0, 0), (17, 5), (63, 15)):
        want = pt.group_scale_center_crop([frames[clip * t + fr].cpu().numpy()], 256, 224)[0]
        assert np.array_equal(u8[clip * t + fr].cpu().numpy(), want)
    assert torch.equal(u8, pre.cropped_u8(frames))
    x = pre(frames, t)
    assert x.shape == (b, 3 * t, 224, 224)
    ref = pt.stack_to_tensor_normalize(u8[3 * t:4 * t].cpu().numpy(), MEAN, STD)       # Stack() order of clip 3
    assert np.array_equal(x[3].cpu().numpy(), ref)


def test_device_transform_rejects_bad_arguments():
    from adafocus_b200.preprocess import FramePreprocessor
    pre = FramePreprocessor(256, 224, MEAN, STD, DEV)
    frames = torch.zeros(3, 256, 340, 3, dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        pre(frames, 2)                                  # 3 frames are not whole clips of 2
    with pytest.raises(ValueError):
        pre.cropped_u8(frames.float())
    with pytest.raises(ValueError):
        FramePreprocessor(64, 224, MEAN, STD, DEV).cropped_u8(frames)      # crop larger than the scaled frame
