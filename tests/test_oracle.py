"""The oracle (oracle/adafocus_oracle.py) against golden vectors produced by the reference's own classes
(tests/golden/make_golden.py) and against the get_patch known-answer table.  CPU only."""
import os

import numpy as np
import pytest
import torch

from adafocus_b200 import synth
from adafocus_b200.models.gfv_net import GFV
from oracle import adafocus_oracle as orc

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ck(args):
    model = GFV(args)
    return synth.synth_checkpoint_act(model, synth.SEED)


def test_get_patch_kat_table(golden_dir):
    kat = np.load(os.path.join(golden_dir, "get_patch_kat.npz"))
    # coordinates of the shipped grids
    expect = {(7, 128): [0, 16, 32, 48, 64, 80, 96], (7, 96): [0, 21, 42, 64, 85, 106, 128],
              (7, 160): [0, 10, 21, 32, 42, 53, 64], (7, 192): [0, 5, 10, 16, 21, 26, 32],
              (5, 144): [0, 20, 40, 60, 80]}
    for (n, p), vals in expect.items():
        grid = orc.standard_actions(n * n)
        c = orc.patch_coordinates(grid, 224, p)
        assert sorted(set(c[:, 0].tolist())) == vals
        assert np.array_equal(c, kat[f"grid{n}_p{p}"])
    for n, p in ((6, 112), (8, 176)):
        assert np.array_equal(orc.patch_coordinates(orc.standard_actions(n * n), 224, p), kat[f"grid{n}_p{p}"])
    assert orc.patch_coordinates(np.array([[0.5, 1.0]], np.float32), 224, 128).tolist() == [[48, 96]]


def test_get_patch_matches_reference_outputs(golden_dir):
    kat = np.load(os.path.join(golden_dir, "get_patch_kat.npz"))
    g = torch.Generator().manual_seed(5)
    img = torch.randn(4, 3, 224, 224, generator=g).numpy()
    acts = kat["rand_actions"]
    for p in (96, 128, 144, 130):
        out = orc.get_patch(img, acts, p)
        assert out.shape == (4, 3, p, p)
        assert np.array_equal(orc.patch_coordinates(acts, 224, p), kat[f"rand_p{p}_coords"])
        assert np.array_equal(out[:, :, 0, :4], kat[f"rand_p{p}_first"])
        assert np.allclose(out.astype(np.float64).sum(axis=(1, 2, 3)), kat[f"rand_p{p}_sum"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag,over,batch", [
    ("t4_p96_b3", dict(num_segments=4, patch_size=96, action_dim=36, num_classes=51), 3),
    ("c3_b2", dict(), 2),
])
def test_act_forward_matches_reference(golden_dir, tag, over, batch):
    gold = np.load(os.path.join(golden_dir, f"act_{tag}.npz"))
    args = synth.act_args(**over)
    ck = _ck(args)
    x = synth.synth_clips(batch, args.num_segments, args.input_size, synth.SEED)
    assert bool(gold["scan_equals_input"])            # F.interpolate to glance_size=224 is the identity
    out = orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
    # integer / byte work: exact
    assert np.array_equal(out["actions"].numpy(), gold["actions"])
    assert np.array_equal(out["coords"], gold["coords"])
    assert np.array_equal(out["patches"][:, :, :, :2, :2].numpy(), gold["patch_corner"])
    assert np.allclose(out["patches"].double().sum(dim=(2, 3, 4)).numpy(), gold["patch_checksum"], rtol=0, atol=1e-9)
    # floating point: same fp32 math up to summation order (oneDNN kernels differ with batch shape / threads)
    np.testing.assert_allclose(out["gvec"].numpy(), gold["gvec"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["fmap"][:, :, ::64].numpy(), gold["fmap_sample"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["lfeat"].numpy(), gold["lfeat"], rtol=1e-4, atol=1e-5)
    # logits pass through a 16-step GRU written out gate by gate here vs torch's fused GRU in the reference
    np.testing.assert_allclose(out["logits"].numpy(), gold["logits"], rtol=1e-3, atol=3e-4)
    np.testing.assert_allclose(out["last_out"].numpy(), gold["last_out"], rtol=1e-3, atol=3e-4)
    assert np.array_equal(out["last_out"].argmax(1).numpy(), gold["last_out"].argmax(1))


# ------------------------------------------------------------------------------------------------ STH tree
STH_CASES = [
    ("r50_p144_b2", dict(), 2),
    ("div2_p96_b4", dict(video_div=2, num_segments_focuser=8, patch_size=96, actorcritic_with_bn=False,
                         num_classes=40), 4),
]


def _sth_ck(args):
    from adafocus_b200.models_sth.gfv_net import GFV as GFV_STH
    model = GFV_STH(args)
    synth.strip_fc_sth(model)
    return synth.synth_checkpoint_sth(model, synth.SEED)


@pytest.mark.parametrize("tag,over,batch", STH_CASES)
def test_sth_forward_matches_reference(golden_dir, tag, over, batch):
    gold = np.load(os.path.join(golden_dir, f"sth_{tag}.npz"))
    args = synth.sth_args(**over)
    ck = _sth_ck(args)
    gi = synth.synth_clips(batch, args.num_segments_glancer, 224, synth.SEED + 1)
    fi = synth.synth_clips(batch, args.num_segments_focuser, 224, synth.SEED + 2)
    out = orc.sth_forward(gi, fi, ck, args.patch_size, args.num_segments_glancer, args.num_segments_focuser,
                          args.video_div, args.shift_div, rand_actions=gold["rand_draws"])
    np.testing.assert_allclose(out["fmap"][:, :, ::64].numpy(), gold["fmap_sample"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["glogit"].numpy(), gold["glogit"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["actions"].numpy(), gold["actions"], rtol=0, atol=2e-6)
    assert np.array_equal(out["coords"], gold["coords"])
    assert np.allclose(out["patches"].double().sum(dim=(2, 3, 4)).numpy(), gold["patch_checksum"], rtol=0, atol=1e-9)
    for d in range(args.video_div):
        np.testing.assert_allclose(out["preds"][d].numpy(), gold["pred_stage3"][d], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(out["preds"][d].numpy(), gold["pred_stage2"][d], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(out["baselines"][d].numpy(), gold["baseline_stage2"][d], rtol=1e-4, atol=2e-5)
    assert np.array_equal(out["pred"].argmax(1).numpy(), gold["pred_stage3"][-1].argmax(1))


def test_temporal_shift_restatement():
    x = torch.arange(2 * 4 * 8 * 1 * 1, dtype=torch.float32).view(8, 8, 1, 1)
    y = orc.temporal_shift(x, 4, 8)
    assert torch.equal(y[0, 0], x[1, 0]) and float(y[3, 0]) == 0.0          # channel 0 comes from t+1, zero at the end
    assert float(y[0, 1]) == 0.0 and torch.equal(y[1, 1], x[0, 1])          # channel 1 comes from t-1, zero at the start
    assert torch.equal(y[:, 2:], x[:, 2:])
    assert float(y[4, 1]) == 0.0 and torch.equal(y[3, 0], torch.zeros(1, 1))   # clips do not leak into each other


def test_metrics_restatement_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "metrics.npz"))
    acc1, acc5 = orc.accuracy_topk(gold["logits"], gold["target"], (1, 5))
    assert abs(acc1 - float(gold["acc1"][0])) < 1e-4 and abs(acc5 - float(gold["acc5"][0])) < 1e-4
    m1, ap1 = orc.cal_map(gold["logits"], gold["target"].reshape(-1, 1))
    m2, ap2 = orc.cal_map(gold["logits"], gold["labels"])
    # two samples carry identical logits: torch.sort(descending=True) orders ties arbitrarily, the restatement stably
    np.testing.assert_allclose(ap1, gold["ap_single"], rtol=1e-4, atol=2e-2)
    np.testing.assert_allclose(ap2, gold["ap_multi"], rtol=1e-4, atol=2e-2)
    assert abs(m1 - float(gold["map_single"])) < 5e-3 and abs(m2 - float(gold["map_multi"])) < 5e-3


# ------------------------------------------------------------------------------------------------ decoded-frame transforms
def _transform_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_transforms",
                                                  os.path.join(GOLDEN, "make_golden_transforms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pil_transform_oracle_vs_reference_golden():
    """oracle/pil_transforms.py (integer restatement of Pillow's bilinear resampler + torchvision's size / crop rules +
    the Stack / ToTorchFormatTensor / GroupNormalize chain) against the outputs of the reference's own transform
    classes (tests/golden/transforms.npz): bit-exact, uint8 frames and fp32 tensors."""
    import hashlib
    from oracle import pil_transforms as pt
    mod = _transform_cases()
    gold = np.load(os.path.join(GOLDEN, "transforms.npz"))
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    for tag, n, h, w, scale, crop in mod.CASES:
        frames = mod.synthetic_frames(tag, n, h, w)
        u8 = pt.group_scale_center_crop(frames, scale, crop)
        assert tuple(u8.shape) == tuple(gold[f"{tag}_shape"])
        assert hashlib.sha256(u8.tobytes()).digest() == gold[f"{tag}_sha256"].tobytes(), tag
        if f"{tag}_u8" in gold:
            assert np.array_equal(u8[:1], gold[f"{tag}_u8"])
        x = pt.stack_to_tensor_normalize(u8, mean, std)
        assert hashlib.sha256(x.tobytes()).digest() == gold[f"{tag}_tensor_sha256"].tobytes(), tag


def test_pil_transform_oracle_vs_pillow_when_available():
    pil = pytest.importorskip("PIL.Image")
    tvt = pytest.importorskip("torchvision.transforms")
    from oracle import pil_transforms as pt
    rng = np.random.default_rng(3)
    for h, w, scale, crop in [(255, 341, 256, 224), (100, 60, 48, 40), (64, 64, 64, 64), (90, 200, 77, 70)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(tvt.CenterCrop(crop)(tvt.Resize(scale, pil.BILINEAR)(pil.fromarray(a))))
        assert np.array_equal(pt.group_scale_center_crop([a], scale, crop)[0], ref)


def test_host_resize_tables_match_the_oracle_coefficients():
    """adafocus_b200.preprocess.resize_tables (product, host side) == the oracle's coefficient tables restricted to the
    centre crop; the scratch row window covers every vertical tap."""
    from adafocus_b200 import preprocess as pp
    from oracle import pil_transforms as pt
    for h, w, scale, crop in [(256, 340, 256, 224), (320, 240, 256, 224), (224, 224, 256, 224), (97, 131, 64, 56)]:
        t = pp.resize_tables(h, w, scale, crop)
        oh, ow = pt.resized_output_size(h, w, scale)
        assert (t["oh"], t["ow"]) == (oh, ow)
        y0, x0 = pt.center_crop_origin(oh, ow, crop, crop)
        assert (t["y0"], t["x0"]) == (y0, x0)
        hb, hk = pt.bilinear_coeffs(w, ow)
        vb, vk = pt.bilinear_coeffs(h, oh)
        assert np.array_equal(t["hb"], hb[x0:x0 + crop]) and np.array_equal(t["hk"], hk[x0:x0 + crop])
        assert np.array_equal(t["vb"], vb[y0:y0 + crop]) and np.array_equal(t["vk"], vk[y0:y0 + crop])
        assert t["row0"] <= int(t["vb"][:, 0].min()) and t["row0"] + t["rows"] >= int((t["vb"][:, 0] + t["vb"][:, 1]).max())
        assert t["row0"] >= 0 and t["row0"] + t["rows"] <= h
