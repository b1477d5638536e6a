"""Something-Something tree on the B200: the CUDA path driven exactly like STH/evaluate.py (strip fc, eval(), glance,
action_stage2 / action_stage3) and the fused forward_eval plan, against the CPU oracle and the reference golden vectors.
Bars: crop coordinates / cropped bytes exact; continuous actions within 2e-3 (they feed a floor()); logits as in
tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STH_TOL = 1e-3

CASES = [
    ("r50_p144_b2", dict(), 2),
    ("div2_p96_b4", dict(video_div=2, num_segments_focuser=8, patch_size=96, actorcritic_with_bn=False,
                         num_classes=40), 4),
]


def _build(over):
    from adafocus_b200 import synth
    from adafocus_b200.models_sth.gfv_net import GFV
    args = synth.sth_args(**over)
    model = GFV(args).to(DEV)
    synth.strip_fc_sth(model)                                # STH/evaluate.py:83
    ck = synth.synth_checkpoint_sth(model, synth.SEED)
    synth.load_checkpoint_sth(model, ck)
    model.focuser.policy.policy.to(DEV)
    model.focuser.policy.policy_old.to(DEV)
    assert model.eval() is None
    model.focuser.policy.policy.eval()
    model.focuser.policy.policy_old.eval()
    return args, model, ck


@pytest.mark.parametrize("tag,over,batch", CASES)
def test_sth_reference_call_pattern(golden_dir, tag, over, batch):
    from adafocus_b200 import synth
    from oracle import adafocus_oracle as orc
    gold = np.load(os.path.join(golden_dir, f"sth_{tag}.npz"))
    args, model, ck = _build(over)
    tg, tf, p = args.num_segments_glancer, args.num_segments_focuser, args.patch_size
    gi = synth.synth_clips(batch, tg, 224, synth.SEED + 1)
    fi = synth.synth_clips(batch, tf, 224, synth.SEED + 2)
    ref = orc.sth_forward(gi, fi, ck, p, tg, tf, args.video_div, args.shift_div, rand_actions=gold["rand_draws"])
    gid, fid = gi.to(DEV), fi.to(DEV)
    fimg = fid.view(batch, tf, 3, 224, 224)
    with torch.no_grad():
        fmap, glogit = model.glance(gid)
        assert fmap.shape == (batch, tg, 1280, 7, 7) and glogit.shape == (batch, tg, args.num_classes)
        scale = max(1.0, float(ref["glogit"].abs().max()))
        assert float((glogit.cpu() - ref["glogit"]).abs().max()) <= STH_TOL * scale
        # the two loops share focuser.memory (the GRU state), so run them one after the other like make_golden_sth.py
        lp = None
        for step in range(args.video_div):
            pred3, lp = model.action_stage3(fimg, fmap, glogit, step, args, prev_local_patch=lp)
            s = max(1.0, float(np.abs(gold["pred_stage3"][step]).max()))
            assert np.abs(pred3.cpu().numpy() - gold["pred_stage3"][step]).max() <= STH_TOL * s
            assert float((pred3.cpu() - ref["preds"][step]).abs().max()) <= STH_TOL * s
        lp2 = None
        for step in range(args.video_div):
            draws = torch.from_numpy(gold["rand_draws"][step])      # replay the reference's CPU draws for the baseline
            real_rand = torch.rand
            torch.rand = lambda *a, **k: draws.clone()
            try:
                pred2, base, lp2 = model.action_stage2(fimg, fmap, glogit, step, args, prev_local_patch=lp2,
                                                       training=False)
            finally:
                torch.rand = real_rand
            s = max(1.0, float(np.abs(gold["pred_stage2"][step]).max()))
            assert np.abs(pred2.cpu().numpy() - gold["pred_stage2"][step]).max() <= STH_TOL * s
            assert np.abs(base.cpu().numpy() - gold["baseline_stage2"][step]).max() <= STH_TOL * s
        assert torch.equal(lp, lp2)
        assert list(lp.shape) == gold["patch_shape"].tolist()
        # cropped bytes: exact (same coordinates -> same patch checksum as the reference)
        assert np.allclose(lp.double().sum(dim=(2, 3, 4)).cpu().numpy(), gold["patch_checksum"], rtol=0, atol=1e-9)
        assert np.array_equal(pred3.argmax(1).cpu().numpy(), gold["pred_stage3"][-1].argmax(1))


@pytest.mark.parametrize("tag,over,batch", CASES)
def test_sth_fused_plan(golden_dir, tag, over, batch):
    from adafocus_b200 import synth
    gold = np.load(os.path.join(golden_dir, f"sth_{tag}.npz"))
    args, model, ck = _build(over)
    gi = synth.synth_clips(batch, args.num_segments_glancer, 224, synth.SEED + 1).to(DEV)
    fi = synth.synth_clips(batch, args.num_segments_focuser, 224, synth.SEED + 2).to(DEV)
    pred = model.forward_eval(gi, fi, args)
    plan = model.last_plan
    assert np.array_equal(plan.yx.view(batch, args.video_div, 2).cpu().numpy(), gold["coords"])
    assert np.abs(plan.action.view(batch, args.video_div, 2).cpu().numpy() - gold["actions"]).max() <= 2e-3
    s = max(1.0, float(np.abs(gold["pred_stage3"][-1]).max()))
    assert np.abs(pred.cpu().numpy() - gold["pred_stage3"][-1]).max() <= STH_TOL * s
    assert np.array_equal(pred.argmax(1).cpu().numpy(), gold["pred_stage3"][-1].argmax(1))
    pred2 = model.forward_eval(gi.clone(), fi.clone(), args)
    assert torch.equal(pred, pred2)


def test_sth_fused_plan_row_streaming_blocks(golden_dir, monkeypatch):
    """Same golden comparison with the glancer's MobileNet-V2 blocks forced through af_mbconv_rows, the kernel that
    bench-size batches take (with the temporal shift applied to the block input, the residual un-shifted)."""
    import adafocus_b200.models.mobilenet as mb
    from adafocus_b200 import synth
    monkeypatch.setattr(mb, "_ROWS_MODE", "force")
    tag, over, batch = CASES[0]
    gold = np.load(os.path.join(golden_dir, f"sth_{tag}.npz"))
    args, model, ck = _build(over)
    gi = synth.synth_clips(batch, args.num_segments_glancer, 224, synth.SEED + 1).to(DEV)
    fi = synth.synth_clips(batch, args.num_segments_focuser, 224, synth.SEED + 2).to(DEV)
    pred = model.forward_eval(gi, fi, args)
    plan = model.last_plan
    assert np.array_equal(plan.yx.view(batch, args.video_div, 2).cpu().numpy(), gold["coords"])
    assert np.abs(plan.action.view(batch, args.video_div, 2).cpu().numpy() - gold["actions"]).max() <= 2e-3
    s = max(1.0, float(np.abs(gold["pred_stage3"][-1]).max()))
    assert np.abs(pred.cpu().numpy() - gold["pred_stage3"][-1]).max() <= STH_TOL * s
    assert np.array_equal(pred.argmax(1).cpu().numpy(), gold["pred_stage3"][-1].argmax(1))


def test_temporal_shift_public_api_bit_exact():
    from adafocus_b200.models_sth.temporal_shift import TemporalShift
    from oracle import adafocus_oracle as orc
    x = torch.randn(24, 64, 9, 9)
    out = TemporalShift.shift(x.to(DEV), 12, fold_div=8)
    assert torch.equal(out.cpu(), orc.temporal_shift(x, 12, 8))


def test_resnet101_tsm_runs():
    """cfg5 architecture (base_model=resnet101, T_f=12, P=144): shapes + determinism + agreement with the oracle."""
    from adafocus_b200 import synth
    from oracle import adafocus_oracle as orc
    args, model, ck = _build(dict(base_model="resnet101"))
    gi = synth.synth_clips(1, 8, 224, synth.SEED + 1)
    fi = synth.synth_clips(1, 12, 224, synth.SEED + 2)
    ref = orc.sth_forward(gi, fi, ck, 144, 8, 12, 1, 8, layers=(3, 4, 23, 3))
    pred = model.forward_eval(gi.to(DEV), fi.to(DEV), args)
    s = max(1.0, float(ref["pred"].abs().max()))
    assert float((pred.cpu() - ref["pred"]).abs().max()) <= STH_TOL * s
    assert np.array_equal(model.last_plan.yx.view(1, 1, 2).cpu().numpy(), ref["coords"])
