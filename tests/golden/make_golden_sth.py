"""Golden vectors for the Something-Something tree: runs the REFERENCE's own STH classes (CPU fp32, imported from
/root/reference) on the seeded synthetic checkpoint of adafocus_b200.synth, driven exactly like STH/evaluate.py
(strip fc :83, eval() on model + both policies :175-177, glance :195, action_stage2 :199, also action_stage3).
Run in the build container only:  python tests/golden/make_golden_sth.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STH = "/root/reference/Experiments on Something-Something V1&V2"
sys.path.insert(0, ROOT)

from adafocus_b200 import synth  # noqa: E402


def import_reference_sth():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, STH)
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "ops")]:
        del sys.modules[m]
    import models.gfv_net as ref
    return ref


def run(ref, args, batch, tag, seed_rand=123):
    torch.manual_seed(0)
    model = ref.GFV(args)
    model.focuser.net.base_model = torch.nn.Sequential(*list(model.focuser.net.base_model.children())[:-1])
    ck = synth.synth_checkpoint_sth(model, synth.SEED)
    synth.load_checkpoint_sth(model, ck)
    model.eval()
    model.focuser.policy.policy.eval()
    model.focuser.policy.policy_old.eval()
    tg, tf = args.num_segments_glancer, args.num_segments_focuser
    gi = synth.synth_clips(batch, tg, 224, synth.SEED + 1)
    fi = synth.synth_clips(batch, tf, 224, synth.SEED + 2)
    rec = {"actions": []}
    old_act = model.focuser.policy.policy_old.act

    def act(*a, **k):
        out = old_act(*a, **k)
        rec["actions"].append(out.clone())
        return out
    model.focuser.policy.policy_old.act = act
    with torch.no_grad():
        gin = torch.nn.functional.interpolate(gi, (args.glance_size, args.glance_size))
        fmap, glogit = model.glance(gin)
        fimg = fi.view(batch, tf, 3, 224, 224)
        preds3, preds2, bases = [], [], []
        lp = None
        for step in range(args.video_div):
            pred, lp = model.action_stage3(fimg, fmap, glogit, step, args, prev_local_patch=lp)
            preds3.append(pred)
        actions3 = torch.stack(rec["actions"], 1)
        rec["actions"].clear()
        torch.manual_seed(seed_rand)                  # CPU generator feeds random_patching (STH/models/gfv_net.py:424)
        rand_draws = []
        real_rand = torch.rand

        def recording_rand(*a, **k):                  # record the (B,2) draws random_patching actually used (under the
            out = real_rand(*a, **k)                  # CPU shim the policy's discarded dist.sample() shares this RNG)
            rand_draws.append(out.clone())
            return out
        torch.rand = recording_rand
        lp2 = None
        try:
            for step in range(args.video_div):
                pred, base, lp2 = model.action_stage2(fimg, fmap, glogit, step, args, prev_local_patch=lp2,
                                                      training=False)
                preds2.append(pred)
                bases.append(base)
        finally:
            torch.rand = real_rand
    out = {
        "fmap_sample": fmap[:, :, ::64].numpy().astype(np.float32),
        "glogit": glogit.numpy().astype(np.float32),
        "actions": actions3.numpy().astype(np.float32),                      # (B, video_div, 2)
        "coords": torch.floor(actions3 * (224 - args.patch_size)).int().numpy(),
        "pred_stage3": torch.stack(preds3, 0).numpy().astype(np.float32),    # (video_div, B, C)
        "pred_stage2": torch.stack(preds2, 0).numpy().astype(np.float32),
        "baseline_stage2": torch.stack(bases, 0).numpy().astype(np.float32),
        "rand_draws": torch.stack(rand_draws, 0).numpy().astype(np.float32),
        "patch_checksum": lp.double().sum(dim=(2, 3, 4)).numpy(),
        "patch_shape": np.array(lp.shape),
        "seed_rand": np.array(seed_rand),
    }
    path = os.path.join(HERE, f"sth_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    return model


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = import_reference_sth()
    m = run(ref, synth.sth_args(), 2, "r50_p144_b2")
    keys = {"glancer": m.glancer.state_dict(), "focuser": m.focuser.state_dict(), "fc": m.classifier.state_dict(),
            "policy": m.focuser.policy.policy.state_dict(), "model": m.state_dict()}
    with open(os.path.join(HERE, "ref_state_keys_sth.json"), "w") as f:
        json.dump({part: [(k, list(v.shape)) for k, v in sd.items()] for part, sd in keys.items()}, f)
    # second configuration: two video divisions, 96^2 patches, discrete-free continuous policy without BN, 8+8 frames
    run(ref, synth.sth_args(video_div=2, num_segments_focuser=8, patch_size=96, actorcritic_with_bn=False,
                            num_classes=40), 4, "div2_p96_b4")
