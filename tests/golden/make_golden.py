"""Generate golden vectors by running the REFERENCE's own classes (imported from /root/reference, CPU fp32) on the
seeded synthetic checkpoint + inputs of adafocus_b200.synth.  Run in the build container only (the GPU box has no
/root/reference); the resulting tests/golden/*.npz are committed.

    python tests/golden/make_golden.py

Shims (all outside the read-only reference tree), as listed in SURVEY.md section 8(c):
  * mobilenet_v2 / resnet50 constructed with pretrained=False (the constructors otherwise download weights),
  * Tensor.cuda / Module.cuda -> identity (the reference hard-codes .cuda()),
  * args = SimpleNamespace with the YAML values.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ACT = "/root/reference/Experiments on ActivityNet, FCVID and Mini-Kinetics"
sys.path.insert(0, ROOT)

from adafocus_b200 import synth  # noqa: E402


def import_reference_act():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, ACT)
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "ops" or k.startswith("ops.")]:
        del sys.modules[m]
    import models.gfv_net as ref
    import models.mobilenet as ref_mb
    import models.resnet as ref_rn
    ref.mobilenet_v2 = lambda pretrained=True, **kw: ref_mb.mobilenet_v2(pretrained=False, **kw)
    ref.resnet50 = lambda pretrained=True, **kw: ref_rn.resnet50(pretrained=False, **kw)
    return ref


def run_reference(ref, args, batch, tag):
    torch.manual_seed(0)
    model = ref.GFV(args)
    ck = synth.synth_checkpoint_act(model, synth.SEED)
    # the reference's loading sequence, ACT/main_dist.py:100-110
    model.glancer.load_state_dict(ck["glancer"])
    model.focuser.load_state_dict(ck["focuser"], strict=False)
    model.classifier.load_state_dict(ck["fc"])
    model.focuser.policy.policy.load_state_dict(ck["policy"])
    model.focuser.policy.policy_old.load_state_dict(ck["policy"])
    model.eval()

    rec = {"actions": [], "std_actions": [], "patches": [], "lfeat": []}
    old_act = model.focuser.policy.policy_old.act

    def act(*a, **k):
        out = old_act(*a, **k)
        rec["actions"].append(out.clone())
        return out
    model.focuser.policy.policy_old.act = act
    old_gp = ref.get_patch

    def gp(images, action_sequence, patch_size):
        out = old_gp(images, action_sequence, patch_size)
        rec["std_actions"].append(action_sequence.clone())
        rec["patches"].append(out.clone())
        return out
    ref.get_patch = gp
    old_fm = model.focuser.net.get_featmap

    def fm(x, pooled=True):
        out = old_fm(x, pooled)
        rec["lfeat"].append(out.flatten(1).clone())
        return out
    model.focuser.net.get_featmap = fm

    x = synth.synth_clips(batch, args.num_segments, args.input_size, synth.SEED)
    scan = torch.nn.functional.interpolate(x, (args.glance_size, args.glance_size))     # ACT/main_dist.py:332
    with torch.no_grad():
        fmap, gvec = model.glance(scan)
        logits, last_out = model(input=x, scan=scan, training=False, backbone_pred=False, one_step=True, gpu=None)
    ref.get_patch = old_gp
    patches = torch.stack(rec["patches"], 1)            # (B,T,3,P,P)
    std = torch.stack(rec["std_actions"], 1)            # (B,T,2)
    coords = torch.floor(std * (args.input_size - args.patch_size)).int()
    out = {
        "scan_equals_input": np.array(torch.equal(scan, x)),
        "fmap_mean": fmap.mean(dim=(3, 4)).numpy().astype(np.float32),          # (B,T,1280), == gvec
        "fmap_sample": fmap[:, :, ::64].numpy().astype(np.float32),             # (B,T,20,7,7)
        "gvec": gvec.numpy().astype(np.float32),
        "actions": torch.stack(rec["actions"], 1).numpy().astype(np.int64),     # (B,T)
        "std_actions": std.numpy().astype(np.float32),
        "coords": coords.numpy().astype(np.int32),
        "patch_checksum": patches.double().sum(dim=(2, 3, 4)).numpy(),          # (B,T)
        "patch_corner": patches[:, :, :, :2, :2].numpy().astype(np.float32),
        "lfeat": torch.stack(rec["lfeat"], 1).numpy().astype(np.float32),       # (B,T,2048)
        "logits": logits.numpy().astype(np.float32),
        "last_out": last_out.numpy().astype(np.float32),
    }
    path = os.path.join(HERE, f"act_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    return model


def run_reference_stage2(ref, args, batch, tag, np_seed=7):
    """The stage-2 (RL) validation loop of ACT/main_dist.py:343-366: glance, then one_step_act per step; the reward
    baseline crops come from the numpy host generator."""
    torch.manual_seed(0)
    model = ref.GFV(args)
    ck = synth.synth_checkpoint_act(model, synth.SEED)
    synth.load_checkpoint_act(model, ck)
    model.eval()
    x = synth.synth_clips(batch, args.num_segments, args.input_size, synth.SEED)
    images = x.view(batch, args.num_segments, 3, args.input_size, args.input_size)
    np.random.seed(np_seed)
    preds, bases, acts = [], [], []
    with torch.no_grad():
        fmap, gvec = model.glance(x)
        for t in range(args.num_segments):
            out, pred, _, action, base = model.one_step_act(images[:, t], fmap[:, t], gvec[:, t], restart_batch=(t == 0),
                                                            training=False)
            preds.append(pred.clone())
            bases.append(base.clone())
            acts.append(action.clone())
    out = {"pred": torch.stack(preds, 0).numpy().astype(np.float32),          # (T, B, C)
           "baseline_logits": torch.stack(bases, 0).numpy().astype(np.float32),
           "std_actions": torch.stack(acts, 0).numpy().astype(np.float32), "np_seed": np.array(np_seed)}
    path = os.path.join(HERE, f"act_stage2_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def metrics_golden():
    """accuracy() / cal_map() of the reference (ACT/ops/utils.py:35-88) on seeded logits, incl. ties and -1 labels."""
    import warnings
    from ops.utils import accuracy, cal_map
    g = torch.Generator().manual_seed(11)
    n, c = 257, 12
    logits = torch.randn(n, c, generator=g)
    logits[5] = logits[9]                     # tied rows
    logits[20, 3] = logits[20, 7]             # tie inside a row
    target = torch.randint(0, c, (n,), generator=g)
    labels = torch.stack([target, torch.randint(-1, c, (n,), generator=g)], 1)
    out = {"logits": logits.numpy(), "target": target.numpy(), "labels": labels.numpy()}
    acc1, acc5 = accuracy(logits, target, topk=(1, 5))
    out["acc1"], out["acc5"] = acc1.numpy(), acc5.numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m1, ap1 = cal_map(logits, target.view(-1, 1))
        m2, ap2 = cal_map(logits, labels)
    out["map_single"], out["ap_single"] = np.array(float(m1)), ap1.numpy()
    out["map_multi"], out["ap_multi"] = np.array(float(m2)), ap2.numpy()
    path = os.path.join(HERE, "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, float(acc1), float(acc5), float(m1), float(m2))


def get_patch_kat(ref):
    """Known-answer table for get_patch straight from the reference function (ACT/models/utils.py:37-51)."""
    from models.utils import get_patch
    rows = {}
    g = torch.Generator().manual_seed(5)
    img = torch.randn(4, 3, 224, 224, generator=g)
    for n, p in ((7, 128), (7, 96), (7, 160), (7, 192), (5, 144), (6, 112), (8, 176)):
        grid = torch.tensor([[iy / (n - 1), ix / (n - 1)] for iy in range(n) for ix in range(n)], dtype=torch.float32)
        coords = torch.floor(grid * (224 - p)).int()
        rows[f"grid{n}_p{p}"] = coords.numpy()
    acts = torch.rand(4, 2, generator=g)
    acts[0] = torch.tensor([0.5, 1.0])
    acts[1] = torch.tensor([0.0, 0.0])
    acts[2] = torch.tensor([1.0, 1.0])
    for p in (96, 128, 144, 130):
        out = get_patch(img, acts, p)
        rows[f"rand_p{p}_sum"] = out.double().sum(dim=(1, 2, 3)).numpy()
        rows[f"rand_p{p}_coords"] = torch.floor(acts * (224 - p)).int().numpy()
        rows[f"rand_p{p}_first"] = out[:, :, 0, :4].numpy()
    rows["rand_actions"] = acts.numpy()
    path = os.path.join(HERE, "get_patch_kat.npz")
    np.savez_compressed(path, **rows)
    print("wrote", path)


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = import_reference_act()
    if "--metrics-only" in sys.argv:
        metrics_golden()
        sys.exit(0)
    if "--stage2-only" in sys.argv:
        run_reference_stage2(ref, synth.act_args(num_segments=4, patch_size=96, action_dim=36, num_classes=51), 3,
                             "t4_p96_b3")
        sys.exit(0)
    get_patch_kat(ref)
    metrics_golden()
    run_reference_stage2(ref, synth.act_args(num_segments=4, patch_size=96, action_dim=36, num_classes=51), 3,
                         "t4_p96_b3")
    # config 3 shape: T=16, P=128, 49 actions, 200 classes; 2 clips
    model = run_reference(ref, synth.act_args(), 2, "c3_b2")
    import json
    keys = {"glancer": model.glancer.state_dict(), "focuser": model.focuser.state_dict(),
            "fc": model.classifier.state_dict(), "policy": model.focuser.policy.policy.state_dict(),
            "model": model.state_dict()}
    with open(os.path.join(HERE, "ref_state_keys.json"), "w") as f:
        json.dump({part: [(k, list(v.shape)) for k, v in sd.items()] for part, sd in keys.items()}, f)
    # a second, differently shaped configuration: T=4, P=96, 36 actions, 51 classes; 3 clips
    run_reference(ref, synth.act_args(num_segments=4, patch_size=96, action_dim=36, num_classes=51), 3, "t4_p96_b3")
