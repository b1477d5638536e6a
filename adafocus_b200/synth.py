"""Deterministic synthetic checkpoints and inputs (there is no network for the reference's Google-Drive checkpoints
or datasets).  Values come from a seeded CPU torch.Generator with per-tensor rules chosen so that activations stay
O(1) through 50-100 layers without a calibration pass -- the result is bit-reproducible on any machine with the same
torch build, which lets golden outputs produced by the *reference's* modules in the build container be compared on
the GPU box where /root/reference does not exist.

The checkpoint dict has the reference's format (ACT/main_dist.py:278-288): keys 'glancer', 'focuser', 'fc', 'policy'.
"""
import math
from types import SimpleNamespace

import torch

SEED = 1007   # ACT/conf/default.yaml:61


def act_args(**over):
    """args namespace with the fields GFV.__init__ reads (ACT/models/gfv_net.py:17-58), README eval values."""
    a = dict(num_segments=16, num_classes=200, reward="random", dataset="actnet", input_size=224, batch_size=64,
             patch_size=128, with_glancer=True, feature_map_channels=1280, glance_size=224, action_dim=49,
             hidden_state_dim=1024, policy_conv=True, gpu=0, continuous=False, gamma=0.7, policy_lr=0.0003,
             random_patch=False, dropout=0.5, consensus="gru", hidden_dim=1024, train_stage=3, evaluate=True,
             seed=SEED)
    a.update(over)
    return SimpleNamespace(**a)


_MBV2_RESIDUAL_BLOCKS = {3, 5, 6, 8, 9, 10, 12, 13, 15, 16}   # stride 1 and cin == cout in the MobileNet-V2 table


def _mbv2_block(name):
    """index N of a '...features.N.conv.K...' MobileNet-V2 entry, else None."""
    parts = name.split(".")
    if "features" in parts:
        i = parts.index("features")
        if i + 1 < len(parts) and parts[i + 1].isdigit():
            return int(parts[i + 1]), parts[i + 2:]
    return None


def _closes_residual_branch(name):
    """BatchNorm that ends a residual branch: ResNet bn3, MobileNet-V2 project BN of a residual block."""
    if ".bn3." in name:
        return True
    mb = _mbv2_block(name)
    if mb is not None and mb[0] in _MBV2_RESIDUAL_BLOCKS and mb[1][0] == "conv" and len(mb[1]) == 3:
        # ACT layout: nested ConvBNReLU -> project BN is conv.2 / conv.3; STH (tonylins) flat layout: conv.7
        return mb[1][1] in ("2", "3", "7")
    return False


def _input_is_linear(name):
    """Convs fed by an un-activated tensor: MobileNet-V2 expand convs (blocks >= 2) and its last 1x1 conv."""
    mb = _mbv2_block(name)
    if mb is None:
        return False
    n, rest = mb
    if n == 18:
        return True
    if n < 2 or rest[0] != "conv" or rest[1] != "0":
        return False
    # ACT: features.N.conv.0.0.weight ; STH flat: features.N.conv.0.weight or (TSM-wrapped) features.N.conv.0.net.weight
    return rest[2:] in (["0", "weight"], ["weight"], ["net", "weight"])


def _fill(name, t, sd, g):
    shape = tuple(t.shape)
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 1)[0] if "." in name else ""
    is_bn = (parent + ".running_mean") in sd
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_var":
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    if is_bn and leaf == "weight":
        if _closes_residual_branch(name):
            return torch.rand(shape, generator=g) * 0.2 + 0.2    # keeps the un-normalised residual stream O(1)
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    if is_bn and leaf == "bias":
        return torch.randn(shape, generator=g) * 0.1
    if "gru." in name or leaf.startswith("weight_ih") or leaf.startswith("weight_hh") or leaf.startswith("bias_ih") \
            or leaf.startswith("bias_hh"):
        hidden = shape[0] // 3
        k = 1.0 / math.sqrt(hidden)
        return (torch.rand(shape, generator=g) * 2 - 1) * k
    if t.dim() == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        gain = 1.0 if _input_is_linear(name) else 2.0             # He init only where the input went through a ReLU
        return torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
    if t.dim() == 2:
        gain = 4.0 if ".actor." in "." + name or name.startswith("actor.") else 1.0
        return torch.randn(shape, generator=g) * (gain / math.sqrt(shape[1]))
    if t.dim() == 1:
        return torch.randn(shape, generator=g) * 0.05
    return torch.zeros(shape)


def synth_state_dict(module, seed):
    """Seeded values for every entry of module.state_dict(), in state_dict order."""
    g = torch.Generator().manual_seed(int(seed))
    sd = module.state_dict()
    out = {}
    for name, t in sd.items():
        v = _fill(name, t, sd, g)
        out[name] = v.to(t.dtype) if v.dtype != t.dtype and t.dtype.is_floating_point else v
    return out


def synth_checkpoint_act(model, seed=SEED):
    """Checkpoint dict in the reference's format for an ACT-tree GFV (ours or the reference's class)."""
    ck = {
        "glancer": synth_state_dict(model.glancer, seed + 1),
        "focuser": synth_state_dict(model.focuser, seed + 2),
        "fc": synth_state_dict(model.classifier, seed + 3),
        "policy": synth_state_dict(model.focuser.policy.policy, seed + 4),
        "best_acc": 0.0, "epoch": 0,
    }
    # the policy tensors also live inside the focuser dict (PPO is an nn.Module in the ACT tree): keep them equal
    for k, v in ck["policy"].items():
        ck["focuser"]["policy.policy." + k] = v.clone()
        ck["focuser"]["policy.policy_old." + k] = v.clone()
    return ck


def load_checkpoint_act(model, ck):
    """The reference's loading sequence (ACT/main_dist.py:100-110)."""
    model.glancer.load_state_dict(ck["glancer"])
    model.focuser.load_state_dict(ck["focuser"], strict=False)
    model.classifier.load_state_dict(ck["fc"])
    model.focuser.policy.policy.load_state_dict(ck["policy"])
    model.focuser.policy.policy_old.load_state_dict(ck["policy"])


def sth_args(**over):
    """args namespace for the STH-tree GFV (STH/models/gfv_net.py:21-69) with evaluate.sh's values."""
    a = dict(num_segments_glancer=8, num_segments_focuser=12, num_classes=174, batch_size=32, patch_size=144,
             with_glancer=True, feature_map_channels=1280, video_div=1, glance_size=224, action_dim=25,
             hidden_state_dim=1024, policy_conv=True, gpu=0, ppo_continuous=True, gamma=0.7, policy_lr=0.0003,
             action_std=0.25, actorcritic_with_bn=True, modality="RGB", base_model="resnet50", partial_bn=False,
             pretrain="imagenet", is_shift=True, shift_div=8, shift_place="blockres", fc_lr5=False,
             temporal_pool=False, non_local=False, random_patch=False, dropout=0.5, train_stage=2, evaluate=True,
             seed=SEED)
    a.update(over)
    return SimpleNamespace(**a)


def synth_checkpoint_sth(model, seed=SEED):
    """Checkpoint dict in the STH format (STH/evaluate.py:136-146): 'glancer', 'focuser', 'fc', 'policy'.  Call it after
    the fc of focuser.net.base_model has been stripped (STH/evaluate.py:83) so the keys match what evaluate.py loads."""
    return {
        "glancer": synth_state_dict(model.glancer, seed + 11),
        "focuser": synth_state_dict(model.focuser, seed + 12),
        "fc": synth_state_dict(model.classifier, seed + 13),
        "policy": synth_state_dict(model.focuser.policy.policy, seed + 14),
        "best_acc": 0.0, "epoch": 0,
    }


def load_checkpoint_sth(model, ck):
    """The reference's loading sequence (STH/evaluate.py:141-146)."""
    model.glancer.load_state_dict(ck["glancer"], strict=True)
    model.focuser.load_state_dict(ck["focuser"], strict=True)
    model.classifier.load_state_dict(ck["fc"], strict=True)
    model.focuser.policy.policy.load_state_dict(ck["policy"])
    model.focuser.policy.policy_old.load_state_dict(ck["policy"])


def strip_fc_sth(model):
    """STH/evaluate.py:83."""
    model.focuser.net.base_model = torch.nn.Sequential(*list(model.focuser.net.base_model.children())[:-1])


def synth_clips(batch, frames=16, size=224, seed=SEED, device="cpu"):
    """(B, 3T, H, W) fp32 ~ N(0,1) (the range of ImageNet-normalised pixels), seeded on CPU for reproducibility."""
    g = torch.Generator().manual_seed(int(seed) + 77)
    coarse = torch.randn(batch * frames, 3, 7, 7, generator=g)
    smooth = torch.nn.functional.interpolate(coarse, size=(size, size), mode="nearest")   # blocky "objects"
    x = smooth + 0.5 * torch.randn(batch * frames, 3, size, size, generator=g)
    return x.view(batch, 3 * frames, size, size).contiguous().to(device)
