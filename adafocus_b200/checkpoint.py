"""Checkpoint ingest in the reference's formats.

The module trees of adafocus_b200.models / models_sth carry the reference's parameter names, so ingest is the
reference's own `load_state_dict` sequence plus the key remaps its scripts apply to TSM-pretrained backbones.
Packed fp16 kernel weights are derived lazily at the first forward and invalidated by any reload.
"""
import os
from collections import OrderedDict

import torch


def _as_dict(ckpt, map_location="cpu"):
    if isinstance(ckpt, (str, os.PathLike)):
        return torch.load(os.path.expanduser(ckpt), map_location=map_location)
    return ckpt


def load_act_checkpoint(model, ckpt, train_stage=3):
    """ACT tree, `resume` handling of ACT/main_dist.py:92-110: keys 'glancer', 'focuser' (strict=False), 'fc', and for
    stage 3 the 'policy' dict into both policy and policy_old.  Returns the checkpoint's bookkeeping fields."""
    ck = _as_dict(ckpt)
    model.glancer.load_state_dict(ck["glancer"])
    model.focuser.load_state_dict(ck["focuser"], strict=False)
    model.classifier.load_state_dict(ck["fc"])
    if train_stage == 3 and ck.get("policy") is not None:
        model.focuser.policy.policy.load_state_dict(ck["policy"])
        model.focuser.policy.policy_old.load_state_dict(ck["policy"])
    return {"epoch": ck.get("epoch"), "best_acc": ck.get("best_acc")}


def remap_tsm_glancer_state_dict(state_dict):
    """Keys of a TSM-MobileNet-V2 training checkpoint -> Glancer.net keys (STH/evaluate.py:43-52):
    'module.base_model.X' -> 'X', 'module.new_fc.X' -> 'classifier.X', everything else unchanged."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if k.startswith("module.base_model."):
            out[k[len("module.base_model."):]] = v
        elif k.startswith("module.new_fc."):
            out["classifier." + k[len("module.new_fc."):]] = v
        else:
            out[k] = v
    return out


def remap_tsm_focuser_state_dict(state_dict):
    """Keys of a TSM-ResNet training checkpoint -> (Focuser.net keys, classifier keys) (STH/evaluate.py:63-72):
    'module.new_fc.X' -> classifier 'X'; other 'module.X' -> 'X' (i.e. 'base_model....')."""
    net, fc = OrderedDict(), OrderedDict()
    for k, v in state_dict.items():
        if k.startswith("module.new_fc."):
            fc[k[len("module.new_fc."):]] = v
        elif k.startswith("module."):
            net[k[len("module."):]] = v
        else:
            net[k] = v
    return net, fc


def load_sth_pretrained(model, glancer_ckpt=None, focuser_ckpt=None):
    """STH tree, `pretrained_glancer` / `pretrained_focuser` handling of STH/evaluate.py:40-81.  Must run BEFORE the fc
    of focuser.net.base_model is stripped (the TSM checkpoint still carries base_model.fc.*; strict=False skips it)."""
    if glancer_ckpt is not None:
        sd = _as_dict(glancer_ckpt)["state_dict"]
        model.glancer.net.load_state_dict(remap_tsm_glancer_state_dict(sd), strict=True)
    if focuser_ckpt is not None:
        net, fc = remap_tsm_focuser_state_dict(_as_dict(focuser_ckpt)["state_dict"])
        model.classifier.load_state_dict(fc, strict=True)
        model.focuser.net.load_state_dict(net, strict=False)


def load_sth_checkpoint(model, ckpt):
    """STH tree, `resume` handling of STH/evaluate.py:136-146 (call after the fc strip of :83)."""
    ck = _as_dict(ckpt)
    model.glancer.load_state_dict(ck["glancer"], strict=True)
    model.focuser.load_state_dict(ck["focuser"], strict=True)
    model.classifier.load_state_dict(ck["fc"], strict=True)
    model.focuser.policy.policy.load_state_dict(ck["policy"])
    model.focuser.policy.policy_old.load_state_dict(ck["policy"])
    return {"epoch": ck.get("epoch"), "best_acc": ck.get("best_acc")}


def save_checkpoint(state, path):
    """Atomic save like */basic_tools/checkpoint.py:47-52: write to a temporary file, then rename over the target."""
    tmp = str(path) + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
