"""ORACLE -- test infrastructure only.  CPU fp32 restatement of the AdaFocus offline-inference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file; the
product (adafocus_b200/) never does.  It restates, with plain torch.nn.functional calls on CPU fp32 tensors and
explicit GRU gate equations, what the reference computes, driven directly by the reference's checkpoint dict
({'glancer','focuser','fc','policy'}); each function cites the reference lines it follows
(ACT/ = "Experiments on ActivityNet, FCVID and Mini-Kinetics/", STH/ = "Experiments on Something-Something V1&V2/").

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against the
reference's own classes executed in the build container: tests/golden/make_golden.py imports /root/reference, loads
the same seeded synthetic checkpoint through the reference's load_state_dict sequence and stores its outputs in
tests/golden/*.npz; tests/test_oracle.py checks this file against those vectors and against the get_patch
known-answer table of SURVEY.md section 8 (a5).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5

# MobileNet-V2 block table (expand, channels, repeats, stride): ACT/models/mobilenet.py:89-98
MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                (6, 320, 1, 1))
RESNET_LAYERS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


# ------------------------------------------------------------------------------------------------ crop
def patch_coordinates(action, image_size, patch_size):
    """floor(action * (image_size - patch_size)).int() on fp32 -- ACT/models/utils.py:42."""
    a = np.asarray(action, dtype=np.float32)
    return np.floor(a * np.float32(image_size - patch_size)).astype(np.int32)


def get_patch(images, action_sequence, patch_size):
    """ACT/models/utils.py:37-51 (= STH/models/utils.py:44-58): per-sample slice at the floored coordinates; rows come
    from action[:,0], columns from action[:,1]; image_size is images.shape[2] for both axes."""
    images = np.asarray(images)
    n = images.shape[0]
    coord = patch_coordinates(action_sequence, images.shape[2], patch_size)
    out = np.empty((n, images.shape[1], patch_size, patch_size), dtype=images.dtype)
    for i in range(n):
        y, x = int(coord[i, 0]), int(coord[i, 1])
        out[i] = images[i, :, y:y + patch_size, x:x + patch_size]
    return out


def standard_actions(action_dim):
    """ACT/models/gfv_net.py:272-307: n x n grid of (row, col) = (iy/(n-1), ix/(n-1)), row-major, stored as fp32."""
    n = int(round(math.sqrt(action_dim)))
    return np.array([[iy / (n - 1), ix / (n - 1)] for iy in range(n) for ix in range(n)], dtype=np.float32)


# ------------------------------------------------------------------------------------------------ CNN pieces
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False,
                        0.0, BN_EPS)


def temporal_shift(x, n_segment, fold_div):
    """STH/ops/temporal_shift.py:29-46."""
    nt, c, h, w = x.shape
    x = x.view(nt // n_segment, n_segment, c, h, w)
    fold = c // fold_div
    out = torch.zeros_like(x)
    out[:, :-1, :fold] = x[:, 1:, :fold]
    out[:, 1:, fold:2 * fold] = x[:, :-1, fold:2 * fold]
    out[:, :, 2 * fold:] = x[:, :, 2 * fold:]
    return out.view(nt, c, h, w)


def mobilenet_v2_features(x, sd, prefix="net.features.", tsm=None):
    """ACT/models/mobilenet.py:32-68,104-118 (ConvBNReLU = conv, BN(eval), ReLU6; InvertedResidual).
    tsm=(n_segment, fold_div) shifts the input of the first conv of residual blocks (STH/models/gfv_net.py:238-241)."""
    def cbr(x, p, stride, groups):
        w = sd[p + "0.weight"]
        x = F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2, 1, groups)
        return F.relu6(_bn(x, sd, p + "1."))

    x = cbr(x, prefix + "0.", 2, 1)
    idx, cin = 1, 32
    for t, c, n, s in MBV2_SETTING:
        for i in range(n):
            stride = s if i == 0 else 1
            p = f"{prefix}{idx}.conv."
            hidden = cin * t
            res = stride == 1 and cin == c
            y = x
            if tsm is not None and res:
                y = temporal_shift(y, tsm[0], tsm[1])
            k = 0
            if t != 1:
                y = cbr(y, f"{p}{k}.", 1, 1)
                k += 1
            y = cbr(y, f"{p}{k}.", stride, hidden)
            k += 1
            y = F.conv2d(y, sd[f"{p}{k}.weight"])
            y = _bn(y, sd, f"{p}{k + 1}.")
            x = x + y if res else y
            cin = c
            idx += 1
    return cbr(x, f"{prefix}{idx}.", 1, 1)


def mobilenet_v2_features_flat(x, sd, prefix="net.features.", tsm=None):
    """STH/models/mobilenetv2.py:7-66 (tonylins layout: InvertedResidual.conv is one flat Sequential) with the temporal
    shift wrapped around conv[0] of the residual blocks that have an expand conv (STH/models/gfv_net.py:238-241)."""
    x = F.relu6(_bn(F.conv2d(x, sd[prefix + "0.0.weight"], None, 2, 1), sd, prefix + "0.1."))
    idx, cin = 1, 32
    for t, c, n, s in MBV2_SETTING:
        for i in range(n):
            stride = s if i == 0 else 1
            p = f"{prefix}{idx}.conv."
            res = stride == 1 and cin == c
            y = x
            k = 0
            if t != 1:
                w = sd.get(p + "0.weight")
                if w is None:                               # TemporalShift wrapper: conv.0.net.weight
                    w = sd[p + "0.net.weight"]
                    y = temporal_shift(y, tsm[0], tsm[1])
                y = F.relu6(_bn(F.conv2d(y, w), sd, p + "1."))
                k = 3
            y = F.relu6(_bn(F.conv2d(y, sd[f"{p}{k}.weight"], None, stride, 1, 1, cin * t), sd, f"{p}{k + 1}."))
            y = _bn(F.conv2d(y, sd[f"{p}{k + 3}.weight"]), sd, f"{p}{k + 4}.")
            x = x + y if res else y
            cin = c
            idx += 1
    return F.relu6(_bn(F.conv2d(x, sd[f"{prefix}{idx}.0.weight"]), sd, f"{prefix}{idx}.1."))


def mobilenet_v2_get_featmap(x, sd, prefix="net.features."):
    """ACT/models/mobilenet.py:146-148: (feature map, mean over H,W)."""
    f = mobilenet_v2_features(x, sd, prefix)
    return f, f.mean([2, 3])


def resnet_trunk(x, sd, prefix="net.", layers=(3, 4, 6, 3), pooled=True, tsm=None, names=None):
    """ACT/models/resnet.py:211-225 with Bottleneck.forward :94-114 (stride on conv2); fc not applied.
    tsm=(n_segment, fold_div, n_round) shifts conv1's input of every n_round-th block of a stage
    (STH/ops/temporal_shift.py:113-135, blockres)."""
    # names: child names of (conv1, bn1, layer1..4); the fc-stripped nn.Sequential of STH/evaluate.py:83 numbers them
    names = names or ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4")
    x = F.conv2d(x, sd[f"{prefix}{names[0]}.weight"], None, 2, 3)
    x = F.relu(_bn(x, sd, f"{prefix}{names[1]}."))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nblocks in enumerate(layers, start=1):
        for bi in range(nblocks):
            p = f"{prefix}{names[1 + li]}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            idn = x
            y = x
            if tsm is not None and bi % tsm[2] == 0:
                y = temporal_shift(y, tsm[0], tsm[1])
            w1 = sd.get(p + "conv1.weight", sd.get(p + "conv1.net.weight"))
            y = F.relu(_bn(F.conv2d(y, w1), sd, p + "bn1."))
            y = F.relu(_bn(F.conv2d(y, sd[p + "conv2.weight"], None, stride, 1), sd, p + "bn2."))
            y = _bn(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3.")
            if (p + "downsample.0.weight") in sd:
                idn = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], None, stride), sd, p + "downsample.1.")
            x = F.relu(y + idn)
    if pooled:
        return F.adaptive_avg_pool2d(x, (1, 1))
    return x


# ------------------------------------------------------------------------------------------------ recurrent pieces
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRU single step, gates ordered r,z,n: n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    hd = h.shape[1]
    r = torch.sigmoid(gi[:, :hd] + gh[:, :hd])
    z = torch.sigmoid(gi[:, hd:2 * hd] + gh[:, hd:2 * hd])
    n = torch.tanh(gi[:, 2 * hd:] + r * gh[:, 2 * hd:])
    return (1 - z) * n + z * h


def policy_encode(state, sd):
    """ACT/models/ppo.py:33-39: 1x1 conv (no bias) -> ReLU -> Flatten (NCHW order) -> Linear -> ReLU."""
    s = F.relu(F.conv2d(state, sd["state_encoder.0.weight"]))
    s = s.flatten(1)
    return F.relu(s @ sd["state_encoder.3.weight"].t() + sd["state_encoder.3.bias"])


def policy_act(state, h, sd):
    """ACT/models/ppo.py:67-96, eval branch: returns (action index via argmax of softmax probs, new hidden, probs)."""
    s = policy_encode(state, sd)
    h = gru_cell(s, h, sd["gru.weight_ih_l0"], sd["gru.weight_hh_l0"], sd["gru.bias_ih_l0"], sd["gru.bias_hh_l0"])
    probs = torch.softmax(h @ sd["actor.0.weight"].t() + sd["actor.0.bias"], dim=-1)
    return probs.max(1)[1], h, probs


def recurrent_classifier(features, sd):
    """ACT/models/gfv_net.py:427-435: GRU over (B,T,F) from h0 = 0, dropout(eval) = identity, fc on every step."""
    b, t, _ = features.shape
    h = torch.zeros(b, sd["gru.weight_hh_l0"].shape[1])
    outs = []
    for i in range(t):
        h = gru_cell(features[:, i], h, sd["gru.weight_ih_l0"], sd["gru.weight_hh_l0"], sd["gru.bias_ih_l0"],
                     sd["gru.bias_hh_l0"])
        outs.append(h)
    out = torch.stack(outs, 1)
    logits = out.reshape(b * t, -1) @ sd["fc.weight"].t() + sd["fc.bias"]
    last_out = logits.reshape(b, t, -1)[:, -1, :].reshape(b, -1)
    return logits, last_out


# ------------------------------------------------------------------------------------------------ whole path (ACT)
SEQ_NAMES = ("0", "1", "4", "5", "6", "7")     # conv1, bn1, layer1..4 inside nn.Sequential(*children[:-1])


def policy_act_continuous(state, h, sd):
    """STH/models/ppo_continuous.py:78-109, eval branch: conv1x1 -> BN2d -> ReLU -> Flatten -> Linear -> BN1d -> ReLU,
    one GRU step, action_mean = sigmoid(Linear)."""
    p = "state_encoder."
    s = F.conv2d(state, sd[p + "0.weight"])
    if (p + "1.running_mean") in sd:
        s = F.relu(_bn(s, sd, p + "1.")).flatten(1)
        s = s @ sd[p + "4.weight"].t() + sd[p + "4.bias"]
        s = F.relu(F.batch_norm(s, sd[p + "5.running_mean"], sd[p + "5.running_var"], sd[p + "5.weight"],
                                sd[p + "5.bias"], False, 0.0, BN_EPS))
    else:
        s = F.relu(s).flatten(1)
        s = F.relu(s @ sd[p + "3.weight"].t() + sd[p + "3.bias"])
    h = gru_cell(s, h, sd["gru.weight_ih_l0"], sd["gru.weight_hh_l0"], sd["gru.bias_ih_l0"], sd["gru.bias_hh_l0"])
    return torch.sigmoid(h @ sd["actor.0.weight"].t() + sd["actor.0.bias"]), h


@torch.no_grad()
def sth_forward(glancer_images, focuser_images, ck, patch_size=144, t_g=8, t_f=12, video_div=1, shift_div=8,
                layers=(3, 4, 6, 3), with_glancer=True, rand_actions=None):
    """STH/evaluate.py:188-201 + STH/models/gfv_net.py:101-225 on CPU fp32: glance, then per video division one
    continuous policy step, get_patch over all frames of the division, TSM-ResNet, classifier, average consensus.
    rand_actions (video_div, B, 2): the torch.rand draws of random_patching -> also returns the baseline logits."""
    gi, fi = glancer_images.float().cpu(), focuser_images.float().cpu()
    b = gi.shape[0]
    g = gi.shape[-1]
    fmap = mobilenet_v2_features_flat(gi.view(b * t_g, 3, g, g), ck["glancer"], "net.features.", (t_g, shift_div))
    glogit = fmap.mean(3).mean(2) @ ck["glancer"]["net.classifier.weight"].t() + ck["glancer"]["net.classifier.bias"]
    fmap5 = fmap.view(b, t_g, *fmap.shape[1:])
    glogit = glogit.view(b, t_g, -1)
    frames = fi.view(b, t_f, 3, fi.shape[-2], fi.shape[-1])
    n_round = 2 if layers[2] >= 23 else 1
    fpg, fpf = t_g // video_div, t_f // video_div
    hid = torch.zeros(b, ck["policy"]["gru.weight_hh_l0"].shape[1])
    patches, base_patches, actions, coords, preds, baselines = None, None, [], [], [], []

    def head(pt, nframes):
        feat = resnet_trunk(pt.reshape(-1, 3, patch_size, patch_size), ck["focuser"], "net.base_model.", layers, True,
                            (t_f, shift_div, n_round), SEQ_NAMES).flatten(1)
        logit = (feat @ ck["fc"]["weight"].t() + ck["fc"]["bias"]).view(b, nframes, -1).mean(1)
        return logit + glogit.mean(1) if with_glancer else logit

    for d in range(video_div):
        cur_img = frames[:, d * fpf:(d + 1) * fpf].reshape(b, -1, *frames.shape[-2:])
        cur_map = fmap5[:, d * fpg:(d + 1) * fpg].reshape(b, -1, *fmap5.shape[-2:])
        a, hid = policy_act_continuous(cur_map, hid, ck["policy"])
        actions.append(a)
        coords.append(patch_coordinates(a.numpy(), cur_img.shape[2], patch_size))
        cur = torch.from_numpy(get_patch(cur_img.numpy(), a.numpy(), patch_size)).view(b, fpf, 3, patch_size, patch_size)
        if rand_actions is not None:
            rnd = torch.from_numpy(get_patch(cur_img.numpy(), np.asarray(rand_actions[d]), patch_size))
            rnd = rnd.view(b, fpf, 3, patch_size, patch_size)
            base_patches = rnd if patches is None else torch.cat([patches, rnd], 1)
            baselines.append(head(base_patches, fpf * (d + 1)))
        patches = cur if patches is None else torch.cat([patches, cur], 1)
        preds.append(head(patches, fpf * (d + 1)))
    return {"fmap": fmap5, "glogit": glogit, "actions": torch.stack(actions, 1), "coords": np.stack(coords, 1),
            "patches": patches, "preds": preds, "pred": preds[-1], "baselines": baselines}


@torch.no_grad()
def act_forward(inp, scan, ck, patch_size=128, action_dim=49, with_glancer=True, actions_override=None,
                layers=(3, 4, 6, 3)):
    """GFV.forward(one_step=True) of the ACT tree -- ACT/models/gfv_net.py:95-133 -- on CPU fp32.

    inp (B,3T,H,W), scan (B,3T,g,g) torch fp32.  Returns a dict with every stage's output.  actions_override
    (B,T) int64 replaces the policy's argmax (used for staged parity tests)."""
    inp, scan = inp.float().cpu(), scan.float().cpu()
    b, tc, h, w = inp.shape
    t = tc // 3
    frames = inp.view(b, t, 3, h, w)
    g = scan.shape[-1]
    fmap, gvec = mobilenet_v2_get_featmap(scan.view(b * t, 3, g, g), ck["glancer"])          # glance(), :152-158
    fmap = fmap.view(b, t, *fmap.shape[1:])
    gvec = gvec.view(b, t, -1)
    table = torch.from_numpy(standard_actions(action_dim))
    hid = torch.zeros(b, ck["policy"]["gru.weight_hh_l0"].shape[1])                           # restart_batch, ppo.py:68-70
    feats, actions, coords, patches, lfeats, probs_all = [], [], [], [], [], []
    for step in range(t):                                                                     # gfv_net.py:110
        a, hid, probs = policy_act(fmap[:, step], hid, ck["policy"])
        if actions_override is not None:
            a = actions_override[:, step]
        std = table[a]                                                                        # :345-347
        img = frames[:, step]
        coords.append(patch_coordinates(std.numpy(), h, patch_size))
        patch = torch.from_numpy(get_patch(img.numpy(), std.numpy(), patch_size))             # utils.py:37-51
        lf = resnet_trunk(patch, ck["focuser"], "net.", layers).view(b, -1)                   # :331
        feats.append(torch.cat([gvec[:, step], lf], 1) if with_glancer else lf)               # :120-122
        actions.append(a)
        patches.append(patch)
        lfeats.append(lf)
        probs_all.append(probs)
    features = torch.stack(feats, 1)                                                          # :132
    logits, last_out = recurrent_classifier(features, ck["fc"])                               # :133
    return {
        "fmap": fmap, "gvec": gvec, "actions": torch.stack(actions, 1), "coords": np.stack(coords, 1),
        "patches": torch.stack(patches, 1), "lfeat": torch.stack(lfeats, 1), "features": features,
        "probs": torch.stack(probs_all, 1), "logits": logits, "last_out": last_out,
    }


# ------------------------------------------------------------------------------------------------ metrics (f-4)
def accuracy_topk(output, target, topk=(1,)):
    """ACT/ops/utils.py:35-49: percentage of rows whose target is among the k largest logits."""
    output, target = np.asarray(output), np.asarray(target)
    order = np.argsort(-output, axis=1, kind="stable")
    res = []
    for k in topk:
        res.append(100.0 * np.mean([target[i] in order[i, :k] for i in range(len(target))]))
    return res


def cal_map(output, labels):
    """ACT/ops/utils.py:51-88: labels re-ranked among the distinct non-negative values (get_multi_hot with
    assumes_starts_zero=False), softmax over classes, per-class AP = mean over positives of precision at their rank in
    the descending (stable) order of the class probability; returns (mAP %, AP % per class)."""
    output = np.asarray(output, dtype=np.float32)
    labels = np.asarray(labels).reshape(output.shape[0], -1).copy()
    n, c = output.shape
    uniq = np.unique(labels[labels >= 0])
    remap = {int(v): i for i, v in enumerate(uniq)}
    gt = np.zeros((n, c + 1), dtype=bool)
    for i in range(n):
        for l in labels[i]:
            gt[i, remap[int(l)] if l >= 0 else -1] = True       # -1 lands in the extra last column, dropped below
    gt = gt[:, :c]
    z = output - output.max(axis=1, keepdims=True)
    probs = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
    ap = np.zeros(c, dtype=np.float64)
    for k in range(c):
        order = np.argsort(-probs[:, k], kind="stable")
        truth = gt[order, k]
        tp = np.cumsum(truth)
        prec = tp / np.arange(1, n + 1)
        ap[k] = prec[truth].sum() / max(float(truth.sum()), 1.0)
    return ap.mean() * 100, ap * 100
