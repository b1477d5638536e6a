"""AdaFocus top module for the ActivityNet / FCVID / Mini-Kinetics tree -- mirror of ACT/models/gfv_net.py.

Same classes, constructor arguments, attribute names and return values as the reference (GFV :13-228, Glancer
:231-252, Focuser :255-351, PatchSampler :354-385, RecurrentClassifier :409-457) so `main_dist.py validate()`
(stage 3, ACT/main_dist.py:367-371) runs on it unchanged and reference checkpoints load by name.  The arithmetic of
the inference path is one recorded launch plan on the B200 engine:

    fG over all B*T frames -> policy rollout for all T steps (no host sync) -> crop fused into the fL stem staging
    -> fL over all B*T patches in ONE batch -> [global | local] features -> GRU classifier -> logits

The reference instead runs fL T times at batch B inside a Python loop with 4*B `.item()` syncs per step
(ACT/models/gfv_net.py:110-131, ACT/models/utils.py:44-49); the policy only ever sees fG features
(ACT/models/gfv_net.py:320), so the reordering is exact.
"""
import math

import torch
from torch import nn

from .. import _lib
from ..engine import get_engine, pack_conv_split
from ..packcache import cached_runner
from .mobilenet import _param_key, invalidate_packed, mobilenet_v2
from .ppo import PPO, Memory
from .resnet import resnet50
from .utils import get_patch


def standard_action_table(action_dim, device=None):
    """The reference's hard-coded n x n grids of (row, col) in [0,1] (ACT/models/gfv_net.py:272-307), row-major."""
    n = int(round(math.sqrt(action_dim)))
    if n * n != action_dim or n < 2:
        raise ValueError(f"action_dim {action_dim} is not a square grid")
    rows = [[iy / (n - 1), ix / (n - 1)] for iy in range(n) for ix in range(n)]
    return torch.tensor(rows, dtype=torch.float32, device=device)


class Glancer(nn.Module):
    """Global network fG (MobileNet-V2)."""

    def __init__(self, skip=False, num_classes=200):
        super().__init__()
        self.net = mobilenet_v2(pretrained=True)
        self.net.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(self.net.last_channel, num_classes))
        self.skip = skip

    def forward(self, input):
        return self.net.get_featmap(input)

    def predict(self, input):
        return self.net(input)

    @property
    def feature_dim(self):
        return self.net.feature_dim


class PatchSampler(nn.Module):
    def __init__(self, size=96, random=True):
        super().__init__()
        self.random, self.size = random, size

    def sample(self, imgs, action=None):
        if self.random:
            raise NotImplementedError("random patch sampling (stage-1 training) is outside the inference hot path")
        assert action is not None
        return get_patch(imgs, action, self.size)

    def random_sample(self, imgs):
        """Crop at uniformly random positions -- ACT/models/gfv_net.py:376-381 with ACT/models/utils.py:24-35: per
        image y then x from np.random.randint(0, size_range) on the HOST generator (same draw order as the reference),
        then one gather kernel instead of per-image slicing + torch.stack."""
        import numpy as np
        n, c, h, w = imgs.shape
        if self.size == h:
            return imgs
        yx = np.empty((n, 2), dtype=np.int32)
        for i in range(n):
            yx[i, 0] = np.random.randint(0, h - self.size)
            yx[i, 1] = np.random.randint(0, w - self.size)
        from ..engine import get_engine
        eng = get_engine(imgs.device)
        return eng.crop(imgs.contiguous(), yx=torch.from_numpy(yx).to(imgs.device), patch=self.size)

    def forward(self, *argv, **kwargs):
        raise NotImplementedError


class Focuser(nn.Module):
    """Local network fL (ResNet-50) + policy + patch sampler."""

    def __init__(self, size=96, random=True, policy_params=None, num_classes=200):
        super().__init__()
        self.net = resnet50(pretrained=True)
        self.net.fc = nn.Linear(self.net.fc.in_features, num_classes)
        self.patch_size, self.random = size, random
        self.patch_sampler = PatchSampler(self.patch_size, self.random)
        self.policy = None
        self.memory = Memory()
        if not self.random:
            assert policy_params is not None
            self.standard_actions_set = {a: standard_action_table(a) for a in (25, 36, 49, 64)}
            self.policy_feature_dim = policy_params["feature_dim"]
            self.policy_state_dim = policy_params["state_dim"]
            self.policy_action_dim = policy_params["action_dim"]
            self.policy_hidden_state_dim = policy_params["hidden_state_dim"]
            self.policy_conv = policy_params["policy_conv"]
            self.gpu = policy_params["gpu"]
            self.policy = PPO(self.policy_feature_dim, self.policy_state_dim, self.policy_action_dim,
                              self.policy_hidden_state_dim, self.policy_conv, self.gpu, gamma=policy_params["gamma"],
                              lr=policy_params["policy_lr"])

    def forward(self, *argv, **kwargs):
        """One focus step, reference call pattern (ACT/models/gfv_net.py:316-334)."""
        if self.random:
            raise NotImplementedError("random-patch focusing is stage-1 training, outside the inference hot path")
        action = self.policy.select_action(kwargs["state"], self.memory, kwargs["restart_batch"], kwargs["training"])
        standard_action, _ = self._get_standard_action(action)
        patch = self.patch_sampler.sample(kwargs["input"], standard_action)
        return self.net.get_featmap(patch, pooled=True), (None, standard_action)

    def random_patching(self, imgs):
        """Baseline feature of a random patch (stage-2 reward, ACT/models/gfv_net.py:336-338)."""
        patch = self.patch_sampler.random_sample(imgs)
        return self.net.get_featmap(patch, pooled=True), None

    def predict(self, input):
        return self.net(input)

    def update(self):
        raise NotImplementedError("PPO training is outside the inference hot path")

    def _get_standard_action(self, action):
        table = self.standard_actions_set[self.policy_action_dim]
        if table.device != action.device:
            table = table.to(action.device)
            self.standard_actions_set[self.policy_action_dim] = table
        return table[action], None

    @property
    def feature_dim(self):
        return self.net.feature_dim


class GRUHeadRunner:
    """RecurrentClassifier weights in kernel layout + the T-step schedule.

    The head is < 1 % of the FLOPs but every logit passes through it, so it runs at ~fp32 operand precision on the
    tensor core: features and the recurrent state stay fp32 and enter the GEMMs as split-precision fp16 rows
    [x_hi | x_lo | x_hi] against weights packed [W_hi | W_hi | W_lo] (engine.pack_conv_split)."""

    def __init__(self, clf, key=None):
        self.key = key
        g = clf.gru
        dev = g.weight_ih_l0.device
        self.hidden = clf.hidden_dim
        self.gru_ih = pack_conv_split(g.weight_ih_l0, g.bias_ih_l0, device=dev)
        self.gru_hh = pack_conv_split(g.weight_hh_l0, g.bias_hh_l0, device=dev, block_n=32)
        self.fc = pack_conv_split(clf.fc.weight, clf.fc.bias, device=dev)
        self.num_classes = clf.fc.weight.shape[0]
        self.logit_stride = (self.num_classes + 7) // 8 * 8

    def sequence(self, eng, feat32, b, t, logits, h0=None, h_out=None):
        """feat32 (B*T, F) fp32 rows b*T+t -> logits fp32 (B*T, logit_stride)."""
        hd = self.hidden
        feat3 = eng.split3(feat32)
        xg = eng.linear(feat3, self.gru_ih, out_f32=True)
        eng.release(feat3)
        hseq3 = eng.empty((b * t, 3 * hd), torch.float16)
        if eng.can_gru_sequence(b, hd):
            # small batch: the whole T-step recurrence is one persistent warp-reduction kernel
            eng.gru_sequence(xg, self.gru_hh, b, t, hseq3, h0=h0, h_out=h_out)
            eng.linear(hseq3, self.fc, out=logits, out_f32=True, out_stride=self.logit_stride)
            eng.release(xg)
            eng.release(hseq3)
            return
        if eng.can_gru_sequence_tc(b, hd, split=True):
            # bench batch sizes: the recurrence stays one persistent tensor-core launch (W_hh resident in smem)
            eng.gru_sequence_tc(xg, self.gru_hh, b, t, hseq3, h0=h0, h_out=h_out)
            eng.linear(hseq3, self.fc, out=logits, out_f32=True, out_stride=self.logit_stride)
            eng.release(xg)
            eng.release(hseq3)
            return
        h = eng.empty((b, hd), torch.float32) if h_out is None else h_out
        if h0 is None:
            eng.fill(h, 0.0)
        elif h0 is not h:
            h.copy_(h0)
        h3 = eng.split3(h)
        hg = eng.empty((b, 3 * hd), torch.float32)
        xg3, hs3 = xg.view(b, t, 3 * hd), hseq3.view(b, t, 3 * hd)
        for step in range(t):
            eng.linear(h3, self.gru_hh, out=hg, out_f32=True, out_stride=3 * hd)
            eng.gru_gates(xg3[:, step], t * 3 * hd, hg, h, h, h3, hs3[:, step], t * 3 * hd, split=True)
        eng.linear(hseq3, self.fc, out=logits, out_f32=True, out_stride=self.logit_stride)
        for tmp in (xg, h3, hg, hseq3) + ((h,) if h_out is None else ()):
            eng.release(tmp)


class RecurrentClassifier(nn.Module):
    """GRU classifier over [global | local] features (ACT/models/gfv_net.py:409-457)."""

    def __init__(self, seq_len, input_dim, batch_size, hidden_dim, num_classes, dropout, bias=True):
        super().__init__()
        self.seq_len, self.input_dim, self.hidden_dim = seq_len, input_dim, hidden_dim
        self.num_classes, self.batch_size = num_classes, batch_size
        self.gru = nn.GRU(input_size=input_dim, hidden_size=hidden_dim, bias=bias, batch_first=True)
        self.fc = nn.Linear(hidden_dim, num_classes)
        self.hx = None
        self.cx = None
        self.dropout = nn.Dropout(dropout)
        self._runner = None

    def runner(self):
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "GRUHeadRunner", lambda: GRUHeadRunner(self, key), key)
        return self._runner

    def _run(self, feature, h0, keep_state):
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("train-mode dropout is outside the inference hot path; call model.eval()")
        b, t, f = feature.shape
        eng = get_engine(feature.device)
        r = self.runner()
        feat32 = feature.contiguous().view(b * t, f).float()
        logits = torch.empty(b * t, r.logit_stride, dtype=torch.float32, device=feature.device)
        h = torch.empty(b, self.hidden_dim, dtype=torch.float32, device=feature.device)
        r.sequence(eng, feat32, b, t, logits, h0=h0, h_out=h)
        logits = logits[:, : self.num_classes].contiguous()
        last_out = logits.reshape(b, t, -1)[:, -1, :].reshape(b, -1)
        return logits, last_out, h

    def forward(self, feature):
        logits, last_out, _ = self._run(feature, None, False)
        return logits, last_out

    def single_forward(self, feature, reset=False, gpu=0):
        if reset:
            self.hx = torch.zeros(1, feature.shape[0], self.hidden_dim, device=feature.device)
        logits, last_out, h = self._run(feature, self.hx[0].contiguous(), True)
        self.hx = h[None]
        return logits, last_out

    def test_single_forward(self, feature, reset=False, gpu=0):
        if reset:
            self.hx = torch.zeros(1, feature.shape[0], self.hidden_dim, device=feature.device)
        logits, last_out, _ = self._run(feature, self.hx[0].contiguous(), False)
        return logits, last_out


class _FusedPlan:
    """Static buffers + recorded launch sequence of one stage-3 forward for a fixed batch size."""

    def __init__(self, model, b, t, h, w, g, device, share_scan):
        eng = get_engine(device)
        self.eng, self.b, self.t = eng, b, t
        self.graph = None
        p = model.patch_size
        self.input = torch.empty(b, 3 * t, h, w, dtype=torch.float32, device=device)
        self.scan = self.input if share_scan else torch.empty(b, 3 * t, g, g, dtype=torch.float32, device=device)
        clf = model.classifier.runner()
        fdim = model.classifier.input_dim
        self.feat32 = torch.zeros(b * t, fdim, dtype=torch.float32, device=device)
        self.logits = torch.zeros(b * t, clf.logit_stride, dtype=torch.float32, device=device)
        self.num_classes = clf.num_classes
        glancer = model.glancer.net.runner()
        focuser = model.focuser.net.runner()
        policy = model.focuser.policy.policy_old.runner()
        gdim = model.glancer.feature_dim if model.with_glancer else 0
        eng.begin_plan()
        try:
            m0 = eng.mark()
            fmap = glancer.run_chunked(eng, self.scan.view(b * t, 3, g, g), model.fg_chunk)
            if model.with_glancer:
                eng.avgpool(fmap, out_f32=self.feat32, out_f32_stride=fdim)
            m1 = eng.mark()
            self.yx, self.action_idx, self.action_yx = policy.rollout(eng, fmap, b, t, h, p)
            eng.release(fmap)
            m2 = eng.mark()
            focuser.run_pooled_chunked(eng, self.input.view(b * t, 3, h, w), self.feat32[:, gdim:], fdim,
                                       model.fl_chunk, yx=self.yx, patch=p, out_f32=True)
            m3 = eng.mark()
            clf.sequence(eng, self.feat32, b, t, self.logits)
            m4 = eng.mark()
            self.marks = {"fG": (m0, m1), "policy": (m1, m2), "fL": (m2, m3), "head": (m3, m4), "total": (m0, m4)}
        finally:
            self.plan = eng.end_plan()
        # name the plan's I/O so that the whole forward is ONE C-ABI call (af_gfv_forward)
        _lib.check(self.plan.lib.af_plan_bind_forward(
            self.plan.handle, self.input.data_ptr(), self.input.numel() * 4,
            None if share_scan else self.scan.data_ptr(), 0 if share_scan else self.scan.numel() * 4,
            self.logits.data_ptr(), b * t, clf.logit_stride, clf.num_classes, t, self.plan.workspace_bytes),
            "af_plan_bind_forward")
        self.keys = (_param_key(model.glancer.net), _param_key(model.focuser.net),
                     _param_key(model.focuser.policy.policy_old), _param_key(model.classifier))

    def run(self):
        if self.graph is not None:
            self.graph.replay()
            return
        self.plan.run(torch.cuda.current_stream(self.input.device).cuda_stream)

    def forward(self, inp, scan, logits, last_out):
        """The whole stage-3 forward as ONE C-ABI call (af_gfv_forward): device-to-device copies of the caller's
        tensors into the plan's static buffers (skipped when they already are those buffers), the recorded launch
        sequence, and contiguous (B*T, C) / (B, C) outputs."""
        from ctypes import c_void_p
        stream = torch.cuda.current_stream(self.input.device).cuda_stream
        _lib.check(self.plan.lib.af_gfv_forward(self.plan.handle, c_void_p(inp.data_ptr()),
                                                c_void_p(scan.data_ptr()) if scan is not None else None,
                                                c_void_p(logits.data_ptr()), c_void_p(last_out.data_ptr()),
                                                c_void_p(stream)), "af_gfv_forward")

    def capture_graph(self):
        """Capture one replay of the plan into a CUDA graph (small batches are launch-bound: ~170 launches of a few
        microseconds each).  Later run() calls replay the graph; stage_ms() is unavailable for graph replays (the
        timing marks are not captured)."""
        dev = self.input.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.plan.run(side.cuda_stream)          # warm-up outside capture (function attributes, lazy init)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.plan.run(torch.cuda.current_stream(dev).cuda_stream)
        self.graph = g
        return g

    def features(self):
        """[global | local] features (B*T, F) of the last replay, as the classifier consumed them."""
        return self.feat32

    def stage_ms(self):
        """Device time of each stage in the last completed replay (call after a synchronize)."""
        return {k: self.plan.elapsed_ms(a, b) for k, (a, b) in self.marks.items()}


class GFV(nn.Module):
    """Top class for adaptive inference on video."""

    def __init__(self, args):
        super().__init__()
        self.num_segments = args.num_segments
        self.num_class = args.num_classes
        self.rew = args.reward
        if args.dataset == "fcvid":
            assert args.num_classes == 239
        self.input_size, self.batch_size, self.patch_size = args.input_size, args.batch_size, args.patch_size
        self.glance_size = args.glance_size
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        self.with_glancer = args.with_glancer
        self.glancer = Glancer(num_classes=self.num_class)
        fm = math.ceil(args.glance_size / 32)
        policy_params = {
            "feature_dim": args.feature_map_channels, "state_dim": args.feature_map_channels * fm * fm,
            "action_dim": args.action_dim, "hidden_state_dim": args.hidden_state_dim,
            "policy_conv": args.policy_conv, "gpu": args.gpu, "continuous": args.continuous, "gamma": args.gamma,
            "policy_lr": args.policy_lr,
        }
        self.focuser = Focuser(args.patch_size, args.random_patch, policy_params, self.num_class)
        self.dropout = nn.Dropout(p=args.dropout)
        feat_dim = self.focuser.feature_dim + (self.glancer.feature_dim if self.with_glancer else 0)
        if args.consensus == "gru":
            self.classifier = RecurrentClassifier(seq_len=args.num_segments, input_dim=feat_dim,
                                                  batch_size=self.batch_size, hidden_dim=args.hidden_dim,
                                                  num_classes=args.num_classes, dropout=args.dropout)
        else:
            raise NotImplementedError("consensus='fc' (LinearCLassifier) is not used by any shipped configuration")
        self._plans = {}
        # sub-batch sizes (frames / patches) of the fused plan: intermediates of one sub-batch stay L2-resident
        import os
        self.fg_chunk = int(os.environ.get("AF_FG_CHUNK", "0")) or None
        self.fl_chunk = int(os.environ.get("AF_FL_CHUNK", "0")) or None

    def train(self, mode=True):
        # the reference's train() returns None, so `model.eval()` returns None (ACT/models/gfv_net.py:60-62)
        super().train(mode)
        return

    def train_mode(self, args):
        raise NotImplementedError("training stages are outside the inference hot path")

    def _require_eval(self):
        """The plans fold BatchNorm running statistics and drop dropout: that is only the reference's arithmetic in
        eval mode (the reference would use batch statistics and classifier dropout in train mode)."""
        if self.training or self.glancer.training or self.focuser.training or self.classifier.training:
            raise NotImplementedError("adafocus_b200 implements inference only: call model.eval() first "
                                      "(ACT/main_dist.py:316)")

    def invalidate_packed(self):
        """Drop every packed (kernel-layout) weight copy and recorded plan; call after editing weights in place through
        `.data` (load_state_dict / .to() are tracked automatically)."""
        for m in (self.glancer.net, self.focuser.net, self.classifier):
            invalidate_packed(m)
        if self.focuser.policy is not None:
            invalidate_packed(self.focuser.policy.policy_old)
            invalidate_packed(self.focuser.policy.policy)
        self._plans.clear()

    # ------------------------------------------------------------------ fused stage-3 inference
    def fused_plan(self, b, t, h, w, g, device, share_scan, slot=0):
        self._require_eval()
        key = (b, t, h, w, g, share_scan, str(device), slot)
        plan = self._plans.get(key)
        if plan is not None:
            cur = (_param_key(self.glancer.net), _param_key(self.focuser.net),
                   _param_key(self.focuser.policy.policy_old), _param_key(self.classifier))
            if cur != plan.keys:
                plan = None
        if plan is None:
            plan = _FusedPlan(self, b, t, h, w, g, device, share_scan)
            self._plans[key] = plan
        return plan

    def input_buffers(self, batch, device, height=None, width=None, glance=None, share_scan=True):
        """Static (input, scan) device buffers of the fused plan for `batch` clips: fill them (e.g. H2D copy) and call
        forward(input=buf, scan=buf, ...) to run with zero extra copies."""
        h = height or self.input_size
        w = width or self.input_size
        g = glance or self.glance_size
        plan = self.fused_plan(batch, self.num_segments, h, w, g, torch.device(device), share_scan)
        return plan.input, plan.scan

    def forward(self, *argv, **kwargs):
        if kwargs.get("backbone_pred"):
            raise NotImplementedError("backbone_pred (stage-0 pre-training) is outside the inference hot path")
        if not kwargs.get("one_step"):
            raise NotImplementedError("one_step=False (stage-1 random-patch training) is outside the hot path")
        if kwargs.get("training"):
            raise NotImplementedError("training=True samples actions for PPO; only inference is implemented")
        if self.focuser.random:
            raise NotImplementedError("random_patch=True has no policy rollout; stage-3 inference needs a policy")
        self._require_eval()
        inp, scan = kwargs["input"], kwargs["scan"]
        if not inp.is_cuda:
            raise RuntimeError("adafocus_b200 has no CPU path: inputs must be CUDA tensors")
        b, tc, h, w = inp.shape
        t = tc // 3
        g = scan.shape[-1]
        share = scan.data_ptr() == inp.data_ptr() and scan.shape == inp.shape
        plan = self.fused_plan(b, t, h, w, g, inp.device, share)
        if plan.graph is not None or not inp.is_contiguous() or not scan.is_contiguous() or inp.dtype != torch.float32:
            if inp.data_ptr() != plan.input.data_ptr():
                plan.input.copy_(inp)
            if not share and scan.data_ptr() != plan.scan.data_ptr():
                plan.scan.copy_(scan)
            plan.run()
            logits = plan.logits[:, : plan.num_classes].contiguous()
            last_out = logits.reshape(b, t, -1)[:, -1, :].reshape(b, -1)
        else:
            logits = torch.empty(b * t, plan.num_classes, dtype=torch.float32, device=inp.device)
            last_out = torch.empty(b, plan.num_classes, dtype=torch.float32, device=inp.device)
            plan.forward(inp, None if share else scan, logits, last_out)      # one C-ABI call: af_gfv_forward
        self.last_plan = plan
        return logits, last_out

    # ------------------------------------------------------------------ reference-style pieces
    def glance(self, input_prime):
        b, tc, h, w = input_prime.shape
        t = tc // 3
        fmap, vec = self.glancer(input_prime.contiguous().view(b * t, 3, h, w))
        _, c, fh, fw = fmap.shape
        return fmap.view(b, t, c, fh, fw), vec.view(b, t, -1)

    def one_step_act(self, img, global_feat_map, global_feat, restart_batch=False, training=True):
        """One focus step of the stage-2 (RL) validation loop -- ACT/models/gfv_net.py:160-210, driven by
        ACT/main_dist.py:343-366: policy step + crop + fL, the reward baseline feature, and the step-wise classifier
        (baseline through test_single_forward, which does not advance the GRU state, then single_forward)."""
        if training:
            raise NotImplementedError("training=True samples actions for PPO; only evaluation is implemented")
        b = img.shape[0]
        local_feat, pack = self.focuser(input=img, state=global_feat_map, restart_batch=restart_batch, training=False)
        patch_size_list, action_list = pack if pack is not None else (None, None)
        local = local_feat.view(b, -1)
        feature = torch.cat([global_feat, local], dim=1) if self.with_glancer else local
        feature = feature.unsqueeze(1)
        if self.rew == "random":
            base_local, _ = self.focuser.random_patching(img)
            base_local = base_local.view(b, -1)
        elif self.rew in ("padding", "prev", "conf"):
            base_local = torch.zeros(b, self.focuser.feature_dim, device=img.device)
        else:
            raise NotImplementedError
        baseline_feature = torch.cat([global_feat, base_local], dim=1) if self.with_glancer else base_local
        baseline_feature = baseline_feature.unsqueeze(1)
        baseline_logits, _ = self.classifier.test_single_forward(baseline_feature, reset=restart_batch)
        logits, last_out = self.classifier.single_forward(feature, reset=restart_batch)
        return logits, last_out, patch_size_list, action_list, baseline_logits

    @property
    def scale_size(self):
        return self.input_size * 256 // 224

    @property
    def crop_size(self):
        return self.input_size

    def get_augmentation(self, flip=True):
        # host-side PIL transforms stay with the reference's ops/ package (data loading is out of scope here)
        import torchvision
        from ops.transforms import GroupMultiScaleCrop, GroupRandomHorizontalFlip
        tf = [GroupMultiScaleCrop(self.input_size, [1, .875, .75, .66])]
        if flip:
            tf.append(GroupRandomHorizontalFlip(is_flow=False))
        return torchvision.transforms.Compose(tf)
