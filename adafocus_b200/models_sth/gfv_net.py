"""AdaFocus top module for the Something-Something V1/V2 tree -- mirror of STH/models/gfv_net.py.

Same classes / attributes / return values as the reference (GFV :16-225, Glancer :228-253, Focuser :256-434) so
`evaluate.py validate()` (STH/evaluate.py:165-226) runs on it unchanged: `glance()`, `action_stage2()` (policy patches +
random baseline patches, as evaluate.py drives it) and `action_stage3()` (policy patches only).  `forward_eval()` is
the fused plan for the whole video: TSM-MobileNet-V2 over B*T_g frames -> one policy step per video division ->
crop fused into the fL stem staging -> TSM-ResNet over B*T_f patches -> per-frame logits -> average consensus
(+ glancer consensus), all replayed natively with no host synchronisation."""
import math

import torch
from torch import nn

from ..engine import get_engine, pack_conv_split
from ..models.gfv_net import standard_action_table
from ..models.mobilenet import _param_key
from .basic_ops import ConsensusModule
from .mobilenetv2 import InvertedResidual, mobilenet_v2
from .ppo import PPO, Memory
from .ppo_continuous import PPO_Continuous
from .temporal_shift import TemporalShift
from .tsn import TSN
from .utils import get_patch


class Glancer(nn.Module):
    """Global network: MobileNet-V2 with the temporal shift plugged into its residual blocks (:238-241)."""

    def __init__(self, args, skip=False):
        super().__init__()
        self.net = mobilenet_v2(n_class=args.num_classes, pretrained=False)
        for m in self.net.modules():
            if isinstance(m, InvertedResidual) and len(m.conv) == 8 and m.use_res_connect:
                m.conv[0] = TemporalShift(m.conv[0], n_segment=args.num_segments_glancer, n_div=args.shift_div)
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

    def forward(self, *argv, **kwargs):
        raise NotImplementedError("Policy driven patch sampler not implemented.")


class Focuser(nn.Module):
    """Local network (TSN/TSM ResNet) + policy."""

    def __init__(self, size=96, random=False, policy_params=None, focuser_base_model_params=None):
        super().__init__()
        p = focuser_base_model_params
        self.net = TSN(num_segments=p["num_segments"], modality=p["modality"], base_model=p["base_model"],
                       partial_bn=p["partial_bn"], pretrain=p["pretrain"], is_shift=p["is_shift"],
                       shift_div=p["shift_div"], shift_place=p["shift_place"], fc_lr5=p["fc_lr5"],
                       temporal_pool=p["temporal_pool"], non_local=p["non_local"], print_spec=False)
        self.patch_size, self.random = size, random
        self.patch_sampler = PatchSampler(self.patch_size, self.random)
        self.policy = None
        self.memory = Memory()
        if not self.random:
            assert policy_params is not None
            self.patch_sizes = torch.Tensor([self.patch_size, 0])
            self.standard_actions_set = {a: standard_action_table(a) for a in (16, 25, 36, 49, 64, 81, 100)}
            self.policy_feature_dim = policy_params["feature_dim"]
            self.policy_state_dim = policy_params["state_dim"]
            self.policy_action_dim = policy_params["action_dim"]
            self.policy_hidden_state_dim = policy_params["hidden_state_dim"]
            self.policy_conv = policy_params["policy_conv"]
            self.gpu = policy_params["gpu"]
            self.ppo_continuous = policy_params["ppo_continuous"]
            if self.ppo_continuous:
                self.policy = PPO_Continuous(self.policy_feature_dim, self.policy_state_dim,
                                             self.policy_hidden_state_dim, self.policy_conv, self.gpu,
                                             gamma=policy_params["gamma"], lr=policy_params["policy_lr"],
                                             action_std=policy_params["action_std"], with_bn=policy_params["with_bn"])
            else:
                self.policy = PPO(self.policy_feature_dim, self.policy_state_dim, self.policy_action_dim,
                                  self.policy_hidden_state_dim, self.policy_conv, self.gpu,
                                  gamma=policy_params["gamma"], lr=policy_params["policy_lr"])

    def forward(self, *argv, **kwargs):
        """Policy step + crop of all frames of the division (STH/models/gfv_net.py:402-422): returns the patches."""
        if self.random:
            return self.random_patching(kwargs["input"])
        action = self.policy.select_action(kwargs["state"], self.memory, kwargs["restart_batch"], kwargs["training"])
        if self.ppo_continuous:
            standard_action = action
        else:
            standard_action, _ = self._get_standard_action(action)
        return get_patch(kwargs["input"], action_sequence=standard_action, patch_size=self.patch_size)

    def random_patching(self, imgs):
        """Baseline patches at uniformly random positions drawn from the CPU generator like the reference (:424-427)."""
        rand_index = torch.rand(imgs.size(0), 2)
        return get_patch(imgs, action_sequence=rand_index, patch_size=self.patch_size)

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


class _SthPlan:
    """Static buffers + recorded launch sequence of one whole-video evaluation for a fixed batch size."""

    def __init__(self, model, args, b, device):
        eng = get_engine(device)
        tg, tf, vd = args.num_segments_glancer, args.num_segments_focuser, args.video_div
        g, s, p = args.glance_size, model.input_size, model.focuser.patch_size
        self.b, self.tg, self.tf = b, tg, tf
        self.glancer_images = torch.empty(b, 3 * tg, g, g, dtype=torch.float32, device=device)
        self.focuser_images = torch.empty(b, 3 * tf, s, s, dtype=torch.float32, device=device)
        gl = model.glancer.net.runner()
        fl = model.focuser.net.runner()
        pol = model.focuser.policy.policy_old.runner()
        head = model._head_pack()
        c = model.num_class
        cs = (c + 7) // 8 * 8          # logits rows are padded to a multiple of 8 floats; the pad columns are never read
        eng.begin_plan()
        try:
            m0 = eng.mark()
            fmap = gl.run(eng, self.glancer_images.view(b * tg, 3, g, g))
            glog = gl.logits(eng, fmap, padded=True)                        # (B*T_g, cs) fp32
            gcons = eng.consensus_avg(glog, b, tg) if model.with_glancer else None
            m1 = eng.mark()
            self.yx, self.action_idx, self.action = pol.rollout(eng, fmap, b, vd, s, p)
            eng.release(fmap)
            m2 = eng.mark()
            lmap = fl.run(eng, self.focuser_images.view(b * tf, 3, s, s), yx=self.yx, patch=p, yx_div=tf // vd)
            lvec = eng.empty((b * tf, lmap.shape[-1]), torch.float32)
            eng.avgpool(lmap, out_f32=lvec, out_f32_stride=lmap.shape[-1])
            eng.release(lmap)
            lvec3 = eng.split3(lvec)                  # split-precision rows for the classifier (see GRUHeadRunner)
            eng.release(lvec)
            llog = eng.empty((b * tf, cs), torch.float32)
            eng.linear(lvec3, head, out=llog, out_f32=True, out_stride=cs)
            eng.release(lvec3)
            m3 = eng.mark()
            self.pred_padded = eng.consensus_avg(llog, b, tf, add=gcons)    # (B, cs)
            m4 = eng.mark()
            self.marks = {"fG": (m0, m1), "policy": (m1, m2), "fL": (m2, m3), "head": (m3, m4), "total": (m0, m4)}
        finally:
            self.plan = eng.end_plan()
        self.pred = self.pred_padded[:, :c]
        self.keys = model._weights_key()

    def run(self):
        self.plan.run(torch.cuda.current_stream(self.pred.device).cuda_stream)

    def stage_ms(self):
        return {k: self.plan.elapsed_ms(a, b_) for k, (a, b_) in self.marks.items()}


class GFV(nn.Module):
    """Top class for adaptive inference on video (Something-Something tree)."""

    def __init__(self, args):
        super().__init__()
        self.num_segments_glancer = args.num_segments_glancer
        self.num_segments_focuser = args.num_segments_focuser
        self.num_class = args.num_classes
        self.input_size = 224
        self.batch_size, self.patch_size = args.batch_size, args.patch_size
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        self.with_glancer = args.with_glancer
        self.glancer = Glancer(args)
        fpd = args.num_segments_glancer // args.video_div
        fm = math.ceil(args.glance_size / 32)
        policy_params = {
            "feature_dim": args.feature_map_channels * fpd, "state_dim": args.feature_map_channels * fpd * fm * fm,
            "action_dim": args.action_dim, "hidden_state_dim": args.hidden_state_dim,
            "policy_conv": args.policy_conv, "gpu": args.gpu, "ppo_continuous": args.ppo_continuous,
            "gamma": args.gamma, "policy_lr": args.policy_lr, "action_std": args.action_std,
            "with_bn": args.actorcritic_with_bn,
        }
        base = {"num_segments": args.num_segments_focuser, "modality": args.modality, "base_model": args.base_model,
                "partial_bn": args.partial_bn, "pretrain": args.pretrain, "is_shift": args.is_shift,
                "shift_div": args.shift_div, "shift_place": args.shift_place, "fc_lr5": args.fc_lr5,
                "temporal_pool": args.temporal_pool, "non_local": args.non_local}
        self.focuser = Focuser(args.patch_size, args.random_patch, policy_params, base)
        self.dropout = nn.Dropout(p=args.dropout)
        self.classifier = nn.Linear(in_features=self.focuser.feature_dim, out_features=args.num_classes)
        self.consensus = ConsensusModule(consensus_type="avg")
        self._plans = {}
        self._head = None

    def train(self, mode=True):
        super().train(mode)
        return

    # ------------------------------------------------------------------ helpers
    def _head_pack(self):
        key = _param_key(self.classifier)
        if self._head is None or self._head[0] != key:
            self._head = (key, pack_conv_split(self.classifier.weight, self.classifier.bias,
                                               device=self.classifier.weight.device))
        return self._head[1]

    def _weights_key(self):
        return (_param_key(self.glancer.net), id(self.focuser.net.base_model), _param_key(self.focuser.net.base_model),
                _param_key(self.focuser.policy.policy_old), _param_key(self.classifier))

    def _local_logits(self, patches):
        """(N,3,P,P) fp32 -> classifier(dropout(fL(patches))) : (N, C) fp32 (eval: dropout is the identity)."""
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("train-mode dropout is outside the inference hot path; call model.eval()")
        eng = get_engine(patches.device)
        fmap = self.focuser.net.runner().run(eng, patches.contiguous())
        n, h, w, c = fmap.shape
        vec = torch.empty(n, c, dtype=torch.float32, device=patches.device)
        eng.avgpool(fmap, out_f32=vec, out_f32_stride=c)
        cs = (self.num_class + 7) // 8 * 8
        out = torch.empty(n, cs, dtype=torch.float32, device=patches.device)
        eng.linear(eng.split3(vec), self._head_pack(), out=out, out_f32=True, out_stride=cs)
        return out[:, : self.num_class]

    # ------------------------------------------------------------------ reference API
    def forward(self, *argv, **kwargs):
        raise NotImplementedError("GFV.forward of the STH tree is the stage-1 (random patch) training path; inference "
                                  "goes through glance() + action_stage2/3() or forward_eval()")

    def glance(self, input_prime):
        b, tc, h, w = input_prime.shape
        t = tc // 3
        fmap, logits = self.glancer(input_prime.contiguous().view(b * t, 3, h, w))
        _, c, fh, fw = fmap.shape
        return fmap.view(b, t, c, fh, fw), logits.reshape(b, t, -1)

    def adjust_patch_size(self, patch_size):
        self.focuser.patch_size = patch_size
        self.focuser.patch_sampler.size = patch_size
        self.patch_size = patch_size

    def _stage(self, focuser_image, global_feat_map, global_feat_logit, focus_time_step, args, prev_local_patch,
               training, with_baseline):
        if training:
            raise NotImplementedError("training=True samples actions for PPO; only inference is implemented")
        fpd_g = args.num_segments_glancer // args.video_div
        fpd_f = args.num_segments_focuser // args.video_div
        b, _, _, h, w = focuser_image.shape
        p = args.patch_size
        cur_image = focuser_image[:, focus_time_step * fpd_f:(focus_time_step + 1) * fpd_f].reshape(b, -1, h, w)
        _, _, c, fh, fw = global_feat_map.shape
        cur_map = global_feat_map[:, focus_time_step * fpd_g:(focus_time_step + 1) * fpd_g].reshape(b, -1, fh, fw)
        cur_patch = self.focuser(input=cur_image, state=cur_map, training=False,
                                 restart_batch=focus_time_step == 0).view(b, fpd_f, 3, p, p)
        frames_so_far = fpd_f * (focus_time_step + 1)
        outs = []
        variants = [cur_patch]
        if with_baseline:
            variants.append(self.focuser.random_patching(cur_image).view(b, fpd_f, 3, p, p))
        gcons = self.consensus(global_feat_logit).squeeze(1) if self.with_glancer else None
        patches_out = None
        for i, cur in enumerate(variants):
            patch = cur if prev_local_patch is None else torch.cat([prev_local_patch, cur], dim=1)
            if i == 0:
                patches_out = patch
            logit = self._local_logits(patch.reshape(-1, 3, p, p)).reshape(b, frames_so_far, -1)
            total = self.consensus(logit).squeeze(1)
            outs.append(total + gcons if gcons is not None else total)
        return outs, patches_out

    def action_stage2(self, focuser_image, global_feat_map, global_feat_logit, focus_time_step, args,
                      prev_local_patch=None, training=True):
        """-> (total_logit (B,C), baseline_logit (B,C), local_patch) -- STH/models/gfv_net.py:136-188."""
        (total, baseline), patch = self._stage(focuser_image, global_feat_map, global_feat_logit, focus_time_step, args,
                                               prev_local_patch, training, True)
        return total, baseline, patch

    def action_stage3(self, focuser_image, global_feat_map, global_feat_logit, focus_time_step, args,
                      prev_local_patch=None):
        """-> (total_logit (B,C), local_patch) -- STH/models/gfv_net.py:190-225."""
        (total,), patch = self._stage(focuser_image, global_feat_map, global_feat_logit, focus_time_step, args,
                                      prev_local_patch, False, False)
        return total, patch

    # ------------------------------------------------------------------ fused whole-video evaluation
    def eval_plan(self, args, batch, device, slot=0):
        key = (batch, args.num_segments_glancer, args.num_segments_focuser, args.video_div, args.glance_size,
               self.focuser.patch_size, str(device), slot)
        plan = self._plans.get(key)
        if plan is not None and plan.keys != self._weights_key():
            plan = None
        if plan is None:
            plan = _SthPlan(self, args, batch, torch.device(device))
            self._plans[key] = plan
        return plan

    def forward_eval(self, glancer_images, focuser_images, args):
        """glancer_images (B, 3*T_g, g, g), focuser_images (B, 3*T_f, 224, 224) or (B, T_f, 3, 224, 224) fp32 CUDA ->
        pred (B, C): the final `pred` of evaluate.py's loop (action_stage3 semantics, no random baseline)."""
        if self.training:
            raise NotImplementedError("call model.eval() first (reference: STH/evaluate.py:175-177)")
        b = glancer_images.shape[0]
        plan = self.eval_plan(args, b, glancer_images.device)
        if glancer_images.data_ptr() != plan.glancer_images.data_ptr():
            plan.glancer_images.copy_(glancer_images.reshape(plan.glancer_images.shape))
        if focuser_images.data_ptr() != plan.focuser_images.data_ptr():
            plan.focuser_images.copy_(focuser_images.reshape(plan.focuser_images.shape))
        plan.run()
        self.last_plan = plan
        return plan.pred.clone()

    @property
    def scale_size(self):
        return self.input_size * 256 // 224

    @property
    def crop_size(self):
        return self.input_size

    def get_augmentation(self, flip=True):
        import torchvision
        from ops.transforms import GroupMultiScaleCrop, GroupRandomHorizontalFlip
        tf = [GroupMultiScaleCrop(self.input_size, [1, .875, .75, .66])]
        if flip:
            tf.append(GroupRandomHorizontalFlip(is_flow=False))
        return torchvision.transforms.Compose(tf)
