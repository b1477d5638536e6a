"""Recurrent patch-selection policy (pi): parameter tree + engine runner.

Mirror of ACT/models/ppo.py (Memory :9-24, ActorCritic :27-96, PPO :125-145).  Only the inference branch of
`act()` is implemented (argmax action, :94); sampling / evaluate / update belong to PPO training and raise.
The engine evaluates the state encoder and the GRU input projection for all T frames as two batched GEMMs, then
runs the T recurrent steps (h W_hh^T GEMM + gate kernel) back to back on the stream without returning to the host.
"""
import math

import torch
from torch import nn

from ..engine import AF_ACT_NONE, AF_ACT_RELU, fold_bn, get_engine, host, pack_conv


class Memory:
    """Rollout buffers with the reference's attribute names (ACT/models/ppo.py:9-24)."""

    def __init__(self):
        self.actions, self.states, self.logprobs, self.rewards, self.is_terminals, self.hidden = [], [], [], [], [], []

    def clear_memory(self):
        for lst in (self.actions, self.states, self.logprobs, self.rewards, self.is_terminals, self.hidden):
            del lst[:]


class ActorCritic(nn.Module):
    def __init__(self, feature_dim, state_dim, action_dim, hidden_state_dim=1024, policy_conv=True):
        super().__init__()
        if policy_conv:
            self.state_encoder = nn.Sequential(
                nn.Conv2d(feature_dim, 32, 1, bias=False), nn.ReLU(), nn.Flatten(),
                nn.Linear(int(state_dim * 32 / feature_dim), hidden_state_dim), nn.ReLU())
        else:
            self.state_encoder = nn.Sequential(nn.Linear(state_dim, 2048), nn.ReLU(),
                                               nn.Linear(2048, hidden_state_dim), nn.ReLU())
        self.gru = nn.GRU(hidden_state_dim, hidden_state_dim, batch_first=False)
        self.actor = nn.Sequential(nn.Linear(hidden_state_dim, action_dim), nn.Softmax(dim=-1))
        self.critic = nn.Sequential(nn.Linear(hidden_state_dim, 1))
        self.hidden_state_dim, self.action_dim = hidden_state_dim, action_dim
        self.policy_conv, self.feature_dim = policy_conv, feature_dim
        self.feature_ratio = int(math.sqrt(state_dim / feature_dim))
        self._runner = None

    def forward(self):
        raise NotImplementedError

    def runner(self):
        from ..packcache import cached_runner
        from .mobilenet import _param_key
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "PolicyRunner", lambda: PolicyRunner(self, key), key)
        return self._runner

    def act(self, state_ini, memory, restart_batch=False, training=True):
        """One policy step on (B, C, h, w) fp32 glance features -> action indices (B,) int64; the hidden state lives
        in memory.hidden like the reference (ACT/models/ppo.py:67-96)."""
        if training:
            raise NotImplementedError("sampling actions (PPO training) is outside the inference hot path")
        eng = get_engine(state_ini.device)
        b = state_ini.shape[0]
        if restart_batch:
            del memory.hidden[:]
            memory.hidden.append(torch.zeros(1, b, self.hidden_state_dim, device=state_ini.device))
        r = self.runner()
        fmap = eng.nchw_to_nhwc_f16(state_ini.contiguous())
        h_prev = memory.hidden[-1][0].contiguous()
        h_new, idx = r.step(eng, fmap, h_prev)
        memory.hidden.append(h_new[None])
        return idx.long()

    def evaluate(self, state, action):
        raise NotImplementedError("PPO training is outside the inference hot path")


class PolicyRunner:
    """Policy weights in kernel layout.  Handles both trees' encoders: ACT (conv 1x1 -> ReLU -> Linear -> ReLU,
    ACT/models/ppo.py:33-39) and STH (conv 1x1 -> BN2d -> ReLU -> Linear -> BN1d -> ReLU over `fps` glance maps
    stacked on channels, STH/models/ppo_continuous.py:33-42, STH/models/gfv_net.py:145-147), and both heads:
    softmax/argmax over a square action grid or the continuous sigmoid head (STH/models/ppo_continuous.py:61-63)."""

    def __init__(self, ac, key=None, map_channels=1280):
        if not ac.policy_conv:
            raise NotImplementedError("only the policy_conv=True encoder (MobileNet-V2 glancer) is implemented")
        self.key = key
        enc = list(ac.state_encoder)
        conv = enc[0]
        dev = conv.weight.device
        self.hidden = ac.hidden_state_dim
        self.enc_c = conv.weight.shape[0]
        cin_total = conv.weight.shape[1]
        self.map_channels = min(map_channels, cin_total)
        self.fps = cin_total // self.map_channels            # glance maps per policy state
        bn2 = next((m for m in enc if isinstance(m, nn.BatchNorm2d)), None)
        lin = next(m for m in enc if isinstance(m, nn.Linear))
        bn1 = next((m for m in enc if isinstance(m, nn.BatchNorm1d)), None)
        # conv over [t*C + c] channels == a (fps x 1) convolution over the frame axis of (M, fps, h*w, C)
        w4 = host(conv.weight).reshape(self.enc_c, self.fps, self.map_channels).permute(0, 2, 1)[..., None]
        sc, bi = (None, None)
        if bn2 is not None:
            sc, bi = fold_bn(bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var, bn2.eps)
        self.enc_conv = pack_conv(w4.contiguous(), sc, bi, act=AF_ACT_RELU, device=dev)
        hw = lin.weight.shape[1] // self.enc_c
        self.hw = hw
        j = torch.arange(hw * self.enc_c)
        perm = (j % self.enc_c) * hw + (j // self.enc_c)         # NHWC-flatten index -> NCHW-flatten index
        if bn1 is not None:
            s1, b1 = fold_bn(bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, bn1.eps)
            lb = host(lin.bias) * s1 + b1 if lin.bias is not None else b1
            self.enc_fc = pack_conv(lin.weight, s1, lb, act=AF_ACT_RELU, device=dev, cin_perm=perm)
        else:
            self.enc_fc = pack_conv(lin.weight, None, lin.bias, act=AF_ACT_RELU, device=dev, cin_perm=perm)
        g = ac.gru
        self.gru_ih = pack_conv(g.weight_ih_l0, None, g.bias_ih_l0, device=dev)
        self.gru_hh = pack_conv(g.weight_hh_l0, None, g.bias_hh_l0, device=dev, block_n=32)
        self.actor = pack_conv(ac.actor[0].weight, None, ac.actor[0].bias, device=dev)
        self.continuous = isinstance(ac.actor[1], nn.Sigmoid)
        self.action_dim = ac.actor[0].weight.shape[0]
        self.logit_stride = (self.action_dim + 7) // 8 * 8

    def encode(self, eng, fmap):
        """fmap (M*fps, h, w, C) NHWC fp16, fps consecutive maps per state -> W_ih s + b_ih, fp32 (M, 3H)."""
        mf, h, w, c = fmap.shape
        m = mf // self.fps
        e = eng.empty((m, 1, h * w, self.enc_c), torch.float16)
        eng.conv(fmap, self.enc_conv, out=e, out_stride=self.enc_c, shape=(m, self.fps, h * w, c, c))
        s = eng.linear(e.view(m, h * w * self.enc_c), self.enc_fc)           # (M,H) fp16
        eng.release(e)
        xg = eng.linear(s, self.gru_ih, out_f32=True)                        # (M,3H) fp32
        eng.release(s)
        return xg

    def head(self, eng, hseq16, rows, img_h, patch, action_idx, action_yx, yx):
        logits = eng.empty((rows, self.logit_stride), torch.float32)
        eng.linear(hseq16, self.actor, out=logits, out_f32=True, out_stride=self.logit_stride)
        if self.continuous:
            eng.policy_head_continuous(logits, img_h, patch, action_yx, yx)
        else:
            eng.policy_head(logits, self.action_dim, img_h, patch, action_idx, action_yx, yx)
        eng.release(logits)

    def rollout(self, eng, fmap, b, t, img_h, patch):
        """All t policy steps for b clips; states are row-major (b*t + step).  Returns (yx int32 (b*t,2),
        action_idx int32 (b*t,) [discrete only], action fp32 (b*t,2))."""
        hd = self.hidden
        xg = self.encode(eng, fmap)
        hseq16 = eng.empty((b * t, hd), torch.float16)
        if eng.can_gru_sequence(b, hd):
            # small batch: the whole T-step recurrence is one persistent warp-reduction kernel
            eng.gru_sequence(xg, self.gru_hh, b, t, hseq16)
            tmps = (xg, hseq16)
        elif eng.can_gru_sequence_tc(b, hd):
            # up to 64 clips: one persistent tensor-core launch, W_hh resident in shared memory
            eng.gru_sequence_tc(xg, self.gru_hh, b, t, hseq16)
            tmps = (xg, hseq16)
        else:
            h = eng.empty((b, hd), torch.float32)
            eng.fill(h, 0.0)
            h16 = eng.f32_to_f16(h)
            hg = eng.empty((b, 3 * hd), torch.float32)
            xg3 = xg.view(b, t, 3 * hd)
            hs3 = hseq16.view(b, t, hd)
            for step in range(t):
                eng.linear(h16, self.gru_hh, out=hg, out_f32=True, out_stride=3 * hd)
                eng.gru_gates(xg3[:, step], t * 3 * hd, hg, h, h, h16, hs3[:, step], t * hd)
            tmps = (xg, h, h16, hg, hseq16)
        yx = eng.empty((b * t, 2), torch.int32)
        idx = None if self.continuous else eng.empty((b * t,), torch.int32)
        ayx = eng.empty((b * t, 2), torch.float32)
        self.head(eng, hseq16, b * t, img_h, patch, idx, ayx, yx)
        for tmp in tmps:
            eng.release(tmp)
        return yx, idx, ayx

    def step(self, eng, fmap, h_prev):
        """Single step (reference-style call pattern): returns (h_new fp32 (B,H), action) where action is the int32
        index (B,) for the discrete head or the fp32 (B,2) action mean for the continuous head."""
        b, hd = h_prev.shape
        xg = self.encode(eng, fmap)
        h16 = eng.f32_to_f16(h_prev)
        hg = eng.linear(h16, self.gru_hh, out_f32=True)
        h_new = torch.empty_like(h_prev)
        hn16 = torch.empty(b, hd, dtype=torch.float16, device=h_prev.device)
        eng.gru_gates(xg, 3 * hd, hg, h_prev, h_new, hn16)
        if self.continuous:
            act = torch.empty(b, 2, dtype=torch.float32, device=h_prev.device)
            self.head(eng, hn16, b, 2, 1, None, act, None)
            return h_new, act
        idx = torch.empty(b, dtype=torch.int32, device=h_prev.device)
        self.head(eng, hn16, b, 2, 1, idx, None, None)
        return h_new, idx


class PPO(nn.Module):
    """Holds `policy` and `policy_old` (ACT/models/ppo.py:125-145); update() is PPO training -> not implemented."""

    def __init__(self, feature_dim, state_dim, action_dim, hidden_state_dim, policy_conv, gpu=0, lr=0.0003,
                 betas=(0.9, 0.999), gamma=0.7, K_epochs=1, eps_clip=0.2):
        super().__init__()
        self.lr, self.betas, self.gamma, self.eps_clip, self.K_epochs = lr, betas, gamma, eps_clip, K_epochs
        self.policy = ActorCritic(feature_dim, state_dim, action_dim, hidden_state_dim, policy_conv)
        self.policy_old = ActorCritic(feature_dim, state_dim, action_dim, hidden_state_dim, policy_conv)
        self.policy_old.load_state_dict(self.policy.state_dict())

    def select_action(self, state, memory, restart_batch=False, training=True):
        return self.policy_old.act(state, memory, restart_batch, training)

    def update(self, memory):
        raise NotImplementedError("PPO training is outside the inference hot path")
