"""Recurrent patch-selection policy (pi): parameter tree + engine runner.

Mirror of ACT/models/ppo.py (Memory :9-24, ActorCritic :27-96, PPO :125-145).  Only the inference branch of
`act()` is implemented (argmax action, :94); sampling / evaluate / update belong to PPO training and raise.
The engine evaluates the state encoder and the GRU input projection for all T frames as two batched GEMMs, then
runs the T recurrent steps (h W_hh^T GEMM + gate kernel) back to back on the stream without returning to the host.
"""
import math

import torch
from torch import nn

from ..engine import AF_ACT_NONE, AF_ACT_RELU, get_engine, pack_conv


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
        from .mobilenet import _param_key
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = PolicyRunner(self, key)
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
    def __init__(self, ac, key=None):
        if not ac.policy_conv:
            raise NotImplementedError("only the policy_conv=True encoder (MobileNet-V2 glancer) is implemented")
        self.key = key
        enc = ac.state_encoder
        dev = enc[0].weight.device
        self.hidden = ac.hidden_state_dim
        self.action_dim = ac.action_dim
        self.enc_c = enc[0].weight.shape[0]                      # 32
        self.enc_conv = pack_conv(enc[0].weight, act=AF_ACT_RELU, device=dev)
        lin = enc[3]
        hw = lin.weight.shape[1] // self.enc_c                   # 49
        self.hw = hw
        j = torch.arange(hw * self.enc_c)
        perm = (j % self.enc_c) * hw + (j // self.enc_c)         # NHWC-flatten index -> NCHW-flatten index
        self.enc_fc = pack_conv(lin.weight, None, lin.bias, act=AF_ACT_RELU, device=dev, cin_perm=perm)
        g = ac.gru
        self.gru_ih = pack_conv(g.weight_ih_l0, None, g.bias_ih_l0, device=dev)
        self.gru_hh = pack_conv(g.weight_hh_l0, None, g.bias_hh_l0, device=dev, block_n=32)
        self.actor = pack_conv(ac.actor[0].weight, None, ac.actor[0].bias, device=dev)
        self.logit_stride = (self.action_dim + 7) // 8 * 8

    def encode(self, eng, fmap):
        """fmap (M,h,w,C) NHWC fp16 -> GRU input pre-activations W_ih s + b_ih, fp32 (M, 3H)."""
        m, h, w, c = fmap.shape
        e = eng.conv(fmap, self.enc_conv)                                    # (M,h,w,32)
        s = eng.linear(e.view(m, h * w * self.enc_c), self.enc_fc)           # (M,H) fp16
        eng.release(e)
        xg = eng.linear(s, self.gru_ih, out_f32=True)                        # (M,3H) fp32
        eng.release(s)
        return xg

    def head(self, eng, hseq16, rows, img_h, patch, action_idx, action_yx, yx):
        logits = eng.empty((rows, self.logit_stride), torch.float32)
        eng.linear(hseq16, self.actor, out=logits, out_f32=True, out_stride=self.logit_stride)
        eng.policy_head(logits, self.action_dim, img_h, patch, action_idx, action_yx, yx)
        eng.release(logits)

    def rollout(self, eng, fmap, b, t, img_h, patch):
        """All T steps for B clips; fmap rows are frame-major (b*T + t).  Returns (yx int32 (B*T,2),
        action_idx int32 (B*T,), standard action fp32 (B*T,2))."""
        hd = self.hidden
        xg = self.encode(eng, fmap)
        h = eng.empty((b, hd), torch.float32)
        eng.fill(h, 0.0)
        h16 = eng.f32_to_f16(h)
        hg = eng.empty((b, 3 * hd), torch.float32)
        hseq16 = eng.empty((b * t, hd), torch.float16)
        xg3 = xg.view(b, t, 3 * hd)
        hs3 = hseq16.view(b, t, hd)
        for step in range(t):
            eng.linear(h16, self.gru_hh, out=hg, out_f32=True, out_stride=3 * hd)
            eng.gru_gates(xg3[:, step], t * 3 * hd, hg, h, h, h16, hs3[:, step], t * hd)
        yx = eng.empty((b * t, 2), torch.int32)
        idx = eng.empty((b * t,), torch.int32)
        ayx = eng.empty((b * t, 2), torch.float32)
        self.head(eng, hseq16, b * t, img_h, patch, idx, ayx, yx)
        for tmp in (xg, h, h16, hg, hseq16):
            eng.release(tmp)
        return yx, idx, ayx

    def step(self, eng, fmap, h_prev):
        """Single step (reference-style call pattern): returns (h_new fp32 (B,H), action index int32 (B,))."""
        b, hd = h_prev.shape
        xg = self.encode(eng, fmap)
        h16 = eng.f32_to_f16(h_prev)
        hg = eng.linear(h16, self.gru_hh, out_f32=True)
        h_new = torch.empty_like(h_prev)
        hn16 = torch.empty(b, hd, dtype=torch.float16, device=h_prev.device)
        eng.gru_gates(xg, 3 * hd, hg, h_prev, h_new, hn16)
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
