"""Continuous-action policy of the Something-Something tree -- mirror of STH/models/ppo_continuous.py
(ActorCritic :27-109, PPO_Continuous :138-164).  Inference branch only: act() returns the action mean
(:106-107).  The reference also draws `dist.sample()` on the CUDA generator in eval (:96-98) and discards it; that
draw does not influence any output (the random baseline patches use the CPU generator, STH/models/gfv_net.py:424) and
is not reproduced."""
import math

import torch
from torch import nn

from ..engine import get_engine
from ..models.mobilenet import _param_key
from ..packcache import cached_runner
from ..models.ppo import Memory, PolicyRunner  # noqa: F401  (Memory re-exported like the reference module)


class ActorCritic(nn.Module):
    def __init__(self, feature_dim, state_dim, hidden_state_dim=1024, policy_conv=True, action_std=0.1, with_bn=False):
        super().__init__()
        if not policy_conv:
            raise NotImplementedError("only the policy_conv=True encoder is implemented")
        flat = int(state_dim * 64 / feature_dim)
        if with_bn:
            self.state_encoder = nn.Sequential(nn.Conv2d(feature_dim, 64, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(),
                                               nn.Flatten(), nn.Linear(flat, hidden_state_dim),
                                               nn.BatchNorm1d(hidden_state_dim), nn.ReLU())
        else:
            self.state_encoder = nn.Sequential(nn.Conv2d(feature_dim, 64, 1, bias=False), nn.ReLU(), nn.Flatten(),
                                               nn.Linear(flat, hidden_state_dim), nn.ReLU())
        self.gru = nn.GRU(hidden_state_dim, hidden_state_dim, batch_first=False)
        self.actor = nn.Sequential(nn.Linear(hidden_state_dim, 2), nn.Sigmoid())
        self.critic = nn.Sequential(nn.Linear(hidden_state_dim, 1))
        self.action_var = torch.full((2,), action_std)
        self.hidden_state_dim, self.policy_conv, self.feature_dim = hidden_state_dim, policy_conv, feature_dim
        self.feature_ratio = int(math.sqrt(state_dim / feature_dim))
        self._runner = None

    def forward(self):
        raise NotImplementedError

    def runner(self):
        key = _param_key(self)
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self, "PolicyRunner", lambda: PolicyRunner(self, key), key)
        return self._runner

    def act(self, state_ini, memory, restart_batch=False, training=False):
        """state_ini (B, fps*1280, h, w) fp32 -> action mean (B, 2) fp32 in (0,1)."""
        if training:
            raise NotImplementedError("sampling actions (PPO training) is outside the inference hot path")
        if self.training:
            raise NotImplementedError("the policy's BatchNorm needs eval mode for inference (STH/evaluate.py:176-177)")
        eng = get_engine(state_ini.device)
        b = state_ini.shape[0]
        if restart_batch:
            del memory.hidden[:]
            memory.hidden.append(torch.zeros(1, b, self.hidden_state_dim, device=state_ini.device))
        r = self.runner()
        _, ctot, h, w = state_ini.shape
        fmap = eng.nchw_to_nhwc_f16(state_ini.contiguous().view(b * r.fps, ctot // r.fps, h, w))
        h_prev = memory.hidden[-1][0].contiguous()
        h_new, action = r.step(eng, fmap, h_prev)
        memory.hidden.append(h_new[None])
        return action.detach()

    def evaluate(self, state, action):
        raise NotImplementedError("PPO training is outside the inference hot path")


class PPO_Continuous:
    """Plain holder of `policy` / `policy_old` (not an nn.Module in the STH tree, so the policy travels under the
    checkpoint's separate 'policy' key)."""

    def __init__(self, feature_dim, state_dim, hidden_state_dim, policy_conv, gpu=0, action_std=0.1, lr=0.0003,
                 betas=(0.9, 0.999), gamma=0.7, K_epochs=1, eps_clip=0.2, with_bn=False):
        self.lr, self.betas, self.gamma, self.eps_clip, self.K_epochs = lr, betas, gamma, eps_clip, K_epochs
        self.policy = ActorCritic(feature_dim, state_dim, hidden_state_dim, policy_conv, action_std, with_bn=with_bn)
        self.policy_old = ActorCritic(feature_dim, state_dim, hidden_state_dim, policy_conv, action_std,
                                      with_bn=with_bn)
        self.policy_old.load_state_dict(self.policy.state_dict())
        if torch.cuda.is_available():
            self.policy.cuda(gpu)
            self.policy_old.cuda(gpu)

    def select_action(self, state, memory, restart_batch=False, training=True):
        return self.policy_old.act(state, memory, restart_batch, training)

    def update(self, memory):
        raise NotImplementedError("PPO training is outside the inference hot path")
