"""Networks of the RL stack (reference: RL/dist_rl/model.py:6-45, RL/network.py)."""
from __future__ import annotations

import torch
from torch import nn

from ..models.relu_nn import FFReLUNet as _FFReLUNet


class FFReLUNet(_FFReLUNet):
    """ReLU MLP that also accepts numpy observations (float32 coercion, reference :42-43)."""

    def __init__(self, shape, dtype=None):
        super().__init__(shape, dtype=dtype, coerce_numpy=True)


class ActorCritic(nn.Module):
    """One graph node's policy and value networks as a single module, so a node's consensus
    variable is the concatenation [actor || critic] — the reference keeps two parameter vectors
    with separate duals but identical rho (RL/dist_rl/dinnoPPO.py:11-19), which is the same
    update on the concatenated vector."""

    def __init__(self, obs_dim, act_dim, hidden=(64, 64, 64), dtype=None):
        super().__init__()
        self.actor = FFReLUNet([obs_dim, *hidden, act_dim], dtype=dtype)
        self.critic = FFReLUNet([obs_dim, *hidden, 1], dtype=dtype)

    def forward(self, obs):
        return self.actor(obs), self.critic(obs)
