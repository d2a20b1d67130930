"""Sine-feature ("Fourier"/SIREN first layer) implicit density network
(reference: models/fourier_nn.py:14-62).

Unlike the reference this module does not switch the global default dtype at
import time; pass ``dtype=`` instead.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .spec import MLPSpec


class SIRENLayer(nn.Module):
    """``sin(scale * (W x + b))`` with W ~ U(-sqrt(6/out), sqrt(6/out))."""

    def __init__(self, in_features, out_features, scale=1.0, dtype=None):
        super().__init__()
        self.in_features, self.out_features, self.scale = in_features, out_features, scale
        kw = {} if dtype is None else {"dtype": dtype}
        self.linear = nn.Linear(in_features, out_features, **kw)
        self.init_weights()

    def init_weights(self):
        bound = math.sqrt(6.0 / self.out_features)
        with torch.no_grad():
            self.linear.weight.uniform_(-bound, bound)

    def forward(self, x):
        return torch.sin(self.scale * self.linear(x))


class FourierNet(nn.Module):
    """SIREN layer -> ReLU -> (Linear -> ReLU)* -> Linear -> Sigmoid."""

    def __init__(self, shape, scale=1.0, dtype=None):
        super().__init__()
        shape = tuple(int(s) for s in shape)
        self.spec = MLPSpec(shape, first="sin_relu", hidden="relu", last="sigmoid", scale=float(scale))
        kw = {} if dtype is None else {"dtype": dtype}
        mods = []
        n_layers = len(shape) - 1
        for li in range(n_layers):
            if li == 0:
                mods.append(SIRENLayer(shape[0], shape[1], scale=scale, dtype=dtype))
            else:
                mods.append(nn.Linear(shape[li], shape[li + 1], **kw))
            mods.append(nn.Sigmoid() if li == n_layers - 1 else nn.ReLU(inplace=True))
        self.seq = nn.Sequential(*mods)

    def forward(self, x):
        return self.seq(x)
