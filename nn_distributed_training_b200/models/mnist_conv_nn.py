"""MNIST convolutional classifier (reference: models/mnist_conv_nn.py:10-28).

The module keeps the reference's ``seq.<idx>`` state_dict keys so checkpoints
are interchangeable; the architecture itself is carried by ``self.spec`` which
is what the fused sm_100a forward/backward kernel (ops/csrc/mnist.cu) consumes.
"""
from __future__ import annotations

import torch
from torch import nn

from .spec import ConvNetSpec


class MNISTConvNet(nn.Module):
    def __init__(self, num_filters, kernel_size, linear_width, dtype=None):
        super().__init__()
        self.spec = ConvNetSpec(int(num_filters), int(kernel_size), int(linear_width))
        s = self.spec
        kw = {} if dtype is None else {"dtype": dtype}
        stages = [
            nn.Conv2d(1, s.num_filters, s.kernel_size, 1, **kw),
            nn.ReLU(inplace=True),
            nn.MaxPool2d(2),
            nn.Flatten(),
            nn.Linear(s.fc1_in, s.linear_width, **kw),
            nn.ReLU(inplace=True),
            nn.Linear(s.linear_width, s.num_classes, **kw),
            nn.LogSoftmax(dim=1),
        ]
        self.seq = nn.Sequential(*stages)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.seq(x)
