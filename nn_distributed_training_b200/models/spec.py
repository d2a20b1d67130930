"""Architecture descriptors shared by the nn.Module shells and the fused
sm_100a kernels.  A spec is pure data: the kernels are selected and
parameterised from it, and the flat-arena offsets of every tensor follow the
``nn.Module.parameters()`` order so a node's parameter row is bit-compatible
with ``torch.nn.utils.parameters_to_vector`` (reference: optimizers/dinno.py:81).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass(frozen=True)
class ConvNetSpec:
    """1 conv (valid, stride 1) -> ReLU -> maxpool 2 -> fc -> ReLU -> fc -> log-softmax."""

    num_filters: int
    kernel_size: int
    linear_width: int
    in_hw: int = 28
    num_classes: int = 10
    kind: str = "mnist_conv"

    @property
    def conv_hw(self) -> int:
        return self.in_hw - (self.kernel_size - 1)

    @property
    def pool_hw(self) -> int:
        return self.conv_hw // 2

    @property
    def fc1_in(self) -> int:
        return self.num_filters * self.pool_hw * self.pool_hw

    def param_shapes(self) -> List[Tuple[int, ...]]:
        f, k, w = self.num_filters, self.kernel_size, self.linear_width
        return [(f, 1, k, k), (f,), (w, self.fc1_in), (w,), (self.num_classes, w), (self.num_classes,)]


@dataclass(frozen=True)
class MLPSpec:
    """Fully connected net.  ``first`` is the first-layer activation
    (``"sin"`` = SIREN/Fourier layer followed by ReLU, as the reference's
    FourierNet composes them, models/fourier_nn.py:48-57), ``hidden`` the
    hidden activation and ``last`` the output activation."""

    shape: Tuple[int, ...]
    first: str = "relu"      # relu | sin_relu | tanh | sigmoid
    hidden: str = "relu"     # relu | tanh | sigmoid
    last: str = "none"       # none | sigmoid | tanh
    scale: float = 1.0       # multiplies the first layer pre-activation when first == sin_relu
    kind: str = "mlp"

    def param_shapes(self) -> List[Tuple[int, ...]]:
        out: List[Tuple[int, ...]] = []
        for i in range(len(self.shape) - 1):
            out.append((self.shape[i + 1], self.shape[i]))
            out.append((self.shape[i + 1],))
        return out

    @property
    def num_layers(self) -> int:
        return len(self.shape) - 1

    def activation(self, layer: int) -> str:
        if layer == self.num_layers - 1:
            return self.last
        if layer == 0:
            return self.first
        return self.hidden
