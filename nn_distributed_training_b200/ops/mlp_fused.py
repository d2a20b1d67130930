"""Python side of the tcgen05 MLP kernels (csrc/mlp_tc.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import load_ext
from ..models.spec import MLPSpec

FIRST = {"relu": 0, "sin_relu": 1}
LAST = {"none": 0, "sigmoid": 1}
LOSS = {"BCELoss": 0, "MSELoss": 1, "L1Loss": 2}
TRAIN_KERNEL_READY = True


def shape_supported(spec: MLPSpec) -> bool:
    s = spec.shape
    return (len(s) == 6 and s[0] <= 4 and s[1] in (64, 128, 256) and tuple(s[2:5]) == (64, 64, 64) and s[5] == 1
            and spec.first in FIRST and spec.hidden == "relu" and spec.last in LAST)


def supports(spec: MLPSpec, base_loss) -> bool:
    return (TRAIN_KERNEL_READY and shape_supported(spec) and type(base_loss).__name__ in LOSS
            and getattr(base_loss, "reduction", "mean") == "mean")


def op_dict(arena, spec: MLPSpec, L: int):
    off = [s.offset for s in arena.layout.slots]
    assert len(off) == 10
    return dict(theta=arena.theta.data_ptr(), n_pad=arena.n_pad, L=L, off=off, d_in=spec.shape[0], h1=spec.shape[1],
                first_act=FIRST[spec.first], last_act=LAST[spec.last], scale=float(spec.scale))


class MlpForward:
    """Forward-only evaluation of every local node's network on shared inputs ``x [M, d_in]``."""

    def __init__(self, arena, spec: MLPSpec, L: int, device):
        self.ext = load_ext(required=True)
        self.arena, self.spec, self.L, self.device = arena, spec, L, device
        self.sms = torch.cuda.get_device_properties(device).multi_processor_count
        self._cache = {}

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(torch.float32).contiguous()
        M = x.shape[0]
        out = torch.empty(self.L, M, dtype=torch.float32, device=self.device)
        d = op_dict(self.arena, self.spec, self.L)
        ctas = max(1, min(-(-M // 128), max(1, (2 * self.sms) // max(1, self.L))))
        d.update(x=x.data_ptr(), n_rows=M, out=out.data_ptr(), fwd_ctas=ctas)
        op = self.ext.MlpOp(d)
        op.forward()
        self._keep = (x, out, op)
        return out
