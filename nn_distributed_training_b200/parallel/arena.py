"""Flat parameter arena.

All three consensus algorithms are elementwise over the concatenated parameter
vector; the reference pays ``parameters_to_vector`` (a cat + copy) on every
primal step and every round (optimizers/dinno.py:81-83,104-106).  Here each
local graph node owns one 512-byte aligned row of ``theta [L, n_pad]`` and the
``nn.Parameter`` objects of its model are *views* of that row, so kernels see
the flat vector and PyTorch sees ordinary modules (state_dict, checkpoints).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn

ROW_ALIGN_ELEMS = 128  # rows are 512 B aligned (fp32)
SLOT_ALIGN_ELEMS = 4    # every parameter tensor starts on a 16 B boundary -> float4 / cp.async.16 everywhere


@dataclass(frozen=True)
class ParamSlot:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int


class FlatLayout:
    """Offsets of a module's parameters inside a flat row, in ``parameters()`` order."""

    def __init__(self, slots: Sequence[ParamSlot]):
        self.slots = list(slots)
        self.n = sum(s.numel for s in self.slots)
        end = max((s.offset + s.numel for s in self.slots), default=0)
        self.n_pad = ((end + ROW_ALIGN_ELEMS - 1) // ROW_ALIGN_ELEMS) * ROW_ALIGN_ELEMS
        self.dense = all(a.offset + a.numel == b.offset for a, b in zip(self.slots, self.slots[1:]))

    @classmethod
    def from_module(cls, module: nn.Module) -> "FlatLayout":
        slots, off = [], 0
        for name, p in module.named_parameters():
            slots.append(ParamSlot(name, tuple(p.shape), off, p.numel()))
            off += -(-p.numel() // SLOT_ALIGN_ELEMS) * SLOT_ALIGN_ELEMS
        return cls(slots)

    def offsets(self) -> List[int]:
        return [s.offset for s in self.slots]

    def flatten(self, module: nn.Module, out: torch.Tensor) -> torch.Tensor:
        for s, p in zip(self.slots, module.parameters()):
            out[s.offset: s.offset + s.numel].copy_(p.detach().reshape(-1))
        return out

    def views(self, row: torch.Tensor) -> List[torch.Tensor]:
        return [row[s.offset: s.offset + s.numel].view(s.shape) for s in self.slots]

    def compact(self, t: torch.Tensor) -> torch.Tensor:
        """``[..., n_pad] -> [..., n]``: drop alignment holes, giving exactly
        ``parameters_to_vector`` order.  Holes hold zeros forever (every update is
        elementwise with zero gradient there), so norms/distances may also be
        taken on padded rows directly."""
        if self.dense:
            return t[..., : self.n]
        return torch.cat([t[..., s.offset: s.offset + s.numel] for s in self.slots], dim=-1)


class NodeArena:
    """Rows of flat state for the ``L`` graph nodes hosted by this rank."""

    def __init__(self, layout: FlatLayout, n_local: int, device, dtype=torch.float32):
        self.layout = layout
        self.L = int(n_local)
        self.n = layout.n
        self.n_pad = layout.n_pad
        self.device = torch.device(device)
        self.dtype = dtype
        self.theta = self.zeros()
        self.grad = self.zeros()

    def zeros(self, *lead: int) -> torch.Tensor:
        return torch.zeros(*lead, self.L, self.n_pad, device=self.device, dtype=self.dtype)

    def attach(self, l: int, module: nn.Module, init_from_module: bool = True) -> nn.Module:
        """Re-point ``module``'s parameters at row ``l`` (optionally seeding the
        row from the module's current values)."""
        row = self.theta[l]
        if init_from_module:
            self.layout.flatten(module, row)
        for p, v in zip(module.parameters(), self.layout.views(row)):
            p.data = v
            p.grad = None
        return module

    def set_row_from_grads(self, l: int, grads: Sequence[torch.Tensor]):
        g = self.grad[l]
        for s, t in zip(self.layout.slots, grads):
            g[s.offset: s.offset + s.numel].copy_(t.reshape(-1))

    def compact(self, t: torch.Tensor) -> torch.Tensor:
        return self.layout.compact(t)
