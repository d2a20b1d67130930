"""Dense, device-resident training shards.

The reference keeps datasets on the host and pushes every minibatch through a
Python DataLoader + H2D copy (problems/dist_mnist_problem.py:83-98).  MNIST is
47 MB as uint8 and a lidar shard is ~15 MB, so each node's shard simply lives in
HBM and the kernels gather rows by (stateless) index.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset, Subset, TensorDataset


@dataclass
class Shard:
    """``x [M, ...]`` (uint8 pixels or float features), ``y [M]``.

    ``norm=(mean, std)`` means ``x`` is uint8 and the model input is
    ``(x / 255 - mean) / std`` — the ToTensor+Normalize transform of the
    reference runner (experiments/dist_mnist_ex.py:98-100), applied in-kernel.
    """

    x: torch.Tensor
    y: torch.Tensor
    norm: Optional[Tuple[float, float]] = None

    def __len__(self) -> int:
        return int(self.x.shape[0])

    def select(self, idx) -> "Shard":
        idx = torch.as_tensor(idx, dtype=torch.long)
        return Shard(self.x[idx], self.y[idx], self.norm)

    def to(self, device) -> "Shard":
        return Shard(self.x.to(device), self.y.to(device), self.norm)

    def inputs(self, idx: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        xb = self.x[idx]
        if self.norm is not None:
            mean, std = self.norm
            return (xb.to(dtype) / 255.0 - mean) / std
        return xb.to(dtype)

    def targets(self, idx: torch.Tensor) -> torch.Tensor:
        return self.y[idx]

    # torch Dataset protocol so a Shard can be handed to reference-style code
    def __getitem__(self, i):
        return self.inputs(torch.as_tensor([i]), torch.get_default_dtype())[0], self.y[i]


def _norm_from_transform(tf) -> Optional[Tuple[float, float]]:
    stack = [tf]
    while stack:
        t = stack.pop()
        if t is None:
            continue
        if hasattr(t, "transforms"):
            stack.extend(t.transforms)
        if t.__class__.__name__ == "Normalize":
            mean = float(torch.as_tensor(t.mean).reshape(-1)[0])
            std = float(torch.as_tensor(t.std).reshape(-1)[0])
            return mean, std
    return None


def as_shard(ds) -> Shard:
    """Materialise a dataset as a ``Shard`` (no copy when it already is one)."""
    if isinstance(ds, Shard):
        return ds
    if isinstance(ds, Subset):
        return as_shard(ds.dataset).select(torch.as_tensor(ds.indices))
    if isinstance(ds, TensorDataset):
        return Shard(ds.tensors[0], ds.tensors[1])
    if hasattr(ds, "shard"):  # our lidar datasets
        return ds.shard
    if hasattr(ds, "tds"):  # reference lidar datasets
        return as_shard(ds.tds)
    if hasattr(ds, "data") and hasattr(ds, "targets"):  # torchvision MNIST-like
        x = torch.as_tensor(ds.data)
        if x.dtype == torch.uint8 and x.dim() == 3:
            x = x.unsqueeze(1)
            norm = _norm_from_transform(getattr(ds, "transform", None)) or (0.0, 1.0)
            return Shard(x, torch.as_tensor(ds.targets), norm)
        return Shard(x, torch.as_tensor(ds.targets))
    xs, ys = [], []
    for i in range(len(ds)):
        x, y = ds[i]
        xs.append(torch.as_tensor(x))
        ys.append(torch.as_tensor(y))
    return Shard(torch.stack(xs), torch.stack(ys))


class ShardSet:
    """The shards of this rank's local nodes, concatenated on the device.

    ``x``/``y`` hold every local shard back to back; ``offsets[l]`` is the first
    row of local node ``l`` — the layout the fused kernels index with
    ``offsets[l] + perm(pos)``.
    """

    def __init__(self, shards: Sequence[Shard], device):
        self.sizes = [len(s) for s in shards]
        self.norm = shards[0].norm if shards else None
        self.offsets = [0]
        for m in self.sizes:
            self.offsets.append(self.offsets[-1] + m)
        self.x = torch.cat([s.x for s in shards]).to(device).contiguous()
        self.y = torch.cat([s.y for s in shards]).to(device).contiguous()
        self.device = torch.device(device)

    def shard(self, l: int) -> Shard:
        a, b = self.offsets[l], self.offsets[l + 1]
        return Shard(self.x[a:b], self.y[a:b], self.norm)
