"""Process / device context: one ``torch.distributed`` rank per GPU, several
*virtual* graph nodes per rank.

The reference has no distributed runtime at all (every node is a
deepcopy in one process, optimizers/dsgd.py:37-46 reads other replicas by
attribute access).  Here the node -> rank placement is explicit and every
consumer (arena, exchange, problems) goes through ``DistContext``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def block_partition(n_items: int, n_parts: int) -> List[int]:
    """Sizes of a contiguous block partition (first ``n_items % n_parts`` parts get one more)."""
    base, rem = divmod(n_items, n_parts)
    return [base + (1 if r < rem else 0) for r in range(n_parts)]


def cpus_from_mask(words: Sequence[int], bits_per_word: int = 64) -> set:
    """CPU indices of an NVML affinity bitmask (array of machine words, CPU ``64 w + b`` = bit ``b`` of word ``w``)."""
    return {w * bits_per_word + b for w, word in enumerate(words) for b in range(bits_per_word) if (int(word) >> b) & 1}


def bind_to_local_cpus(device_index: int, min_cpus: int = 4, nvml=None) -> Optional[set]:
    """NUMA placement for the host-fed input pipeline: restrict this process to the CPUs NVML reports as local to its GPU
    *before* any pinned host buffer is allocated, so first-touch places the staging dataset on the GPU's own socket — with
    8 ranks on a two-socket node, device-initiated pulls from the far socket cross the inter-socket link and share it.
    Returns the CPU set applied, or None when nothing was changed: NVML missing, ``NNDT_NUMA_BIND=0``, the mask does not
    narrow the current affinity, or the cgroup leaves fewer than ``min_cpus`` of the local CPUs."""
    if os.environ.get("NNDT_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        if nvml is None:
            import pynvml as nvml
            nvml.nvmlInit()
        handle = None
        try:        # CUDA_VISIBLE_DEVICES may renumber the devices: match by UUID
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = nvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:  # noqa: BLE001
            handle = nvml.nvmlDeviceGetHandleByIndex(device_index)
        allowed = set(os.sched_getaffinity(0))
        n_words = (max(allowed) // 64) + 1 if allowed else 1
        local = cpus_from_mask(nvml.nvmlDeviceGetCpuAffinity(handle, n_words)) & allowed
        if len(local) < min_cpus or local == allowed:
            return None
        os.sched_setaffinity(0, local)
        return local
    except Exception:  # noqa: BLE001  (placement is an optimisation, never a failure)
        return None


class DistContext:
    """Rank/world info plus the collectives the non-fused paths need."""

    def __init__(self, rank: int = 0, world_size: int = 1, device: torch.device | str = "cpu",
                 group=None):
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.device = torch.device(device)
        self.group = group
        self.local_cpus = None

    # -- construction --------------------------------------------------
    @classmethod
    def single(cls, device="cpu") -> "DistContext":
        return cls(0, 1, device)

    @classmethod
    def from_env(cls, use_cuda: bool = True, backend: Optional[str] = None) -> "DistContext":
        """Initialise from torchrun-style env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*)."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
        cuda = use_cuda and torch.cuda.is_available()
        device = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
        if cuda:
            torch.cuda.set_device(device)
            bound = bind_to_local_cpus(local_rank)
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            be = backend or ("nccl" if cuda else "gloo")
            kw = {"device_id": device} if (cuda and be == "nccl") else {}
            dist.init_process_group(be, rank=rank, world_size=world, **kw)
        ctx = cls(rank, world, device)
        ctx.local_cpus = None if not cuda or not bound else len(bound)      # CPUs this rank was bound to (NUMA-local to its GPU)
        return ctx

    @property
    def is_distributed(self) -> bool:
        return self.world_size > 1

    @property
    def is_main(self) -> bool:
        return self.rank == 0

    # -- collectives (plumbing only; the hot path is the fused P2P kernel) --
    def barrier(self):
        if self.is_distributed:
            dist.barrier(group=self.group)

    def all_gather_cat(self, t: torch.Tensor, sizes: Sequence[int]) -> torch.Tensor:
        """Concatenate per-rank tensors along dim 0 (``sizes[r]`` rows on rank r)."""
        if not self.is_distributed:
            return t
        maxs = max(sizes)
        pad = t
        if t.shape[0] != maxs:
            pad = t.new_zeros((maxs,) + tuple(t.shape[1:]))
            pad[: t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad.contiguous(), group=self.group)
        return torch.cat([o[: sizes[r]] for r, o in enumerate(out)], dim=0)

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.is_distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.is_distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast_object(self, obj, src: int = 0):
        if not self.is_distributed:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]


@dataclass
class Placement:
    """Contiguous block placement of N graph nodes over ``world_size`` ranks."""

    N: int
    world_size: int
    rank: int
    counts: List[int] = field(init=False)
    offsets: List[int] = field(init=False)

    def __post_init__(self):
        self.counts = block_partition(self.N, self.world_size)
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)
        self.node_rank = np.repeat(np.arange(self.world_size), self.counts)
        self.node_local = np.concatenate([np.arange(c) for c in self.counts]) if self.N else np.zeros(0, int)

    @property
    def L(self) -> int:
        return self.counts[self.rank]

    @property
    def lo(self) -> int:
        return self.offsets[self.rank]

    @property
    def local_nodes(self) -> List[int]:
        return list(range(self.lo, self.lo + self.L))

    def is_local(self, g: int) -> bool:
        return self.lo <= g < self.lo + self.L

    def local_index(self, g: int) -> int:
        return g - self.lo
