"""Symmetric (peer-mapped) device buffers over NVLink.

Each rank allocates the same-shaped buffer; after ``rendezvous`` every rank holds
the device address of every peer's copy, valid for plain ``ld.global`` /
``st.global`` inside our kernels (P2P over NVLink 5 / NVSwitch).  Primary
implementation: ``torch.distributed._symmetric_memory`` (CUDA backend: VMM
allocations exchanged as POSIX fds, with an NVLS multicast mapping when the
fabric supports it).  Fallback: legacy CUDA IPC handles of a caching-allocator
block exchanged through the process group and opened by our runtime
(csrc/runtime.cpp).  The reference has no communication backend at all
(SURVEY §5.8); NCCL is used here only for bootstrap and in the baseline.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class SymmetricBuffer:
    """``local`` tensor + ``peer_ptrs[r]`` device address of rank r's copy."""

    def __init__(self, shape, dtype, ctx, zero: bool = True):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.multicast_ptr: int = 0
        self.how = "local"
        self._keep = []
        dev = ctx.device
        if not ctx.is_distributed:
            self.local = torch.zeros(self.shape, dtype=dtype, device=dev)
            self.peer_ptrs = [self.local.data_ptr()]
            return
        numel = 1
        for s in self.shape:
            numel *= s
        try:
            self._alloc_symm_mem(numel, dev)
        except Exception as e:  # noqa: BLE001
            self._symm_err = repr(e)
            self._alloc_ipc(numel, dev)
        if zero:
            self.local.zero_()
        torch.cuda.synchronize(dev)
        ctx.barrier()

    def _alloc_symm_mem(self, numel, dev):
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(numel, dtype=self.dtype, device=dev)
        group = self.ctx.group or dist.group.WORLD
        hdl = symm_mem.rendezvous(t, group.group_name)
        self.local = t.view(self.shape)
        self.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
        try:
            self.multicast_ptr = int(hdl.multicast_ptr or 0)
        except Exception:  # noqa: BLE001
            self.multicast_ptr = 0
        self._keep.append(hdl)
        self.how = "symm_mem"

    def _alloc_ipc(self, numel, dev):
        from ..ops import load_ext

        ext = load_ext(required=True)
        t = torch.zeros(numel, dtype=self.dtype, device=dev)
        handle = ext.ipc_get_handle(t.data_ptr())  # (bytes handle, offset from allocation base)
        gathered: List[Optional[tuple]] = [None] * self.ctx.world_size
        dist.all_gather_object(gathered, handle, group=self.ctx.group)
        ptrs = []
        for r, (hb, off) in enumerate(gathered):
            if r == self.ctx.rank:
                ptrs.append(t.data_ptr())
            else:
                ptrs.append(ext.ipc_open_handle(hb) + off)
        self.local = t.view(self.shape)
        self.peer_ptrs = ptrs
        self.how = "cuda_ipc"
