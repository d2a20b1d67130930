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

import os
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
            if os.environ.get("NNDT_SYMM", "symm_mem") == "ipc":      # A/B switch: plain cudaMalloc memory exported by CUDA IPC
                raise RuntimeError("NNDT_SYMM=ipc")
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


class DeviceBarrier:
    """All-rank barrier executed ON THE DEVICE (csrc/consensus.cu: rank_barrier_kernel): every rank stores an epoch into
    its slot of every peer's array and spins until all peers did.  Enqueued on the current stream, so whatever is enqueued
    behind it starts within an NVLink flag latency on every GPU — host-side barrier exit skew (tens of microseconds, and
    milliseconds when NVML or Python run between the barrier and the first launch) never reaches the device timeline.
    ``gate=True`` makes the kernel first wait for ``open_gate()``: the host enqueues the whole timed region behind the
    barrier and only then releases it, so launch latency is not timed either."""

    def __init__(self, ctx):
        from ..ops import load_ext

        self.ctx = ctx
        self.ext = load_ext(required=True)
        w = max(ctx.world_size, 1)
        self.buf = SymmetricBuffer((w,), torch.int32, ctx)
        self.peer_slot = torch.tensor([p + 4 * ctx.rank for p in self.buf.peer_ptrs] + [0] * (w - len(self.buf.peer_ptrs)),
                                      dtype=torch.int64, device=ctx.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=ctx.device)
        self.gate = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.epoch = 0

    def enqueue(self, gate: bool = False):
        self.epoch += 1
        if gate:
            self.gate.zero_()
        self.ext.rank_barrier(self.buf.local.data_ptr(), self.peer_slot.data_ptr(), max(self.ctx.world_size, 1), self.ctx.rank,
                              self.epoch, self.gate.data_ptr() if gate else 0, self.err.data_ptr())

    def open_gate(self):
        self.gate.fill_(1)

    def check(self):
        if int(self.err.item()) != 0:
            raise RuntimeError("device rank barrier timed out")
