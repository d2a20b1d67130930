"""Python side of the fused MNIST kernels (csrc/mnist.cu).

``FusedMnist`` owns the device buffers the kernels read/write for one problem
instance: the flattened uint8/float shard rows, the device draw counters of the
stateless sampler, the per-slice gradient partials and the validation outputs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import load_ext

SPB = 8  # samples per CTA (template instantiation in mnist.cu)


class FusedMnist:
    def __init__(self, problem):
        self.pr = problem
        self.ext = load_ext(required=True)
        dev = problem.device
        a, pl = problem.arena, problem.placement
        self.L, self.n_pad = pl.L, a.n_pad
        self.B = problem.train_batch_size
        self.S = -(-self.B // SPB)
        sh = problem.shards
        self.x = sh.x.reshape(sh.x.shape[0], -1).contiguous()
        assert self.x.shape[1] == 784
        self.x_is_u8 = self.x.dtype == torch.uint8
        if not self.x_is_u8:
            self.x = self.x.to(torch.float32).contiguous()
        self.y = sh.y.to(torch.int64).contiguous()
        mean, std = sh.norm if sh.norm is not None else (0.0, 1.0)
        self.shard_off = torch.tensor(sh.offsets[:-1], dtype=torch.int32, device=dev)
        self.shard_len = torch.tensor(sh.sizes, dtype=torch.int32, device=dev)
        self.calls = torch.zeros(self.L, dtype=torch.int32, device=dev)
        self.grad_part = torch.zeros(self.L, self.S, self.n_pad, dtype=torch.float32, device=dev)
        self.loss_part = torch.zeros(self.L, self.S, dtype=torch.float32, device=dev)
        off = {s.name: s.offset for s in a.layout.slots}
        names = [s.name for s in a.layout.slots]
        self.base = dict(
            theta=a.theta.data_ptr(), n_pad=a.n_pad, L=self.L,
            off_wc=off[names[0]], off_bc=off[names[1]], off_w1=off[names[2]],
            off_b1=off[names[3]], off_w2=off[names[4]], off_b2=off[names[5]],
            x=self.x.data_ptr(), y=self.y.data_ptr(), x_is_u8=int(self.x_is_u8),
            mean=float(mean), inv_std=1.0 / float(std),
            direct=0, batch=self.B, seed=problem.seed, node0=pl.lo,
            shard_off=self.shard_off.data_ptr(), shard_len=self.shard_len.data_ptr(),
            calls=self.calls.data_ptr(),
            grad_part=self.grad_part.data_ptr(), loss_part=self.loss_part.data_ptr(),
            spb=SPB, S=self.S)
        self.train_op = self.ext.MnistOp(self.base)
        self._setup_eval()
        self._direct_op = None

    # ---- training ---------------------------------------------------------
    def launch(self):
        """Enqueue fwd+bwd of the next batch of every local node (graph-capturable).
        The draw counter is advanced by the consensus kernel that consumes the partials."""
        self.train_op.train()

    def compute_grads(self) -> torch.Tensor:
        """Eager API: fills ``arena.grad`` and advances the counters itself."""
        pr = self.pr
        self.launch()
        torch.sum(self.grad_part, dim=1, out=pr.arena.grad)
        self.calls += 1
        pr.count_draws_all(1)
        pr.last_losses = self.loss_part.sum(1)
        return pr.last_losses

    def sync_calls_from_host(self):
        pl = self.pr.placement
        self.calls.copy_(torch.as_tensor(self.pr.calls[pl.lo: pl.lo + pl.L].astype(np.int32)))

    # ---- host-fed batches (end-to-end mode) --------------------------------
    def direct_op(self, x_stage: torch.Tensor, y_stage: torch.Tensor):
        """Kernel op reading one staged batch per node from ``x_stage [L,B,784]`` /
        ``y_stage [L,B]`` (filled by an H2D copy) instead of the resident shards."""
        d = dict(self.base)
        d.update(direct=1, x=x_stage.data_ptr(), y=y_stage.data_ptr(),
                 x_is_u8=int(x_stage.dtype == torch.uint8))
        return self.ext.MnistOp(d)

    # ---- validation ---------------------------------------------------------
    def _setup_eval(self):
        pr = self.pr
        if pr.val is None:
            self.eval_op = None
            return
        dev = pr.device
        vx = pr.val.x.reshape(len(pr.val), -1).contiguous()
        self.vx = vx if vx.dtype == torch.uint8 else vx.to(torch.float32).contiguous()
        self.vy = pr.val.y.to(torch.int64).contiguous()
        V = len(pr.val)
        self.val_loss = torch.zeros(self.L, V, dtype=torch.float32, device=dev)
        self.val_correct = torch.zeros(self.L, V, dtype=torch.uint8, device=dev)
        mean, std = pr.val.norm if pr.val.norm is not None else (0.0, 1.0)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        ctas = max(1, min(-(-V // SPB), sms // max(1, self.L)))
        d = dict(self.base)
        d.update(x=self.vx.data_ptr(), y=self.vy.data_ptr(), x_is_u8=int(self.vx.dtype == torch.uint8),
                 mean=float(mean), inv_std=1.0 / float(std), n_val=V,
                 val_loss=self.val_loss.data_ptr(), val_correct=self.val_correct.data_ptr(),
                 eval_ctas=ctas)
        self.eval_op = self.ext.MnistOp(d)

    def validate(self):
        self.eval_op.eval()
        return self.val_loss, self.val_correct.bool()
