"""Python side of the fused MNIST kernels (csrc/mnist.cu).

``FusedMnist`` owns the device buffers the kernels read/write for one problem
instance: the flattened uint8/float shard rows, the device draw counters of the
stateless sampler, the per-slice gradient partials and the validation outputs.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import load_ext, mnist_kernel_is_paper_shape

SPB = 8  # samples per CTA of the evaluation kernel / upper bound for training (template instantiations in mnist.cu)


def choose_spb(batch: int, n_local: int, sms: int) -> int:
    """Samples per training CTA (4..8).  A node's batch is cut into ceil(batch/spb) CTAs; per-CTA time grows
    with spb (conv / fc work) on top of a fixed part (weight staging, dW1 write-out), so the best choice is
    the smallest spb whose L x S CTAs still fit in ONE wave of the SMs (1 CTA/SM: ~200 KB smem, 768 threads)."""
    for spb in range(4, SPB + 1):
        if n_local * -(-batch // spb) <= sms:
            return spb
    return SPB


class FusedMnist:
    def __init__(self, problem):
        self.pr = problem
        self.ext = load_ext(required=True)
        dev = problem.device
        a, pl = problem.arena, problem.placement
        self.L, self.n_pad = pl.L, a.n_pad
        self.B = problem.train_batch_size
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        spec = problem.base_model.spec
        self.dtype = a.dtype
        # the paper's (3, 5, 64) net in fp32 runs the specialised kernels; every other shape and all of fp64 run the
        # generic CUDA-core kernel (csrc/mnist_generic.cu)
        self.generic = not (self.dtype == torch.float32 and mnist_kernel_is_paper_shape(spec)) \
            or os.environ.get("NNDT_MNIST_GENERIC") == "1"
        if self.generic:
            d64 = int(self.dtype == torch.float64)
            fits8 = self.ext.convnet_generic_smem_bytes(spec.num_filters, spec.kernel_size, spec.linear_width, d64, 8) <= 200 * 1024
            want = int(os.environ.get("NNDT_GENERIC_SPB", "0")) or (8 if fits8 and self.L * -(-self.B // 8) >= sms // 2 else 4)
            self.spb = 8 if (want == 8 and fits8) else 4
        else:
            self.spb = (int(os.environ.get("NNDT_SPB", "0")) or int(problem.conf.get("samples_per_cta", 0))
                        or choose_spb(self.B, self.L, sms))
            assert 4 <= self.spb <= SPB
        self.S = -(-self.B // self.spb)
        # paper shape, fp32, batch <= 64: the tcgen05 / TMEM K-split cluster kernel (csrc/mnist_tc.cu) — one gradient
        # row per node instead of S per-slice partials.  NNDT_MNIST_TC=0 keeps the batch-split mma.sync kernel (A/B).
        self.tc = (not self.generic and self.B <= 64 and os.environ.get("NNDT_MNIST_TC", "1") != "0"
                   and str(problem.conf.get("mnist_kernel", "tc")) == "tc" and self.ext.mnist_tc_max_clusters() >= 1)
        if self.tc:
            # batch splits per node (1, 2 or 4 clusters of 6 CTAs, M = 64 / 32 / 16 samples each): as many as keep all
            # 6 * nsplit * L CTAs in one wave — the conv / conv-grad phases are CUDA-core work that scales with SMs
            want = int(os.environ.get("NNDT_TC_SPLIT", "0"))
            self.S = want if want in (1, 2, 4) else max([n for n in (1, 2, 4) if 6 * n * self.L <= sms] or [1])
        # float64, paper shape, batch <= 64: the same K-split cluster decomposition on the fp64 CUDA cores (csrc/mnist_cl64.cu);
        # NNDT_MNIST_CL64=0 keeps the batch-split generic kernel (A/B)
        self.cl64 = (self.generic and self.dtype == torch.float64 and mnist_kernel_is_paper_shape(spec) and self.B <= 64
                     and os.environ.get("NNDT_MNIST_CL64", "1") != "0" and self.ext.mnist_cl64_max_clusters() >= 1)
        if self.cl64:
            want = int(os.environ.get("NNDT_TC_SPLIT", "0"))
            self.S = want if want in (1, 2, 4) else max([n for n in (1, 2, 4) if 6 * n * self.L <= sms] or [1])
        self.kernel_name = (f"mnist_cl64_train_kernel<{64 // self.S}> (fp64 CUDA cores, {self.S} x 6-CTA cluster per node, DSMEM reduce)" if self.cl64 else
                            f"mnist_tc_train_kernel<{64 // self.S}> (tcgen05 kind::tf32 3xTF32, TMEM, TMA tensor map, {self.S} x 6-CTA cluster per node)" if self.tc
                            else "convnet_generic_kernel (CUDA cores)" if self.generic else "mnist_kernel (mma.sync 3xTF32)")
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
        self.arrive = torch.zeros(self.L, dtype=torch.int32, device=dev)
        self.owns_calls = True      # the training kernel advances the draw counters (not the consensus kernels)
        self.grad_part = torch.zeros(self.L, self.S, self.n_pad, dtype=self.dtype, device=dev)
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
            calls=self.calls.data_ptr(), arrive=self.arrive.data_ptr(),
            grad_part=self.grad_part.data_ptr(), loss_part=self.loss_part.data_ptr(),
            spb=self.spb, S=self.S, tune=int(os.environ.get("NNDT_MNIST_TUNE", "1")),
            generic=int(self.generic), num_filters=spec.num_filters, kernel_size=spec.kernel_size,
            linear_width=spec.linear_width, dtype64=int(self.dtype == torch.float64), cl64=int(self.cl64))
        if self.tc:
            self.base.update(tc=1, w1_map=self.ext.make_w1_tensor_map(a.theta.data_ptr(), a.n_pad, self.L, off[names[2]]))
        if os.environ.get("NNDT_STEP_PROF") == "1":     # scripts/profile_round_phases.py --per-step
            self.step_prof = torch.zeros(self.L * self.S * (6 if (self.tc or self.cl64) else 1), 64, dtype=torch.int64, device=dev)
            self.base["step_prof"] = self.step_prof.data_ptr()
        self.train_op = self.ext.MnistOp(self.base)
        self._setup_eval()
        self.host_feed = None

    # ---- training ---------------------------------------------------------
    def launch(self):
        """Enqueue fwd+bwd of the next batch of every local node (graph-capturable).
        The draw counter is advanced by the consensus kernel that consumes the partials."""
        self.train_op.train()

    # ---- one launch per DiNNO round (csrc/dinno_round.cu) -----------------------
    MAX_ROUND_STEPS = 8

    def supports_round_kernel(self, opt) -> bool:
        """The cluster kernel needs the node's S batch slices in one cluster (S <= 8 CTAs) and fp32 state."""
        return (not self.generic and not self.tc and opt.alg_name == "dinno" and self.spb == SPB and self.S <= 8 and 1 <= opt.pits <= self.MAX_ROUND_STEPS
                and self.pr.arena.dtype == torch.float32 and self.ext.dinno_round_max_clusters(self.S) >= 1)

    def round_op(self, cons_dict, stage_set=None):
        """``DinnoRoundOp`` running all primal iterations of a round; ``stage_set`` selects the host-fed
        staging buffers (one batch source per step) instead of the resident shards."""
        P = int(cons_dict["pits"])
        md = dict(self.base)
        if getattr(self, "round_prof", None) is not None:
            md["prof"] = self.round_prof.data_ptr()     # scripts/profile_round_phases.py
        if stage_set is None:
            steps = [dict(x=self.x.data_ptr(), y=self.y.data_ptr(), direct_bs=None) for _ in range(P)]
        else:
            md.update(direct=1)
            if getattr(self, "loss_mode", "memcpy") == "mirror":
                md.update(loss_mirror=self.loss_host.data_ptr())
            b = stage_set
            steps = [dict(x=self.x_stage[b, p].data_ptr(), y=self.y_stage[b, p].data_ptr(),
                          direct_bs=self.bs_stage[b, p].data_ptr()) for p in range(P)]
        return self.ext.DinnoRoundOp(md, cons_dict, steps)

    def compute_grads(self) -> torch.Tensor:
        """Eager API: fills ``arena.grad`` and advances the counters itself."""
        pr = self.pr
        self.launch()
        torch.sum(self.grad_part, dim=1, out=pr.arena.grad)
        pr.count_draws_all(1)
        pr.last_losses = self.loss_part.sum(1).to(self.dtype)
        return pr.last_losses

    def sync_calls_from_host(self):
        pl = self.pr.placement
        self.calls.copy_(torch.as_tensor(self.pr.calls[pl.lo: pl.lo + pl.L].astype(np.int32)))

    # ---- host-fed batches (end-to-end input pipeline) -----------------------
    def enable_host_feed(self, steps_per_round: int, nslots: int = 4, threads: int = 4, mode: str = "gpu_pull"):
        """Switch to the host-fed input pipeline: the dataset stays in (pinned) host memory and
        every round's minibatches cross PCIe; each round also reads its losses back (D2H).
        This is the path ``bench.py`` times end to end; the default keeps shards in HBM.

        ``mode="gpu_pull"``: a staging kernel on a side stream pulls the *next* round's rows
        straight out of the pinned host dataset (device-initiated H2D, in-kernel sampler) while
        the current round computes — no CPU work per round.
        ``mode="cpu_loader"``: native loader threads (csrc/runtime.cpp) assemble rounds into a
        ring of pinned slots and the runner issues one ``cudaMemcpyAsync`` per round."""
        if mode in ("gpu_pull", "staged"):
            return self._enable_gpu_pull(steps_per_round, source="device" if mode == "staged" else "host")
        pr, dev = self.pr, self.pr.device
        P, L, B = int(steps_per_round), self.L, self.B
        xb = 1 if self.x_is_u8 else 4
        self.host_x = self.x.cpu().contiguous()
        self.host_y = self.y.cpu().contiguous()
        kw = dict(pin_memory=True)
        self.x_pin = torch.empty(nslots, P, L, B, 784, dtype=self.x.dtype, **kw)
        self.y_pin = torch.empty(nslots, P, L, B, dtype=torch.int64, **kw)
        self.bs_pin = torch.empty(nslots, P, L, dtype=torch.int32, **kw)
        # two device staging sets: the H2D copy of round r+1 overlaps the kernels of round r
        self.x_stage = torch.zeros(2, P, L, B, 784, dtype=self.x.dtype, device=dev)
        self.y_stage = torch.zeros(2, P, L, B, dtype=torch.int64, device=dev)
        self.bs_stage = torch.zeros(2, P, L, dtype=torch.int32, device=dev)
        self.loss_host = torch.zeros(L, self.S, dtype=torch.float32, **kw)
        self.direct_ops = []
        for b in range(2):
            ops = []
            for p in range(P):
                d = dict(self.base)
                d.update(direct=1, x=self.x_stage[b, p].data_ptr(), y=self.y_stage[b, p].data_ptr(),
                         direct_bs=self.bs_stage[b, p].data_ptr())
                ops.append(self.ext.MnistOp(d))
            self.direct_ops.append(ops)
        pl = pr.placement
        calls0 = [int(c) for c in pr.calls[pl.lo: pl.lo + pl.L]]
        self.loader = self.ext.HostBatchLoader(
            self.host_x.data_ptr(), self.host_y.data_ptr(), 784 * xb,
            [int(o) for o in pr.shards.offsets[:-1]], [int(m) for m in pr.shards.sizes], calls0,
            B, P, pr.seed, pl.lo,
            [self.x_pin[s].data_ptr() for s in range(nslots)],
            [self.y_pin[s].data_ptr() for s in range(nslots)],
            [self.bs_pin[s].data_ptr() for s in range(nslots)], threads)
        self.host_feed = dict(P=P, nslots=nslots, mode="cpu_loader",
                              h2d_bytes=P * L * B * (784 * xb + 8) + P * L * 4,
                              d2h_bytes=L * self.S * 4)
        return self.host_feed

    def _enable_gpu_pull(self, steps_per_round: int, source: str = "host"):
        """``source="host"``: rows come out of the pinned host copy of the dataset (PCIe).  ``source="device"``
        (``input_pipeline: staged``): the same staging kernel gathers the next round's rows from the HBM-resident
        shards into the compact staging set, so the training kernel never runs the sampler chain or a random HBM
        gather on its critical path."""
        pr, dev = self.pr, self.pr.device
        P, L, B = int(steps_per_round), self.L, self.B
        xb = 1 if self.x_is_u8 else 4
        if source == "host":
            self.host_x = self.x.cpu().contiguous().pin_memory()
            self.host_y = self.y.cpu().contiguous().pin_memory()
        else:
            self.host_x, self.host_y = self.x, self.y
        self.x_stage = torch.zeros(2, P, L, B, 784, dtype=self.x.dtype, device=dev)
        self.y_stage = torch.zeros(2, P, L, B, dtype=torch.int64, device=dev)
        self.bs_stage = torch.zeros(2, P, L, dtype=torch.int32, device=dev)
        self.loss_host = torch.zeros(L, self.S, dtype=torch.float32, pin_memory=True)
        # result read-back: "mirror" = the training kernel stores each step's losses straight into this pinned host
        # buffer over PCIe (no copy node on the round's critical path); "memcpy" = a D2H copy node per round
        self.loss_mode = str(pr.conf.get("host_loss", os.environ.get("NNDT_HOST_LOSS", "mirror")))
        if source == "device":
            self.loss_mode = "none"          # nothing crosses PCIe in the staged-resident pipeline
        pl = pr.placement
        self.calls0 = torch.as_tensor(pr.calls[pl.lo: pl.lo + pl.L].astype(np.int32), device=dev)
        self.stage_round = torch.zeros(1, dtype=torch.int32, device=dev)
        self.stage_done = torch.zeros(1, dtype=torch.int32, device=dev)
        # a training CTA fills an SM's register file (768 threads x 80 registers): a staging block that lands on
        # an SM evicts a training CTA into a second wave, so the staging grid is sized to the SMs left over
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        ctas = self.L * self.S * (6 if (self.tc or self.cl64) else 1)
        free = sms - ctas if ctas <= sms else 0      # multi-wave grids leave no SM idle
        gather_blocks = int(os.environ.get("NNDT_GATHER_BLOCKS", "0")) or max(8, min(24, free - 2))
        self.direct_ops, self.gather_ops = [], []
        for b in range(2):
            ops = []
            for p in range(P):
                d = dict(self.base)
                d.update(direct=1, x=self.x_stage[b, p].data_ptr(), y=self.y_stage[b, p].data_ptr(),
                         direct_bs=self.bs_stage[b, p].data_ptr())
                if self.loss_mode == "mirror":
                    d.update(loss_mirror=self.loss_host.data_ptr())
                ops.append(self.ext.MnistOp(d))
            self.direct_ops.append(ops)
            self.gather_ops.append(self.ext.GatherOp(dict(
                x_host=self.host_x.data_ptr(), y_host=self.host_y.data_ptr(), row_bytes=784 * xb,
                x_stage=self.x_stage[b].data_ptr(), y_stage=self.y_stage[b].data_ptr(),
                bs_stage=self.bs_stage[b].data_ptr(), P=P, L=L, batch=B, seed=pr.seed, node0=pl.lo,
                shard_off=self.shard_off.data_ptr(), shard_len=self.shard_len.data_ptr(),
                calls0=self.calls0.data_ptr(), stage_round=self.stage_round.data_ptr(),
                done_ctr=self.stage_done.data_ptr(), max_blocks=gather_blocks)))
        self.loader = None
        self.host_feed = dict(P=P, nslots=2, mode="gpu_pull", source=source,
                              h2d_bytes=P * L * B * (784 * xb + 8) if source == "host" else 0,
                              d2h_bytes=L * self.S * 4 if source == "host" else 0)
        return self.host_feed

    def make_pull_runner(self, round_graphs):
        """Capture the two staging graphs and build the native two-stream driver."""
        copy_graphs = []
        side = torch.cuda.Stream(device=self.pr.device)
        for b in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self.gather_ops[b].launch()
            copy_graphs.append(g)
        self._graphs = (copy_graphs, round_graphs)
        stream = torch.cuda.current_stream(self.pr.device).cuda_stream
        self.runner = self.ext.PullRunner([g.raw_cuda_graph_exec() for g in copy_graphs],
                                          [g.raw_cuda_graph_exec() for g in round_graphs], stream)
        return self.runner

    def make_runner(self, graphs, nslots):
        """Native per-round driver (csrc/runtime.cpp: HostFedRunner) over the two captured round graphs."""
        self._graphs = graphs
        stream = torch.cuda.current_stream(self.pr.device).cuda_stream
        self.runner = self.ext.HostFedRunner(
            self.loader, [g.raw_cuda_graph_exec() for g in graphs], stream,
            [self.x_pin[s].data_ptr() for s in range(nslots)], [self.y_pin[s].data_ptr() for s in range(nslots)],
            [self.bs_pin[s].data_ptr() for s in range(nslots)],
            [self.x_stage[b].data_ptr() for b in range(2)], [self.y_stage[b].data_ptr() for b in range(2)],
            [self.bs_stage[b].data_ptr() for b in range(2)],
            self.x_pin[0].numel() * self.x_pin.element_size(), self.y_pin[0].numel() * 8, self.bs_pin[0].numel() * 4)
        return self.runner

    def loss_readback(self):
        """Enqueue the D2H read of the round's per-node losses (nothing to enqueue when the kernel mirrors them
        into the pinned host buffer itself)."""
        if getattr(self, "loss_mode", "memcpy") == "memcpy":
            self.loss_host.copy_(self.loss_part, non_blocking=True)

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
        self.val_loss = torch.zeros(self.L, V, dtype=self.dtype, device=dev)
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
