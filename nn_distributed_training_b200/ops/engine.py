"""Consensus engine: device tables + fused kernel ops for one optimizer run.

Built once per ``train()``; afterwards a communication round is a fixed sequence
of kernel launches that reads every per-round scalar from device memory, so the
sequence is captured in a CUDA graph and replayed (``round_program.py``).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np
import torch

from . import load_ext
from ..parallel.symm import SymmetricBuffer
from ..utils.graph_generation import Topology

OPT_CODE = {"sgd": 0, "adam": 1, "adamw": 2}


class ConsensusEngine:
    def __init__(self, opt, graphs_per_round: List, forked_graphs: bool = False):
        self.opt = opt
        pr = self.pr = opt.pr
        self.ext = load_ext(required=True)
        dev, a, pl, ctx = pr.device, pr.arena, pr.placement, pr.ctx
        self.dtype = a.dtype
        npdt = np.float32 if self.dtype == torch.float32 else np.float64
        self.C = 2 if opt.alg_name == "dsgt" else 1
        L, n_pad, oits = pl.L, a.n_pad, opt.oits

        # ---- published rows (double buffered, peer mapped when multi-GPU) -----
        Lmax = max(pl.counts)
        self.pub_buf = SymmetricBuffer((2, self.C, Lmax, n_pad), self.dtype, ctx)
        self.pub = self.pub_buf.local
        self.Lpub = Lmax
        k0 = opt.k
        # round k0 (0, or the round a checkpoint resumed at) is "published" in the parity it will be read from
        self.pub[k0 & 1, 0, :L].copy_(a.theta)
        if opt.alg_name == "dsgt" and getattr(opt, "_initialised", False):
            self.pub[k0 & 1, 1, :L].copy_(opt.y)

        # ---- schedules ----------------------------------------------------------
        rho = np.zeros(oits); lr = np.zeros(oits); alpha = np.zeros(oits)
        if opt.alg_name == "dinno":
            rho[:] = [opt.rho_at(k) for k in range(oits)]
            lr[:] = [opt.lr_at(k) for k in range(oits)]
        elif opt.alg_name == "dsgd":
            alpha[:] = opt.alpha_table()
        else:
            alpha[:] = opt.alpha
        self.rho = torch.as_tensor(rho.astype(npdt), device=dev)
        self.lr = torch.as_tensor(lr.astype(npdt), device=dev)
        self.alpha = torch.as_tensor(alpha.astype(npdt), device=dev)

        # ---- topology tables ------------------------------------------------------
        topos: List[Topology] = []
        key_to_id: Dict[bytes, int] = {}
        gid = np.zeros(oits, dtype=np.int32)
        for k, g in enumerate(graphs_per_round):
            t = pr._topo_cache.get(g) if hasattr(pr, "_topo_cache") else Topology(g)
            if t.key not in key_to_id:
                key_to_id[t.key] = len(topos)
                topos.append(t)
            gid[k] = key_to_id[t.key]
        self.topos = topos
        G = len(topos)
        dmax = max(1, max(t.max_degree for t in topos))
        itemsize = a.theta.element_size()
        nbr_ptr = np.zeros((G, L, dmax, 2, self.C), dtype=np.int64)
        nbr_w = np.zeros((G, L, dmax), dtype=npdt)
        self_w = np.zeros((G, L), dtype=npdt)
        deg = np.zeros((G, L), dtype=np.int32)
        nbr_rank = -np.ones((G, L, dmax), dtype=np.int32)
        for gi, t in enumerate(topos):
            for l, g in enumerate(pl.local_nodes):
                nb = t.neighbors_noself[g]
                deg[gi, l] = len(nb)
                self_w[gi, l] = t.W[g, g]
                for e, j in enumerate(nb):
                    r, lj = int(pl.node_rank[j]), int(pl.node_local[j])
                    nbr_w[gi, l, e] = t.W[g, j]
                    if r != ctx.rank:
                        nbr_rank[gi, l, e] = r
                    for par in range(2):
                        for ch in range(self.C):
                            row = ((par * self.C + ch) * self.Lpub + lj) * n_pad
                            nbr_ptr[gi, l, e, par, ch] = self.pub_buf.peer_ptrs[r] + row * itemsize
        self.dmax = dmax
        self.t_nbr_ptr = torch.as_tensor(nbr_ptr, device=dev)
        self.t_nbr_w = torch.as_tensor(nbr_w, device=dev)
        self.t_self_w = torch.as_tensor(self_w, device=dev)
        self.t_deg = torch.as_tensor(deg, device=dev)
        self.t_nbr_rank = torch.as_tensor(nbr_rank, device=dev)
        self.t_gid = torch.as_tensor(gid, device=dev)

        # ---- optional protocol self-check: published rows carry their round, neighbor reads verify it ----
        self.seq_buf = None
        self.t_nbr_seq = None
        if opt.conf.get("debug_sequence_check", False) or os.environ.get("NNDT_SEQ_CHECK") == "1":
            self.seq_buf = SymmetricBuffer((2, Lmax), torch.int32, ctx)
            self.seq_buf.local.fill_(k0 - 1)
            self.seq_buf.local[k0 & 1].fill_(k0)
            nbr_seq = np.zeros((G, L, dmax, 2), dtype=np.int64)
            for gi, t in enumerate(topos):
                for l, g in enumerate(pl.local_nodes):
                    for e, j in enumerate(t.neighbors_noself[g]):
                        r, lj = int(pl.node_rank[j]), int(pl.node_local[j])
                        for par in range(2):
                            nbr_seq[gi, l, e, par] = self.seq_buf.peer_ptrs[r] + (par * Lmax + lj) * 4
            self.t_nbr_seq = torch.as_tensor(nbr_seq, device=dev)

        # ---- counters / flags -----------------------------------------------------
        self.round_ctr = torch.full((1,), k0, dtype=torch.int32, device=dev)
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.flag_buf = SymmetricBuffer((max(ctx.world_size, 1),), torch.int32, ctx)
        self.flag_buf.local.fill_(k0)
        peer_flag = np.zeros(max(ctx.world_size, 1), dtype=np.int64)
        for r in range(ctx.world_size):
            peer_flag[r] = self.flag_buf.peer_ptrs[r] + 4 * ctx.rank
        self.t_peer_flag = torch.as_tensor(peer_flag, device=dev)
        # pull transport: rank r's own counter is slot r of its own array; readers poll it over NVLink
        peer_pub = np.zeros(max(ctx.world_size, 1), dtype=np.int64)
        for r in range(ctx.world_size):
            peer_pub[r] = self.flag_buf.peer_ptrs[r] + 4 * r
        self.t_peer_pub = torch.as_tensor(peer_pub, device=dev)
        self.flag_mode = os.environ.get("NNDT_FLAG_MODE", str(opt.conf.get("flag_transport", pr.conf.get("flag_transport", "push"))))
        if self.flag_mode not in ("push", "pull"):
            raise ValueError(f"flag_transport must be push or pull, got {self.flag_mode!r}")
        # ranks that own a neighbor of a local node in ANY round's graph: the only ones that need this rank's flags
        notify = 0
        remote_node = np.zeros(L, dtype=bool)
        for r in np.unique(nbr_rank[nbr_rank >= 0]):
            notify |= 1 << int(r)
        for l in range(L):
            remote_node[l] = bool((nbr_rank[:, l, :] >= 0).any())
        self.notify_mask = notify
        # launch order of the local nodes: the ones pulling over NVLink first (their CTAs become resident while the
        # preceding forward/backward kernel still runs, which hides the link latency)
        order = np.argsort(~remote_node, kind="stable").astype(np.int32)
        self.t_node_order = torch.as_tensor(order, device=dev)
        if ctx.is_distributed:
            torch.cuda.synchronize(dev)
            ctx.barrier()

        # ---- complete graph: uniform Metropolis weights -> aggregates are functions of the network sum ----
        self.sum_mode = (G == 1 and topos[0].is_complete() and pr.N > 1
                         and opt.conf.get("complete_graph_mode", "sum") == "sum")
        self.sum_buf = self.sum_flag_buf = None
        sum_mc = None
        if self.sum_mode:
            self.sum_buf = SymmetricBuffer((2, self.C, n_pad), torch.float64, ctx)   # fp64: S - N*theta cancels in fp32
            self.sum_flag_buf = SymmetricBuffer((max(ctx.world_size, 1),), torch.int32, ctx)
            self.sum_flag_buf.local.fill_(k0)
            if ctx.is_distributed:
                if self.sum_buf.multicast_ptr:
                    sum_mc = self.sum_buf.multicast_ptr      # NVLS: one in-switch reduction per element
                else:
                    self.sum_mode = False                    # no multicast mapping on this fabric: pointer table
                torch.cuda.synchronize(dev)
                ctx.barrier()
        peer_sum_flag = np.zeros(max(ctx.world_size, 1), dtype=np.int64)
        if self.sum_mode:
            for r in range(ctx.world_size):
                peer_sum_flag[r] = self.sum_flag_buf.peer_ptrs[r] + 4 * ctx.rank
        self.t_peer_sum_flag = torch.as_tensor(peer_sum_flag, device=dev)

        # ---- gradient source --------------------------------------------------------
        if pr.fused is not None:
            grad_part, S, calls = pr.fused.grad_part, pr.fused.S, pr.fused.calls
            if getattr(pr.fused, "owns_calls", False):
                calls = None       # the forward/backward kernel advances its own draw counters
        else:
            grad_part, S, calls = a.grad, 1, None
        self.S = S

        d = dict(L=L, n_pad=n_pad, S=S, theta=a.theta.data_ptr(), grad_part=grad_part.data_ptr(),
                 pub=self.pub.data_ptr(), C=self.C, nbr_ptr=self.t_nbr_ptr.data_ptr(),
                 nbr_w=self.t_nbr_w.data_ptr(), self_w=self.t_self_w.data_ptr(), deg=self.t_deg.data_ptr(),
                 nbr_rank=self.t_nbr_rank.data_ptr(), dmax=dmax, round_ctr=self.round_ctr.data_ptr(),
                 rho=self.rho.data_ptr(), lr=self.lr.data_ptr(), alpha=self.alpha.data_ptr(),
                 graph_id=self.t_gid.data_ptr(), calls=None if calls is None else calls.data_ptr(),
                 flags=self.flag_buf.local.data_ptr(), peer_flag=self.t_peer_flag.data_ptr(),
                 world=ctx.world_size, rank=ctx.rank, done_ctr=self.done_ctr.data_ptr(), err=self.err.data_ptr(),
                 flag_pull=int(self.flag_mode == "pull"), peer_pub=self.t_peer_pub.data_ptr(),
                 notify_mask=int(self.notify_mask) if ctx.world_size > 1 else 0,
                 node_order=self.t_node_order.data_ptr() if ctx.world_size > 1 else None)
        self.timeline = None
        if os.environ.get("NNDT_TIMELINE") == "1":       # debug: %globaltimer stamps of the update kernels (scripts/timeline_rounds.py)
            self.timeline = torch.zeros(4096, 16, dtype=torch.int64, device=dev)
            d["timeline"] = self.timeline.data_ptr()
        # the C++ side indexes pub rows with stride L; when ranks host different node counts the
        # published buffer is allocated with the max count, so pass that as the row count of pub
        d["L"] = L
        d["pub_L"] = self.Lpub
        # multi-GPU: the "round published" flags can be written by publish_round_kernel on a forked graph branch
        # (round_program.py) instead of the round's last kernel.
        # Off by default.  The one-CTA publish kernel becomes ready together with the next round's forward/backward
        # kernel, whose PDL-launched update kernel fills every remaining register file with CTAs that spin on the
        # peers' flags: with the fp64 cluster kernel nothing is left for the publish CTA until the forward/backward
        # CTAs exit, i.e. every rank announces its round ~34 us late (scripts/timeline_rounds.py, 2 GPUs: flag wait
        # 33.9 us vs 2.5 us with the announcement inside the round's last kernel).  (In round 1 the fork gained
        # 1.9 us / round on the host-fed fp32 graphs and cost 2.9 us on the resident ones.)
        sp = opt.conf.get("separate_publish", pr.conf.get("separate_publish", False))
        if os.environ.get("NNDT_SEPARATE_PUBLISH") in ("0", "1"):       # A/B switch
            sp = os.environ["NNDT_SEPARATE_PUBLISH"] == "1"
        if sp == "auto":
            sp = forked_graphs
        self.separate_publish = bool(ctx.world_size > 1 and sp)
        # in-kernel announcement: "start" = by the first consensus kernel of the round that READS the rows (its block 0,
        # under the forward/backward kernel), "end" = by the last kernel of the round that wrote them (3-4 us of system
        # fence + NVLink stores on the critical path of every round)
        self.announce = os.environ.get("NNDT_ANNOUNCE", str(opt.conf.get("announce", pr.conf.get("announce", "start"))))
        if self.announce not in ("start", "end"):
            raise ValueError(f"announce must be 'start' or 'end', got {self.announce!r}")
        d["flags_in_kernel"] = 0 if self.separate_publish else (2 if self.announce == "start" else 1)
        if pr.fused is not None and getattr(pr, "track_tloss", False) and self.dtype == torch.float32:
            # the kernel that consumes a gradient also folds that step's loss into the EMA tracker
            d.update(loss_part=pr.fused.loss_part.data_ptr(), tloss=pr.tloss_local.data_ptr(),
                     tdecay=float(pr.tloss_decay), loss_S=int(pr.fused.loss_part.shape[1]))
            pr.fused.ema_in_kernel = True
        if self.seq_buf is not None:
            d.update(pub_seq=self.seq_buf.local.data_ptr(), nbr_seq=self.t_nbr_seq.data_ptr())
        if self.sum_mode:
            d.update(sum_mode=1, n_total=pr.N, sum_local=self.sum_buf.local.data_ptr(), sum_mc=sum_mc,
                     sum_flags=self.sum_flag_buf.local.data_ptr(), peer_sum_flag=self.t_peer_sum_flag.data_ptr())
        if opt.alg_name == "dinno":
            d.update(dual=opt.duals.data_ptr(), delta=opt.delta.data_ptr(),
                     m=None if opt.m is None else opt.m.data_ptr(),
                     v=None if opt.v is None else opt.v.data_ptr(),
                     pits=opt.pits, opt=OPT_CODE[opt.opt_kind], persistent=int(opt.persistent))
        if opt.alg_name == "dsgt":
            d.update(g_old=opt.g.data_ptr())
        cls = self.ext.ConsensusOpF32 if self.dtype == torch.float32 else self.ext.ConsensusOpF64
        self.op = cls(d)
        self._keep = d

    def consensus_metric(self, k: int):
        """Fused consensus-error metric (csrc/consensus.cu: consensus_metric_kernel) on the rows published
        for round ``k``: every rank pulls all N rows (local or NVLink peer) and produces the distance rows of
        its own nodes; returns (pairwise [N, N], to-mean [N, 1]) gathered over ranks, as float64 CPU tensors."""
        pr, pl, ctx, dev = self.pr, self.pr.placement, self.pr.ctx, self.pr.device
        N, L, n_pad = pr.N, pl.L, self.pr.arena.n_pad
        par = k & 1
        itemsize = self.pub.element_size()
        rows = np.zeros(N, dtype=np.int64)
        for g in range(N):
            r, lj = int(pl.node_rank[g]), int(pl.node_local[g])
            rows[g] = self.pub_buf.peer_ptrs[r] + ((par * self.C) * self.Lpub + lj) * n_pad * itemsize
        t_rows = torch.as_tensor(rows, device=dev)
        inv = torch.empty(N, dtype=torch.float64, device=dev)
        pair = torch.empty(L, N, dtype=torch.float64, device=dev)
        mean = torch.empty(L, dtype=torch.float64, device=dev)
        if ctx.is_distributed:
            torch.cuda.synchronize(dev)
            ctx.barrier()              # every rank's rows of round k are published and quiescent
        self.ext.consensus_metric(self.dtype == torch.float64, t_rows.data_ptr(), N, n_pad, pl.lo, L,
                                  inv.data_ptr(), pair.data_ptr(), mean.data_ptr())
        d_all = pr.gather_rows(pair).cpu()
        d_mean = pr.gather_rows(mean.reshape(-1, 1)).cpu()
        if ctx.is_distributed:
            ctx.barrier()              # nobody starts overwriting published rows while a peer still reads
        return d_all, d_mean

    def check(self):
        e = int(self.err.item())
        if e == 2:
            raise RuntimeError("sequence check failed: a neighbor row was read that is not tagged with the current round")
        if e != 0:
            raise RuntimeError("consensus kernel timed out waiting for a peer's published round")
