"""Fused execution of the consensus optimizers: CUDA-graph round programs.

A round is a short, fixed kernel sequence
    DiNNO:  [fwd/bwd, dinno_update(p)] x primal_iterations
            (MNIST: ONE cluster kernel per round, csrc/dinno_round.cu — fwd/bwd and the update of
             every primal iteration separated by cluster barriers instead of kernel boundaries)
    DSGD :  dsgd_mix, fwd/bwd, dsgd_step
    DSGT :  dsgt_mix, fwd/bwd, dsgt_track
whose per-round scalars come from device schedules indexed by a device round
counter, so ``R`` consecutive rounds are captured once as a CUDA graph and
replayed between evaluation points with no host work (the reference issues
~800 serial micro-launches per round from Python).  When the
model has no fused forward/backward kernel the same consensus kernels run
eagerly around an autograd step.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from .engine import ConsensusEngine

MAX_ROUNDS_PER_GRAPH = 64
PULL_ROUNDS_PER_GRAPH = 64     # host-fed / staged rounds captured per graph (staging kernel forked inside the graph); a
                               # production chunk (evaluate_frequency rounds, 20 in the PAPER configs) is ONE graph launch


def _nvtx(name):
    """NVTX range per phase (visible in Nsight / torch.profiler traces; no-op cost when not profiling)."""
    return torch.cuda.nvtx.range(name)


def _round_ops(opt, eng, grads, round_op=None, publish=None):
    with _nvtx(f"consensus_round/{opt.alg_name}"):
        _round_ops_impl(opt, eng, grads, round_op)
        if eng.separate_publish:
            (publish or eng.op.publish)()      # announce the round to the peers (forked branch under capture)


def _round_ops_impl(opt, eng, grads, round_op=None):
    alg = opt.alg_name
    if eng.sum_mode:
        eng.op.local_sum()   # complete graph: per-rank partial sums feeding the NVLS reduction
    if round_op is not None:
        round_op.launch()    # whole DiNNO round (all primal iterations) in one cluster launch
    elif alg == "dinno":
        for p in range(opt.pits):
            grads(p)
            eng.op.dinno_update(p)
    elif alg == "dsgd":
        eng.op.dsgd_mix()
        grads(0)
        eng.op.dsgd_step()
    elif alg == "dsgt":
        eng.op.dsgt_mix()
        grads(0)
        eng.op.dsgt_track()
    else:  # pragma: no cover
        raise NameError("Unknown distributed opt algorithm.")


def draws_per_round(opt) -> int:
    return opt.pits if opt.alg_name == "dinno" else 1


class RoundProgram:
    """Owns the engine and the captured graphs for one optimizer."""

    def __init__(self, opt):
        self.opt = opt
        pr = self.pr = opt.pr
        self.dpr = draws_per_round(opt)
        init_draws = 1 if (opt.alg_name == "dsgt" and opt.init_grads and not opt._initialised) else 0
        graphs = pr.plan_graphs(opt.oits, opt.k, self.dpr, init_draws,
                                refresh=getattr(opt, "refresh_graph", True))
        self.capturable = pr.fused is not None and os.environ.get("NNDT_NO_GRAPH", "0") != "1"
        # ---- input pipeline of the fused MNIST problem (decided first: the engine's peer-announcement mode depends on
        #      whether the rounds run in forked multi-round graphs) -----------------------------------------------------
        pipeline = "resident"
        if pr.fused is not None:
            pipeline = pr.conf.get("input_pipeline", "auto")
            can_stage = hasattr(pr.fused, "enable_host_feed") and self.capturable
            if pipeline == "auto":
                # staged-resident is the faster way to run resident shards (4 % on the headline round): the training
                # kernel reads a compact, L2-resident batch instead of chasing the sampler through HBM
                pipeline = "staged" if can_stage else "resident"
            if not hasattr(pr.fused, "enable_host_feed"):
                pipeline = "resident"
        forked = (pipeline == "staged" or (pipeline == "host" and pr.conf.get("host_gather", "gpu_pull") == "gpu_pull")) \
            and os.environ.get("NNDT_PULL_DRIVER", pr.conf.get("host_pull_driver", "graph")) == "graph"
        self.eng = ConsensusEngine(opt, graphs, forked_graphs=forked)
        self.graph_plan = graphs
        # evaluation between rounds can use the fused consensus-metric kernel on the published rows
        pr._metric_engine = (self.eng, lambda: opt.k)
        self._graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.host_mode = False
        self._round_ops = None
        self.pipeline = "resident"
        self._pub_side = None
        self._pub_pending = False
        if pr.fused is not None:
            pr.fused.sync_calls_from_host()
            self._deferred_pipeline = None
            if pipeline in ("host", "staged") and hasattr(pr.fused, "enable_host_feed"):
                if init_draws:
                    # DSGT init_grads (optimizers/dsgt.py:33-46 of the reference): the ONE initial gradient draw runs on the
                    # resident shards; the host-fed / staged stream starts at the draw after it (dsgt_init switches over)
                    self._deferred_pipeline = pipeline
                else:
                    self._enable_pipeline(pipeline)
            # opt-in: measured slower than the PDL-overlapped per-step kernels (docs/perf_notes.md), kept as the
            # in-kernel phase profiler (scripts/profile_round_phases.py) and for launch-bound environments
            want = pr.conf.get("fused_round", opt.conf.get("fused_round", False)) or os.environ.get("NNDT_FUSED_ROUND") == "1"
            if (want and os.environ.get("NNDT_NO_FUSED_ROUND", "0") != "1"
                    and getattr(pr.fused, "supports_round_kernel", lambda o: False)(opt)):
                sets = [0, 1] if self.host_mode else [None]
                self._round_ops = {b: pr.fused.round_op(self.eng._keep, b) for b in sets}

    def _enable_pipeline(self, pipeline: str):
        pr = self.pr
        pr.fused.enable_host_feed(self.dpr, nslots=int(pr.conf.get("host_slots", 4)),
                                  threads=int(pr.conf.get("host_threads", 4)),
                                  mode="staged" if pipeline == "staged" else pr.conf.get("host_gather", "gpu_pull"))
        self.host_mode = True
        self.pipeline = pipeline
        self._runner = None
        self._stage_set = 0
        self._pull_graphs: Dict = {}
        self._pull_parity = 0
        self._pull_primed = False
        self._side = None

    # ---- peer announcement off the critical path ---------------------------------------------------------------
    def _publish_forked(self):
        """Under capture: run publish_round_kernel on a side stream that forks after the round's last kernel, so the
        system fence + remote flag stores overlap the next round's forward/backward; joined at the end of the graph."""
        if self._pub_side is None:
            self._pub_side = torch.cuda.Stream(device=self.pr.device)
        main = torch.cuda.current_stream(self.pr.device)
        self._pub_side.wait_stream(main)
        with torch.cuda.stream(self._pub_side):
            self.eng.op.publish()
        self._pub_pending = True

    def _join_publish(self):
        if self._pub_pending:
            torch.cuda.current_stream(self.pr.device).wait_stream(self._pub_side)
            self._pub_pending = False

    def round_op(self):
        if self._round_ops is None:
            return None
        return self._round_ops[self._stage_set if self.host_mode else None]

    def launches_per_round(self) -> int:
        """Kernel launches of one communication round (the staging kernel of the host-fed / staged pipelines and the
        forked peer announcement included)."""
        n = 1 if self.eng.sum_mode else 0
        if self.host_mode and self.pr.fused.host_feed["mode"] == "gpu_pull":
            n += 1
        if self.eng.separate_publish:
            n += 1
        if self._round_ops is not None:
            return n + 1
        return n + (2 * self.opt.pits if self.opt.alg_name == "dinno" else 3)

    def grads(self, p: int = 0):
        pr = self.pr
        if pr.fused is None:
            pr.compute_grads()
        elif self.host_mode:
            pr.fused.direct_ops[self._stage_set][p].train()
        else:
            pr.fused.launch()

    def _count(self, rounds: int):
        """Host mirror of the device-side draw counters."""
        if self.pr.fused is not None:
            self.pr.count_draws_all(rounds * self.dpr)

    def dsgt_init(self):
        self.grads()
        self.eng.op.dsgt_init()
        if self.pr.fused is not None:
            self.pr.count_draws_all(1)
            if getattr(self, "_deferred_pipeline", None):
                torch.cuda.synchronize(self.pr.device)      # the kernel-owned draw counters have advanced
                self._enable_pipeline(self._deferred_pipeline)
                self._deferred_pipeline = None

    def _capture_pull_graph(self, R: int, parity: int):
        """``R`` host-fed rounds as ONE graph with two branches: while round i computes on the capture stream,
        the staging kernel pulls round i+1's rows out of pinned host memory on a forked stream (device-initiated
        H2D over PCIe); every round ends with the D2H read of its losses.  No CPU work per round at all."""
        fz = self.pr.fused
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.pr.device)
        side = self._side
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream(self.pr.device)
            for i in range(R):
                b = (parity + i) & 1
                # set b^1 was last read by round i-1, which precedes this point on `main`
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    fz.gather_ops[b ^ 1].launch()
                self._stage_set = b
                _round_ops(self.opt, self.eng, self.grads, self.round_op(), self._publish_forked)
                fz.loss_readback()
                main.wait_stream(side)     # round i+1 consumes what was just staged (also joins the fork)
            self._join_publish()
        return g

    def _pull_graph(self, r: int, parity: int):
        key = (r, parity)
        g = self._pull_graphs.get(key)
        if g is None:
            g = self._pull_graphs[key] = self._capture_pull_graph(r, parity)
        return g

    def _run_pull_graphs(self, rounds: int, capture_only: bool = False):
        fz = self.pr.fused
        if not self._pull_primed:
            fz.gather_ops[self._pull_parity].launch()      # stage the very first round
            self._pull_primed = True
        left, parity = rounds, self._pull_parity
        while left > 0:
            r = min(left, PULL_ROUNDS_PER_GRAPH)
            g = self._pull_graph(r, parity)
            parity = (parity + r) & 1
            left -= r
            if not capture_only:
                g.replay()
                self._pull_parity = parity
                self._count(r)

    def _run_host_fed(self, rounds: int):
        """Host-fed rounds.  ``gpu_pull`` (default): multi-round graphs with the staging kernel forked inside
        (``_capture_pull_graph``); ``host_pull_driver: runner`` keeps the native two-stream driver
        (csrc/runtime.cpp: PullRunner) that launches one staging graph + one round graph per round.
        ``cpu_loader``: the native runner issues, per round, the H2D copy of that round's inputs, the captured
        round graph (kernels + D2H loss read) and the slot hand-back to the loader threads."""
        fz = self.pr.fused
        if fz.host_feed["mode"] == "gpu_pull" and os.environ.get("NNDT_PULL_DRIVER", self.pr.conf.get("host_pull_driver", "graph")) == "graph":
            return self._run_pull_graphs(rounds)
        if self._runner is None:
            graphs = []
            for b in range(2):
                self._stage_set = b
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    _round_ops(self.opt, self.eng, self.grads, self.round_op())
                    fz.loss_readback()
                graphs.append(g)
            if fz.host_feed["mode"] == "gpu_pull":
                self._runner = fz.make_pull_runner(graphs)
            else:
                self._runner = fz.make_runner(graphs, fz.host_feed["nslots"])
        # graphs were captured on torch's capture stream but are launched on the current stream
        self._runner.run(rounds)
        self._count(rounds)

    def _resident_graph(self, r: int):
        g = self._graphs.get(r)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(r):
                    _round_ops(self.opt, self.eng, self.grads, self.round_op(), self._publish_forked)
                self._join_publish()
            self._graphs[r] = g
        return g

    def prepare(self, rounds: int):
        """Capture (without executing) every CUDA graph that ``run(rounds)`` will replay from the current state, so a
        following ``run`` issues graph launches only."""
        if not self.capturable:
            return
        if self.host_mode:
            fz = self.pr.fused
            if fz.host_feed["mode"] == "gpu_pull" and os.environ.get("NNDT_PULL_DRIVER", self.pr.conf.get("host_pull_driver", "graph")) == "graph":
                self._run_pull_graphs(rounds, capture_only=True)
            return
        left = rounds
        while left > 0:
            r = min(left, MAX_ROUNDS_PER_GRAPH)
            self._resident_graph(r)
            left -= r

    def run(self, rounds: int):
        """Execute ``rounds`` consecutive rounds starting at the device round counter."""
        if self.host_mode:
            return self._run_host_fed(rounds)
        left = rounds
        while left > 0:
            r = min(left, MAX_ROUNDS_PER_GRAPH)
            if self.capturable:
                self._resident_graph(r).replay()
            else:
                for _ in range(r):
                    _round_ops(self.opt, self.eng, self.grads, self.round_op())
            self._count(r)
            left -= r

    def sync_back(self):
        """Mirror device-resident optimizer state into the optimizer object."""
        opt, eng = self.opt, self.eng
        L = self.pr.placement.L
        if opt.alg_name == "dsgt":
            par = opt.k & 1
            opt.y.copy_(eng.pub[par, 1, :L])
        if opt.alg_name == "dinno" and opt.k > 0:
            opt.rho = opt.rho_at(opt.k - 1)
        if opt.alg_name == "dsgd" and opt.k > 0:
            opt.alph = opt.alpha_table()[opt.k - 1]


def run_fused_training(opt, profiler=None):
    pr = opt.pr
    prog = getattr(opt, "_program", None)
    if prog is None:
        prog = opt._program = RoundProgram(opt)
    if opt.alg_name == "dsgt" and opt.init_grads and not opt._initialised:
        prog.dsgt_init()
    if opt.alg_name == "dsgt":
        opt._initialised = True
    every = opt._eval_every()
    oits = opt.oits
    while opt.k < oits:
        k = opt.k
        opt._maybe_eval(k)
        if k >= oits - 1:
            nxt = oits
        else:
            nxt = min((k // every + 1) * every, oits - 1)
        if profiler is not None or opt.checkpointer is not None:
            nxt = k + 1 if profiler is not None else min(nxt, opt.checkpointer.next_save_after(k))
        prog.run(nxt - k)
        opt.k = nxt
        if profiler is not None:
            profiler.step()
        if opt.checkpointer is not None:
            prog.sync_back()
            opt.checkpointer.maybe_save(opt)
    torch.cuda.synchronize(pr.device)
    prog.eng.check()
    prog.sync_back()
