"""Checkpoint / resume of a consensus-optimizer run.

The reference cannot resume a supervised run: only end-of-run metrics (and, online, the
models) are written; duals, trackers, Adam moments, rho/alpha, data cursors are lost
(SURVEY §5.4).  A checkpoint here is one file per rank,
``<dir>/<problem>_ckpt_rank<r>.pt``, holding the optimizer state of the rank's nodes
(theta, dual | y,g, Adam m/v, step counters), the round index, the problem's draw counters
(which *are* the data cursors: sampling is a pure function of them) and the metrics so far.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch


class Checkpointer:
    def __init__(self, directory: str, name: str, every: int, ctx):
        self.dir, self.name, self.every, self.ctx = directory, name, int(every), ctx
        self._last_saved = -1

    def path(self) -> str:
        return os.path.join(self.dir, f"{self.name}_ckpt_rank{self.ctx.rank}.pt")

    def next_save_after(self, k: int) -> int:
        """First round > k at which a checkpoint is due (a huge sentinel when periodic saving is off, e.g. a
        ``resume: true`` run without ``checkpoint_every``)."""
        if self.every <= 0:
            return 2 ** 62
        return (k // self.every + 1) * self.every

    def maybe_save(self, opt):
        if self.every > 0 and opt.k % self.every == 0 and opt.k != self._last_saved:
            self.save(opt)

    def save(self, opt):
        os.makedirs(self.dir, exist_ok=True)
        pr = opt.pr
        prog = getattr(opt, "_program", None)
        if prog is not None:
            torch.cuda.synchronize(pr.device)
            prog.sync_back()
        payload = {"optimizer": opt.state_dict(), "alg": opt.alg_name,
                   "calls": np.asarray(pr.calls).copy(), "forward_cnt": pr.forward_cnt,
                   "metrics": pr.metrics, "world_size": self.ctx.world_size,
                   "graph_round": int(getattr(pr, "_graph_round", 0))}
        if hasattr(pr, "tloss_local"):
            payload["tloss_local"] = pr.tloss_local.cpu()
        tmp = self.path() + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, self.path())
        self._last_saved = opt.k

    def load(self, opt) -> bool:
        p = self.path()
        if not os.path.exists(p):
            return False
        payload = torch.load(p, map_location="cpu", weights_only=False)
        if payload["alg"] != opt.alg_name or payload["world_size"] != self.ctx.world_size:
            raise RuntimeError("checkpoint was written by a different algorithm / world size")
        opt.load_state_dict(payload["optimizer"])
        pr = opt.pr
        pr.calls[:] = payload["calls"]
        pr.forward_cnt = payload["forward_cnt"]
        pr.metrics = payload["metrics"]
        if "tloss_local" in payload and hasattr(pr, "tloss_local"):
            pr.tloss_local.copy_(payload["tloss_local"].to(pr.device))
        if hasattr(pr, "_graph_round"):
            pr._graph_round = int(payload.get("graph_round", opt.k))     # fault-injection schedule of the PyTorch path
        # the fused forward/backward kernels sample from DEVICE draw counters: mirror the restored host counters, also when
        # the consensus ops run on the PyTorch path (RoundProgram does this itself when it is built)
        fused = getattr(pr, "fused", None)
        if fused is not None and hasattr(fused, "sync_calls_from_host"):
            fused.sync_calls_from_host()
        self._last_saved = opt.k
        return True


def attach(opt, directory: str, name: str, every: int, ctx, resume: bool = False) -> Checkpointer:
    cp = Checkpointer(directory, name, every, ctx)
    opt.checkpointer = cp
    if resume:
        if cp.load(opt):
            if ctx.is_main:
                print(f"resumed {name} at round {opt.k} from {cp.path()}")
        elif ctx.is_main:
            print(f"[nndt] WARNING: resume requested but no checkpoint at {cp.path()}: starting {name} from round 0", flush=True)
    return cp


def attach_from_conf(opt, opt_conf, output_dir: str, name: str, ctx) -> Optional[Checkpointer]:
    """YAML keys (extensions): ``checkpoint_every`` (rounds, 0 = off), ``checkpoint_dir``, ``resume``.

    The run's output directory is time-stamped (``<metadir>/<Y-m-d_H-M>_<name>``), so a checkpoint written there cannot
    be found by a second invocation.  ``checkpoint_dir`` therefore defaults to the STABLE sibling
    ``<metadir>/<name>_ckpt`` (run name = the part of the output directory after the time stamp): the two-invocation
    flow is simply "same YAML with ``resume: true``"."""
    every = int(opt_conf.get("checkpoint_every", 0) or 0)
    resume = bool(opt_conf.get("resume", False))
    if every <= 0 and not resume:
        return None
    directory = opt_conf.get("checkpoint_dir")
    if not directory:
        meta, leaf = os.path.split(os.path.normpath(output_dir))
        parts = leaf.split("_", 2)          # <date>_<time>_<name>
        run = parts[2] if len(parts) == 3 else leaf
        directory = os.path.join(meta, run + "_ckpt")
    return attach(opt, directory, name, max(every, 0), ctx, resume=resume)
