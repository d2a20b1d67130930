"""Online lidar density mapping with a time-varying communication graph
(reference: problems/dist_online_dense_problem.py).

Each robot streams its trajectory through a sliding window of scans; the graph is the
Euclidean disk graph over the robots' *current* window poses (:141-155).  Because window
position is a pure function of how many samples a node has drawn, every rank can compute
every robot's pose — and hence the whole graph schedule of a run — locally:
``plan_graphs`` turns it into device tables indexed by the round counter, so a dynamic
topology costs nothing on the fused path.
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from .density_common import DensityProblemBase
from ..data.sampler import BatchSchedule
from ..data.shards import ShardSet, as_shard
from ..utils import graph_generation


class DistOnlineDensityProblem(DensityProblemBase):
    def __init__(self, base_model, base_loss, train_sets, val_set, device, conf, **kw):
        self.comm_radius = conf["comm_radius"]
        self.dynamic_graph = conf["dynamic_graph"]   # stored, never consulted (as in the reference :27)
        self._datasets = train_sets
        super().__init__(None, base_model, base_loss, train_sets, val_set, device, conf, **kw)
        self.update_graph()
        self.track_tloss = "train_loss_moving_average" in self.metrics
        if self.track_tloss:
            self.tloss_decay = conf["metrics_config"]["tloss_decay"]
            self.tloss_local = torch.zeros(self.placement.L, device=self.device, dtype=self.dtype)
        if "mesh_grid_density" in self.metrics:
            self._setup_mesh(val_set)

    # ---- streaming sampler --------------------------------------------------------
    def draws_after(self, g: int, calls: int) -> int:
        """Samples node ``g`` has consumed after ``calls`` minibatch draws (DataLoader epochs
        over ``len(dataset)`` points with a partial last batch)."""
        sched = self.schedules[g]
        epoch, b = divmod(int(calls), sched.batches_per_epoch)
        return epoch * sched.m + min(b * sched.batch_size, sched.m)

    def _draw_indices(self, g: int) -> torch.Tensor:
        sched = self.schedules[g]
        c = int(self.calls[g])
        _, _, size = sched.locate(c)
        ds = self._datasets[g]
        idx = ds.schedule.indices(self.draws_after(g, c), size, self.seed, g, device=self.device)
        self._count_draw(g)
        return idx

    def positions(self, calls=None) -> np.ndarray:
        calls = self.calls if calls is None else calls
        return np.vstack([self._datasets[g].pos_after(self.draws_after(g, calls[g])).reshape(1, 2)
                          for g in range(self.N)])

    # ---- graph ----------------------------------------------------------------------
    def update_graph(self):
        self.graph, connected = graph_generation.euclidean_disk_graph(self.positions(), self.comm_radius)
        if self._faults is not None:
            self.graph = self.faulted_graph(self.graph, self._graph_round)
            self._graph_round += 1
        if not connected and self.ctx.is_main:
            print("** WARNING: the communication graph is not connected. **")
        return

    def plan_graphs(self, oits, k0, draws_per_round, init_draws=0, refresh=True):
        out = []
        base = self.calls.copy()
        frozen = None
        for k in range(oits):
            if refresh or frozen is None:
                calls = base + init_draws + max(0, k - k0) * draws_per_round
                g, _ = graph_generation.euclidean_disk_graph(self.positions(calls), self.comm_radius)
                g = self.faulted_graph(g, k)
                frozen = g
            out.append(frozen if not refresh else g)
        return out

    # ---- loss hooks --------------------------------------------------------------------
    def _loss(self, model, x, y):
        yh = model(x)
        if torch.isnan(yh).any():   # fail fast (reference :118-126)
            print(torch.norm(torch.nn.utils.parameters_to_vector(model.parameters())))
            raise NameError("NaN again")
        return self.base_loss(torch.squeeze(yh), y.to(yh.dtype))

    def _after_loss(self, i, loss):
        if self.track_tloss:
            l = self.placement.local_index(i)
            self._ema_update(self.tloss_local[l: l + 1], loss.detach().reshape(1))

    def _ema_update(self, tracker, loss):
        d = self.tloss_decay
        tracker.copy_(torch.where(tracker != 0.0, (1 - d) * tracker + d * loss, loss))

    @property
    def tloss_tracker(self) -> torch.Tensor:
        return self.gather_rows(self.tloss_local).cpu()

    # ---- outputs -------------------------------------------------------------------------
    def save_metrics(self, output_dir):
        super().save_metrics(output_dir)
        if self.conf["save_models"]:
            self.save_models(output_dir)
        return

    def evaluate_metrics(self, at_end=False):
        line = "| "
        for name in self.conf["metrics"]:
            if name == "consensus_error":
                d_all, d_mean = self._consensus_metric()
                d_all, d_mean = d_all.cpu(), d_mean.cpu()
                self.metrics[name].append((d_all, d_mean))
                line += "Consensus: {:.4f} - {:.4f} | ".format(d_mean.min().item(), d_mean.max().item())
            elif name == "validation_loss":
                vl = self.gather_rows(self._val_losses_local()).cpu()
                self.metrics[name].append(vl)
                line += "Val Loss: {:.4f} - {:.4} - {:.4f} | ".format(vl.min().item(), vl.mean().item(), vl.max().item())
            elif name == "train_loss_moving_average":
                tl = self.tloss_tracker
                self.metrics[name].append(tl.clone())
                line += "Train Loss MA: {:.4f} - {:.4f} | ".format(tl.min().item(), tl.max().item())
            elif name == "mesh_grid_density":
                if not self.conf["metrics_config"]["mesh_only_at_end"] or at_end:
                    self.metrics[name].append(self._mesh_all())
            elif name == "forward_pass_count":
                self.metrics[name].append(self.forward_cnt)
                line += "Num Forward: {} | ".format(self.forward_cnt)
            elif name == "current_epoch":
                ep = self.epoch_tracker
                self.metrics[name].append(copy.deepcopy(ep))
                line += "Ep Range: {} - {} | ".format(int(ep.min().item()), int(ep.max().item()))
            elif name == "current_position":
                self.metrics[name].append(self.positions())
            elif name == "current_graph":
                self.metrics[name].append(copy.deepcopy(self.graph))
            else:
                raise NameError("Unknown metric.")
        if self.ctx.is_main:
            print(line)
        return
