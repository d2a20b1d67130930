"""Offline lidar density mapping on a static graph (reference: problems/dist_dense_problem.py).

The reference module is stale against its own optimizers (``evaluate_metrics`` lacks the
``at_end`` argument every optimizer passes — SURVEY Q12); this implementation is on the current
problem API.  Metric formats follow the reference: ``consensus_error`` stores only the pairwise
distance matrix (:159-180), ``mesh_grid_density`` is evaluated at every evaluation.
"""
from __future__ import annotations

import copy

import torch

from .density_common import DensityProblemBase


class DistDensityProblem(DensityProblemBase):
    def __init__(self, graph, base_model, base_loss, train_sets, val_set, device, conf, **kw):
        super().__init__(graph, base_model, base_loss, train_sets, val_set, device, conf, **kw)
        if "mesh_grid_density" in self.metrics:
            self._setup_mesh(val_set)

    def evaluate_metrics(self, at_end=False):
        line = "| "
        for name in self.conf["metrics"]:
            if name == "consensus_error":
                d_all, _ = self._consensus_metric()
                d_all = d_all.cpu()
                davg = d_all.sum(dim=1) / self.N
                self.metrics[name].append(d_all)
                line += "Consensus: {:.4f} - {:.4f} | ".format(davg.min().item(), davg.max().item())
            elif name == "validation_loss":
                vl = self.gather_rows(self._val_losses_local()).cpu()
                self.metrics[name].append(vl)
                line += "Val Loss: {:.4f} - {:.4f} | ".format(vl.min().item(), vl.max().item())
            elif name == "mesh_grid_density":
                self.metrics[name].append(self._mesh_all())
            elif name == "forward_pass_count":
                self.metrics[name].append(self.forward_cnt)
                line += "Num Forward: {} | ".format(self.forward_cnt)
            elif name == "current_epoch":
                ep = self.epoch_tracker
                self.metrics[name].append(copy.deepcopy(ep))
                line += "Ep Range: {} - {} | ".format(int(ep.min().item()), int(ep.max().item()))
            else:
                raise NameError("Unknown metric.")
        if self.ctx.is_main:
            print(line)
        return
