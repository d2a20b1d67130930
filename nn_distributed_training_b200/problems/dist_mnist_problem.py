"""Distributed MNIST classification (reference: problems/dist_mnist_problem.py).

Metrics and their stored types follow SURVEY Appendix B:
``forward_pass_count`` int, ``validation_loss``/``top1_accuracy``/``current_epoch``
Tensor[N], ``consensus_error`` (Tensor[N,N], Tensor[N,1]), ``validation_as_vector``
dict node -> BoolTensor[|val|,1].
"""
from __future__ import annotations

import copy

import torch

from .base import ConsensusProblem, per_sample_loss, sum_of_batch_means
from ..models.spec import ConvNetSpec


class DistMNISTProblem(ConsensusProblem):
    def __init__(self, graph, base_model, base_loss, train_sets, val_set, device, conf, **kw):
        super().__init__(graph, base_model, base_loss, train_sets, val_set, device, conf, **kw)

    # ---- validation ------------------------------------------------------
    def _validate_local(self):
        """Per-sample NLL ``[L, V]`` and correctness ``[L, V]`` of every local node."""
        if self.fused is not None:
            return self.fused.validate()
        ps_fn = per_sample_loss(self.base_loss)
        V = len(self.val)
        L = self.placement.L
        ps = torch.zeros(L, V, device=self.device, dtype=self.dtype)
        ok = torch.zeros(L, V, device=self.device, dtype=torch.bool)
        chunk = 2048
        with torch.no_grad():
            for l, g in enumerate(self.placement.local_nodes):
                model = self.models[g]
                for a in range(0, V, chunk):
                    idx = torch.arange(a, min(V, a + chunk), device=self.device)
                    x, y = self.val.inputs(idx, self.dtype), self.val.targets(idx)
                    yh = model(x)
                    if ps_fn is not None:
                        ps[l, idx] = ps_fn(yh, y)
                    else:  # exotic loss object: fall back to batch-of-one evaluation
                        ps[l, idx] = torch.stack([self.base_loss(yh[k:k + 1], y[k:k + 1]) for k in range(len(idx))])
                    ok[l, idx] = yh.argmax(dim=1).eq(y)
        return ps, ok

    def validate(self, i):
        """(avg_loss, accuracy, correctness [|val|,1]) of node ``i``
        (dist_mnist_problem.py:111-132; ``avg_loss`` keeps the reference's
        sum-of-batch-means / |val| normalisation)."""
        ps, ok = self._validate_local()
        l = self.placement.local_index(i)
        loss = sum_of_batch_means(ps[l], self.val_batch_size) / ps.shape[1]
        return float(loss), float(ok[l].float().mean()), ok[l].reshape(-1, 1)

    # ---- metrics ---------------------------------------------------------
    def evaluate_metrics(self, at_end=False):
        need_val = any(k in self.metrics for k in ("validation_loss", "top1_accuracy", "validation_as_vector"))
        if need_val:
            ps, ok = self._validate_local()
            V = ps.shape[1]
            losses = self.gather_rows(sum_of_batch_means(ps, self.val_batch_size) / V).cpu()
            accs = self.gather_rows(ok.to(self.dtype).mean(1)).cpu()
            self.true_val_loss = self.gather_rows(ps.mean(1)).cpu()  # plain mean, logged alongside (Q9)
            if "validation_as_vector" in self.metrics:
                ok_all = self.gather_rows(ok).cpu()
                valid_vecs = {g: ok_all[g].reshape(-1, 1) for g in range(self.N)}

        line = "| "
        for name in self.conf["metrics"]:
            if name == "consensus_error":
                d_all, d_mean = self._consensus_metric()
                d_all, d_mean = d_all.cpu(), d_mean.cpu()
                self.metrics[name].append((d_all, d_mean))
                line += "Consensus: {:.4f} - {:.4f} | ".format(d_mean.min().item(), d_mean.max().item())
            elif name == "validation_loss":
                self.metrics[name].append(losses)
                line += "Val Loss: {:.4f} - {:.4f} | ".format(losses.min().item(), losses.max().item())
            elif name == "top1_accuracy":
                self.metrics[name].append(accs)
                line += "Top1: {:.2f} - {:.2f} |".format(accs.min().item(), accs.max().item())
            elif name == "forward_pass_count":
                self.metrics[name].append(self.forward_cnt)
                line += "Num Forward: {} | ".format(self.forward_cnt)
            elif name == "current_epoch":
                ep = self.epoch_tracker
                self.metrics[name].append(copy.deepcopy(ep))
                line += "Ep Range: {} - {} | ".format(int(ep.min().item()), int(ep.max().item()))
            elif name == "validation_as_vector":
                self.metrics[name].append(valid_vecs)
            else:
                raise NameError("Unknown metric.")
        if self.ctx.is_main:
            print(line)
        return

    # ---- fused backend -----------------------------------------------------
    def _fused_supported(self) -> bool:
        spec = getattr(self.base_model, "spec", None)
        if not isinstance(spec, ConvNetSpec) or not isinstance(self.base_loss, torch.nn.NLLLoss):
            return False
        from ..ops import fused_available, mnist_kernel_supports
        return fused_available() and mnist_kernel_supports(spec, self.train_batch_size, self.dtype)

    def _setup_fused(self):
        from ..ops.mnist_fused import FusedMnist
        self.fused = FusedMnist(self)
