"""Shared plumbing of the experiment runners: output layout, device/rank setup, the
per-problem optimizer loop, profiler and checkpoint hooks.

Output layout (reference experiments/dist_mnist_ex.py:73-95,224-225):
    <output_metadir>/<YYYY-MM-DD_HH-MM>_<name>/{<time>.yaml, graph.gpickle, solo_results.pt,
    <problem_name>_results.pt, <problem_name>_models.pt, <problem_name>opt_profile/}
Only rank 0 writes.  When launched through ``torchrun`` every rank hosts a block of graph
nodes on its own GPU; launched plainly, all nodes share one device.
"""
from __future__ import annotations

import contextlib
import os
import pickle
from datetime import datetime
from shutil import copyfile

import torch

from ..optimizers import build_optimizer
from ..parallel.context import DistContext
from ..utils import checkpoint as ckpt


def make_context(exp_conf) -> DistContext:
    use_cuda = bool(exp_conf.get("use_cuda", True)) and torch.cuda.is_available()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        ctx = DistContext.from_env(use_cuda=use_cuda)
    else:
        ctx = DistContext.single(torch.device("cuda") if use_cuda else torch.device("cpu"))
    if ctx.is_main:
        print("Device is set to GPU" if ctx.device.type == "cuda" else "Device is set to CPU")
    return ctx


def setup_output(exp_conf, yaml_pth, ctx: DistContext) -> str:
    metadir = exp_conf["output_metadir"]
    time_now = datetime.now().strftime("%Y-%m-%d_%H-%M")
    time_now = ctx.broadcast_object(time_now)
    output_dir = os.path.join(metadir, time_now + "_" + exp_conf["name"])
    if exp_conf["writeout"] and ctx.is_main:
        os.makedirs(output_dir, exist_ok=True)
        copyfile(yaml_pth, os.path.join(output_dir, time_now + ".yaml"))
    exp_conf["output_dir"] = output_dir
    ctx.barrier()
    return output_dir


def write_gpickle(graph, path):
    """networkx>=3 dropped ``nx.write_gpickle``; the file was a plain pickle."""
    with open(path, "wb") as f:
        pickle.dump(graph, f, pickle.HIGHEST_PROTOCOL)


def read_gpickle(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def make_loss(name: str):
    table = {"NLL": torch.nn.NLLLoss, "BCE": torch.nn.BCELoss, "MSE": torch.nn.MSELoss, "L1": torch.nn.L1Loss}
    if name not in table:
        raise NameError("Unknown loss function.")
    return table[name]()


def make_solo_optimizer(model, conf):
    table = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD, "adamw": torch.optim.AdamW}
    if conf["optimizer"] not in table:
        raise NameError("Unknown individual optimizer.")
    return table[conf["optimizer"]](model.parameters(), lr=conf["lr"])


@contextlib.contextmanager
def maybe_profiler(enabled: bool, out_dir: str, name: str):
    """``optimizer_config.profile: true`` wraps training in the same torch.profiler schedule
    as the reference (experiments/dist_mnist_ex.py:207-220); optimizers call ``step()`` once
    per round."""
    if not enabled:
        yield None
        return
    with torch.profiler.profile(
        schedule=torch.profiler.schedule(wait=1, warmup=1, active=3, repeat=3),
        on_trace_ready=torch.profiler.tensorboard_trace_handler(os.path.join(out_dir, name + "opt_profile")),
        record_shapes=True, with_stack=True,
    ) as prof:
        yield prof


def run_problem(prob, prob_conf, exp_conf, ctx: DistContext, label=None):
    """Build the optimizer named by ``alg_name``, train, save metrics."""
    opt_conf = prob_conf["optimizer_config"]
    dopt = build_optimizer(prob, ctx.device, opt_conf)
    out_dir = exp_conf["output_dir"]
    ckpt.attach_from_conf(dopt, opt_conf, out_dir, prob_conf["problem_name"], ctx)
    if ctx.is_main:
        print("-------------------------------------------------------")
        print("-------------------------------------------------------")
        print(label or ("Running problem: " + prob_conf["problem_name"]))
    with maybe_profiler(bool(opt_conf.get("profile", False)) and ctx.is_main, out_dir, prob_conf["problem_name"]) as prof:
        dopt.train(profiler=prof)
    if exp_conf["writeout"]:
        prob.save_metrics(out_dir)
    return dopt
