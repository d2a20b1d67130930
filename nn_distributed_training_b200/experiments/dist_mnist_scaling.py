"""MNIST scaling study: sweep the number of nodes at constant Fiedler value, or the Fiedler
value at constant size (reference: experiments/dist_mnist_scaling.py)."""
from __future__ import annotations

import copy
import os
import random
import sys

import torch

from . import common
from ..data.mnist import load_mnist
from ..models import MNISTConvNet
from ..problems import DistMNISTProblem
from ..utils import graph_generation
from ..utils.config import load_experiment


def make_graphs(scale_conf, rng=None):
    if scale_conf["const"] == "fiedler":
        Ns = torch.linspace(scale_conf["min_N"], scale_conf["max_N"], scale_conf["num_trials"]).int().tolist()
        return [graph_generation.disk_with_fied(N, scale_conf["target_fied"], rng=rng) for N in Ns]
    if scale_conf["const"] == "num_nodes":
        fieds = torch.linspace(scale_conf["min_fied"], scale_conf["max_fied"], scale_conf["num_trials"])
        return [graph_generation.disk_with_fied(scale_conf["num_nodes"], float(f), rng=rng) for f in fieds]
    raise NameError("Unknown const factor in scaling")


def label_sorted_shards(train, N):
    """Label-sorted indices chunked into N shards (reference :122-129)."""
    order = torch.argsort(train.y, stable=True)
    return [train.select(idx) for idx in order.chunk(N)]


def experiment(yaml_pth):
    conf_dict = load_experiment(yaml_pth, "mnist_scaling")
    exp_conf = conf_dict["experiment"]
    ctx = common.make_context(exp_conf)
    output_dir = common.setup_output(exp_conf, yaml_pth, ctx)
    train, src = load_mnist(exp_conf["data_dir"], train=True, source=exp_conf.get("data_source", "auto"))
    val, _ = load_mnist(exp_conf["data_dir"], train=False, source=exp_conf.get("data_source", "auto"))
    model_conf = exp_conf["model"]
    torch.manual_seed(int(exp_conf.get("seed", 0)))
    base_model = MNISTConvNet(model_conf["num_filters"], model_conf["kernel_size"], model_conf["linear_width"])
    base_loss = common.make_loss(exp_conf["loss"])

    rng = random.Random(int(exp_conf["seed"])) if "seed" in exp_conf else None
    graphs = ctx.broadcast_object(make_graphs(exp_conf["scaling"], rng))
    if ctx.is_main:
        print("Graph generation successful!")

    prob_conf = conf_dict["problem"]
    for trial, graph in enumerate(graphs):
        N = len(graph.nodes)
        pc = copy.deepcopy(prob_conf)
        pc["problem_name"] = str(trial)
        if exp_conf["writeout"] and ctx.is_main:
            common.write_gpickle(graph, os.path.join(output_dir, str(trial) + ".gpickle"))
        shards = label_sorted_shards(train, N)
        prob = DistMNISTProblem(graph, base_model, base_loss, shards, val, ctx.device, pc, ctx=ctx,
                                seed=int(exp_conf.get("seed", 0)))
        label = "Running problem:  {}  /  {}\nNum Nodes:  {}\nDS size:  {}".format(trial, len(graphs), N, len(shards[-1]))
        common.run_problem(prob, pc, exp_conf, ctx, label=label)
    return conf_dict


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if not os.path.exists(argv[1]):
        raise NameError("YAML configuration file does not exist, exiting!")
    experiment(argv[1])


if __name__ == "__main__":
    main()
