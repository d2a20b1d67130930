"""Oracle tests: our optimizers vs the unmodified reference, fp64, CPU.

Each node's shard holds exactly one batch, so both frameworks see the same
samples every step regardless of their (different) shuffling machinery.
"""
import copy

import networkx as nx
import pytest
import torch

from nn_distributed_training_b200.models import MNISTConvNet
from nn_distributed_training_b200.optimizers import DiNNO, DSGD, DSGT
from nn_distributed_training_b200.problems.dist_mnist_problem import DistMNISTProblem

N, B, ROUNDS = 5, 12, 6


@pytest.fixture(autouse=True)
def _fp64_default():
    """The reference only works with float64 as the global default dtype
    (experiments/dist_mnist_ex.py:19); ours takes dtype from the model."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _data(seed=0):
    g = torch.Generator().manual_seed(seed)
    train = [torch.utils.data.TensorDataset(torch.randn(B, 1, 28, 28, generator=g, dtype=torch.float64),
                                            torch.randint(0, 10, (B,), generator=g)) for _ in range(N)]
    val = torch.utils.data.TensorDataset(torch.randn(40, 1, 28, 28, generator=g, dtype=torch.float64),
                                         torch.randint(0, 10, (40,), generator=g))
    return train, val


def _prob_conf(opt_conf):
    return {"problem_name": "t", "train_batch_size": B, "val_batch_size": 16,
            "metrics": ["forward_pass_count", "validation_loss", "consensus_error", "top1_accuracy", "current_epoch"],
            "metrics_config": {"evaluate_frequency": 3}, "optimizer_config": opt_conf}


def _graph(kind):
    if kind == "cycle":
        return nx.cycle_graph(N)
    if kind == "wheel":
        return nx.wheel_graph(N)
    return nx.erdos_renyi_graph(N, 0.6, seed=3)


def _run_pair(reference, ref_cls, our_cls, opt_conf, graph):
    torch.manual_seed(0)
    base = MNISTConvNet(3, 5, 64, dtype=torch.float64)
    ref_base = reference.mnist_model.MNISTConvNet(3, 5, 64).double()
    ref_base.load_state_dict(base.state_dict())
    train, val = _data()
    conf = _prob_conf(opt_conf)

    rp = reference.mnist_problem.DistMNISTProblem(graph, ref_base, torch.nn.NLLLoss(), train, val,
                                                  torch.device("cpu"), copy.deepcopy(conf))
    ro = ref_cls(rp, torch.device("cpu"), copy.deepcopy(opt_conf))
    ro.train()

    ours_conf = copy.deepcopy(conf)
    op = DistMNISTProblem(graph, base, torch.nn.NLLLoss(), train, val, "cpu", ours_conf, backend="torch")
    oo = our_cls(op, "cpu", copy.deepcopy(opt_conf))
    oo.train()

    for i in range(N):
        ref_vec = torch.nn.utils.parameters_to_vector(rp.models[i].parameters()).detach()
        our_vec = torch.nn.utils.parameters_to_vector(op.models[i].parameters()).detach()
        _assert_close(our_vec, ref_vec, strict=opt_conf.get("primal_optimizer", "sgd") == "sgd")
    return rp, op, ro, oo


def _assert_close(ours, ref, strict, rtol=1e-6, atol=1e-8):
    """Adam divides by sqrt(v)+eps: a parameter whose loss gradient is exactly 0
    on a node (dead ReLU unit, 12 samples) sees only ~1e-17 rounding noise, which
    Adam amplifies by lr/eps per step.  Those few coordinates are chaotic in the
    reference itself, so Adam runs must match on >= 99.9% of coordinates."""
    if strict:
        torch.testing.assert_close(ours, ref, rtol=rtol, atol=atol)
        return
    bad = (ours - ref).abs() > atol + rtol * ref.abs()
    assert bad.double().mean().item() < 1e-3, f"{int(bad.sum())} coordinates differ"


@pytest.mark.parametrize("kind", ["cycle", "wheel", "random"])
@pytest.mark.parametrize("opt", ["adam", "sgd", "adamw"])
def test_dinno_matches_reference(reference, kind, opt):
    conf = {"alg_name": "dinno", "rho_init": 0.5, "rho_scaling": 1.01, "outer_iterations": ROUNDS,
            "primal_iterations": 2, "primal_optimizer": opt, "persistant_primal_opt": False,
            "primal_lr_start": 0.005, "primal_lr_finish": 0.0005, "lr_decay_type": "log", "profile": False}
    rp, op, ro, oo = _run_pair(reference, reference.dinno.DiNNO, DiNNO, conf, _graph(kind))
    for i in range(N):
        _assert_close(op.arena.compact(oo.duals[i]), ro.duals[i].double(), strict=opt == "sgd")
    # metric parity: same evaluation cadence and values
    assert len(op.metrics["validation_loss"]) == len(rp.metrics["validation_loss"])
    for a, b in zip(op.metrics["validation_loss"], rp.metrics["validation_loss"]):
        torch.testing.assert_close(a.double(), b.double(), rtol=1e-5, atol=1e-7)
    for a, b in zip(op.metrics["top1_accuracy"], rp.metrics["top1_accuracy"]):
        torch.testing.assert_close(a.double(), b.double())
    for (a0, a1), (b0, b1) in zip(op.metrics["consensus_error"], rp.metrics["consensus_error"]):
        torch.testing.assert_close(a0.double(), b0.double(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(a1.double(), b1.double(), rtol=1e-4, atol=1e-5)
    assert op.metrics["forward_pass_count"] == rp.metrics["forward_pass_count"]
    for a, b in zip(op.metrics["current_epoch"], rp.metrics["current_epoch"]):
        assert torch.equal(a.double(), b.double())


def test_dinno_persistent_matches_reference(reference):
    conf = {"alg_name": "dinno", "rho_init": 0.3, "rho_scaling": 1.0, "outer_iterations": ROUNDS,
            "primal_iterations": 3, "primal_optimizer": "adam", "persistant_primal_opt": True,
            "primal_lr_start": 0.004, "primal_lr_finish": 0.001, "lr_decay_type": "linear", "profile": False}
    _run_pair(reference, reference.dinno.DiNNO, DiNNO, conf, _graph("cycle"))


@pytest.mark.parametrize("kind", ["cycle", "wheel", "random"])
def test_dsgd_reference_order(reference, kind):
    conf = {"alg_name": "dsgd", "alpha0": 0.05, "mu": 0.01, "outer_iterations": ROUNDS, "profile": False,
            "mixing_order": "reference"}
    _run_pair(reference, reference.dsgd.DSGD, DSGD, conf, _graph(kind))


@pytest.mark.parametrize("kind", ["cycle", "wheel", "random"])
@pytest.mark.parametrize("init_grads", [True, False])
def test_dsgt_reference_order(reference, kind, init_grads):
    conf = {"alg_name": "dsgt", "alpha": 0.02, "init_grads": init_grads, "outer_iterations": ROUNDS,
            "profile": False, "mixing_order": "reference"}
    rp, op, ro, oo = _run_pair(reference, reference.dsgt.DSGT, DSGT, conf, _graph(kind))
    for i in range(N):
        ref_y = torch.cat([t.reshape(-1) for t in ro.ylists[i]])
        torch.testing.assert_close(op.arena.compact(oo.y[i]), ref_y.double(), rtol=1e-6, atol=1e-8)
