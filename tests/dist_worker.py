"""Multi-process worker (launched by torch.distributed.run from test_distributed.py).

Every rank hosts N/world graph nodes; after a few rounds the gathered parameters must match
a single-process run of the same problem (rank 0 recomputes it locally)."""
import argparse
import copy
import os
import sys

import networkx as nx
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nn_distributed_training_b200.data.mnist import synthetic_mnist  # noqa: E402
from nn_distributed_training_b200.models import MNISTConvNet  # noqa: E402
from nn_distributed_training_b200.optimizers import build_optimizer  # noqa: E402
from nn_distributed_training_b200.parallel.context import DistContext  # noqa: E402
from nn_distributed_training_b200.problems.dist_mnist_problem import DistMNISTProblem  # noqa: E402

CONFS = {
    "dinno": {"alg_name": "dinno", "rho_init": 0.5, "rho_scaling": 1.01, "outer_iterations": 6,
              "primal_iterations": 2, "primal_optimizer": "adam", "persistant_primal_opt": False,
              "primal_lr_start": 0.005, "primal_lr_finish": 0.0005, "lr_decay_type": "log", "profile": False},
    "dsgd": {"alg_name": "dsgd", "alpha0": 0.05, "mu": 0.01, "outer_iterations": 6, "profile": False},
    "dsgt": {"alg_name": "dsgt", "alpha": 0.02, "init_grads": True, "outer_iterations": 6, "profile": False},
}
METRICS = ["forward_pass_count", "validation_loss", "consensus_error", "top1_accuracy", "current_epoch"]


def build(ctx, N, graph, conf, backend, M=200, pipeline="resident", separate_publish=False):
    data = synthetic_mnist(M * N, seed=3)
    val = synthetic_mnist(128, seed=4)
    shards = [data.select(torch.arange(i * M, (i + 1) * M)) for i in range(N)]
    pconf = {"problem_name": "t", "train_batch_size": 32, "val_batch_size": 64, "metrics": METRICS,
             "metrics_config": {"evaluate_frequency": 3}, "optimizer_config": conf, "input_pipeline": pipeline,
             "separate_publish": separate_publish}
    torch.manual_seed(5)
    base = MNISTConvNet(3, 5, 64)
    return DistMNISTProblem(graph, base, torch.nn.NLLLoss(), shards, val, ctx.device, pconf, ctx=ctx,
                            backend=backend, seed=11)


def delayed(ctx, N, G, backend):
    """Buffer-reuse protocol under stress: the graph changes EVERY round (fault-injected link drops),
    every neighbor read is checked against its round tag (debug_sequence_check), rounds are launched one at a time and one
    rank is held back by a spin kernel before every other round, so its peers run ahead as far as the protocol lets them.
    The gathered parameters must equal the single-process run."""
    from nn_distributed_training_b200.ops import load_ext
    ext = load_ext(required=True)
    ok = True
    R = 14
    for alg in ("dinno", "dsgd", "dsgt"):
        conf = dict(copy.deepcopy(CONFS[alg]), outer_iterations=R, debug_sequence_check=True)
        extra = {"fault_injection": {"link_drop_prob": 0.45, "seed": 3, "from_round": 0, "to_round": R}}

        def make(c):
            data = synthetic_mnist(200 * N, seed=3)
            val = synthetic_mnist(128, seed=4)
            shards = [data.select(torch.arange(i * 200, (i + 1) * 200)) for i in range(N)]
            # same samples-per-CTA split in the distributed and the single-process run: identical fp32 partial sums, so the
            # comparison is exact and not blurred by Adam amplifying summation-order round-off
            pconf = {"problem_name": "t", "train_batch_size": 32, "val_batch_size": 64, "metrics": METRICS, "samples_per_cta": 8,
                     "metrics_config": {"evaluate_frequency": 10 ** 6}, "optimizer_config": conf, **extra}
            torch.manual_seed(5)
            return DistMNISTProblem(G, MNISTConvNet(3, 5, 64), torch.nn.NLLLoss(), shards, val, c.device, pconf, ctx=c,
                                    backend=backend, seed=11)
        pr = make(ctx)
        opt = build_optimizer(pr, ctx.device, copy.deepcopy(conf))
        slow = ctx.world_size - 1
        for r in range(R):
            if ctx.rank == slow and r % 2 == 1:
                ext.spin(600_000)            # ~0.3 ms: many round times
            if ctx.rank == 0 and r % 3 == 2:
                ext.spin(300_000)
            opt.run_rounds(1)
        torch.cuda.synchronize()
        opt._program.eng.check()             # raises on a stale tag (err == 2) or a spin timeout
        ngraphs = len(opt._program.eng.topos)
        theta = pr.gather_rows(pr.arena.theta).cpu()
        if ctx.is_main:
            solo = DistContext.single(ctx.device)
            pr1 = make(solo)
            opt1 = build_optimizer(pr1, solo.device, copy.deepcopy(conf))
            opt1.run_rounds(R)
            torch.cuda.synchronize()
            ref = pr1.arena.theta.cpu()
            rel = ((theta - ref).norm() / ref.norm()).item()
            good = rel < 1e-5
            print(f"[delayed] {alg} world={ctx.world_size} distinct_graphs={ngraphs} rel={rel:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good and ngraphs > 3
        ctx.barrier()
    if ctx.is_main:
        print("DIST_RESULT", "PASS" if ok else "FAIL", flush=True)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cuda", type=int, default=0)
    ap.add_argument("--nodes", type=int, default=6)
    ap.add_argument("--graph", default="cycle")
    ap.add_argument("--pipeline", default="resident")   # host: device-initiated staging + forked peer announcement
    ap.add_argument("--separate-publish", type=int, default=0)   # 1: peers announced by publish_round_kernel on a forked branch
    ap.add_argument("--delayed", type=int, default=0)   # 1: time-varying graphs (link drops) + a deliberately slow rank
    args = ap.parse_args()
    ctx = DistContext.from_env(use_cuda=bool(args.cuda))
    N = args.nodes
    G = {"cycle": nx.cycle_graph(N), "wheel": nx.wheel_graph(N), "complete": nx.complete_graph(N)}[args.graph]
    backend = "fused" if args.cuda else "torch"
    if args.delayed:
        return delayed(ctx, N, G, backend)
    ok = True
    for alg, conf in CONFS.items():
        pr = build(ctx, N, G, conf, backend, pipeline=args.pipeline, separate_publish=bool(args.separate_publish))
        opt = build_optimizer(pr, ctx.device, copy.deepcopy(conf))
        opt.train()
        theta = pr.gather_rows(pr.arena.theta).cpu()
        vl = pr.metrics["validation_loss"][-1]
        if ctx.is_main:
            solo = DistContext.single(ctx.device)
            pr1 = build(solo, N, G, conf, backend)
            opt1 = build_optimizer(pr1, solo.device, copy.deepcopy(conf))
            opt1.train()
            ref = pr1.arena.theta.cpu()
            bad = ((theta - ref).abs() > 2e-5 + 2e-3 * ref.abs()).float().mean().item()
            rel = ((theta - ref).norm() / ref.norm()).item()
            vd = (vl - pr1.metrics["validation_loss"][-1]).abs().max().item()
            good = bad < 5e-3 and rel < 1e-2 and vd < 1e-3
            eng = getattr(getattr(opt, "_program", None), "eng", None)
            how = "" if eng is None else (f" sum_mode={eng.sum_mode} mc={bool(eng.sum_buf and eng.sum_buf.multicast_ptr)}"
                                          f" separate_publish={eng.separate_publish}")
            print(f"[dist] {alg} world={ctx.world_size} graph={args.graph}{how} bad={bad:.2e} rel={rel:.2e} "
                  f"val_diff={vd:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
        ctx.barrier()
    if ctx.is_main:
        print("DIST_RESULT", "PASS" if ok else "FAIL", flush=True)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
