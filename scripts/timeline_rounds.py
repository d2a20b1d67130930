#!/usr/bin/env python
"""Per-round device timeline of the headline DiNNO round (debug aid, not a benchmark).

``NNDT_TIMELINE=1`` makes ``dinno_update_kernel`` stamp ``%globaltimer`` (first and last block of its grid) at
kernel entry (0), after the neighbor-flag wait (1), before (2) / after (3) the programmatic-dependency wait on the
forward/backward kernel, and at exit (4).  From those this script prints, averaged over the timed rounds, where a round's
time goes on every rank — used to locate multi-GPU coupling costs.  Run alone (1 GPU) or under torchrun.

    python scripts/timeline_rounds.py --dtype fp64 [--steps 40]
"""
import argparse
import json
import os
import sys

os.environ["NNDT_TIMELINE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="fp64")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--out", default="gpurun_out/timeline")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from nn_distributed_training_b200.optimizers import DiNNO
    from nn_distributed_training_b200.parallel.context import DistContext

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    ctx = DistContext.from_env(use_cuda=True)
    h = bench.Harness(ctx)
    n_nodes = bench.NODES_PER_GPU * ctx.world_size
    W, K = args.warmup, args.steps
    pipeline = os.environ.get("NNDT_BENCH_PIPELINE", "auto")
    pr = bench.build_problem(ctx, bench._cycle(n_nodes), bench.opt_conf(4096), 10 ** 9, extra={"input_pipeline": pipeline}, dtype=args.dtype)
    opt = DiNNO(pr, ctx.device, pr.conf["optimizer_config"])
    opt.run_rounds(W)
    ms, _ = h.time_rounds(opt, K)
    tl = opt._program.eng.timeline.cpu().numpy().astype(np.int64)      # [4096][16]
    pits = 2
    k0 = W + 2                                   # skip the first rounds after the gate
    rows = []
    for k in range(k0, W + K):
        r = {}
        for s in range(pits):
            a = tl[(k * 4 + s) & 4095]
            f, l = a[0:8], a[8:16]
            r[f"s{s}"] = dict(first=f[:5].tolist(), last=l[:5].tolist())
        rows.append(r)

    def us(x):
        return float(np.mean(x)) / 1e3

    out = {"rank": ctx.rank, "world": ctx.world_size, "dtype": args.dtype, "ms_per_round": ms / K, "kernel": pr.fused.kernel_name}
    # per round: everything relative to the entry of the first block of step 0's update kernel
    t00 = np.array([r["s0"]["first"][0] for r in rows])
    period = np.diff(t00)
    out["period_us"] = us(period)
    segs = {}
    for s in range(pits):
        F = np.array([r[f"s{s}"]["first"] for r in rows]); L = np.array([r[f"s{s}"]["last"] for r in rows])
        segs[f"U{s} first-block entry (rel. round start)"] = us(F[:, 0] - t00)
        segs[f"U{s} first-block flag wait"] = us(F[:, 1] - F[:, 0])
        segs[f"U{s} first-block prologue loads (incl. neighbor pulls)"] = us(F[:, 2] - F[:, 1])
        segs[f"U{s} first-block PDL wait (= forward/backward still running)"] = us(F[:, 3] - F[:, 2])
        segs[f"U{s} first-block tail"] = us(F[:, 4] - F[:, 3])
        segs[f"U{s} last-block entry (rel. round start)"] = us(L[:, 0] - t00)
        segs[f"U{s} last-block flag wait"] = us(L[:, 1] - L[:, 0])
        segs[f"U{s} last-block prologue loads"] = us(L[:, 2] - L[:, 1])
        segs[f"U{s} last-block PDL wait"] = us(L[:, 3] - L[:, 2])
        segs[f"U{s} last-block exit (rel. round start)"] = us(L[:, 4] - t00)
    out["segments_us"] = {k: round(v, 2) for k, v in segs.items()}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(f"{args.out}_{args.dtype}_w{ctx.world_size}_r{ctx.rank}.json", "w") as f:
        json.dump(out, f, indent=1)
    if ctx.rank == 0:
        print(json.dumps(out, indent=1))
    opt._program.eng.check()
    ctx.barrier()


if __name__ == "__main__":
    main()
