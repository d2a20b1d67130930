"""Short eager (no CUDA graph) run of the bench configuration for ncu.
usage: python scripts/profile_round.py [rounds]"""
import os, sys
os.environ.setdefault("NNDT_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nn_distributed_training_b200.optimizers import DiNNO
from nn_distributed_training_b200.parallel.context import DistContext

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = DistContext.single(torch.device("cuda", 0))
pr = bench.build_problem(ctx, bench._cycle(10), bench.opt_conf(2000), 10 ** 9, samples_per_node=int(os.environ.get("SPN", 6000)),
                        dtype=os.environ.get("DTYPE", "fp32"))
opt = DiNNO(pr, ctx.device, pr.conf["optimizer_config"])
opt.run_rounds(rounds)
torch.cuda.synchronize()
if os.environ.get("EVAL", "0") == "1":
    pr.evaluate_metrics()
print("done", opt.k)
