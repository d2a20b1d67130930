"""Distributed PPO with DSGT over a 3-predator wheel graph (reference: RL/dist_rl/train_dsgt_multi.py)."""
import torch

from .consensus_ppo import DSGTPPO
from .train_common import common_conf, make_problem, parse_args


def main(argv=None):
    args = parse_args(argv, default_id=0)
    dppo, hyper = make_problem(args)
    confs = dict(common_conf(args), alpha_actor=hyper["lr"], alpha_critic=hyper["lr"], init_grads=False)
    print("running dsgt")
    DSGTPPO(dppo, torch.device(args.device), confs).train()


if __name__ == "__main__":
    main()
