"""Distributed PPO with DSGD over a 3-predator wheel graph (reference: RL/dist_rl/train_dsgd_multi.py)."""
import torch

from .consensus_ppo import DSGDPPO
from .train_common import common_conf, make_problem, parse_args


def main(argv=None):
    args = parse_args(argv, default_id=0)
    dppo, hyper = make_problem(args)
    confs = dict(common_conf(args), alpha0=hyper["lr"], mu=0.0)
    print("running dsgd")
    DSGDPPO(dppo, torch.device(args.device), confs).train()


if __name__ == "__main__":
    main()
