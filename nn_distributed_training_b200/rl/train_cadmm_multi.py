"""Distributed PPO with DiNNO over a 3-predator wheel graph
(reference: RL/dist_rl/train_cadmm_multi.py — same hyper-parameters, batched environment)."""
import torch

from .consensus_ppo import DiNNOPPO
from .train_common import common_conf, make_problem, parse_args


def main(argv=None):
    args = parse_args(argv, default_id=50)
    dppo, hyper = make_problem(args)
    confs = dict(common_conf(args), rho_init=1.0, rho_scaling=1.0, primal_lr_start=hyper["lr"], primal_lr_finish=0.001,
                 lr_decay_type="constant", persistant_primal_opt=False, primal_iterations=hyper["n_updates_per_iteration"],
                 outer_iterations=10_000_000)
    print("running cadmm")
    DiNNOPPO(dppo, torch.device(args.device), confs).train()


if __name__ == "__main__":
    main()
