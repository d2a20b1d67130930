"""Distributed PPO with DSGD over a 3-predator wheel graph
(reference: RL/dist_rl/train_dsgd_multi.py — same hyper-parameters, batched environment)."""
import argparse

import networkx as nx
import torch

from .consensus_ppo import DSGDPPO
from .dist_ppo import DistPPOProblem
from .model import FFReLUNet
from .simple_tag import SimpleTagEnv


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--max_rl_timesteps", type=int, default=10_000_000)
    ap.add_argument("--num_envs", type=int, default=16)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--ID", type=int, default=50)
    args = ap.parse_args(argv)
    steps = 200
    env = SimpleTagEnv(num_envs=args.num_envs, num_good=1, num_adversaries=3, num_obstacles=8, max_cycles=steps,
                       device=args.device)
    hyper = {"timesteps_per_batch": 2000, "max_timesteps_per_episode": steps, "gamma": 0.99,
             "n_updates_per_iteration": 5, "lr": 3e-4, "clip": 0.2, "render": False, "render_every_i": 1, "save_freq": 10}
    obs_dim = env.observation_spaces["adversary_0"].shape[0]
    act_dim = env.action_spaces["adversary_0"].shape[0]
    base_actor = FFReLUNet([obs_dim, 64, 64, 64, act_dim])
    base_critic = FFReLUNet([obs_dim, 64, 64, 64, 1])
    graph = nx.wheel_graph(3)
    dppo = DistPPOProblem(base_actor, base_critic, graph, env, **hyper)
    confs = {"alpha0": hyper["lr"], "mu": 0.0, "max_rl_timesteps": args.max_rl_timesteps, "ID": args.ID}
    print("running dsgd")
    DSGDPPO(dppo, torch.device(args.device), confs).train()


if __name__ == "__main__":
    main()
