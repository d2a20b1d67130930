"""Multi-agent RL workload: PPO predators in a batched simple_tag world, trained with the
generic consensus optimizers (reference: RL/)."""
from .simple_tag import SimpleTagEnv, heuristic_prey_action
from .model import FFReLUNet, ActorCritic
from .dist_ppo import DistPPOProblem
from .consensus_ppo import DiNNOPPO, DSGDPPO, DSGTPPO, agreement
from .ppo import PPO
