"""nn_distributed_training_b200 — a Blackwell (sm_100a) native decentralized
neural-network training engine.

Capabilities mirror ``javieryu/nn_distributed_training`` (DiNNO / DSGD / DSGT
consensus optimizers, the dist_mnist / dist_dense / dist_online_dense problems,
the graph_generation topologies, the YAML experiment runners and the PPO
multi-agent stack) but the architecture is B200-first:

* every graph node owns a row of a flat, 16-byte aligned fp32 parameter arena
  (``parallel.arena``); several *virtual* nodes may share one GPU and one GPU is
  one ``torch.distributed`` rank,
* the per-round neighbor exchange + mixing + optimizer update is a single fused
  CUDA kernel that pulls neighbor rows through a pointer table (local rows or
  NVLink peer mappings from symmetric memory) — ``ops/csrc/consensus.cu``,
* the local forward/backward of the three model families are hand-written
  kernels (``ops/csrc/mnist.cu`` CUDA-core fused conv net, ``ops/csrc/mlp_tc.cu``
  tcgen05/TMEM MLP),
* a whole communication round is replayed from a CUDA graph with all per-round
  scalars (rho_k, lr_k, alpha_k, graph id) read from device-side schedules.

A pure PyTorch implementation of every op (any dtype, any device) is kept as
the numerical oracle and as the CPU / gloo path.
"""

__version__ = "0.1.0"

from . import utils  # noqa: F401
