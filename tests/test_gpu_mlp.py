"""GPU numerics of the tcgen05 MLP kernels vs fp32 PyTorch."""
import copy

import networkx as nx
import pytest
import torch

from nn_distributed_training_b200.models import FourierNet
from nn_distributed_training_b200.parallel.arena import FlatLayout, NodeArena

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _arena(shape, scale, L, seed=0):
    torch.manual_seed(seed)
    models = []
    base = FourierNet(shape, scale=scale)
    arena = NodeArena(FlatLayout.from_module(base), L, DEV, torch.float32)
    for l in range(L):
        m = FourierNet(shape, scale=scale).to(DEV)
        arena.attach(l, m)
        models.append(m)
    return arena, models, base.spec


@pytest.mark.parametrize("h1", [256, 64, 128])
@pytest.mark.parametrize("M", [1000, 128, 77])
def test_mlp_forward_matches_torch(h1, M):
    from nn_distributed_training_b200.ops.mlp_fused import MlpForward
    arena, models, spec = _arena([2, h1, 64, 64, 64, 1], 0.05, 3)
    x = (torch.rand(M, 2, device=DEV) - 0.5) * 1500
    out = MlpForward(arena, spec, 3, torch.device(DEV))(x)
    torch.cuda.synchronize()
    for l, m in enumerate(models):
        with torch.no_grad():
            ref = m(x).reshape(-1)
        # bf16 operands, fp32 accumulation
        assert (out[l] - ref).abs().max().item() < 2e-2, (out[l] - ref).abs().max().item()
        assert (out[l] - ref).abs().mean().item() < 3e-3
