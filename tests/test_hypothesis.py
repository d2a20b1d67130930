"""Randomised (hypothesis) versions of the structural invariants the kernels rely on: the sampler really is a keyed
bijection with DataLoader geometry for ANY shard size, the flat layout is aligned for ANY parameter shapes, Metropolis
weights are symmetric doubly stochastic on ANY graph, and the three update rules keep their conservation laws on
random connected graphs."""
import networkx as nx
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from nn_distributed_training_b200.data.sampler import BatchSchedule, OnlineWindowSchedule, feistel_permute, mix_key
from nn_distributed_training_b200.ops import consensus_ref
from nn_distributed_training_b200.parallel.arena import ROW_ALIGN_ELEMS, SLOT_ALIGN_ELEMS, FlatLayout
from nn_distributed_training_b200.utils import graph_generation
from nn_distributed_training_b200.utils.graph_generation import Topology

FAST = dict(max_examples=int(__import__("os").environ.get("HYP_EXAMPLES", "30")), deadline=None, derandomize=True,
            suppress_health_check=[HealthCheck.too_slow])


@settings(**FAST)
@given(m=st.integers(1, 5000), key=st.integers(0, 2 ** 32 - 1))
def test_feistel_is_a_bijection_for_any_size_and_key(m, key):
    out = feistel_permute(torch.arange(m, dtype=torch.int64), m, key)
    assert out.min() >= 0 and out.max() < m and out.unique().numel() == m


@settings(**FAST)
@given(m=st.integers(1, 700), B=st.integers(1, 130), seed=st.integers(0, 1000), node=st.integers(0, 99))
def test_every_epoch_visits_every_sample_once(m, B, seed, node):
    sch = BatchSchedule(m, B)
    bpe = sch.batches_per_epoch
    assert bpe == -(-m // B)
    for epoch in (0, 1):
        idx = torch.cat([sch.indices(epoch * bpe + b, seed, node) for b in range(bpe)])
        assert idx.numel() == m and idx.unique().numel() == m                       # a permutation of the shard
        sizes = [sch.locate(epoch * bpe + b)[2] for b in range(bpe)]
        assert sizes[:-1] == [B] * (bpe - 1) and sizes[-1] == m - B * (bpe - 1)       # short last batch, like DataLoader
    assert sch.epochs_completed(bpe) == 0 and sch.epochs_completed(bpe + 1) == 1      # bumped on the StopIteration draw
    if m > 8:                                                                         # different epochs / nodes: different keys
        assert mix_key(seed, node, 0) != mix_key(seed, node, 1) != mix_key(seed, node + 1, 1)


@settings(**FAST)
@given(shapes=st.lists(st.lists(st.integers(1, 9), min_size=1, max_size=3), min_size=1, max_size=6))
def test_flat_layout_is_aligned_and_lossless_for_any_shapes(shapes):
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes])
    mod = M()
    lay = FlatLayout.from_module(mod)
    assert all(o % SLOT_ALIGN_ELEMS == 0 for o in lay.offsets()) and lay.n_pad % ROW_ALIGN_ELEMS == 0
    assert lay.n == sum(int(np.prod(s)) for s in shapes) <= lay.n_pad
    row = lay.flatten(mod, torch.zeros(lay.n_pad))
    for v, p in zip(lay.views(row), mod.parameters()):
        assert torch.equal(v, p.detach())
    assert lay.compact(row).numel() == lay.n


def _random_connected_graph(n, p, seed):
    g = nx.gnp_random_graph(n, p, seed=seed)
    comps = [sorted(c) for c in nx.connected_components(g)]
    for a, b in zip(comps, comps[1:]):          # stitch components together
        g.add_edge(a[0], b[0])
    return g


@settings(**FAST)
@given(n=st.integers(2, 12), p=st.floats(0.1, 0.9), seed=st.integers(0, 10 ** 6))
def test_metropolis_is_symmetric_doubly_stochastic_on_any_graph(n, p, seed):
    g = _random_connected_graph(n, p, seed)
    W = graph_generation.get_metropolis(g, dtype=torch.float64)
    assert torch.allclose(W, W.T) and (W >= 0).all()
    assert torch.allclose(W.sum(0), torch.ones(n, dtype=torch.float64)) and torch.allclose(W.sum(1), torch.ones(n, dtype=torch.float64))
    t = Topology(g)
    assert t.max_degree == max(d for _, d in g.degree()) and all(len(t.neighbors_noself[i]) == g.degree(i) for i in range(n))


@settings(**FAST)
@given(n=st.integers(2, 9), p=st.floats(0.2, 0.9), seed=st.integers(0, 10 ** 6), dim=st.integers(1, 12))
def test_consensus_ops_keep_their_conservation_laws(n, p, seed, dim):
    g = _random_connected_graph(n, p, seed)
    t = Topology(g)
    gen = torch.Generator().manual_seed(seed)
    theta = torch.randn(n, dim, generator=gen, dtype=torch.float64)
    # DiNNO: the dual ascent dual_i -= rho * sum_j (theta_j - theta_i) keeps sum_i dual_i = 0 on an undirected graph
    adj = torch.as_tensor(graph_generation.adjacency(g), dtype=torch.float64)
    adj.fill_diagonal_(0.0)
    deg = adj.sum(1)
    duals, delta = torch.zeros_like(theta), torch.zeros_like(theta)
    for _ in range(3):
        consensus_ref.dinno_exchange_(theta, theta, adj, deg, 0.7, duals, delta)
        theta = theta + 0.1 * torch.randn(theta.shape, generator=gen, dtype=torch.float64)
        assert duals.sum(0).abs().max() < 1e-9
    # Metropolis mixing (DSGD / DSGT) preserves the network mean; the DSGT tracker keeps sum y = sum g
    W = torch.as_tensor(t.W, dtype=torch.float64)
    assert torch.allclose(consensus_ref.dsgd_mix(theta, W).mean(0), theta.mean(0))
    g_old = torch.randn(n, dim, generator=gen, dtype=torch.float64)
    g_new = torch.randn(n, dim, generator=gen, dtype=torch.float64)
    y = g_old.clone()
    y = consensus_ref.dsgt_track(y, W, g_new, g_old)
    assert torch.allclose(y.sum(0), g_new.sum(0))


@settings(**FAST)
@given(T=st.integers(3, 40), S=st.integers(1, 7), Wn=st.integers(1, 12), draws=st.integers(1, 300))
def test_online_window_never_leaves_the_trajectory(T, S, Wn, draws):
    Wn = min(Wn, T)
    sch = OnlineWindowSchedule(T, S, Wn)
    idx = sch.indices(0, draws, seed=1, node=0)
    assert idx.numel() > 0 and idx.min() >= 0 and idx.max() < T * S
    assert 0 <= sch.scan_cursor_at(draws) <= T
