"""CPU tests of the native runtime (ops/csrc/runtime.cpp): the multi-threaded HostBatchLoader assembles, round after
round, exactly the rows the Python sampler (data/sampler.py) — and hence the device sampler — defines.  The extension
is built by ``__graft_entry__.build()``; importing it needs no GPU."""
import numpy as np
import pytest
import torch

from nn_distributed_training_b200.data.sampler import BatchSchedule
from nn_distributed_training_b200.ops import load_ext


@pytest.fixture(scope="module")
def ext():
    e = load_ext()
    if e is None:
        pytest.skip("extension not built (python -m nn_distributed_training_b200.ops.build)")
    return e


@pytest.mark.parametrize("threads", [1, 3])
def test_host_batch_loader_matches_python_sampler(ext, threads):
    rng = np.random.default_rng(0)
    sizes, B, P, seed, node0, row = [37, 50, 64], 16, 2, 11, 4, 24
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).tolist()
    total = int(sum(sizes))
    x = torch.from_numpy(rng.integers(0, 256, size=(total, row), dtype=np.uint8))
    y = torch.arange(total, dtype=torch.int64) * 3 + 1
    calls0 = [0, 5, 2]                                   # nodes resume at different draw counters
    L, nslots = len(sizes), 3
    sx = torch.zeros(nslots, P, L, B, row, dtype=torch.uint8)
    sy = torch.zeros(nslots, P, L, B, dtype=torch.int64)
    sb = torch.zeros(nslots, P, L, dtype=torch.int32)
    loader = ext.HostBatchLoader(x.data_ptr(), y.data_ptr(), row, offs, sizes, calls0, B, P, seed, node0,
                                 [sx[s].data_ptr() for s in range(nslots)], [sy[s].data_ptr() for s in range(nslots)],
                                 [sb[s].data_ptr() for s in range(nslots)], threads)
    try:
        for rnd in range(9):                              # 3 times around the ring, across epoch boundaries
            slot = loader.acquire()
            assert slot == rnd % nslots
            for p in range(P):
                for l in range(L):
                    sch = BatchSchedule(sizes[l], B)
                    idx = sch.indices(calls0[l] + rnd * P + p, seed, node0 + l) + offs[l]
                    n = idx.numel()
                    assert int(sb[slot, p, l]) == n
                    assert torch.equal(sx[slot, p, l, :n], x[idx]) and torch.equal(sy[slot, p, l, :n], y[idx])
            loader.release(slot)
        assert loader.rounds_assembled() >= 9
    finally:
        loader.stop()
