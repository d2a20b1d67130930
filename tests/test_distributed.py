"""One rank per GPU (or per CPU process): the distributed run must reproduce the
single-process run — gloo/world_size 2 here on CPU, NCCL + peer-mapped kernels on GPUs."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _launch(nproc, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + extra
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)


@pytest.mark.parametrize("graph", ["cycle", "wheel"])
def test_gloo_two_ranks_match_single_process(graph):
    r = _launch(2, ["--cuda", "0", "--nodes", "4", "--graph", graph], 29611)
    assert "DIST_RESULT PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("graph,pipeline", [("cycle", "auto"), ("complete", "resident"), ("cycle", "host")])
def test_nccl_peer_mapped_ranks_match_single_process(graph, pipeline):
    """``host``: rows pulled from pinned host memory by the staging kernel inside multi-round graphs, peers
    announced by publish_round_kernel on a forked branch; the single-process oracle uses resident shards."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2
    r = _launch(nproc, ["--cuda", "1", "--nodes", "6", "--graph", graph, "--pipeline", pipeline], 29612)
    assert "DIST_RESULT PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    if pipeline in ("host", "auto"):      # auto = staged-resident multi-round graphs for DiNNO / DSGD
        assert "separate_publish=True" in r.stdout
