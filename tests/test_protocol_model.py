"""Model check of the cross-GPU publication protocol (ops/csrc/consensus_device.cuh: begin_round / wait_neighbors /
finish_round) — CPU only, no kernels: every rank is a small state machine and ALL interleavings of a few rounds on a
time-varying graph are explored.

Protocol of rank r in round k (published rows are double buffered by round parity):
  announce   flag[r] = k                      (its rows of round k were written at the end of round k-1)
  wait       until flag[j] >= k for every j in N_k(r)  [+ N_{k-1}(r): the write-after-read fix]
  read       pub[j][k & 1] for every j in N_k(r), one neighbor at a time
  write      pub[r][(k+1) & 1] = rows of round k+1
Safety: every read returns the rows of round k (never rows of k+2 written early, never stale ones); liveness: no deadlock.
The reference has no counterpart (one process, optimizers/dinno.py:103-110 snapshots all replicas at once); the
hazard only exists because ranks run free of each other between flag waits."""
import itertools
import random


def explore(graphs, n_ranks, wait_prev, announce_at_start=True, max_states=400_000):
    """DFS over all interleavings.  graphs[k][r] = set of neighbor ranks of r in round k.  Returns (violation, deadlock, n_states)."""
    K = len(graphs)
    # per-rank program: list of atomic steps
    def program(r):
        steps = []
        for k in range(K):
            if announce_at_start:
                steps.append(("announce", k))
            need = set(graphs[k][r]) | (set(graphs[k - 1][r]) if (wait_prev and k > 0) else set())
            steps.append(("wait", k, tuple(sorted(need))))
            for j in sorted(graphs[k][r]):
                steps.append(("read", k, j))
            steps.append(("write", k))
            if not announce_at_start:
                steps.append(("announce", k + 1))
        return steps

    progs = [program(r) for r in range(n_ranks)]
    init = (tuple(0 for _ in range(n_ranks)),                      # pc per rank
            tuple(0 for _ in range(n_ranks)),                      # flag per rank
            tuple((0, -1) for _ in range(n_ranks)))                # pub[r] = (tag of parity 0, tag of parity 1)
    seen = {init}
    stack = [init]
    violation = deadlock = None
    while stack and len(seen) < max_states:
        pcs, flags, pubs = stack.pop()
        progressed = False
        done = True
        for r in range(n_ranks):
            if pcs[r] >= len(progs[r]):
                continue
            done = False
            st = progs[r][pcs[r]]
            nflags, npubs = flags, pubs
            if st[0] == "announce":
                nflags = flags[:r] + (max(flags[r], st[1]),) + flags[r + 1:]
            elif st[0] == "wait":
                if any(flags[j] < st[1] for j in st[2]):
                    continue                                        # blocked
            elif st[0] == "read":
                k, j = st[1], st[2]
                if pubs[j][k & 1] != k:
                    violation = (r, k, j, pubs[j][k & 1])
                    return violation, None, len(seen)
            elif st[0] == "write":
                k = st[1]
                p = list(pubs[r]); p[(k + 1) & 1] = k + 1
                npubs = pubs[:r] + (tuple(p),) + pubs[r + 1:]
            progressed = True
            nxt = (pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], nflags, npubs)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not done and not progressed:
            deadlock = (pcs, flags)
            return None, deadlock, len(seen)
    return violation, deadlock, len(seen)


def _sym(n, edges):
    g = [set() for _ in range(n)]
    for a, b in edges:
        g[a].add(b); g[b].add(a)
    return g


# rank 1 is a neighbor of rank 0 in round 0 only: the case VERDICT weak #4 describes
DYNAMIC = [_sym(3, [(0, 1), (0, 2)]), _sym(3, [(0, 2)]), _sym(3, [(0, 2), (1, 2)]), _sym(3, [(0, 1)])]


def test_waiting_only_on_current_neighbors_is_unsafe_on_time_varying_graphs():
    v, d, n = explore(DYNAMIC, 3, wait_prev=False)
    assert d is None
    assert v is not None, "the model must reproduce the write-after-read hazard of the round-1 protocol"
    r, k, j, tag = v
    assert tag == k + 2          # the reader found rows written two rounds ahead in the buffer it was still reading


def test_union_wait_closes_the_window_for_both_announcement_points():
    for at_start in (True, False):
        v, d, n = explore(DYNAMIC, 3, wait_prev=True, announce_at_start=at_start)
        assert v is None and d is None, (at_start, v, d)
        assert 300 < n < 400_000     # the search branched and was exhaustive (not cut off by max_states)


def test_static_graphs_need_no_extra_wait():
    ring = [_sym(4, [(0, 1), (1, 2), (2, 3), (3, 0)])] * 3
    for wait_prev in (False, True):
        v, d, n = explore(ring, 4, wait_prev=wait_prev, max_states=300_000)
        assert v is None and d is None and n < 300_000


def test_random_time_varying_graphs_random_schedules():
    """Larger instances than the exhaustive search can cover: random graphs per round, the union wait, both announcement
    points; isolated ranks (no neighbors in a round) included."""
    rng = random.Random(3)
    for trial in range(40):
        n = rng.choice([3, 4, 5])
        K = 3
        pairs = list(itertools.combinations(range(n), 2))
        graphs = [_sym(n, [e for e in pairs if rng.random() < 0.5]) for _ in range(K)]
        v, d, _ = explore(graphs, n, wait_prev=True, announce_at_start=bool(trial & 1), max_states=60_000)
        assert v is None and d is None, (trial, graphs, v, d)
