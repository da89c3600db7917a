"""The list formulation K3 uses for `search_layer`, checked against the heap formulation of the reference on CPU.

`csrc/hnsw_search.cu` keeps `best` as ONE sorted list (an unexpanded entry == a frontier entry), merges the fresh
neighbors of an expansion by rank, and replays an expansion sequentially only when a group of bit-identical
distances straddles the cut at ef, with the evicted-but-still-tied entries parked in a `limbo` array
(src/index/lockfree_hnsw.rs:165-189: strict `d < worst` admission, strict `dist > worst` break, eviction by
(distance, id)). This file restates exactly that algorithm in Python (small graphs only) and requires it to return
the oracle's ids, distance bits AND work counters on corpora full of exact duplicates, where ties at the boundary
are routine. The CUDA kernel itself is checked against the oracle in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

import oracle


def total_keys(d):
    b = np.ascontiguousarray(d, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(b & np.uint64(0x80000000), ~b & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))


class ListModel:
    """search_layer as the kernel runs it: sorted list + rank merge + tie replay + limbo."""

    def __init__(self, neighbors, top_layer, entry, chunk=32, exact_ties=True):
        self.neighbors, self.top_layer, self.entry, self.chunk = neighbors, top_layer, entry, chunk
        self.exact_ties = exact_ties            # False: the round-1 kernel (rank merge + truncate, nothing else)
        self.evals = self.expansions = self.entries = self.replays = 0

    def search_layer(self, dk, ep, ef, layer):
        visited = {ep}
        L = [[int(dk[ep]), ep, False]]          # ascending (distance key, id); [.., expanded]
        limbo = []
        self.evals += 1
        while True:
            pos = next((i for i, e in enumerate(L) if not e[2]), None)
            if pos is not None:
                L[pos][2] = True
                cur = L[pos][1]
            elif limbo:
                limbo.sort()
                cur = limbo.pop(0)[1]
            else:
                break
            self.expansions += 1
            nl = self.neighbors(cur, layer)
            self.entries += len(nl)
            for c0 in range(0, len(nl), self.chunk):
                fresh = []
                for nb in nl[c0:c0 + self.chunk]:
                    if nb not in visited:
                        visited.add(nb)
                        fresh.append(int(nb))
                if not fresh:
                    continue
                self.evals += len(fresh)
                cand = [[int(dk[nb]), nb, False] for nb in fresh]   # neighbor-list order
                union = sorted(L + cand, key=lambda e: (e[0], e[1]))
                if self.exact_ties and len(union) > ef and union[ef - 1][0] == union[ef][0]:
                    self.replays += 1                               # tie across the cut: sequential replay
                    for c in cand:
                        if len(L) >= ef:
                            worst = L[-1]
                            if not c[0] < worst[0]:
                                continue
                            if not worst[2]:
                                limbo.append((worst[0], worst[1]))
                            L.pop()
                        L.append(c)
                        L.sort(key=lambda e: (e[0], e[1]))
                    limbo = [e for e in limbo if e[0] == L[-1][0]]
                else:
                    L = union[:ef]
                    if limbo and L[-1][0] != limbo[0][0]:
                        limbo = []
        return L

    def search(self, dk, ef):
        ep, top = self.entry
        for layer in range(top, 0, -1):
            ep = self.search_layer(dk, ep, 1, layer)[0][1]
        return self.search_layer(dk, ep, ef, 0)


def duplicate_corpus(n, dim, dup_frac, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, size=(n, dim)).astype(np.float32)
    n_dup = int(n * dup_frac)
    dst = rng.choice(np.arange(n // 4, n), size=n_dup, replace=False)
    src = rng.integers(0, n // 4, size=n_dup)
    x[dst] = x[src]                                # exact copies of earlier rows (same bits after preprocess)
    return x, rng.uniform(-1, 1, size=(24, dim)).astype(np.float32)


@pytest.mark.parametrize("n,dim,m,dup,ef", [(900, 6, 4, 0.2, 10), (900, 6, 4, 0.2, 24), (1200, 8, 8, 0.05, 10),
                                            (1200, 8, 8, 0.35, 64), (700, 4, 16, 0.2, 64)])
def test_list_model_equals_reference_heaps_on_duplicates(n, dim, m, dup, ef):
    x, q = duplicate_corpus(n, dim, dup, seed=n + ef)
    orc = oracle.OracleHnsw(m, 40, 50, 0xD0B1E)
    orc.insert_matrix(x)
    assert orc.check_invariants() == 0
    lists = {}

    def neighbors(v, layer):
        if (v, layer) not in lists:
            lists[(v, layer)] = [int(t) for t in orc.neighbors(v, layer)]
        return lists[(v, layer)]

    prepared = oracle.preprocess_rows(x)
    ids_all, d_all = oracle.brute_force(prepared, q, n, mode="prepared")
    total_replays = 0
    for qi in range(q.shape[0]):
        dk = np.empty(n, np.uint64)
        dk[ids_all[qi]] = total_keys(d_all[qi])
        dist = np.empty(n, np.float32)
        dist[ids_all[qi]] = d_all[qi]
        model = ListModel(neighbors, None, orc.entry())
        got = model.search(dk, ef)
        ctr = np.zeros(3, np.uint64)
        want_ids, want_d = orc.search(q[qi], ef, counters=ctr)
        assert [e[1] for e in got] == want_ids.tolist(), "query %d: list model ids differ from the heap reference" % qi
        assert np.array_equal(dist[[e[1] for e in got]].view(np.uint32), want_d.view(np.uint32))
        assert (model.evals, model.expansions, model.entries) == tuple(int(c) for c in ctr), "query %d: work counters differ" % qi
        total_replays += model.replays
    assert total_replays > 0, "the workload never put a tie on the cut: it does not test what it is meant to"


def test_plain_rank_merge_is_not_enough_on_duplicates():
    """The round-1 kernel merged by (distance, id) and truncated; on duplicates that is not the reference's traversal
    (ids or work counters differ) — this pins WHY the tie replay exists, and that the workload above would catch its
    absence."""
    differs = 0
    for ef in (10, 24):
        x, q = duplicate_corpus(900, 6, 0.2, seed=900 + ef)
        orc = oracle.OracleHnsw(4, 40, 50, 0xD0B1E)
        orc.insert_matrix(x)
        ids_all, d_all = oracle.brute_force(oracle.preprocess_rows(x), q, 900, mode="prepared")
        for qi in range(q.shape[0]):
            dk = np.empty(900, np.uint64)
            dk[ids_all[qi]] = total_keys(d_all[qi])
            model = ListModel(lambda v, l: [int(t) for t in orc.neighbors(v, l)], None, orc.entry(), exact_ties=False)
            got = model.search(dk, ef)
            ctr = np.zeros(3, np.uint64)
            want_ids, _ = orc.search(q[qi], ef, counters=ctr)
            differs += ([e[1] for e in got] != want_ids.tolist()
                        or (model.evals, model.expansions, model.entries) != tuple(int(c) for c in ctr))
    assert differs > 0
