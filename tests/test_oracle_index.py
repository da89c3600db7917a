"""Pins the oracle's HNSW restatement to the reference's index tests:
`src/index/lockfree_hnsw.rs:502-573`, `tests/property_test.rs:29-47,146-249`,
`tests/recall_test.rs:140-214`, `src/index/sharded_rwlock.rs:101-146`."""
import math
import os
import sys

import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets


def index(seed=42, **kw):
    return oracle.OracleHnsw(seed=seed, **kw)


def test_empty_index_returns_nothing():                    # lockfree_hnsw.rs:511-516
    idx = index()
    ids, d = idx.search([1.0, 0.0], 10)
    assert len(ids) == 0 and len(idx) == 0


def test_single_vector_is_found():                         # :518-526
    idx = index()
    h = idx.insert([1.0, 0.0])
    ids, d = idx.search([1.0, 0.0], 10)
    assert list(ids) == [h] and d[0] < 1e-6


def test_nearest_is_ranked_first():                        # :528-538
    idx = index()
    near = idx.insert([1.0, 0.05])
    idx.insert([0.0, 1.0])
    idx.insert([0.5, 0.5])
    ids, d = idx.search([1.0, 0.0], 3)
    assert ids[0] == near and np.all(np.diff(d) >= 0)


def test_tombstoned_vectors_disappear_from_results():      # :540-553
    idx = index()
    a = idx.insert([1.0, 0.0])
    b = idx.insert([0.9, 0.1])
    assert idx.remove(a) and not idx.remove(a)
    assert len(idx) == 1
    ids, _ = idx.search([1.0, 0.0], 10)
    assert list(ids) == [b]


def test_seeded_single_threaded_inserts_are_deterministic():   # :555-573
    def build():
        idx = index()
        for i in range(50):
            ang = np.float32(i) * np.float32(0.13)
            idx.insert([np.cos(ang, dtype=np.float32), np.sin(ang, dtype=np.float32)])
        return idx.search([1.0, 0.0], 10)
    (ia, da), (ib, db) = build(), build()
    assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32))


def test_ef_zero_returns_nothing():                        # :470-472
    idx = index()
    idx.insert([1.0, 0.0])
    assert len(idx.search([1.0, 0.0], 0)[0]) == 0


def test_mismatched_query_length_degrades_to_max_distance():   # metric.rs:17-18,220-222
    idx = index()
    idx.insert([1.0, 0.0])
    ids, d = idx.search([1.0, 0.0, 0.0], 5)
    assert list(ids) == [0] and d[0] == 2.0


def test_caps_and_entry_tracks_max_layer():                # rwlock_hnsw.rs:409-452 (same algorithm)
    rng = np.random.default_rng(11)
    idx = index(seed=7, max_connections=4, ef_construction=16)
    x = datasets.unit_vectors(rng, 400, 6)
    for v in x:
        idx.insert(v)
    assert idx.check_invariants() == 0
    eid, etop = idx.entry()
    tops = [idx.top_layer(i) for i in range(400)]
    assert etop == max(tops) and tops[eid] == etop
    assert eid == tops.index(etop)                          # first node to reach the max layer
    for i in range(400):
        assert len(idx.neighbors(i, 0)) <= 8
        for l in range(1, tops[i] + 1):
            assert len(idx.neighbors(i, l)) <= 4


def _arb_vectors(rng, n, dim=8):                           # tests/property_test.rs:13-20
    out = []
    while len(out) < n:
        v = rng.uniform(-100.0, 100.0, dim).astype(np.float32)
        if np.sqrt((v * v).sum()) > 1e-3:
            out.append(v)
    return out


@pytest.mark.parametrize("case", range(24))
def test_index_matches_the_linear_scan_model_exactly(case):    # tests/property_test.rs:146-229
    rng = np.random.default_rng(1000 + case)
    n_ops = int(rng.integers(1, 48))
    kinds = rng.integers(0, 2, n_ops)
    vecs = _arb_vectors(rng, n_ops)
    queries = _arb_vectors(rng, int(rng.integers(1, 4)))
    idx = index(seed=31)
    model = []
    for i, (kind, v) in enumerate(zip(kinds, vecs)):
        if kind == 0:
            assert idx.insert(v) == len(model)              # handles follow insertion order
            model.append([v, True])
        elif model:
            pick = i % len(model)
            assert idx.remove(pick) == model[pick][1]
            model[pick][1] = False
    live = [(h, v) for h, (v, alive) in enumerate(model) if alive]
    for q in queries:
        qn = oracle.preprocess(q)
        exp = [(oracle.distance_prepared(oracle.preprocess(v), qn), h) for h, v in live]
        exp.sort(key=lambda t: oracle.total_key(t[0]))      # stable sort by distance only
        ids, d = idx.search(q, len(model) + 8)              # exhaustive ef
        assert len(ids) == len(exp)
        for (gid, gd), (wd, wid) in zip(zip(ids, d), exp):
            assert gid == wid
            assert abs(gd - wd) < 1e-6
        if live:
            x = np.stack([oracle.preprocess(v) for v, _ in model]) if model else None
            dele = np.array([0 if a else 1 for _, a in model], dtype=np.uint8)
            bi, bd = oracle.brute_force(x, q[None, :], len(live), mode="prepared", deleted=dele)
            assert list(bi[0]) == [h for _, h in exp]
            assert np.array_equal(bd[0], np.array([w for w, _ in exp], dtype=np.float32))


def test_self_query_distance_is_zero():                    # tests/property_test.rs:29-47
    rng = np.random.default_rng(9)
    vs = _arb_vectors(rng, 40)
    idx = index(seed=3)
    for v in vs:
        idx.insert(v)
    for h, v in enumerate(vs):
        ids, d = idx.search(v, 50)
        assert d[0] < 1e-5


def test_results_sorted_and_unique():                      # tests/property_test.rs:233-249
    rng = np.random.default_rng(10)
    vs = _arb_vectors(rng, 49)
    idx = index(seed=11)
    for v in vs:
        idx.insert(v)
    ids, d = idx.search(vs[0], 20)
    assert np.all(np.diff(d) >= 0) and len(set(ids.tolist())) == len(ids)


K = 10


def _recall(got, expected):
    hits = 0
    for g, e in zip(got, expected):
        hits += len(set(g[:K].tolist()) & set(e[:K].tolist()))
    return hits / (K * len(expected))


def _measure(data, queries, ef, seed):                     # tests/recall_test.rs:86-110
    idx = index(seed=seed)
    assert idx.insert_matrix(data) == len(data)
    assert idx.check_invariants() == 0
    exp, _ = oracle.brute_force(data, queries, K, mode="distance", threads=4)
    got, _, _, _ = idx.search_batch(queries, max(K, ef), K, threads=4)
    return _recall(got, exp)


def test_recall_gate_dim_32():                             # tests/recall_test.rs:140-145
    data, queries = datasets.uniform_dataset(2000, 20, 32, 0xC0FFEE)
    assert _measure(data, queries, 50, 0xC0FFEE) >= 0.95


def test_recall_gate_dim_768_embedding_like():             # :147-152
    data, queries = datasets.embedding_dataset(2000, 20, 0xBEEF)
    assert _measure(data, queries, 50, 0xBEEF) >= 0.95


def test_recall_gate_dim_768_uniform_high_ef():            # :158-162
    data, queries = datasets.uniform_dataset(2000, 20, 768, 0xBEEF)
    assert _measure(data, queries, 200, 0xBEEF) >= 0.95


def test_recall_survives_tombstones():                     # :166-214
    rng = np.random.default_rng(0xDEAD)
    data = datasets.unit_vectors(rng, 2000, 32)
    idx = index(seed=0xDEAD)
    idx.insert_matrix(data)
    dele = np.zeros(2000, dtype=np.uint8)
    for h in range(0, 2000, 3):
        assert idx.remove(h)
        dele[h] = 1
    queries = datasets.unit_vectors(rng, 20, 32)
    exp, _ = oracle.brute_force(data, queries, K, mode="distance", deleted=dele)
    got, _, _, _ = idx.search_batch(queries, 50, K)
    assert np.all(got % 3 != 0)                             # no tombstone leaks
    assert _recall(got, exp) >= 0.90


def test_sharded_merge_matches_global_scan():              # src/index/sharded_rwlock.rs:81-94
    rng = np.random.default_rng(21)
    data = oracle.preprocess_rows(datasets.unit_vectors(rng, 640, 16))
    q = datasets.unit_vectors(rng, 5, 16)
    S, k = 4, 10
    ids = np.empty((S, 5, k), np.uint32)
    ds = np.empty((S, 5, k), np.float32)
    for s in range(S):
        ids[s], ds[s] = oracle.brute_force(data[s::S], q, k, mode="prepared")
    mi, md = oracle.sharded_merge(ids, ds, k)
    gi, gd = oracle.brute_force(data, q, k, mode="prepared")
    assert np.array_equal(mi, gi) and np.array_equal(md, gd)
