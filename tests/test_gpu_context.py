"""`ChronoMind::search_in_context` (src/store.rs:353-377) and the CLI `query` front-end (src/main.rs:195-230) on the GPU
engine, against the oracle's restatement: exact scan of the context's live members with `metric.distance` on the RAW
stored rows (full cosine with norms), `combined_score`, sort by score, truncate."""
import subprocess
import sys
import os

import numpy as np
import pytest

import oracle
from oracle import snapshot_codec as sc
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, CmError, default_config

from parity_helpers import assert_order_matches_up_to_ties

pytestmark = pytest.mark.gpu
NOW = (1_790_000_000, 250_000_000)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _store(n, dim, seed, n_ctx=4):
    rng = np.random.default_rng(seed)
    raw = (rng.normal(size=(n, dim)) * rng.uniform(0.2, 5.0, size=(n, 1))).astype(np.float32)   # NOT unit length
    raw[7] = 0.0                                                                                  # zero vector -> 2.0
    ctx = rng.integers(0, n_ctx, size=n).astype(np.uint32)
    dele = (rng.uniform(size=n) < 0.1).astype(np.uint8)
    ts_s, ts_n, dec = datasets.temporal_attrs(n, seed, NOW[0])
    st = ChronoMindB200.from_matrix(raw, default_config(dimensions=dim), ts_s, ts_n, dec, dele).set_contexts(raw, ctx).upload(0)
    return st, raw, ctx, dele, (ts_s, ts_n, dec)


@pytest.mark.parametrize("n,nq,dim,k,w", [(3000, 33, 64, 10, 0.0), (3000, 33, 64, 10, 0.3), (777, 5, 19, 40, 0.3),
                                          (20000, 40, 768, 10, 0.3)])
def test_search_in_context_matches_oracle(n, nq, dim, k, w):
    st, raw, ctx, dele, (ts_s, ts_n, dec) = _store(n, dim, n + dim)
    rng = np.random.default_rng(99)
    q = (rng.normal(size=(nq, dim)) * 3.0).astype(np.float32)
    want = rng.integers(0, 5, size=nq).astype(np.uint32)           # context 4 has no members
    ids, score = st.search_in_context(want, q, k, now=NOW, temporal_weight=w, base_decay_rate=0.1)
    rc, wi, ws = oracle.search_in_context(raw, ctx, q, want, k, w, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1], deleted=dele, threads=8)
    assert rc == 0
    assert np.all(ids[want == 4] == NO_ID) and np.all(np.isinf(score[want == 4]))
    live = ids != NO_ID
    assert np.array_equal(live, wi != NO_ID)
    assert np.all(ctx[ids[live]] == np.broadcast_to(want[:, None], ids.shape)[live]) and not dele[ids[live]].any()
    if w == 0.0:      # score = distance / 2 exactly: the metric arithmetic is restated bit for bit
        assert np.array_equal(ids, wi) and np.array_equal(bits(score), bits(ws))
    else:             # expf may differ in the last ulp between libm and the device's correctly rounded exp
        assert assert_order_matches_up_to_ties(ids, wi, score, ws) <= 0.001 * ids.size
    assert np.all(np.diff(score, axis=1)[live[:, 1:]] >= 0)


@pytest.mark.parametrize("n,nq,dim,k,w,n_ctx", [(20000, 300, 768, 10, 0.3, 4), (30000, 257, 128, 50, 0.7, 200), (12000, 64, 64, 10, 0.0, 3),
                                                (50000, 600, 96, 20, 0.5, 1000), (9000, 70, 768, 100, 0.3, 2)])
def test_search_in_context_tensor_core_engine_equals_fp32_engine_and_oracle(n, nq, dim, k, w, n_ctx):
    """The tensor-core context path (biased threshold scan + exact decision on the stored rows) must return what the
    fp32 context scan returns BIT FOR BIT — both end in the same arithmetic on the same rows, ordered by (score, handle)
    — and that equals the oracle up to last-ulp expf ties."""
    st, raw, ctx, dele, (ts_s, ts_n, dec) = _store(n, dim, n + k, n_ctx=n_ctx)
    rng = np.random.default_rng(7)
    q = (rng.normal(size=(nq, dim)) * 2.0).astype(np.float32)
    q[5] = raw[11] * 3.0                                         # a stored direction: near-zero distances
    want = rng.integers(0, n_ctx + 1, size=nq).astype(np.uint32)   # id n_ctx has no members
    from chronomind_b200.engine import last_counters
    ia, sa = st.set_exact_engine(2).search_in_context(want, q, k, now=NOW, temporal_weight=w, base_decay_rate=0.1)
    ca = last_counters()
    ib, sb = st.set_exact_engine(1).search_in_context(want, q, k, now=NOW, temporal_weight=w, base_decay_rate=0.1)
    cb = last_counters()
    assert ca["exact_engine"] == 2 and cb["exact_engine"] == 1 and ca["fallback_queries"] == 0
    assert np.array_equal(ia, ib) and np.array_equal(bits(sa), bits(sb))
    rc, wi, ws = oracle.search_in_context(raw, ctx, q, want, k, w, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1], deleted=dele, threads=8)
    assert rc == 0
    if w == 0.0:
        assert np.array_equal(ia, wi) and np.array_equal(bits(sa), bits(ws))
    else:
        assert assert_order_matches_up_to_ties(ia, wi, sa, ws) <= 0.001 * ia.size
    live = ia != NO_ID
    assert np.all(ctx[ia[live]] == np.broadcast_to(want[:, None], ia.shape)[live]) and not dele[ia[live]].any()


def test_search_in_context_errors_and_states():
    st, raw, ctx, dele, _ = _store(500, 8, 1)
    with pytest.raises(CmError) as e:
        st.search_in_context(0, np.ones((1, 9), np.float32), 3, now=NOW)
    assert e.value.kind == "InvalidDimensions"
    bad = np.ones((3, 8), np.float32)
    bad[2, 4] = np.inf
    with pytest.raises(CmError) as e:
        st.search_in_context(0, bad, 3, now=NOW)
    assert e.value.kind == "InvalidVector" and "row 2" in str(e.value)
    plain = ChronoMindB200.from_matrix(raw, default_config(dimensions=8)).upload(0)
    with pytest.raises(CmError) as e:
        plain.search_in_context(0, np.ones((1, 8), np.float32), 3, now=NOW)
    assert e.value.kind == "State"


def test_snapshot_contexts_and_cli_query(tmp_path):
    """The reference's CLI e2e (tests/cli_test.rs:39-55) over its only fixture's records (examples/sample_vectors.json
    values restated here): `query` for vector1's own vector lists vector1; `--context context_b` lists only b's."""
    sample = [("vector1", [0.1, 0.2, 0.3, 0.4], 0.8, "context_a", 0.1), ("vector2", [0.2, 0.3, 0.4, 0.5], 0.6, "context_b", 0.2),
              ("vector3", [0.3, 0.4, 0.5, 0.6], 0.7, "context_a", 0.15), ("vector4", [0.4, 0.5, 0.6, 0.7], 0.9, "context_c", 0.05),
              ("vector5", [0.5, 0.6, 0.7, 0.8], 0.5, "context_b", 0.25)]
    p = str(tmp_path / "s.chrono")
    sc.write_snapshot(p, sc.Config(dimensions=4), [sc.Memory(id=i, data=d, ts_secs=NOW[0] - 3600, importance=imp, context=c,
                                                             decay_rate=r) for i, d, imp, c, r in sample])
    st = ChronoMindB200.from_snapshot(p).upload(0)
    assert st.context_id("context_a") == 0 and st.context_id("context_b") == 1 and st.context_id("nope") is None
    assert st.context_label(3) == "context_c" and abs(st.importance(3) - 0.9) < 1e-6
    ids, score = st.search_in_context("context_b", np.array([[0.1, 0.2, 0.3, 0.4]], np.float32), 5, now=NOW)
    assert [st.external_id(int(h)) for h in ids[0] if h != NO_ID] == ["vector2", "vector5"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "chronomind_b200", "query", "-f", p, "-v", "[0.1, 0.2, 0.3, 0.4]", "-l", "3"],
                         capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr
    assert "Found 3 result(s):" in out.stdout and "id: vector1" in out.stdout
    out = subprocess.run([sys.executable, "-m", "chronomind_b200", "query", "-f", p, "-v", "0.1,0.2,0.3,0.4", "-c", "context_b"],
                         capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l[:1].isdigit()]
    assert len(lines) == 2 and all("context: context_b" in l for l in lines) and "importance: 0.60" in out.stdout
    out = subprocess.run([sys.executable, "-m", "chronomind_b200", "query", "-f", p, "-v", "0.1,0.2", "-l", "3"],
                         capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 1 and "InvalidDimensions" in out.stderr


def test_search_in_context_sorted_layout_large_batches_sparse_ids_and_tombstones():
    """The context-sorted layout (rows ordered by context on the device, queries grouped by context per 16,384-query
    chunk, per-pair tile ranges): more queries than one sort chunk, context ids that are neither dense nor small, an
    unknown context, and tombstones that reach the sorted copy through `cm_index_sync_deleted` — bit for bit the fp32
    context scan, and the oracle on a slice."""
    from chronomind_b200.engine import last_counters
    n, dim, k, nq = 40000, 64, 10, 17000
    rng = np.random.default_rng(2024)
    raw = (rng.normal(size=(n, dim)) * rng.uniform(0.2, 5.0, size=(n, 1))).astype(np.float32)
    labels = np.array([7, 3_000_000_000, 12, 99_999, 0, 4_000_000], dtype=np.uint32)          # sparse ids, uneven sizes
    ctx = labels[rng.choice(len(labels), size=n, p=[0.5, 0.2, 0.15, 0.1, 0.04, 0.01])]
    dele = (rng.uniform(size=n) < 0.05).astype(np.uint8)
    ts_s, ts_n, dec = datasets.temporal_attrs(n, 5, NOW[0])
    st = ChronoMindB200.from_matrix(raw, default_config(dimensions=dim), ts_s, ts_n, dec, dele).set_contexts(raw, ctx).upload(0)
    q = (rng.normal(size=(nq, dim)) * 2.0).astype(np.float32)
    want = np.concatenate([labels, [555]]).astype(np.uint32)[rng.integers(0, len(labels) + 1, size=nq)]   # 555: no members
    for rnd in range(2):
        ia, sa = st.set_exact_engine(2).search_in_context(want, q, k, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
        ca = last_counters()
        ib, sb = st.set_exact_engine(1).search_in_context(want, q, k, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
        assert ca["exact_engine"] == 2 and ca["fallback_queries"] == 0
        assert np.array_equal(ia, ib) and np.array_equal(bits(sa), bits(sb)), rnd
        assert np.all(ia[want == 555] == NO_ID)
        live = ia != NO_ID
        assert np.all(ctx[ia[live]] == np.broadcast_to(want[:, None], ia.shape)[live]) and not dele[ia[live]].any()
        sl = slice(16300, 16500)                                    # straddles the sort-chunk boundary
        rc, wi, ws = oracle.search_in_context(raw, ctx, q[sl], want[sl], k, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1],
                                              deleted=dele, threads=8)
        assert rc == 0 and assert_order_matches_up_to_ties(ia[sl], wi, sa[sl], ws) <= 0.001 * wi.size
        if rnd == 0:                                                # tombstone the current winners, sync, search again
            gone = np.unique(ia[:500, 0][ia[:500, 0] != NO_ID])
            for h in gone:
                assert st.remove(int(h))
            dele[gone] = 1
            st.sync_deleted()
