"""Tensor-core exact scan (tcgen05 bf16 scan + exact fp32 rescore), forced with engine = 2, against the oracle:
the result must be the EXACT top-k — ids bit-exact, distances bit-identical — not a bf16 approximation."""
import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, default_config, last_counters

pytestmark = pytest.mark.gpu


def bits(a):
    """Bit patterns, with every NaN mapped to one pattern (x86 keeps the operand's NaN payload, CUDA returns the
    canonical NaN; the reference only promises that NaN propagates, src/metric.rs:223)."""
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def store(data, deleted=None, engine=2):
    st = ChronoMindB200.from_matrix(data, default_config(dimensions=data.shape[1]), deleted=deleted)
    return st.upload(0).set_exact_engine(engine)


@pytest.mark.parametrize("n,nq,dim,k,kind", [(20000, 300, 768, 10, "emb"), (20000, 300, 768, 10, "uni"),
                                             (5000, 77, 100, 10, "uni"), (33000, 129, 768, 32, "emb"),
                                             (4097, 1, 64, 1, "uni"), (9000, 500, 32, 64, "uni"),
                                             (3000, 40, 771, 10, "emb"), (300000, 700, 64, 10, "uni"),
                                             (200000, 257, 768, 10, "emb"), (150000, 520, 128, 50, "uni"),
                                             (400000, 300, 32, 100, "uni"), (60000, 260, 768, 100, "uni"),
                                             (9000, 70, 4096, 10, "uni"), (8500, 40, 8192, 10, "uni"),
                                             (8200, 24, 12800, 10, "emb")])   # longest rows the engines accept
def test_tc_scan_is_exact(n, nq, dim, k, kind):
    if kind == "emb":
        data, q = datasets.embedding_dataset(n, nq, 0xBEEF + n, dim=dim)
    else:
        data, q = datasets.uniform_dataset(n, nq, dim, 0xC0FFEE + n)
    st = store(data)
    ids, d = st.search_exact(q, k)
    c = last_counters()
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data, threads=8), q, k, mode="prepared", threads=8)
    assert np.array_equal(ids, wi)
    assert np.array_equal(bits(d), bits(wd))
    assert c["rescored"] >= nq * min(k, n)               # the tensor-core path actually ran
    assert c["fallback_queries"] == 0


def test_tc_scan_tombstones_duplicates_and_short_rows():
    rng = np.random.default_rng(12)
    data = datasets.unit_vectors(rng, 6000, 96)
    data[100:140] = data[50]                            # 41 identical rows: ties resolved by handle
    dele = (rng.uniform(size=6000) < 0.3).astype(np.uint8)
    dele[50] = 1
    q = np.vstack([data[50:51], datasets.unit_vectors(rng, 70, 96)])
    st = store(data, dele)
    ids, d = st.search_exact(q, 20)
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 20, mode="prepared", deleted=dele)
    assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
    assert not dele[ids[ids != NO_ID]].any()
    # fewer live rows than k: padded
    small = datasets.unit_vectors(rng, 9000, 16)
    dele = np.ones(9000, np.uint8)
    dele[[5, 77, 8000]] = 0
    st = store(small, dele)
    ids, d = st.search_exact(datasets.unit_vectors(rng, 64, 16), 10)
    assert np.all(np.sort(ids[:, :3], axis=1) == np.array([5, 77, 8000])) and np.all(ids[:, 3:] == NO_ID)


def test_tc_scan_large_k_threshold_ladder_edge_cases():
    """k > 32 takes the histogram threshold (Ladder) in K1's epilogue instead of the exact running top-k list: the
    result must still be the exact top-k where that structure is stressed — every score negative (lowest buckets),
    scores packed into one bucket (near-duplicates of the query), exact duplicates straddling the k-th place with
    tombstones, fewer live rows than k, and rows identical to the query (top bucket, score 1.0)."""
    rng = np.random.default_rng(41)
    dim = 48
    # (a) rows in a cone around +e0, queries around -e0: all scores in [-1, -0.5]
    rows = datasets.unit_vectors(rng, 20000, dim) * 0.35
    rows[:, 0] += 1.0
    qs = datasets.unit_vectors(rng, 90, dim) * 0.35
    qs[:, 0] -= 1.0
    st = store(rows.astype(np.float32))
    for k in (33, 100):
        ids, d = st.search_exact(qs.astype(np.float32), k)
        wi, wd = oracle.brute_force(oracle.preprocess_rows(rows.astype(np.float32)), qs.astype(np.float32), k, mode="prepared")
        assert last_counters()["fallback_queries"] == 0
        assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
    # (b) tight cluster around each query (scores within one ladder step of each other), exact duplicates across the
    # k-th place, 30 % tombstones
    base = datasets.unit_vectors(rng, 12000, dim)
    centers = datasets.unit_vectors(rng, 8, dim)
    for c in range(8):
        base[c * 300:(c + 1) * 300] = centers[c] + 0.01 * rng.standard_normal((300, dim)).astype(np.float32)
    base[5000:5100] = base[10]          # 101 identical rows inside cluster 0
    base[6000:6040] = centers[3]        # rows identical to a query
    dele = (rng.uniform(size=12000) < 0.3).astype(np.uint8)
    q = np.vstack([centers, datasets.unit_vectors(rng, 60, dim)]).astype(np.float32)
    st = store(base.astype(np.float32), dele)
    for k in (48, 128):
        ids, d = st.search_exact(q, k)
        wi, wd = oracle.brute_force(oracle.preprocess_rows(base.astype(np.float32)), q, k, mode="prepared", deleted=dele)
        assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
    # (c) fewer live rows than k
    dele = np.ones(12000, np.uint8)
    dele[[3, 700, 11999]] = 0
    ids, d = store(base.astype(np.float32), dele).search_exact(q, 64)
    assert np.all(np.sort(ids[:, :3], axis=1) == np.array([3, 700, 11999])) and np.all(ids[:, 3:] == NO_ID)


def test_tc_scan_nonfinite_query_falls_back_to_fp32():
    """A NaN query yields NaN scores on the tensor cores (nothing passes the threshold): K2 flags the query and the
    in-stream fp32 scan reproduces the reference's result for it; the other queries are unaffected."""
    rng = np.random.default_rng(3)
    data = datasets.unit_vectors(rng, 9000, 64)
    q = datasets.unit_vectors(rng, 70, 64)
    q[13, 5] = np.nan
    st = store(data)
    ids, d = st.search_exact(q, 10)
    c = last_counters()
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared")
    assert c["fallback_queries"] == 1
    assert np.array_equal(ids, wi)
    assert np.array_equal(bits(d), bits(wd))


def test_auto_engine_matches_forced_engines():
    data, q = datasets.embedding_dataset(12000, 128, 5)
    a = store(data, engine=0).search_exact(q, 10)
    b = store(data, engine=1).search_exact(q, 10)
    c = store(data, engine=2).search_exact(q, 10)
    for x in (b, c):
        assert np.array_equal(a[0], x[0]) and np.array_equal(bits(a[1]), bits(x[1]))


def test_tc_scan_full_size_matches_fp32_engine():
    """BASELINE configs[1] size (1M x 768, k = 10): the tensor-core path must return exactly what the bit-exact
    fp32 CUDA-core scan returns (itself checked against the oracle at oracle-sized cases), with no query routed
    to the fallback; a 48-query slice is also checked against the oracle directly."""
    rng = np.random.default_rng(0x5EED)
    basis = datasets.unit_vectors(rng, 16, 768)
    corpus = datasets.embedding_rows(rng, basis, 1_000_000)
    q = datasets.embedding_rows(np.random.default_rng(99), basis, 2048)
    st = ChronoMindB200.from_matrix(corpus, default_config(dimensions=768)).upload(0)
    ia, da = st.set_exact_engine(2).search_exact(q, 10)
    c = last_counters()
    assert c["fallback_queries"] == 0 and c["rescored"] >= 2048 * 10
    ib, db = st.set_exact_engine(1).search_exact(q, 10)
    assert np.array_equal(ia, ib) and np.array_equal(bits(da), bits(db))
    wi, wd = oracle.brute_force(oracle.preprocess_rows(corpus, threads=16), q[:48], 10, mode="prepared", threads=16)
    assert np.array_equal(ia[:48], wi) and np.array_equal(bits(da[:48]), bits(wd))
    assert np.all(np.diff(da, axis=1) >= 0)          # ascending


def test_tc_scan_more_flagged_queries_than_one_fallback_chunk():
    """200 NaN queries (> one 64-query fallback chunk): every one of them is resolved by the in-stream fp32 scan."""
    rng = np.random.default_rng(8)
    data = datasets.unit_vectors(rng, 9000, 64)
    q = datasets.unit_vectors(rng, 400, 64)
    q[::2, 3] = np.nan
    st = store(data)
    ids, d = st.search_exact(q, 10)
    c = last_counters()
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared")
    assert c["fallback_queries"] == 200
    assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))


def test_tc_scan_batches_larger_than_one_query_chunk():
    """70,000 queries (> the 32,768-query chunk that bounds the candidate buffer): chunks are scanned back to back in
    one call; counters are the call's totals."""
    data, q = datasets.uniform_dataset(20000, 70000, 32, 77)
    st = store(data)
    ids, d = st.search_exact(q, 5)
    c = last_counters()
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 5, mode="prepared", threads=16)
    assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
    assert c["fallback_queries"] == 0 and c["rescored"] >= 70000 * 5
