"""GPU parity tests proper: every call goes through the C ABI of libchronomind_b200.so and is compared with the
CPU oracle on the same seeded inputs. Bars (BASELINE.json north_star): exact-scan ids bit-exact, distances and
temporal scores within 1e-5 relative (they are in fact bit-identical: the kernels replay the reference's fp32
operation order), HNSW results identical to the oracle's search over the same graph."""
import numpy as np
import pytest

import oracle
from oracle import snapshot_codec as sc
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, CmError, Index, default_config, last_counters, native
from parity_helpers import assert_order_matches_up_to_ties

pytestmark = pytest.mark.gpu

NOW = (1_790_000_000, 250_000_000)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def make_store(data, cfg=None, attrs=True, graph_seed=None, threads=1, deleted=None, seed=3):
    cfg = cfg or default_config(dimensions=data.shape[1])
    ts_s = ts_n = dec = None
    if attrs:
        ts_s, ts_n, dec = datasets.temporal_attrs(len(data), seed, NOW[0])
    st = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec, deleted)
    if graph_seed is not None:
        st.build_graph(graph_seed, threads)
    return st.upload(0), (ts_s, ts_n, dec)


# ---- metric surface ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 7, 8, 9, 15, 16, 17, 768, 771])
def test_metric_kernels_bit_exact(dim):
    """preprocess / distance_prepared on the device == src/metric.rs:208-224 bit for bit (lengths around the
    8-lane boundary as in src/metric.rs:267-287)."""
    import torch
    rng = np.random.default_rng(dim)
    a = rng.uniform(-100, 100, (33, dim)).astype(np.float32)
    b = rng.uniform(-100, 100, (33, dim)).astype(np.float32)
    a[3] = 0.0
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    pa, pb = torch.empty_like(ta), torch.empty_like(tb)
    L = native()
    assert L.cm_metric_preprocess_dev(ta.data_ptr(), 33, dim, pa.data_ptr(), 0, None) == 0
    assert L.cm_metric_preprocess_dev(tb.data_ptr(), 33, dim, pb.data_ptr(), 0, None) == 0
    out = torch.empty(33, dtype=torch.float32, device="cuda")
    assert L.cm_metric_distance_prepared_dev(pa.data_ptr(), pb.data_ptr(), 33, dim, out.data_ptr(), 0, None) == 0
    torch.cuda.synchronize()
    wa, wb = oracle.preprocess_rows(a), oracle.preprocess_rows(b)
    assert np.array_equal(bits(pa.cpu().numpy()), bits(wa))
    want = np.array([oracle.distance_prepared(wa[i], wb[i]) for i in range(33)], np.float32)
    assert np.array_equal(bits(out.cpu().numpy()), bits(want))


def test_metric_known_answers_on_device():
    """src/metric.rs:233-256 through the device kernels: identical -> 0, orthogonal -> 1, opposite -> 2."""
    st, _ = make_store(np.array([[1, 0.5, -0.25, 2], [0, 0, 1, 0], [-1, -0.5, 0.25, -2]], np.float32), attrs=False)
    ids, d = st.search_exact(np.array([[1, 0.5, -0.25, 2]], np.float32), 3)
    assert list(ids[0]) == [0, 1, 2]
    assert abs(d[0][0]) < 1e-5 and abs(d[0][2] - 2.0) < 1e-5


# ---- exact scan ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nq,dim,k,kind", [(2000, 20, 768, 10, "emb"), (2000, 20, 32, 10, "uni"),
                                             (10000, 100, 768, 10, "emb"), (5000, 33, 768, 100, "uni"),
                                             (777, 5, 13, 10, "uni"), (300, 3, 8, 1, "uni")])
def test_exact_scan_matches_oracle(n, nq, dim, k, kind):
    """tests/property_test.rs:201-206 arithmetic; ids exact, distances bit-identical, ascending (dist, handle)."""
    if kind == "emb":
        data, q = datasets.embedding_dataset(n, nq, 0xBEEF, dim=dim)
    else:
        data, q = datasets.uniform_dataset(n, nq, dim, 0xC0FFEE)
    st, _ = make_store(data, attrs=False)
    ids, d = st.search_exact(q, k)
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, k, mode="prepared", threads=8)
    assert np.array_equal(ids, wi)
    assert np.array_equal(bits(d), bits(wd))


def test_exact_scan_edge_cases():
    rng = np.random.default_rng(4)
    data = datasets.unit_vectors(rng, 50, 16)
    data[7] = data[3]                                   # exact duplicate: tie broken by the lower handle
    data[11] = 0.0                                      # degenerate row: stays zero, distance 1.0
    dele = np.zeros(50, np.uint8)
    dele[[0, 3, 20]] = 1
    st, _ = make_store(data, attrs=False, deleted=dele)
    q = np.vstack([data[3:4], datasets.unit_vectors(rng, 2, 16)])
    ids, d = st.search_exact(q, 60)                     # k > live count: padded rows
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 60, mode="prepared", deleted=dele)
    assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
    assert ids[0][0] == 7 and np.all(ids[:, 47:] == NO_ID) and np.all(np.isinf(d[:, 47:]))
    assert not np.isin(ids, [0, 3, 20]).any()
    # query length != dim: every distance is exactly 2.0 (src/metric.rs:220-222), handles ascending
    ids, d = st.search_exact(np.ones((1, 5), np.float32), 4)
    assert list(ids[0]) == [1, 2, 4, 5] and np.all(d == 2.0)
    # nq == 0
    ids, d = st.search_exact(np.zeros((0, 16), np.float32), 3)
    assert ids.shape == (0, 3)


def test_nan_query_orders_like_total_cmp_on_both_engines():
    """Index-level search never fails on a degenerate query: a NaN component makes every distance NaN
    (src/metric.rs:223 — clamp passes NaN through), and (TotalF32, handle) ordering then returns the lowest handles.
    (Regression: a float-domain rewrite of total_key by the compiler once turned those NaN keys into -0.0.)"""
    rng = np.random.default_rng(5)
    data = datasets.unit_vectors(rng, 9000, 64)
    q = datasets.unit_vectors(rng, 70, 64)
    q[13, 5] = np.nan
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared")
    assert np.isnan(wd[13]).all()
    for engine in (1, 2):
        st, _ = make_store(data, attrs=False)
        ids, d = st.set_exact_engine(engine).search_exact(q, 10)
        assert np.array_equal(ids, wi)
        assert np.isnan(d[13]).all()
        keep = np.arange(70) != 13
        assert np.array_equal(bits(d[keep]), bits(wd[keep]))


def test_empty_index():
    st = ChronoMindB200.from_matrix(np.zeros((0, 8), np.float32), default_config(dimensions=8))
    st.build_graph(1).upload(0)
    ids, d = st.search_exact(np.ones((2, 8), np.float32), 3)
    assert np.all(ids == NO_ID) and np.all(np.isinf(d))
    ids, d = st.search_index(np.ones((2, 8), np.float32), 10, 3)       # lockfree_hnsw.rs:511-516
    assert np.all(ids == NO_ID)


# ---- HNSW ------------------------------------------------------------------------------------------------------------
def _oracle_for(data, M, efc, seed, deleted=None):
    orc = oracle.OracleHnsw(M, efc, 50, seed)
    orc.insert_matrix(data)
    if deleted is not None:
        for h in np.nonzero(deleted)[0]:
            orc.remove(int(h))
    return orc


@pytest.mark.parametrize("n,nq,dim,ef,kind,seed", [(2000, 20, 768, 50, "emb", 0xBEEF), (2000, 20, 32, 50, "uni", 0xC0FFEE),
                                                   (2000, 20, 768, 200, "uni", 0xBEEF), (10000, 200, 768, 64, "emb", 5),
                                                   (10000, 100, 768, 128, "emb", 5), (3000, 50, 100, 200, "uni", 9),
                                                   (500, 10, 13, 10, "uni", 2), (1500, 12, 4096, 40, "uni", 3),
                                                   (1200, 8, 300, 1000, "uni", 4)])
def test_hnsw_search_identical_to_oracle_on_same_graph(n, nq, dim, ef, kind, seed):
    """Same graph (single-threaded build == the oracle's), equal ef: ids identical, distances bit-identical,
    and the work counters (dist evals, expansions) equal the CPU's — the traversal is the same traversal."""
    if kind == "emb":
        data, q = datasets.embedding_dataset(n, nq, seed, dim=dim)
    else:
        data, q = datasets.uniform_dataset(n, nq, dim, seed)
    cfg = default_config(dimensions=dim, ef_construction=100)
    st, _ = make_store(data, cfg, attrs=False, graph_seed=seed)
    orc = _oracle_for(data, 16, 100, seed)
    ids, d = st.search_index(q, ef, ef)
    c = last_counters()
    wi, wd, cnt, octr = orc.search_batch(q, ef, ef)
    assert np.array_equal(ids, wi)
    assert np.array_equal(bits(d), bits(wd))
    assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1]) and c["list_entries"] == int(octr[2])
    # recall gates of tests/recall_test.rs (>= 0.95 on these distributions at these ef)
    exp, _ = oracle.brute_force(data, q, 10, mode="distance", threads=8)
    rec = np.mean([len(set(a[:10].tolist()) & set(b.tolist())) / 10 for a, b in zip(ids, exp)])
    if (kind, ef) in (("emb", 50), ("uni", 200)) or dim == 32:
        assert rec >= 0.95


def _with_duplicates(data, frac, seed):
    """Overwrite `frac` of the rows with exact copies of earlier rows (ordinary in a memory store)."""
    rng = np.random.default_rng(seed)
    n = len(data)
    dst = rng.choice(np.arange(n // 4, n), size=int(n * frac), replace=False)
    data = data.copy()
    data[dst] = data[rng.integers(0, n // 4, size=len(dst))]
    return data


@pytest.mark.parametrize("n,dim,M,frac,ef", [(3000, 8, 16, 0.2, 10), (3000, 8, 16, 0.2, 64), (4000, 16, 16, 0.05, 10),
                                             (4000, 16, 16, 0.05, 64), (2500, 6, 4, 0.35, 24), (3000, 12, 32, 0.2, 64),
                                             (6000, 768, 16, 0.2, 64)])
def test_hnsw_exact_duplicates_at_the_admission_boundary(n, dim, M, frac, ef):
    """`search_layer` admits on distance only, strictly (`d < worst`, src/index/lockfree_hnsw.rs:182), breaks on
    `dist > worst` (:166-169) and evicts by (distance, id): with bit-identical distances on the cut at ef the order of
    arrival decides. Corpora with 5-35 % exact duplicate rows, k = ef (the whole list is compared): ids, distance bits
    and work counters must still be the oracle's, and the kernel must actually have taken its tie replay."""
    rng = np.random.default_rng(n + ef)
    base = rng.uniform(-1, 1, size=(n, dim)).astype(np.float32) if dim < 100 else datasets.embedding_dataset(n, 1, 17)[0]
    data = _with_duplicates(base, frac, n)
    q = np.vstack([rng.uniform(-1, 1, size=(120, dim)).astype(np.float32) if dim < 100 else datasets.embedding_dataset(n, 120, 17)[1],
                   data[rng.integers(0, n, size=40)]])            # some queries ARE stored rows (distance-0 ties)
    cfg = default_config(dimensions=dim, max_connections=M, ef_construction=60)
    st, _ = make_store(data, cfg, attrs=False, graph_seed=11)
    orc = _oracle_for(data, M, 60, 11)
    ids, d = st.search_index(q, ef, ef)
    c = last_counters()
    wi, wd, _, octr = orc.search_batch(q, ef, ef)
    assert np.array_equal(ids, wi)
    assert np.array_equal(bits(d), bits(wd))
    assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1]) and c["list_entries"] == int(octr[2])
    assert c["failures"] == 0
    if dim < 100:
        assert c["tie_replays"] > 0, "no tie ever reached the cut: the workload does not test the replay"


def test_hnsw_all_rows_identical():
    """Degenerate extreme: every row is the same vector, so every distance ties. The reference fills `best` in arrival
    order and then rejects everything (`d < worst` is never true)."""
    data = np.tile(np.array([[0.3, -0.2, 0.9, 0.1]], np.float32), (400, 1))
    cfg = default_config(dimensions=4, max_connections=8, ef_construction=40)
    st, _ = make_store(data, cfg, attrs=False, graph_seed=5)
    orc = _oracle_for(data, 8, 40, 5)
    q = np.array([[0.3, -0.2, 0.9, 0.1], [1.0, 0.0, 0.0, 0.0]], np.float32)
    for ef in (1, 10, 50):
        ids, d = st.search_index(q, ef, ef)
        c = last_counters()
        wi, wd, _, octr = orc.search_batch(q, ef, ef)
        assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
        assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1])


def test_hnsw_tombstones_and_small_caps():
    """tests/recall_test.rs:166-214: delete id % 3 == 0; none leak; identical to the oracle. M=4 exercises
    list caps != 32."""
    rng = np.random.default_rng(0xDEAD)
    data = datasets.unit_vectors(rng, 2000, 32)
    q = datasets.unit_vectors(rng, 20, 32)
    dele = (np.arange(2000) % 3 == 0).astype(np.uint8)
    for M, efc in ((16, 200), (4, 16), (32, 64)):
        cfg = default_config(dimensions=32, max_connections=M, ef_construction=efc)
        st, _ = make_store(data, cfg, attrs=False, graph_seed=0xDEAD, deleted=dele)
        orc = _oracle_for(data, M, efc, 0xDEAD, dele)
        ids, d = st.search_index(q, 50, 10)
        wi, wd, _, _ = orc.search_batch(q, 50, 10)
        assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))
        assert np.all(ids[ids != NO_ID] % 3 != 0)


def test_hnsw_degenerate_query_and_single_vector():
    st, _ = make_store(np.array([[1.0, 0.0]], np.float32), default_config(dimensions=2), attrs=False, graph_seed=42)
    ids, d = st.search_index(np.array([[1.0, 0.0]], np.float32), 10, 10)    # lockfree_hnsw.rs:518-526
    assert ids[0][0] == 0 and d[0][0] < 1e-6 and np.all(ids[0][1:] == NO_ID)
    rng = np.random.default_rng(8)
    data = datasets.unit_vectors(rng, 300, 8)
    st, _ = make_store(data, default_config(dimensions=8), attrs=False, graph_seed=1)
    orc = _oracle_for(data, 16, 200, 1)
    ids, d = st.search_index(np.ones((2, 5), np.float32), 20, 20)           # wrong length: all distances 2.0
    wi, wd, _, _ = orc.search_batch(np.ones((2, 5), np.float32), 20, 20)
    assert np.all(d == 2.0) and np.all(wd == 2.0) and np.all(ids != NO_ID)  # src/metric.rs:220-222


def test_hnsw_multithread_graph_parity(tmp_path):
    """A concurrently built graph (non-deterministic linkage) searched by the oracle and the GPU: identical."""
    data, q = datasets.embedding_dataset(20000, 300, 77, dim=128, intrinsic=12)
    cfg = default_config(dimensions=128, ef_construction=100)
    st, _ = make_store(data, cfg, attrs=False, graph_seed=77, threads=8)
    assert st.check_invariants() == 0
    p = str(tmp_path / "g.cmgraph")
    st.save_graph(p)
    orc = oracle.OracleHnsw(16, 100, 50, 77)
    orc.import_graph(oracle.preprocess_rows(data), oracle.read_graph_sidecar(p))
    for ef in (64, 200):
        ids, d = st.search_index(q, ef, 10)
        wi, wd, _, _ = orc.search_batch(q, ef, 10, threads=8)
        assert np.array_equal(ids, wi) and np.array_equal(bits(d), bits(wd))


def test_pyo3_index_surface():
    """bindings/python/src/lib.rs:35-123: fit / set_ef_search / query / batch_query / __len__."""
    data, q = datasets.embedding_dataset(1500, 8, 21, dim=96, intrinsic=8)
    idx = Index(96, max_connections=16, ef_construction=100, ef_search=50, seed=0x5EED)
    idx.fit(data)
    assert len(idx) == 1500
    orc = _oracle_for(data, 16, 100, 0x5EED)
    got = idx.batch_query(q, 10)
    wi, _, _, _ = orc.search_batch(q, 50, 10)
    assert got == [list(map(int, r)) for r in wi]
    idx.set_ef_search(5)
    assert idx.query(q[0], 10) == list(map(int, orc.search_batch(q[:1], 10, 10)[0][0]))   # ef = max(5, n)


# ---- store-level search -----------------------------------------------------------------------------------------------
def _correctly_rounded_scores(dist, ts_s, ts_n, dec, ids, w, base):
    """combined_score with exp evaluated in float64 and rounded once — what the kernel computes."""
    f = np.float32
    out = np.empty(ids.shape, np.float32)
    for i in np.ndindex(ids.shape):
        h = ids[i]
        if h == NO_ID:
            out[i] = np.inf
            continue
        ds, dn = NOW[0] - int(ts_s[h]), NOW[1] - int(ts_n[h])
        if ds < 0 or (ds == 0 and dn < 0):
            ds = dn = 0
        elif dn < 0:
            ds, dn = ds - 1, dn + 1_000_000_000
        age = f(f(f(ds) + f(dn) / f(1e9)) / f(3600.0))
        rate = dec[h] if dec[h] > 0 else f(base)
        t = f(np.exp(np.float64(f(-rate) * age)))
        out[i] = f(f(f(1.0) - f(w)) * f(dist[i] / f(2.0))) + f(f(w) * f(f(1.0) - t))
    return out


def test_temporal_search_with_tensor_core_pool():
    """`cm_search_temporal(exact=1)` on a corpus large enough for the tensor-core engine: the exact pool of
    ef = max(ef_search, 3k) = 64 candidates comes from K1/K2, then the fused rerank."""
    data, q = datasets.embedding_dataset(20000, 130, 0xFEED)
    cfg = default_config(dimensions=768, ef_search=64, temporal_weight=0.3)
    ts_s, ts_n, dec = datasets.temporal_attrs(20000, 4, NOW[0])
    st = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec).upload(0)
    ids, score, dist = st.search(q, 10, now=NOW, exact=True, want_dist=True)
    assert last_counters()["exact_engine"] == 2
    pi, pd = oracle.brute_force(oracle.preprocess_rows(data, threads=8), q, 64, mode="prepared", threads=8)
    wi, ws = oracle.rerank_pool(pi, pd, 10, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert assert_order_matches_up_to_ties(ids, wi, score, ws) <= 0.001 * ids.size
    assert np.array_equal(bits(score), bits(_correctly_rounded_scores(dist, ts_s, ts_n, dec, ids, 0.3, 0.1)))


@pytest.mark.parametrize("exact", [False, True])
def test_temporal_search_matches_oracle(exact):
    """ChronoMind::search (src/store.rs:321-346): pool = max(ef_search, 3k) candidates, combined_score, stable
    sort, truncate. Order identical to the oracle; scores within 1e-5 relative (north_star) and bit-identical to the
    correctly rounded formula."""
    data, q = datasets.embedding_dataset(4000, 60, 0xFACE)
    cfg = default_config(dimensions=768, ef_construction=100, ef_search=64, temporal_weight=0.3)
    st, (ts_s, ts_n, dec) = make_store(data, cfg, graph_seed=0xFACE)
    k = 10
    ids, score, dist = st.search(q, k, now=NOW, exact=exact, want_dist=True)
    if exact:
        pi, pd = oracle.brute_force(oracle.preprocess_rows(data), q, 64, mode="prepared", threads=8)
        wi, ws = oracle.rerank_pool(pi, pd, k, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    else:
        orc = _oracle_for(data, 16, 100, 0xFACE)
        rc, wi, ws, _ = orc.store_search_batch(q, k, 768, 64, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1], threads=8)
        assert rc == 0
    # Rows whose neighbouring scores are separated by > 1e-5 relative must match in order exactly; every other
    # mismatch must sit on a near-tie of the oracle's own list.
    for r in range(len(q)):
        if (np.diff(ws[r]) > 1e-5 * np.abs(ws[r][1:]) + 1e-7).all():
            assert np.array_equal(ids[r], wi[r]), r
    assert assert_order_matches_up_to_ties(ids, wi, score, ws) <= 0.001 * ids.size
    assert np.array_equal(bits(score), bits(_correctly_rounded_scores(dist, ts_s, ts_n, dec, ids, 0.3, 0.1)))
    assert np.all(np.diff(score, axis=1) >= 0)


def test_temporal_orderings_of_the_reference_store_tests():
    # tests/store_test.rs:38-51 — nearest first: x, z, y
    st, _ = make_store(np.array([[1, 0, 0], [0, 1, 0], [0.7, 0.7, 0]], np.float32), default_config(dimensions=3),
                       attrs=False, graph_seed=1)
    st.set_attrs(np.full(3, NOW[0], np.uint64), np.full(3, NOW[1], np.uint32))
    st.upload(0)
    ids, sc_ = st.search(np.array([[1.0, 0.1, 0.0]], np.float32), 3, now=NOW)
    assert list(ids[0]) == [0, 2, 1] and np.all(np.diff(sc_[0]) >= 0)
    # :66-92 — identical vectors, week-old vs fresh at w = 0.5: fresh first
    cfg = default_config(dimensions=2, temporal_weight=0.5)
    st = ChronoMindB200.from_matrix(np.array([[1, 0], [1, 0]], np.float32), cfg,
                                    np.array([NOW[0] - 7 * 24 * 3600, NOW[0]], np.uint64),
                                    np.array([NOW[1], NOW[1]], np.uint32)).build_graph(1).upload(0)
    ids, sc_ = st.search(np.array([[1.0, 0.0]], np.float32), 2, now=NOW)
    assert list(ids[0]) == [1, 0] and sc_[0][0] == 0.0
    # :95-117 — w = 0: purely by distance
    cfg = default_config(dimensions=2, temporal_weight=0.0)
    st = ChronoMindB200.from_matrix(np.array([[1, 0], [0, 1]], np.float32), cfg,
                                    np.array([NOW[0] - 30 * 24 * 3600, NOW[0]], np.uint64),
                                    np.zeros(2, np.uint32)).build_graph(1).upload(0)
    ids, _ = st.search(np.array([[1.0, 0.0]], np.float32), 2, now=NOW)
    assert ids[0][0] == 0
    # :54-63 — k truncates
    vs = np.array([[np.cos(i * 0.1), np.sin(i * 0.1)] for i in range(10)], np.float32)
    st, _ = make_store(vs, default_config(dimensions=2), graph_seed=1)
    assert st.search(np.array([[1.0, 0.0]], np.float32), 3, now=NOW)[0].shape == (1, 3)


def test_invalid_queries_return_typed_errors():           # tests/store_test.rs:135-164
    st, _ = make_store(np.eye(3, dtype=np.float32), default_config(dimensions=3), graph_seed=1)
    with pytest.raises(CmError) as e:
        st.search(np.array([[1.0, 0.0]], np.float32), 1, now=NOW)
    assert e.value.kind == "InvalidDimensions"
    for bad in (np.nan, np.inf):
        with pytest.raises(CmError) as e:
            st.search(np.array([[1.0, 0.0, 0.0], [1.0, bad, 0.0]], np.float32), 1, now=NOW)
        assert e.value.kind == "InvalidVector" and "row 1" in str(e.value)


def test_snapshot_to_search_end_to_end(tmp_path):         # tests/persistence_test.rs:48-56, tests/cli_test.rs:39-55
    mems = [sc.Memory(id="m%d" % i, data=[float(i), i + 1.0, i + 2.0, i + 3.0], ts_secs=NOW[0] - 60, importance=0.5,
                      last_access_secs=NOW[0]) for i in range(20)]
    p = str(tmp_path / "t.chrono")
    sc.write_snapshot(p, sc.Config(dimensions=4), mems)
    st = ChronoMindB200.from_snapshot(p).build_graph(3).upload(0)
    ids, _ = st.search(np.array([[0.0, 1.0, 2.0, 3.0]], np.float32), 1, now=NOW)
    assert st.external_id(int(ids[0][0])) == "m0"
    sample = [("vector1", [0.1, 0.2, 0.3, 0.4], 0.1), ("vector2", [0.2, 0.3, 0.4, 0.5], 0.2),
              ("vector3", [0.3, 0.4, 0.5, 0.6], 0.15), ("vector4", [0.4, 0.5, 0.6, 0.7], 0.05),
              ("vector5", [0.5, 0.6, 0.7, 0.8], 0.25)]                          # examples/sample_vectors.json
    sc.write_snapshot(p, sc.Config(dimensions=4), [sc.Memory(id=i, data=d, ts_secs=NOW[0], decay_rate=r)
                                                   for i, d, r in sample])
    st = ChronoMindB200.from_snapshot(p).build_graph(3).upload(0)
    ids, _ = st.search(np.array([[0.1, 0.2, 0.3, 0.4]], np.float32), 3, now=NOW)
    assert "vector1" in [st.external_id(int(h)) for h in ids[0]]


# ---- shard merge ---------------------------------------------------------------------------------------------------------
def test_shard_merge_equals_global_scan():
    """src/index/sharded_rwlock.rs:81-94 with G shards by id (global = local * G + shard): per-shard exact top-k,
    device merge == the single-index scan."""
    import torch
    from chronomind_b200.engine import merge_topk_cuda
    rng = np.random.default_rng(31)
    data = datasets.unit_vectors(rng, 4000, 64)
    q = datasets.unit_vectors(rng, 40, 64)
    G, k = 4, 10
    ids = torch.empty((G, 40, k), dtype=torch.int32, device="cuda")
    ds = torch.empty((G, 40, k), dtype=torch.float32, device="cuda")
    tq = torch.from_numpy(q).cuda()
    for s in range(G):
        st, _ = make_store(data[s::G], attrs=False)
        st.set_shard(s, G)
        st.search_exact_cuda(tq, k, ids[s], ds[s])
        torch.cuda.synchronize()
    mi, md = merge_topk_cuda(ids, ds, k)
    torch.cuda.synchronize()
    whole, _ = make_store(data, attrs=False)
    gi, gd = whole.search_exact(q, k)
    assert np.array_equal(mi.cpu().numpy().view(np.uint32), gi)
    assert np.array_equal(bits(md.cpu().numpy()), bits(gd))
    oi, od = oracle.sharded_merge(ids.cpu().numpy().view(np.uint32), ds.cpu().numpy(), k)
    assert np.array_equal(oi, gi) and np.array_equal(bits(od), bits(gd))
    # the packed form (one exchanged array per step): same result, padding included
    from chronomind_b200.engine import merge_topk_keys_cuda, pack_topk_cuda
    ids[1, 3, 7:] = -1                                     # a short list: CM_NO_ID padding must survive the packing
    ki, kd = merge_topk_keys_cuda(pack_topk_cuda(ids, ds), k)
    mi2, md2 = merge_topk_cuda(ids, ds, k)
    torch.cuda.synchronize()
    assert torch.equal(ki, mi2) and np.array_equal(bits(kd.cpu().numpy()), bits(md2.cpu().numpy()))


def test_sharded_store_search_is_merge_then_rerank():
    """with the corpus sharded by id, `ChronoMind::search` = per-shard `VectorIndex::search(q, ef)` ->
    merge to the GLOBAL top-ef (src/index/sharded_rwlock.rs:81-94) -> rerank (src/store.rs:327-344) with attributes
    indexed by global handle. Three shards emulated on one GPU; the oracle does the same three steps on the CPU."""
    import torch
    from chronomind_b200.engine import merge_topk_cuda, rerank_cuda
    data, q = datasets.embedding_dataset(3000, 40, 0xD1CE)
    ts_s, ts_n, dec = datasets.temporal_attrs(3000, 3, NOW[0])
    G, ef, k, w, base = 3, 64, 10, 0.3, 0.1
    cfg = default_config(dimensions=768, ef_construction=100)
    tq = torch.from_numpy(q).cuda()
    pool_i = torch.empty((G, len(q), ef), dtype=torch.int32, device="cuda")
    pool_d = torch.empty((G, len(q), ef), dtype=torch.float32, device="cuda")
    want_i, want_d = [], []
    for s in range(G):
        st = ChronoMindB200.from_matrix(data[s::G], cfg).build_graph(0xD1CE + s, 1).set_shard(s, G).upload(0)
        st.search_index_cuda(tq, ef, ef, pool_i[s], pool_d[s])
        torch.cuda.synchronize()
        orc = oracle.OracleHnsw(16, 100, 50, 0xD1CE + s)
        orc.insert_matrix(data[s::G])
        oi, od, _, _ = orc.search_batch(q, ef, ef)
        want_i.append(oi)
        want_d.append(od)
    mi, md = merge_topk_cuda(pool_i, pool_d, ef)
    ri, rs, rd = rerank_cuda(mi, md, k, torch.from_numpy(ts_s.view(np.int64)).cuda(), torch.from_numpy(ts_n.view(np.int32)).cuda(),
                             torch.from_numpy(dec).cuda(), w, base, NOW)
    torch.cuda.synchronize()
    gi, gd = oracle.sharded_merge(np.stack(want_i), np.stack(want_d), ef)
    assert np.array_equal(mi.cpu().numpy().view(np.uint32), gi) and np.array_equal(bits(md.cpu().numpy()), bits(gd))
    wi, ws = oracle.rerank_pool(gi, gd, k, w, base, ts_s, ts_n, dec, NOW[0], NOW[1])
    got_i, got_s = ri.cpu().numpy().view(np.uint32), rs.cpu().numpy()
    assert assert_order_matches_up_to_ties(got_i, wi, got_s, ws) <= 0.001 * wi.size
    assert np.all(np.diff(got_s, axis=1) >= 0)


def test_gpu_assisted_build_is_sound_searchable_and_oracle_identical(tmp_path):
    """candidates from the device's all-pairs kNN, select/link on the host. Same layer draw and entry
    as the host builder; passes check_invariants; recall@10 at ef=64 no worse than the host-built graph's minus 0.01;
    and the traversal of THAT graph is still the oracle's traversal (ids, distance bits, work counters)."""
    data, q = datasets.embedding_dataset(30000, 256, 0xB111D)
    cfg = default_config(dimensions=768, ef_construction=100)
    host = ChronoMindB200.from_matrix(data, cfg).build_graph(0xB111D, 8).upload(0)
    gpu = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(0xB111D, 0, 8).upload(0)
    assert gpu.check_invariants() == 0
    assert gpu.entry() == host.entry()
    assert [gpu.top_layer(i) for i in range(0, 30000, 97)] == [host.top_layer(i) for i in range(0, 30000, 97)]
    truth, _ = host.search_exact(q, 10)

    def recall(st):
        ids, _ = st.search_index(q, 64, 10)
        return float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ids.tolist(), truth.tolist())]))
    r_host, r_gpu = recall(host), recall(gpu)
    assert r_gpu >= 0.95 and r_gpu >= r_host - 0.01, (r_host, r_gpu)
    p = str(tmp_path / "g.cmg")
    gpu.save_graph(p)
    orc = oracle.OracleHnsw(16, 100, 50, 0xB111D)
    orc.import_graph(oracle.preprocess_rows(data, threads=8), oracle.read_graph_sidecar(p))
    assert orc.check_invariants() == 0
    oi, od, _, octr = orc.search_batch(q, 64, 10, threads=8)
    gi, gd = gpu.search_index(q, 64, 10)
    c = last_counters()
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1])


def test_concurrent_callers_share_one_index():
    """`&self` + `Send + Sync` (src/index/mod.rs:59-81): every cm_search_* is re-entrant on a shared uploaded index —
    each in-flight call leases its own workspace and stream. Eight host threads mix exact (both engines), HNSW and
    temporal searches; every result must equal the single-threaded one."""
    import threading
    data, q = datasets.embedding_dataset(12000, 96, 0xC0C0)
    ts_s, ts_n, dec = datasets.temporal_attrs(12000, 2, NOW[0])
    cfg = default_config(dimensions=768, ef_construction=100)
    st = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec).build_graph(7, 8).upload(0)
    want = {"exact": st.search_exact(q, 10), "hnsw": st.search_index(q, 64, 10), "temporal": st.search(q, 5, now=NOW)}
    errors = []

    def worker(kind, reps):
        try:
            for _ in range(reps):
                got = {"exact": lambda: st.search_exact(q, 10), "hnsw": lambda: st.search_index(q, 64, 10),
                       "temporal": lambda: st.search(q, 5, now=NOW)}[kind]()
                for a, b in zip(got, want[kind]):
                    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                        errors.append(kind)
        except Exception as e:                           # noqa: BLE001
            errors.append("%s: %r" % (kind, e))

    threads = [threading.Thread(target=worker, args=(k, 6)) for k in ("exact", "hnsw", "temporal", "exact", "hnsw", "temporal", "exact", "hnsw")]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_graph_sidecar_load_upload_search(tmp_path):
    """a graph saved next to a snapshot is loaded into a FRESH index (no build), uploaded and searched:
    results equal the index it was saved from and the oracle over the same sidecar."""
    data, q = datasets.embedding_dataset(8000, 100, 0x51DE, dim=256, intrinsic=12)
    cfg = default_config(dimensions=256, ef_construction=100)
    built = ChronoMindB200.from_matrix(data, cfg).build_graph(0x51DE, 8).upload(0)
    p = str(tmp_path / "graph.cmg")
    built.save_graph(p)
    fresh = ChronoMindB200.from_matrix(data, cfg).load_graph(p).upload(0)
    assert fresh.check_invariants() == 0 and fresh.entry() == built.entry()
    for h in range(0, 8000, 7):
        for layer in range(built.top_layer(h) + 1):
            assert np.array_equal(fresh.neighbors(h, layer), built.neighbors(h, layer))
    orc = oracle.OracleHnsw(16, 100, 50, 0x51DE)
    orc.import_graph(oracle.preprocess_rows(data), oracle.read_graph_sidecar(p))
    for ef in (50, 128):
        gi, gd = fresh.search_index(q, ef, ef)
        c = last_counters()
        bi, bd = built.search_index(q, ef, ef)
        oi, od, _, octr = orc.search_batch(q, ef, ef, threads=8)
        assert np.array_equal(gi, bi) and np.array_equal(bits(gd), bits(bd))
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
        assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1])
    # a sidecar of another index is rejected, not searched
    other = ChronoMindB200.from_matrix(data[:4000], cfg)
    with pytest.raises(CmError) as e:
        other.load_graph(p)
    assert e.value.kind == "InvalidSnapshot"


def test_hnsw_1m_graph_identical_to_oracle(tmp_path):
    """BASELINE configs[2] at full size: 1M x 768 embedding-like rows, GPU-assisted graph -> sidecar -> oracle import;
    256 queries at ef = 64 and 200 must come back with the oracle's ids, distance bits and work counters, and the
    asynchronous status word must stay clean."""
    gen = datasets.ChunkedCorpus("emb", 1_000_000, 768, 0x5EED)
    data = gen.shard(0, 1)
    q = gen.queries(256)
    cfg = default_config(dimensions=768, ef_construction=100)
    st = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(0x5EED, 0, 0).upload(0)
    assert st.check_invariants() == 0
    p = str(tmp_path / "g1m.cmg")
    st.save_graph(p)
    orc = oracle.OracleHnsw(16, 100, 50, 0x5EED)
    orc.import_graph(oracle.preprocess_rows(data, threads=16), oracle.read_graph_sidecar(p))
    del data
    for ef in (64, 200):
        gi, gd = st.search_index(q, ef, 10)
        c = last_counters()
        oi, od, _, octr = orc.search_batch(q, ef, 10, threads=16)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od)), ef
        assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1]) and c["list_entries"] == int(octr[2])
        assert c["failures"] == 0
    assert st.dev_status() == 0


def test_remove_then_sync_deleted_reaches_every_engine():
    """`VectorIndex::remove` is O(1) in the reference (src/index/lockfree_hnsw.rs:451-467). Here: remove on the host,
    `cm_index_sync_deleted` pushes a few bytes per handle, and the fp32 scan, the tensor-core scan, the HNSW search
    and the store search all drop the handles — identical to the oracle with the same tombstones, and to a full
    re-upload."""
    data, q = datasets.embedding_dataset(12000, 120, 0xDE1)
    ts_s, ts_n, dec = datasets.temporal_attrs(12000, 9, NOW[0])
    cfg = default_config(dimensions=768, ef_construction=100, ef_search=64)
    st = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec).build_graph(0xDE1, 1).upload(0)
    orc = _oracle_for(data, 16, 100, 0xDE1)
    first, _ = st.search_exact(q, 10)
    victims = np.unique(first[:, :3].ravel())                 # the best hits of every query
    dele = np.zeros(12000, np.uint8)
    dele[victims] = 1
    for h in victims:
        assert st.remove(int(h)) and not st.remove(int(h))    # the second call finds it gone (:455-466)
        orc.remove(int(h))
    assert len(st) == 12000 - len(victims)
    before, _ = st.search_exact(q, 10)
    assert np.array_equal(before, first)                      # the replica has not been told yet
    st.sync_deleted()
    prepared = oracle.preprocess_rows(data)
    wi, wd = oracle.brute_force(prepared, q, 10, mode="prepared", deleted=dele, threads=8)
    for engine_id in (1, 2):
        gi, gd = st.set_exact_engine(engine_id).search_exact(q, 10)
        assert np.array_equal(gi, wi) and np.array_equal(bits(gd), bits(wd)), engine_id
    st.set_exact_engine(0)
    hi, hd = st.search_index(q, 64, 10)
    oi, od, _, _ = orc.search_batch(q, 64, 10)
    assert np.array_equal(hi, oi) and np.array_equal(bits(hd), bits(od))
    assert not np.isin(hi, victims).any()
    ti, tsc = st.search(q, 5, now=NOW)
    rc, ri, rs, _ = orc.store_search_batch(q, 5, 768, 64, cfg.temporal_weight, cfg.base_decay_rate, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert rc == 0 and assert_order_matches_up_to_ties(ti, ri, tsc, rs) <= 0.001 * ti.size
    again = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec, dele).build_graph(0xDE1, 1).upload(0)
    ai, ad = again.search_exact(q, 10)
    assert np.array_equal(ai, wi) and np.array_equal(bits(ad), bits(wd))


def test_gpu_assisted_build_on_clustered_data_is_navigable():
    """Tight, well separated clusters are where a nearest-neighbor graph falls apart into islands (every candidate of a
    row lies in its own cluster). The GPU-assisted build takes PREDECESSOR neighbors like `LockFreeHnsw::insert` does, so
    early rows link across clusters: every node must be reachable and queries at the cluster centres must find their
    cluster — recall no worse than the host replay of the reference's insertion minus 0.02."""
    rng = np.random.default_rng(0xC1)
    centres = rng.normal(size=(200, 48)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    data = (np.repeat(centres, 50, axis=0) + 0.02 * rng.normal(size=(10000, 48))).astype(np.float32)
    data = data[rng.permutation(len(data))]
    q = (centres + 0.01 * rng.normal(size=centres.shape)).astype(np.float32)
    cfg = default_config(dimensions=48, ef_construction=64)
    host = ChronoMindB200.from_matrix(data, cfg).build_graph(9, 1).upload(0)
    gpu = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(9, 0, 8, 24).upload(0)   # 24 candidates < a cluster's 50 rows
    assert gpu.check_invariants() == 0 and gpu.unreachable() == 0
    truth, _ = host.search_exact(q, 10)

    def recall(st, ef):
        ids, _ = st.search_index(q, ef, 10)
        return float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ids.tolist(), truth.tolist())]))
    for ef in (32, 64):
        r_host, r_gpu = recall(host, ef), recall(gpu, ef)
        assert r_gpu >= 0.95 and r_gpu >= r_host - 0.02, (ef, r_host, r_gpu)


def test_async_calls_on_different_streams_do_not_share_scratch():
    """The `_dev` entry points return while their kernels still run; the workspace they used carries a completion event,
    so a call on ANOTHER stream (or thread) either gets a different workspace or waits for that event on the device.
    Interleave exact, HNSW and temporal searches over three streams without synchronising in between: every result must
    equal the one computed alone, and the status word must stay clean."""
    import torch
    data, q = datasets.embedding_dataset(16000, 3000, 0x57E)
    ts_s, ts_n, dec = datasets.temporal_attrs(16000, 2, NOW[0])
    cfg = default_config(dimensions=768, ef_construction=100)
    st = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec).build_graph(3, 8).upload(0)
    tq = [torch.from_numpy(q[i * 1000:(i + 1) * 1000]).cuda() for i in range(3)]
    want = []
    for t in tq:
        e = st.search_exact_cuda(t, 10)
        h = st.search_index_cuda(t, 64, 10)
        s = st.search_cuda(t, 5, NOW, ef_search=64)
        torch.cuda.synchronize()
        want.append([x.clone() for x in (*e, *h, *s)])
    streams = [torch.cuda.Stream() for _ in range(3)]
    for rep in range(4):
        got = []
        for i, (t, sm) in enumerate(zip(tq, streams)):          # enqueue everything before waiting for anything
            with torch.cuda.stream(sm):
                e = st.search_exact_cuda(t, 10)
                h = st.search_index_cuda(t, 64, 10)
                s = st.search_cuda(t, 5, NOW, ef_search=64)
            got.append((*e, *h, *s))
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32)), rep
    assert st.dev_status() == 0
    c = st.dev_counters()
    assert c["fallback_queries"] == 0 and c["rescored"] >= 10 * 1000 * 3


@pytest.mark.parametrize("n,dim,M,kind", [(30000, 768, 16, "emb"), (20000, 64, 8, "uni"), (12000, 100, 16, "clustered")])
def test_device_select_builds_the_same_graph_as_the_host_select(n, dim, M, kind):
    """SURVEY.md §8f N2 on the device: `select_diverse` and the reverse-edge re-selection (src/index/lockfree_hnsw.rs:
    205-280) run as one kernel over (center, candidates) items with the reference's exact fp32 arithmetic. From the same
    predecessor candidates the device must produce BIT-FOR-BIT the lists the host code produces: every list of every
    layer of every node, the entry point and the reachability verdict."""
    import os
    if kind == "emb":
        data, _ = datasets.embedding_dataset(n, 1, 0x5E1)
    elif kind == "uni":
        data, _ = datasets.uniform_dataset(n, 1, dim, 0x5E1)
        data[::8] = data[3]                                       # 2,500 copies of one row: targets with more incoming edges
    else:                                                         # than the kernel sorts go to the host (too_big)
        rng = np.random.default_rng(5)
        centres = rng.normal(size=(150, dim)).astype(np.float32)
        data = (np.repeat(centres, n // 150, axis=0) + 0.05 * rng.normal(size=(n // 150 * 150, dim))).astype(np.float32)
        data = data[rng.permutation(len(data))]
        data[::50] = data[7]                                      # a hub: hundreds of rows identical to one row
    cfg = default_config(dimensions=dim, max_connections=M, ef_construction=64)
    os.environ["CM_BUILD_SELECT"] = "host"
    try:
        host = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(77, 0, 8)
    finally:
        del os.environ["CM_BUILD_SELECT"]
    dev = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(77, 0, 8)
    assert dev.check_invariants() == 0 and host.check_invariants() == 0
    assert dev.entry() == host.entry() and dev.unreachable() == host.unreachable() == 0
    for h in range(len(data)):
        t = host.top_layer(h)
        assert dev.top_layer(h) == t
        for layer in range(t + 1):
            a, b = dev.neighbors(h, layer), host.neighbors(h, layer)
            assert np.array_equal(a, b), (h, layer, a, b)
