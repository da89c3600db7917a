"""BASELINE configs[2] says "from reference snapshot": a synthetic store written as a CHRONO1 v2 file
(`save_snapshot`, src/persistence.rs:41-64) reaches the engine only through `cm_index_from_snapshot`
(`load_snapshot`, :73-123) — 60,000 x 768 records (185 MB) — and every read-path entry point must then answer like
the oracle over the same records: exact scan, HNSW over the GPU-built graph (ids, distance bits, work counters),
`ChronoMind::search` with the file's timestamps and decay rates, `search_in_context` with the file's contexts."""
import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, last_counters

from parity_helpers import assert_order_matches_up_to_ties

pytestmark = pytest.mark.gpu
NOW = (1_790_000_000, 0)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_snapshot_file_to_every_search_entry_point(tmp_path):
    n, dim, n_ctx = 60_000, 768, 12
    gen = datasets.ChunkedCorpus("emb", n, dim, 0x5A7)
    rng = np.random.default_rng(11)
    rows = gen.shard(0, 1) * rng.uniform(0.5, 3.0, size=(n, 1)).astype(np.float32)      # stored rows are not unit length
    q = gen.queries(200) * 1.5
    ts_s, ts_n, dec = datasets.temporal_attrs(n, 4, NOW[0])
    ctx = rng.integers(0, n_ctx, size=n).astype(np.uint32)
    p = str(tmp_path / "store.chrono")
    info = datasets.write_snapshot(p, rows, ts_s, ts_n, dec, ctx, n_ctx, config={"ef_construction": 100, "ef_search": 64})
    assert info["records"] == n and info["bytes"] > n * dim * 4

    st = ChronoMindB200.from_snapshot(p)
    assert len(st) == n and st.config.dimensions == dim and st.external_id(n - 1) == "m%09d" % (n - 1)
    st.build_graph_gpu(0x5A7, 0, 0).upload(0)
    assert st.check_invariants() == 0

    prepared = oracle.preprocess_rows(rows, threads=8)
    wi, wd = oracle.brute_force(prepared, q, 10, mode="prepared", threads=8)
    for engine_id in (1, 2):
        gi, gd = st.set_exact_engine(engine_id).search_exact(q, 10)
        assert np.array_equal(gi, wi) and np.array_equal(bits(gd), bits(wd)), engine_id
    st.set_exact_engine(0)

    g = str(tmp_path / "store.cmg")
    st.save_graph(g)
    orc = oracle.OracleHnsw(16, 100, 64, 0x5A7)
    orc.import_graph(prepared, oracle.read_graph_sidecar(g))
    hi, hd = st.search_index(q, 64, 10)
    c = last_counters()
    oi, od, _, octr = orc.search_batch(q, 64, 10, threads=8)
    assert np.array_equal(hi, oi) and np.array_equal(bits(hd), bits(od))
    assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1]) and c["failures"] == 0

    ti, tsc = st.search(q, 5, now=NOW)                                                      # config from the file: w 0.3, base 0.1, ef 64
    rc, si, ss, _ = orc.store_search_batch(q, 5, dim, 64, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert rc == 0 and assert_order_matches_up_to_ties(ti, si, tsc, ss) <= 0.001 * ti.size

    want = rng.integers(0, n_ctx, size=len(q)).astype(np.uint32)
    labels = ["ctx%05d" % c_ for c_ in want]
    assert [st.context_id(l) for l in labels[:5]] == [st.context_id("ctx%05d" % c_) for c_ in want[:5]]
    want_ids = np.array([st.context_id(l) for l in labels], dtype=np.uint32)              # ids in order of first appearance in the file
    ci, cs = st.search_in_context(want_ids, q, 10, now=NOW)
    rc, xi, xs = oracle.search_in_context(rows, ctx, q, want, 10, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1], threads=8)
    assert rc == 0 and assert_order_matches_up_to_ties(ci, xi, cs, xs) <= 0.001 * ci.size
    live = ci != NO_ID
    assert np.all(ctx[ci[live]] == np.broadcast_to(want[:, None], ci.shape)[live])
    st.close()
