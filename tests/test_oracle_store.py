"""Pins the oracle's store-level rerank (`src/store.rs:321-346,395-415`) to the reference's
`tests/store_test.rs` orderings, and the snapshot codec to `tests/persistence_test.rs`."""
import math
import struct

import numpy as np
import pytest

import oracle
from oracle import snapshot_codec as sc

NOW = (1_790_000_000, 123_456_789)


def _store(vectors, dims, ts=None, decay=None, w=0.3, base=0.1, ef_search=50, seed=1):
    idx = oracle.OracleHnsw(seed=seed, ef_search=ef_search)
    for v in vectors:
        idx.insert(v)
    n = len(vectors)
    ts_secs = np.array([t[0] for t in ts] if ts else [NOW[0]] * n, dtype=np.uint64)
    ts_nanos = np.array([t[1] for t in ts] if ts else [NOW[1]] * n, dtype=np.uint32)
    dec = np.array(decay if decay else [0.0] * n, dtype=np.float32)

    def search(q, k):
        rc, ids, scores, d = idx.store_search_batch(np.array([q], np.float32), k, dims, ef_search, w, base,
                                                   ts_secs, ts_nanos, dec, NOW[0], NOW[1])
        return rc, ids[0], scores[0], d[0]
    return search


def test_search_returns_nearest_first():                   # tests/store_test.rs:38-51
    s = _store([[1, 0, 0], [0, 1, 0], [0.7, 0.7, 0]], 3)   # x, y, z
    rc, ids, scores, _ = s([1.0, 0.1, 0.0], 3)
    assert rc == 0 and list(ids) == [0, 2, 1]               # x, z, y
    assert scores[0] <= scores[1] <= scores[2]


def test_search_k_truncates():                             # :54-63
    vs = [[math.cos(i * 0.1), math.sin(i * 0.1)] for i in range(10)]
    rc, ids, _, _ = _store(vs, 2)([1.0, 0.0], 3)
    assert np.sum(ids != 0xFFFFFFFF) == 3


def test_temporal_weight_prefers_recent_memories():        # :66-92
    old = (NOW[0] - 7 * 24 * 3600, NOW[1])
    s = _store([[1, 0], [1, 0]], 2, ts=[old, NOW], w=0.5)
    rc, ids, scores, _ = s([1.0, 0.0], 2)
    assert list(ids) == [1, 0]                               # fresh, old
    assert scores[0] == 0.0
    assert abs(scores[1] - 0.5 * (1 - math.exp(-0.1 * 168))) < 1e-6


def test_zero_temporal_weight_ranks_purely_by_distance():  # :95-117
    old = (NOW[0] - 30 * 24 * 3600, 0)
    s = _store([[1, 0], [0, 1]], 2, ts=[old, NOW], w=0.0)
    rc, ids, _, _ = s([1.0, 0.0], 2)
    assert ids[0] == 0                                       # old_near


def test_invalid_queries_are_typed_errors():               # :135-164, store.rs:379-392
    s = _store([[1, 0, 0]], 3)
    assert s([1.0, 0.0], 1)[0] == 1                          # InvalidDimensions
    assert s([1.0, float("nan"), 0.0], 1)[0] == 2            # InvalidVector
    assert s([1.0, float("inf"), 0.0], 1)[0] == 2


def test_combined_score_formula_and_rounding():            # store.rs:395-415
    f = np.float32
    # decay numbers of tests/store_test.rs:241-273: 24 h at rate 0.1 -> exp(-2.4)
    ts = (NOW[0] - 24 * 3600, NOW[1])
    got = oracle.combined_score(0.5, ts[0], ts[1], 0.0, NOW[0], NOW[1], 1.0, 0.1)
    assert abs(got - (1 - math.exp(-2.4))) < 1e-6
    # per-memory rate overrides base when > 0
    got = oracle.combined_score(0.5, ts[0], ts[1], 0.25, NOW[0], NOW[1], 1.0, 0.1)
    assert abs(got - (1 - math.exp(-6.0))) < 1e-6
    # timestamp in the future -> age 0 (duration_since(..).unwrap_or_default())
    assert oracle.combined_score(0.4, NOW[0] + 10, 0, 0.0, NOW[0], NOW[1], 0.3, 0.1) == float(f(0.7) * f(0.2))
    # nanosecond borrow: Duration::as_secs_f32 = secs as f32 + nanos as f32 / 1e9
    got = oracle.combined_score(1.0, 100, 900_000_000, 0.0, 103, 100_000_000, 0.3, 0.1)
    age = f(f(2.0) + f(200_000_000) / f(1e9)) / f(3600.0)
    want = f(f(1.0) - f(0.3)) * f(0.5) + f(0.3) * (f(1.0) - f(np.exp(np.float64(-(f(0.1) * age)))))
    assert abs(got - float(want)) <= 1e-9


def test_rerank_is_stable_on_score_ties():                 # store.rs:338 (stable sort_by total_cmp)
    ids = np.array([[5, 3, 9, 1]], dtype=np.uint32)
    d = np.array([[0.1, 0.1, 0.2, 0.2]], dtype=np.float32)
    n = 10
    ts_s = np.full(n, NOW[0], np.uint64)
    ts_n = np.full(n, NOW[1], np.uint32)
    oi, osc = oracle.rerank_pool(ids, d, 4, 0.0, 0.1, ts_s, ts_n, np.zeros(n, np.float32), NOW[0], NOW[1])
    assert list(oi[0]) == [5, 3, 9, 1]                       # ties keep pool order


# ---- snapshot container (src/persistence.rs) -----------------------------------------------------
def _sample():                                              # tests/persistence_test.rs:8-29
    cfg = sc.Config(dimensions=4)
    mems = [sc.Memory(id="m%d" % i, data=[float(i), i + 1.0, i + 2.0, i + 3.0], ts_secs=1_700_000_000 + i,
                      ts_nanos=i, importance=i / 20.0, context="even" if i % 2 == 0 else "odd",
                      last_access_secs=1_700_000_000 + i, last_access_nanos=i) for i in range(20)]
    return cfg, mems


def test_snapshot_roundtrip_preserves_everything():        # :31-45
    cfg, mems = _sample()
    mems[3].relationships = ["m1", "m2"]
    mems[3].decay_rate = 0.25
    blob = sc.encode_snapshot(cfg, mems)
    cfg2, mems2 = sc.decode_snapshot(blob)
    assert sc.encode_body(cfg2, []) == sc.encode_body(cfg, [])   # f32 fields compare as bits
    assert (cfg2.dimensions, cfg2.index) == (cfg.dimensions, cfg.index)
    for a, b in zip(mems, mems2):
        assert a.id == b.id and a.data == b.data and a.context == b.context
        assert (a.ts_secs, a.ts_nanos, a.relationships) == (b.ts_secs, b.ts_nanos, b.relationships)
        assert struct.pack("<f", a.importance) == struct.pack("<f", b.importance)


def test_snapshot_layout_known_bytes():
    """Byte layout derived from bincode 1.3.3 defaults: config = 60 bytes."""
    cfg = sc.Config(dimensions=2)
    blob = sc.encode_snapshot(cfg, [sc.Memory(id="a", data=[1.0, 2.0])])
    assert blob[:8] == b"CHRONO1\x02"
    body = blob[12:]
    assert len(sc.encode_body(cfg, [])) == 60 + 8
    assert body[:8] == struct.pack("<Q", 2)
    assert body[60:68] == struct.pack("<Q", 1)
    assert body[68:77] == struct.pack("<Q", 1) + b"a"
    assert body[77:85] == struct.pack("<Q", 2) and body[85:93] == struct.pack("<ff", 1.0, 2.0)


def test_bad_inputs_are_rejected():                        # :59-113
    cfg, mems = _sample()
    with pytest.raises(sc.InvalidSnapshot):
        sc.decode_snapshot(b"definitely not a chronomind snapshot")
    with pytest.raises(sc.InvalidSnapshot):
        sc.decode_snapshot(b"CHR")
    with pytest.raises(sc.InvalidSnapshot, match="99"):
        sc.decode_snapshot(b"CHRONO1" + bytes([99]))
    blob = bytearray(sc.encode_snapshot(cfg, mems))
    blob[len(blob) // 2] ^= 0xFF
    with pytest.raises(sc.InvalidSnapshot, match="checksum"):
        sc.decode_snapshot(bytes(blob))


def test_loaded_store_is_searchable():                     # :48-56
    cfg, mems = _sample()
    _, mems2 = sc.decode_snapshot(sc.encode_snapshot(cfg, mems))
    s = _store([m.data for m in mems2], 4)
    rc, ids, _, _ = s([0.0, 1.0, 2.0, 3.0], 1)
    assert mems2[ids[0]].id == "m0"


def test_decay_sweep_matches_the_documented_curve():       # tests/store_test.rs:241-273 (exp(-2.4) within 0.01)
    h = 3600 * 10**9
    imp, dt = oracle.decay_sweep([1.0, 1.0, 0.5], [0, 0, 0], [0, 0, 30 * h], [0.0, 0.2, 0.1], 0.1, 24 * h)
    assert abs(imp[0] - np.exp(-2.4)) < 0.01 and abs(imp[1] - np.exp(-4.8)) < 0.01
    assert imp[2] == 0.5 and dt[2] == 0 and dt[0] == 24 * h          # accessed after `now`: interval not claimed
    imp2, dt2 = oracle.decay_sweep(imp, dt, [0, 0, 30 * h], [0.0, 0.2, 0.1], 0.1, 24 * h)
    assert np.array_equal(imp2, imp)                                  # the same interval is never applied twice


def test_similar_pairs_is_the_consolidate_pair_test():     # src/store.rs:531-534
    rng = np.random.default_rng(4)
    base = rng.normal(size=(40, 16)).astype(np.float32)
    raw = np.vstack([base, base[:5] * 3.0 + 1e-3 * rng.normal(size=(5, 16)).astype(np.float32)])   # 5 near-duplicates
    i, j, sim = oracle.similar_pairs(raw, 0.95)
    assert list(zip(i.tolist(), j.tolist())) == [(a, 40 + a) for a in range(5)]
    assert np.all(sim > 0.95) and all(abs(s - oracle.similarity(raw[a], raw[b])) == 0 for a, b, s in zip(i, j, sim))
    dele = np.zeros(45, np.uint8)
    dele[2] = 1
    i, j, _ = oracle.similar_pairs(raw, 0.95, deleted=dele)
    assert (2, 42) not in list(zip(i.tolist(), j.tolist())) and len(i) == 4
