"""CPU-side tests of the product's host runtime (no GPU compute): ABI surface, snapshot reader, graph builder
against the oracle, sidecar round trip, and "no CPU fallback" behaviour."""
import os
import re
import struct

import numpy as np
import pytest

import oracle
from oracle import snapshot_codec as sc
import chronomind_b200 as cb
from chronomind_b200 import datasets
from chronomind_b200.engine import ABI, ChronoMindB200, CmError, default_config, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "chronomind_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = native()
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
    assert declared == set(ABI), "engine.py's ABI table and the header disagree: %s" % (declared ^ set(ABI))


def test_default_config_matches_reference():            # src/config.rs:73-85,27-35
    c = default_config()
    assert (c.dimensions, c.max_memories, c.max_relationships) == (768, 100_000, 50)
    assert (c.index.max_connections, c.index.ef_construction, c.index.ef_search) == (16, 200, 50)
    assert abs(c.base_decay_rate - 0.1) < 1e-7 and abs(c.temporal_weight - 0.3) < 1e-7
    assert native().cm_metric_name() == b"cosine"       # src/metric.rs:201-203


def test_config_validation_errors():                    # src/config.rs:97-143
    x = np.zeros((4, 8), np.float32)
    for kw in ({"max_connections": 1}, {"ef_construction": 4}, {"ef_search": 0}, {"temporal_weight": 1.5},
               {"base_decay_rate": 0.0}, {"max_memories": 0}, {"similarity_threshold": 1.0}):
        with pytest.raises(CmError) as e:
            ChronoMindB200.from_matrix(x, default_config(**kw))
        assert e.value.kind == "Config"


def test_rows_are_preprocessed_like_the_reference():    # src/metric.rs:208-214
    rng = np.random.default_rng(1)
    x = rng.uniform(-50, 50, (37, 13)).astype(np.float32)
    x[5] = 0.0
    st = ChronoMindB200.from_matrix(x, default_config(dimensions=13))
    want = oracle.preprocess_rows(x)
    got = np.stack([st.vector(i) for i in range(37)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert len(st) == 37 and st.slots == 37 and st.dim == 13


def test_search_without_upload_fails_loudly():
    st = ChronoMindB200.from_matrix(np.eye(4, dtype=np.float32), default_config(dimensions=4))
    with pytest.raises(CmError) as e:
        st.search_exact(np.ones((1, 4), np.float32), 1)
    assert e.value.kind == "Cuda"                       # no CPU path behind the ABI


def _graphs_equal(st, orc, n):
    assert st.entry() == orc.entry()
    for i in range(n):
        t = st.top_layer(i)
        assert t == orc.top_layer(i)
        for l in range(t + 1):
            assert np.array_equal(st.neighbors(i, l), orc.neighbors(i, l)), (i, l)


@pytest.mark.parametrize("n,dim,M,efc,seed", [(600, 32, 16, 200, 0xC0FFEE), (400, 6, 4, 16, 7), (300, 768, 16, 100, 5),
                                              (250, 11, 8, 32, 99)])
def test_single_thread_build_equals_oracle_graph(n, dim, M, efc, seed):
    """threads == 1 must reproduce `LockFreeHnsw::insert`'s deterministic graph (lockfree_hnsw.rs:555-573)."""
    rng = np.random.default_rng(seed)
    x = datasets.unit_vectors(rng, n, dim)
    cfg = default_config(dimensions=dim, max_connections=M, ef_construction=efc)
    st = ChronoMindB200.from_matrix(x, cfg).build_graph(seed, threads=1)
    assert st.check_invariants() == 0
    orc = oracle.OracleHnsw(M, efc, 50, seed)
    orc.insert_matrix(x)
    _graphs_equal(st, orc, n)


def test_multi_thread_build_is_sound_and_searchable(tmp_path):
    """Concurrent build: invariants hold and the oracle's CPU search over THAT graph meets the recall gate
    (stress_test.rs:41-175 semantics: invariants + recall >= 0.90 at ef = 100)."""
    data, queries = datasets.embedding_dataset(3000, 50, 0xBEEF, dim=64, intrinsic=8)
    cfg = default_config(dimensions=64, ef_construction=100)
    st = ChronoMindB200.from_matrix(data, cfg).build_graph(11, threads=4)
    assert st.check_invariants() == 0
    p = str(tmp_path / "g.cmgraph")
    st.save_graph(p)
    g = oracle.read_graph_sidecar(p)
    orc = oracle.OracleHnsw(16, 100, 50, 11)
    orc.import_graph(oracle.preprocess_rows(data), g)
    assert orc.check_invariants() == 0
    got, _, _, _ = orc.search_batch(queries, 100, 10)
    exp, _ = oracle.brute_force(data, queries, 10, mode="distance")
    rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(got, exp)])
    assert rec >= 0.90
    # sidecar round trip
    st2 = ChronoMindB200.from_matrix(data, cfg).load_graph(p)
    assert st2.entry() == st.entry()
    assert st2.check_invariants() == 0
    for i in range(3000):                                            # every list of every layer, byte for byte
        assert st2.top_layer(i) == st.top_layer(i)
        for layer in range(st.top_layer(i) + 1):
            assert np.array_equal(st2.neighbors(i, layer), st.neighbors(i, layer)), (i, layer)
    p2 = str(tmp_path / "g2.cmgraph")                                # save(load(save(g))) is the same file
    st2.save_graph(p2)
    assert open(p, "rb").read() == open(p2, "rb").read()
    with pytest.raises(CmError):
        ChronoMindB200.from_matrix(data[:100], cfg).load_graph(p)     # wrong index


def _sample_memories(n=20, dim=4):
    return [sc.Memory(id="m%d" % i, data=[float(i), i + 1.0, i + 2.0, i + 3.0][:dim], ts_secs=1_700_000_000 + i,
                      ts_nanos=5 * i, importance=i / 20.0, context="even" if i % 2 == 0 else "odd",
                      decay_rate=0.05 * (i % 3), relationships=["m0"] if i % 4 == 0 else [],
                      last_access_secs=1_700_000_100 + i) for i in range(n)]


def test_snapshot_reader_matches_codec(tmp_path):       # src/persistence.rs:73-123
    cfg = sc.Config(dimensions=4, temporal_weight=0.5)
    mems = _sample_memories()
    p = str(tmp_path / "s.chrono")
    sc.write_snapshot(p, cfg, mems)
    st = ChronoMindB200.from_snapshot(p)
    assert st.slots == 20 and len(st) == 20 and st.dim == 4
    c = st.config
    assert c.dimensions == 4 and abs(c.temporal_weight - 0.5) < 1e-7 and c.index.ef_construction == 200
    for i, m in enumerate(mems):
        assert st.external_id(i) == m.id
        assert np.array_equal(st.vector(i), oracle.preprocess(m.data))


def test_snapshot_duplicate_id_tombstones_older_handle(tmp_path):   # src/store.rs:252-262
    mems = _sample_memories(6)
    mems.append(sc.Memory(id="m2", data=[9.0, 9.0, 9.0, 1.0]))
    p = str(tmp_path / "dup.chrono")
    sc.write_snapshot(p, sc.Config(dimensions=4), mems)
    st = ChronoMindB200.from_snapshot(p)
    assert st.slots == 7 and len(st) == 6
    assert not st.remove(2)                              # already tombstoned by the replacement
    assert st.remove(6) and not st.remove(6)


def test_snapshot_errors_map_to_reference_variants(tmp_path):       # tests/persistence_test.rs:59-150
    def load(b):
        p = str(tmp_path / "x.bin")
        open(p, "wb").write(b)
        return ChronoMindB200.from_snapshot(p)
    good = sc.encode_snapshot(sc.Config(dimensions=4), _sample_memories())
    for blob, kind, needle in ((b"definitely not a chronomind snapshot", "InvalidSnapshot", "magic"),
                               (b"CHR", "InvalidSnapshot", "short"),
                               (b"CHRONO1" + bytes([99]), "InvalidSnapshot", "99")):
        with pytest.raises(CmError) as e:
            load(blob)
        assert e.value.kind == kind and needle in str(e.value)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0xFF
    with pytest.raises(CmError) as e:
        load(bytes(bad))
    assert e.value.kind == "InvalidSnapshot" and "checksum" in str(e.value)
    with pytest.raises(CmError) as e:
        ChronoMindB200.from_snapshot(str(tmp_path / "does_not_exist.chrono"))
    assert e.value.kind == "Io"
    # truncated body with a fixed-up CRC -> Serialization (bincode error)
    body = good[12:-9]
    import zlib
    with pytest.raises(CmError) as e:
        load(good[:8] + struct.pack("<I", zlib.crc32(body)) + body)
    assert e.value.kind == "Serialization"
    # Config::validate / Memory::validate on load
    with pytest.raises(CmError) as e:
        load(sc.encode_snapshot(sc.Config(dimensions=4, temporal_weight=2.0), _sample_memories()))
    assert e.value.kind == "Config"
    with pytest.raises(CmError) as e:
        load(sc.encode_snapshot(sc.Config(dimensions=5), _sample_memories()))
    assert e.value.kind == "InvalidDimensions"
    m = _sample_memories()
    m[3].data[1] = float("nan")
    with pytest.raises(CmError) as e:
        load(sc.encode_snapshot(sc.Config(dimensions=4), m))
    assert e.value.kind == "InvalidVector"
    m = _sample_memories()
    m[3].importance = 1.5
    with pytest.raises(CmError) as e:
        load(sc.encode_snapshot(sc.Config(dimensions=4), m))
    assert e.value.kind == "InvalidImportance"
    with pytest.raises(CmError) as e:
        load(sc.encode_snapshot(sc.Config(dimensions=4, max_memories=5), _sample_memories()))
    assert e.value.kind == "CapacityExceeded"


def test_snapshot_error_precedence_follows_load_order(tmp_path):
    """`load_snapshot` decodes the WHOLE body first (src/persistence.rs:114), then inserts record by record, each insert
    validating before its capacity check (src/store.rs:226-252): a truncated tail beats an invalid record, and among
    records the first failing one in file order decides."""
    import zlib

    def load(b):
        p = str(tmp_path / "y.bin")
        open(p, "wb").write(b)
        return ChronoMindB200.from_snapshot(p)

    def kind_of(blob):
        with pytest.raises(CmError) as e:
            load(blob)
        return e.value.kind

    m = _sample_memories()
    m[3].data = m[3].data[:3]                            # wrong length in record 3 ...
    blob = sc.encode_snapshot(sc.Config(dimensions=4), m)
    assert kind_of(blob) == "InvalidDimensions"
    body = blob[12:-9]                                   # ... but a truncated tail is found first, while decoding
    assert kind_of(blob[:8] + struct.pack("<I", zlib.crc32(body)) + body) == "Serialization"
    m = _sample_memories()
    m[7].importance = 1.5                                # record 7 invalid, capacity (5) runs out at record 5
    assert kind_of(sc.encode_snapshot(sc.Config(dimensions=4, max_memories=5), m)) == "CapacityExceeded"
    m[2].importance = -0.5                               # an earlier invalid record wins over the capacity error
    assert kind_of(sc.encode_snapshot(sc.Config(dimensions=4, max_memories=5), m)) == "InvalidImportance"
    m = _sample_memories()
    m[1].id = ""                                         # id check precedes the dimension check within a record
    m[1].data = m[1].data[:2]
    assert kind_of(sc.encode_snapshot(sc.Config(dimensions=4), m)) == "InvalidVector"


def test_snapshot_empty_store(tmp_path):
    p = str(tmp_path / "e.chrono")
    sc.write_snapshot(p, sc.Config(dimensions=3), [])
    st = ChronoMindB200.from_snapshot(p)
    assert st.slots == 0 and len(st) == 0


def test_attr_validation():
    st = ChronoMindB200.from_matrix(np.eye(3, dtype=np.float32), default_config(dimensions=3))
    with pytest.raises(CmError):
        st.set_attrs(decay_rate=np.array([0.1, -1.0, 0.0], np.float32))    # src/types.rs:112-118
    st.set_attrs(deleted=np.array([0, 1, 0], np.uint8))
    assert len(st) == 2 and st.slots == 3


def test_graph_from_knn_candidates_is_sound_and_searchable(tmp_path):
    """The host half of the GPU-assisted construction (select_diverse + batched reverse edges over per-layer kNN
    candidates), fed here with the oracle's exact kNN instead of the device's: same layer draw and entry as the
    insertion builder, check_invariants clean, and the CPU oracle searching THAT graph reaches the recall gate."""
    data, q = datasets.embedding_dataset(3000, 40, 0xCAFE, dim=64)
    cfg = default_config(dimensions=64, ef_construction=100)
    seed = 0xCAFE
    ref = ChronoMindB200.from_matrix(data, cfg).build_graph(seed, 1)
    st = ChronoMindB200.from_matrix(data, cfg)
    prepared = oracle.preprocess_rows(data)
    members = st.layer_members(seed)
    assert len(members[0]) == 3000 and all(len(members[l]) >= len(members[l + 1]) for l in range(len(members) - 1))
    cands = []
    for m in members:
        k = min(49, len(m))
        ids, _ = oracle.brute_force(prepared[m], prepared[m], k, mode="prepared", threads=4)   # local ids within the layer
        cands.append(np.where(ids == 0xFFFFFFFF, 0xFFFFFFFF, m[np.minimum(ids, len(m) - 1)]).astype(np.uint32))
    st.build_graph_from_candidates(seed, cands, threads=4)
    assert st.check_invariants() == 0
    assert st.entry() == ref.entry()
    assert [st.top_layer(i) for i in range(0, 3000, 37)] == [ref.top_layer(i) for i in range(0, 3000, 37)]
    assert all(len(st.neighbors(i, 0)) <= 32 for i in range(0, 3000, 11))
    p = str(tmp_path / "g.cmg")
    st.save_graph(p)
    orc = oracle.OracleHnsw(16, 100, 50, seed)
    orc.import_graph(prepared, oracle.read_graph_sidecar(p))
    assert orc.check_invariants() == 0
    got, _, _, _ = orc.search_batch(q, 64, 10, threads=4)
    want, _ = oracle.brute_force(prepared, q, 10, mode="prepared")
    recall = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(got.tolist(), want.tolist())])
    assert recall >= 0.95, recall


def _clustered(n_clusters, per, dim, seed, spread=0.02):
    """Tight, well separated clusters: the kNN of a row never leaves its cluster."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(n_clusters, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    x = np.repeat(centres, per, axis=0) + spread * rng.normal(size=(n_clusters * per, dim)).astype(np.float32)
    perm = rng.permutation(len(x))
    return x[perm].astype(np.float32), centres


def _predecessor_knn(prepared, members, k):
    """Candidates as `LockFreeHnsw::insert` sees them: for member j, its k nearest among members 0..j-1."""
    x = prepared[members]
    sim = x @ x.T
    out = np.full((len(members), k), 0xFFFFFFFF, np.uint32)
    for j in range(1, len(members)):
        kk = min(k, j)
        idx = np.argsort(-sim[j, :j], kind="stable")[:kk]
        out[j, :kk] = members[idx]
    return out


@pytest.mark.parametrize("clusters,per,width,bridging_needed,predecessors",
                         [(24, 120, 33, False, False), (160, 18, 65, False, False), (160, 18, 13, True, False),
                          (160, 18, 13, False, True)])
def test_knn_graph_on_clustered_data_is_bridged_and_searchable(tmp_path, clusters, per, width, bridging_needed, predecessors):
    """Nearest-neighbor candidates alone can leave whole clusters without a way in (the reference's incremental insert
    cannot: every new node is linked into what the entry already reaches). The candidate-based builder must bridge
    them: no layer-0 node unreachable from the entry (the way a search moves), invariants intact, and the oracle
    searching THAT graph meets the recall gate on queries at the cluster centres. The last case feeds fewer candidates
    than a cluster has rows — every cluster without an upper-layer member starts out isolated — and only requires what
    the repair guarantees: reachability. With PREDECESSOR candidates (what the GPU-assisted build feeds) the same starved
    configuration is navigable again: early rows have no near predecessor and link across clusters, as in the reference."""
    data, centres = _clustered(clusters, per, 32, 0xC1)
    cfg = default_config(dimensions=32, ef_construction=64)
    seed = 5
    st = ChronoMindB200.from_matrix(data, cfg)
    prepared = oracle.preprocess_rows(data)
    cands = []
    for m in st.layer_members(seed):
        k = min(width, len(m))
        if predecessors:
            cands.append(_predecessor_knn(prepared, m, k))
            continue
        ids, _ = oracle.brute_force(prepared[m], prepared[m], k, mode="prepared", threads=4)
        cands.append(np.where(ids == 0xFFFFFFFF, 0xFFFFFFFF, m[np.minimum(ids, len(m) - 1)]).astype(np.uint32))
    st.build_graph_from_candidates(seed, cands, threads=4)
    assert st.check_invariants() == 0
    assert st.unreachable() == 0
    p = str(tmp_path / "c.cmg")
    st.save_graph(p)
    orc = oracle.OracleHnsw(16, 64, 50, seed)
    orc.import_graph(prepared, oracle.read_graph_sidecar(p))
    assert orc.check_invariants() == 0
    q = centres + 0.01 * np.random.default_rng(3).normal(size=centres.shape).astype(np.float32)
    got, _, _, _ = orc.search_batch(q, 64, 10, threads=4)
    want, _ = oracle.brute_force(prepared, q, 10, mode="prepared")
    recall = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(got.tolist(), want.tolist())])
    assert recall >= (0.85 if bridging_needed else 0.95), recall


def test_candidates_outside_their_layer_are_ignored():
    """Caller-supplied candidates are untrusted input: handles that did not draw the layer (or are out of range for
    it) must neither corrupt memory nor create links that break invariant 7 (neighbor.top_layer >= layer)."""
    data, _ = datasets.uniform_dataset(600, 1, 16, 9)
    cfg = default_config(dimensions=16, max_connections=4, ef_construction=32)
    st = ChronoMindB200.from_matrix(data, cfg)
    members = st.layer_members(77)
    assert len(members) >= 2
    rng = np.random.default_rng(1)
    cands = [rng.integers(0, 600, size=(len(m), 12)).astype(np.uint32) for m in members]   # any handle, any layer
    st.build_graph_from_candidates(77, cands, threads=2)
    assert st.check_invariants() == 0
    for layer, m in enumerate(members[1:], start=1):
        inside = set(m.tolist())
        for h in m[:50]:
            assert set(st.neighbors(int(h), layer).tolist()) <= inside
    with pytest.raises(CmError):
        st.build_graph_from_candidates(77, [np.full((len(m), 3), 600, np.uint32) for m in members])   # out of range


def test_group_needs_cuda_devices():
    """`cm_group_create` on a box without a CUDA device: a typed error, not a crash (and nothing falls back to the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from chronomind_b200.engine import ShardedChronoMindB200
    with pytest.raises(CmError) as e:
        ShardedChronoMindB200([0])
    assert e.value.kind == "Cuda"


def test_fast_snapshot_writer_is_byte_identical_to_the_codec(tmp_path):
    """`datasets.write_snapshot` (the vectorised writer the bench and the scale tests feed the snapshot reader with) against
    the checker's record-by-record bincode restatement: same bytes, and the product's reader decodes it to the same store."""
    n, dim = 300, 12
    rng = np.random.default_rng(5)
    rows = rng.normal(size=(n, dim)).astype(np.float32)
    ts_s, ts_n, dec = datasets.temporal_attrs(n, 3, 1_790_000_000)
    ctx = rng.integers(0, 7, size=n).astype(np.uint32)
    imp = rng.uniform(0, 1, size=n).astype(np.float32)
    p = str(tmp_path / "fast.chrono")
    info = datasets.write_snapshot(p, rows, ts_s, ts_n, dec, ctx, 7, imp, config={"ef_construction": 100, "max_memories": 1000})
    mems = [sc.Memory(id="m%09d" % i, data=rows[i].tolist(), ts_secs=int(ts_s[i]), ts_nanos=int(ts_n[i]), importance=float(imp[i]),
                      context="ctx%05d" % ctx[i], decay_rate=float(dec[i]), last_access_secs=int(ts_s[i]),
                      last_access_nanos=int(ts_n[i])) for i in range(n)]
    want = sc.encode_snapshot(sc.Config(dimensions=dim, max_memories=1000, index=sc.IndexParams(16, 100, 50)), mems)
    got = open(p, "rb").read()
    assert got == want and info["bytes"] == len(want) and info["records"] == n
    st = ChronoMindB200.from_snapshot(p)
    assert len(st) == n and st.config.dimensions == dim and st.config.index.ef_construction == 100
    assert st.external_id(0) == "m000000000" and st.external_id(n - 1) == "m%09d" % (n - 1)
    assert st.context_label(5) == "ctx%05d" % ctx[5] and abs(st.importance(17) - imp[17]) < 1e-7
    st.close()


def test_large_snapshot_takes_the_parallel_reader_paths(tmp_path):
    """Files beyond 8 MB are checksummed in slices (CRC-32 combine) and their payloads copied by all host threads: the
    result must be what the small-file path gives — same store, and a flipped byte anywhere is still a checksum error."""
    n, dim = 12_000, 256
    rng = np.random.default_rng(8)
    rows = rng.normal(size=(n, dim)).astype(np.float32)
    ts_s, ts_n, dec = datasets.temporal_attrs(n, 8, 1_790_000_000)
    ctx = rng.integers(0, 5, size=n).astype(np.uint32)
    p = str(tmp_path / "big.chrono")
    info = datasets.write_snapshot(p, rows, ts_s, ts_n, dec, ctx, 5)
    assert info["bytes"] > (8 << 20)
    st = ChronoMindB200.from_snapshot(p)
    assert len(st) == n and st.external_id(n - 1) == "m%09d" % (n - 1) and st.context_label(n - 1) == "ctx%05d" % ctx[n - 1]
    want = oracle.preprocess_rows(rows)
    for h in (0, 1, 4095, 4096, n // 2, n - 1):                   # across the reader's thread slices
        assert np.array_equal(st.vector(h).view(np.uint32), want[h].view(np.uint32))
    st.close()
    blob = bytearray(open(p, "rb").read())
    for pos in (12 + 100, len(blob) // 3, len(blob) - 5):
        bad = bytearray(blob)
        bad[pos] ^= 0x40
        q = str(tmp_path / "bad.chrono")
        open(q, "wb").write(bad)
        with pytest.raises(CmError) as e:
            ChronoMindB200.from_snapshot(q)
        assert e.value.kind == "InvalidSnapshot" and "checksum" in str(e.value)
    nan_rows = rows.copy()
    nan_rows[7777, 3] = np.nan                                     # found by the parallel payload pass, reported per record
    datasets.write_snapshot(p, nan_rows, ts_s, ts_n, dec, ctx, 5)
    with pytest.raises(CmError) as e:
        ChronoMindB200.from_snapshot(p)
    assert e.value.kind == "InvalidVector" and "m000007777" in str(e.value)
