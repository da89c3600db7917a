"""The sharded read path behind the C ABI (`cm_group_*`): one object, one call per batch, G GPUs of one process —
the role of `ShardedRwLockHnsw` (src/index/sharded_rwlock.rs:43-94). Checked against the unsharded engine and against
the oracle doing the same three steps on the CPU (per-shard search, merge by (distance, global handle), rerank of the
merged pool). The multi-device cases skip below two GPUs (`gpurun --gpus 2`); the single-device group runs everywhere
and exercises everything except the NCCL exchange."""
import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, CmError, ShardedChronoMindB200, default_config
from parity_helpers import assert_order_matches_up_to_ties

pytestmark = pytest.mark.gpu
NOW = (1_790_000_000, 250_000_000)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def n_devices():
    import torch
    return torch.cuda.device_count()


def _oracle_sharded(data, q, G, ef, pool, seed, efc=100):
    """Per-shard oracle HNSW search (sub-graph per shard, seed + shard) merged like sharded_rwlock.rs:81-94."""
    ids, ds = [], []
    for r in range(G):
        orc = oracle.OracleHnsw(16, efc, 50, seed + r)
        orc.insert_matrix(data[r::G])
        oi, od, _, _ = orc.search_batch(q, ef, pool, threads=8)
        ids.append(oi)
        ds.append(od)
    return oracle.sharded_merge(np.stack(ids), np.stack(ds), pool)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_group_search_equals_unsharded_and_oracle(G):
    if n_devices() < G:
        pytest.skip("needs %d CUDA devices" % G)
    data, q = datasets.embedding_dataset(24000, 300, 0x6A0, dim=256, intrinsic=12)
    ts_s, ts_n, dec = datasets.temporal_attrs(len(data), 5, NOW[0])
    dele = (np.arange(len(data)) % 11 == 3).astype(np.uint8)
    cfg = default_config(dimensions=256, ef_construction=100, ef_search=64, temporal_weight=0.3)
    grp = ShardedChronoMindB200.from_matrix(list(range(G)), data, cfg, ts_s, ts_n, dec, dele)
    assert grp.slots == len(data) and len(grp) == int((dele == 0).sum())
    grp.build_graphs(0x6A0, threads=8, gpu=False).upload()
    # exact: == the unsharded index == the oracle, global handles
    whole = ChronoMindB200.from_matrix(data, cfg, ts_s, ts_n, dec, dele).upload(0)
    gi, gd = grp.search_exact(q, 10)
    wi, wd = whole.search_exact(q, 10)
    oi, od = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared", deleted=dele, threads=8)
    assert np.array_equal(gi, wi) and np.array_equal(bits(gd), bits(wd))
    assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
    assert not dele[gi[gi != NO_ID]].any()
    # the exact pool of a store search, then the rerank of the merged pool
    si, ss, sd = grp.search(q, 10, now=NOW, exact=True, want_dist=True)
    pi, pd = oracle.brute_force(oracle.preprocess_rows(data), q, 64, mode="prepared", deleted=dele, threads=8)
    ri, rs = oracle.rerank_pool(pi, pd, 10, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert assert_order_matches_up_to_ties(si, ri, ss, rs) <= 0.001 * si.size
    # HNSW: independent sub-graph per shard (host build, one thread per shard => deterministic)
    grp1 = ShardedChronoMindB200.from_matrix(list(range(G)), data, cfg, ts_s, ts_n, dec).build_graphs(0x6A0, threads=G, gpu=False).upload()
    hi, hd = grp1.search_index(q, 64, 10)
    mi, md = _oracle_sharded(data, q, G, 64, 10, 0x6A0)
    assert np.array_equal(hi, mi) and np.array_equal(bits(hd), bits(md))
    ti, tsc = grp1.search(q, 10, now=NOW, ef_search=64)
    pi, pd = _oracle_sharded(data, q, G, 64, 64, 0x6A0)
    ri, rs = oracle.rerank_pool(pi, pd, 10, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert assert_order_matches_up_to_ties(ti, ri, tsc, rs) <= 0.001 * ti.size
    for g in (grp, grp1):
        g.close()


def test_group_errors_and_edge_cases():
    data, q = datasets.uniform_dataset(3000, 5, 32, 2)
    cfg = default_config(dimensions=32)
    grp = ShardedChronoMindB200.from_matrix([0], data, cfg)
    with pytest.raises(CmError) as e:                      # not resident yet: there is no CPU path
        grp.search_exact(q, 3)
    assert e.value.kind == "Cuda"
    grp.upload()
    with pytest.raises(CmError) as e:                      # validate_query: dimensions first ...
        grp.search(np.ones((2, 31), np.float32), 3, now=NOW, exact=True)
    assert e.value.kind == "InvalidDimensions"
    bad = q.copy()
    bad[3, 7] = np.nan
    with pytest.raises(CmError) as e:                      # ... then finiteness, first bad row named
        grp.search(bad, 3, now=NOW, exact=True)
    assert e.value.kind == "InvalidVector" and "row 3" in str(e.value)
    with pytest.raises(CmError) as e:                      # no graph was built
        grp.search_index(q, 16, 3)
    assert e.value.kind == "State"
    ids, d = grp.search_exact(np.zeros((0, 32), np.float32), 3)
    assert ids.shape == (0, 3)
    with pytest.raises(CmError) as e:
        ShardedChronoMindB200([0, 0])                      # a device twice
    assert e.value.kind == "InvalidArgument"
    with pytest.raises(CmError) as e:
        ShardedChronoMindB200([n_devices()])               # ordinal out of range
    assert e.value.kind == "InvalidArgument"
