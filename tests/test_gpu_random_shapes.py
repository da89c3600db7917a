"""Seeded random shapes through the C ABI against the oracle — the shapes nobody picked by hand: odd dimensions,
k / ef / M combinations, tombstones, duplicate rows, zero rows. Mirrors the reference's property tests
(tests/property_test.rs:146-229: exhaustive-ef search == linear scan, exactly) at GPU speed; every comparison is bit for bit
(ids, distance bits, and for the traversal the work counters)."""
import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import NO_ID, ChronoMindB200, default_config, last_counters

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def _case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([257, 900, 3000, 9000, 20000]))
    dim = int(rng.choice([3, 8, 13, 31, 64, 100, 257, 768]))
    M = int(rng.choice([4, 8, 16, 24]))
    k = int(rng.choice([1, 5, 10, 33, 100]))
    ef = int(rng.choice([1, 7, 10, 50, 130]))
    nq = int(rng.choice([1, 17, 130]))
    x = rng.normal(size=(n, dim)).astype(np.float32) * rng.uniform(0.1, 10.0, size=(n, 1)).astype(np.float32)
    if rng.uniform() < 0.5:                                     # duplicates
        dst = rng.choice(n, size=n // 6, replace=False)
        x[dst] = x[rng.integers(0, n, size=len(dst))]
    if rng.uniform() < 0.3:
        x[rng.integers(0, n, size=3)] = 0.0                     # degenerate rows stay zero (distance 1.0)
    dele = (rng.uniform(size=n) < rng.choice([0.0, 0.1, 0.5])).astype(np.uint8)
    q = rng.normal(size=(nq, dim)).astype(np.float32)
    q[0] = x[rng.integers(0, n)]                               # a stored row as a query
    return x, dele, q, M, k, ef


@pytest.mark.parametrize("seed", list(range(100, 116)))
def test_random_shapes_exact_and_hnsw_match_the_oracle(seed):
    x, dele, q, M, k, ef = _case(seed)
    n, dim = x.shape
    cfg = default_config(dimensions=dim, max_connections=M, ef_construction=40)
    st = ChronoMindB200.from_matrix(x, cfg, deleted=dele).build_graph(seed, 1).upload(0)
    prepared = oracle.preprocess_rows(x)
    wi, wd = oracle.brute_force(prepared, q, k, mode="prepared", deleted=dele, threads=8)
    for engine_id in (1, 2):                                     # fp32 scan, tensor-core scan + rescore
        gi, gd = st.set_exact_engine(engine_id).search_exact(q, k)
        assert np.array_equal(gi, wi) and np.array_equal(bits(gd), bits(wd)), (seed, engine_id)
    orc = oracle.OracleHnsw(M, 40, 50, seed)
    orc.insert_matrix(x)
    for h in np.nonzero(dele)[0]:
        orc.remove(int(h))
    hi, hd = st.search_index(q, ef, ef)
    c = last_counters()
    oi, od, _, octr = orc.search_batch(q, ef, ef, threads=8)
    assert np.array_equal(hi, oi) and np.array_equal(bits(hd), bits(od)), seed
    assert c["dist_evals"] == int(octr[0]) and c["expansions"] == int(octr[1]) and c["list_entries"] == int(octr[2]), seed
    assert c["failures"] == 0
    assert not dele[hi[hi != NO_ID]].any()
    # exhaustive ef (tests/property_test.rs:196-229): the traversal must equal the oracle's, and — whenever the graph the
    # reference's own insertion built reaches every live node (duplicate rows can leave a node without an incoming edge:
    # `select_diverse` rejects a candidate at distance 0 from a chosen one) — the linear scan exactly
    if n <= 1000:
        live = int((dele == 0).sum())
        efx = min(1024, n + 8)
        xi, xd = st.search_index(q[:3], efx, efx)
        ei, ed, _, _ = orc.search_batch(q[:3], efx, efx)
        assert np.array_equal(xi, ei) and np.array_equal(bits(xd), bits(ed)), seed
        reach = min(live, efx)
        if np.all((ei != NO_ID).sum(axis=1) == reach):
            li, ld = oracle.brute_force(prepared, q[:3], efx, mode="prepared", deleted=dele)
            assert np.array_equal(xi[:, :reach], li[:, :reach]) and np.array_equal(bits(xd[:, :reach]), bits(ld[:, :reach])), seed
