"""Committed golden vectors (tests/golden/read_path_small.npz, made by tests/golden/make_golden.py from the pinned
oracle — the Rust reference cannot run here): the oracle must still reproduce them on the CPU, and the CUDA path must
match them through the C ABI on the GPU."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "read_path_small.npz"))
NOW = (int(G["now"][0]), int(G["now"][1]))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_oracle_reproduces_golden_vectors():
    assert np.array_equal(bits(oracle.preprocess_rows(G["data"])), bits(G["prepared"]))
    ids, d = oracle.brute_force(G["prepared"], G["queries"], 10, mode="prepared")
    assert np.array_equal(ids, G["exact_ids"]) and np.array_equal(bits(d), bits(G["exact_dist"]))
    orc = oracle.OracleHnsw(8, 40, 20, 0x601D)
    orc.insert_matrix(G["data"])
    hi, hd, _, ctr = orc.search_batch(G["queries"], 24, 10)
    assert np.array_equal(hi, G["hnsw_ids"]) and np.array_equal(bits(hd), bits(G["hnsw_dist"]))
    assert np.array_equal(ctr, G["hnsw_counters"])
    rc, ti, tsc, _ = orc.store_search_batch(G["queries"], 5, 48, 20, 0.3, 0.1, G["ts_secs"], G["ts_nanos"], G["decay"],
                                            NOW[0], NOW[1])
    assert rc == 0 and np.array_equal(ti, G["temporal_ids"])
    assert np.allclose(tsc, G["temporal_score"], rtol=1e-6, atol=0.0)      # libm expf may differ in the last ulp


@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors():
    from chronomind_b200.engine import ChronoMindB200, default_config, last_counters
    cfg = default_config(dimensions=48, max_connections=8, ef_construction=40, ef_search=20)
    st = ChronoMindB200.from_matrix(G["data"], cfg, G["ts_secs"], G["ts_nanos"], G["decay"]).build_graph(0x601D, 1).upload(0)
    ids, d = st.search_exact(G["queries"], 10)
    assert np.array_equal(ids, G["exact_ids"]) and np.array_equal(bits(d), bits(G["exact_dist"]))
    hi, hd = st.search_index(G["queries"], 24, 10)
    c = last_counters()
    assert np.array_equal(hi, G["hnsw_ids"]) and np.array_equal(bits(hd), bits(G["hnsw_dist"]))
    assert c["dist_evals"] == int(G["hnsw_counters"][0]) and c["expansions"] == int(G["hnsw_counters"][1])
    ti, tsc = st.search(G["queries"], 5, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
    assert np.array_equal(ti, G["temporal_ids"])
    assert np.allclose(tsc, G["temporal_score"], rtol=1e-5, atol=0.3 * 2.0 ** -24)
