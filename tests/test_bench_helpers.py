"""bench.py's host-side helpers: chunked corpus generation (shards == strided views of the whole) and the streaming
CPU scan used as cpu_baseline / parity checker (== one oracle scan over the concatenated corpus)."""
import numpy as np

import bench
import oracle
from chronomind_b200.datasets import ChunkedCorpus


def test_chunked_corpus_shards_partition_the_corpus():
    for kind in ("emb", "uniform"):
        gen = ChunkedCorpus(kind, 1000, 24, 5, chunk=128)
        full = gen.shard()
        assert full.shape == (1000, 24) and np.allclose(np.linalg.norm(full, axis=1), 1.0, atol=1e-5)
        assert np.array_equal(full, np.vstack([gen.chunk(c) for c in range(gen.n_chunks)]))
        for world in (2, 3, 8):
            for rank in range(world):
                assert np.array_equal(gen.shard(rank, world), full[rank::world])
        assert np.array_equal(ChunkedCorpus(kind, 1000, 24, 5, chunk=128).queries(7), gen.queries(7))


def test_streaming_cpu_scan_equals_one_scan():
    gen = ChunkedCorpus("uniform", 700, 16, 9, chunk=100)
    q = gen.queries(13)
    ids, d, secs = bench.cpu_scan_streaming(gen, q, 10, threads=2)
    wi, wd = oracle.brute_force(oracle.preprocess_rows(gen.shard()), q, 10, mode="prepared")
    assert np.array_equal(ids, wi) and np.array_equal(d.view(np.uint32), wd.view(np.uint32)) and secs > 0


def test_total_keys_order_is_total_cmp():
    d = np.array([[-0.0, 0.0, 1.5, -2.0, np.inf]], np.float32)
    k = bench.total_keys(d, np.zeros_like(d, dtype=np.uint32))
    assert list(np.argsort(k[0])) == [3, 0, 1, 2, 4]


def test_frontier_is_the_smallest_ef_that_meets_the_recall_gate():
    per_ef = {"16": {"recall_at_10": 0.91}, "24": {"recall_at_10": 0.949}, "32": {"recall_at_10": 0.97},
              "64": {"recall_at_10": 0.99}}
    assert bench.frontier_ef(per_ef) == 32
    assert bench.frontier_ef({"16": {"recall_at_10": 0.5}}) is None


def test_both_arms_name_the_same_workload():
    """The driver compares the `config.workload` strings of the two arms: the task is one string, engines differ."""
    a = bench.workload_string(1_000_000, 768, "emb", 10_000, 10)
    assert "1000000x768" in a and "recall@10>=0.95" in a and "configs[1 and 2]" in a
    assert bench.workload_string(10_000_000, 768, "emb", 10_000, 10).endswith("configs[3])")


def test_order_mismatches_beyond_ties_counts_only_unexplained_positions():
    want_ids = np.array([[1, 2, 3, 4]])
    want_sc = np.array([[0.1, 0.2, 0.2 + 1e-9, 0.5]])
    assert bench.order_mismatches_beyond_ties(want_ids, want_ids, want_sc, want_sc) == 0
    swapped = np.array([[1, 3, 2, 4]])                       # a swap inside the near-tie: explained
    assert bench.order_mismatches_beyond_ties(swapped, want_ids, want_sc, want_sc) == 0
    wrong = np.array([[9, 2, 3, 4]])                         # a different id with a clear gap around it
    assert bench.order_mismatches_beyond_ties(wrong, want_ids, want_sc, want_sc) == 1
    off = want_sc * (1 + 1e-3)                               # scores off by more than 1e-5 relative
    assert bench.order_mismatches_beyond_ties(want_ids, want_ids, off, want_sc) == 4
