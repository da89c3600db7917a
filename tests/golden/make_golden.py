"""Regenerates tests/golden/read_path_small.npz.

The reference is a Rust crate and cannot be built or imported in this image (no rustc/cargo), so these vectors are
NOT outputs of the reference: they are outputs of the pinned CPU oracle (oracle/chrono_oracle.cpp, itself checked
against the reference's known-answer tests) on small seeded inputs, frozen so that
  * a change of the oracle that alters any result is caught on the CPU (tests/test_golden.py, `not gpu`), and
  * the CUDA path is compared with committed expectations, not only with whatever oracle was built on the GPU box.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from chronomind_b200 import datasets  # noqa: E402

NOW = (1_790_000_000, 250_000_000)


def build():
    out = {}
    data, q = datasets.embedding_dataset(600, 12, 0x601D, dim=48)
    prepared = oracle.preprocess_rows(data)
    out["data"], out["queries"], out["prepared"] = data, q, prepared
    ids, d = oracle.brute_force(prepared, q, 10, mode="prepared")
    out["exact_ids"], out["exact_dist"] = ids, d
    orc = oracle.OracleHnsw(8, 40, 20, 0x601D)
    orc.insert_matrix(data)
    hi, hd, _, ctr = orc.search_batch(q, 24, 10)
    out["hnsw_ids"], out["hnsw_dist"], out["hnsw_counters"] = hi, hd, ctr
    ts_s, ts_n, dec = datasets.temporal_attrs(600, 5, NOW[0])
    out["ts_secs"], out["ts_nanos"], out["decay"] = ts_s, ts_n, dec
    rc, ti, tsc, td = orc.store_search_batch(q, 5, 48, 20, 0.3, 0.1, ts_s, ts_n, dec, NOW[0], NOW[1])
    assert rc == 0
    out["temporal_ids"], out["temporal_score"], out["temporal_dist"] = ti, tsc, td
    out["now"] = np.array(NOW, dtype=np.uint64)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "read_path_small.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
