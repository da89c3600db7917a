#!/usr/bin/env python
"""K1 time against corpus size at a FIXED number of splits (CM_SCAN_SPLITS): the slope is the per-tile cost, the
intercept what a launch pays whatever the corpus (launch, cold thresholds of the first wave, unit boundaries).
10,000 x 768 queries, k = 10; kernel time from the library's CUDA-event bracket. One JSON line per corpus size."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chronomind_b200 import datasets, engine                      # noqa: E402
from chronomind_b200.engine import ChronoMindB200, default_config  # noqa: E402


def main():
    dim, nq, k = 768, 10_000, 10
    sizes = [int(v) for v in os.environ.get("CM_SIZES", "62500,125000,250000,500000,1000000").split(",")]
    gen = datasets.ChunkedCorpus("emb", max(sizes), dim, 0x5EED)
    full = gen.shard(0, 1)
    qs = [gen.queries(nq, stream=i) for i in range(3)]
    if os.environ.get("CM_SAME_BATCH"):
        qs = [qs[0]] * 3
    for n in sizes:
        st = ChronoMindB200.from_matrix(full[:n], default_config(dimensions=dim)).upload(0).set_exact_engine(2)
        for i in range(3):
            st.search_exact(qs[i % 3], k)
        engine.profile_enable(True)
        engine.profile_read()
        for i in range(20):
            st.search_exact(qs[i % 3], k)
        ms, cnt = engine.profile_read()
        engine.profile_enable(False)
        print(json.dumps({"rows": n, "splits": os.environ.get("CM_SCAN_SPLITS"), "epi2": os.environ.get("CM_SCAN_EPI2"), "warm": os.environ.get("CM_SCAN_WARM"), "fallback": engine.last_counters()["fallback_queries"],
                          "kernel_ms": ms / cnt, "tiles": (n + 255) // 256}), flush=True)
        del st


if __name__ == "__main__":
    main()
