#!/usr/bin/env python
"""K1 epilogue floor on a short-row shape: time `scan_tc_kernel` with CM_SCAN_DIAG=1 (every threshold +inf: the epilogue
only reads the accumulators out of TMEM and takes chunk maxima — results are INVALID, every query goes to the fp32
fallback) next to the normal kernel. Measurement tool only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chronomind_b200 import datasets, engine                      # noqa: E402
from chronomind_b200.engine import ChronoMindB200, default_config  # noqa: E402


def main():
    n, dim, nq, k = int(os.environ.get("CM_ROWS", 4_000_000)), int(os.environ.get("CM_DIM", 32)), 10_000, int(os.environ.get("CM_K", 100))
    data, q = datasets.uniform_dataset(n, nq, dim, 0xC0FFEE)
    st = ChronoMindB200.from_matrix(data, default_config(dimensions=dim)).upload(0).set_exact_engine(2)
    for i in range(2):
        st.search_exact(q, k)
    engine.profile_enable(True)
    engine.profile_read()
    for i in range(4):
        st.search_exact(q, k)
    ms, cnt = engine.profile_read()
    engine.profile_enable(False)
    c = engine.last_counters()
    print(json.dumps({"rows": n, "dim": dim, "k": k, "diag": os.environ.get("CM_SCAN_DIAG"), "epi2": os.environ.get("CM_SCAN_EPI2"),
                      "kernel_ms": ms / cnt, "fallback": c["fallback_queries"]}), flush=True)


if __name__ == "__main__":
    main()
