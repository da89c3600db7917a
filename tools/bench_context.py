#!/usr/bin/env python
"""`ChronoMind::search_in_context` (src/store.rs:353-377) throughput: 10,000-query batches against 1M x 768 stored rows
partitioned into {4, 1000} contexts, through the host-buffer C ABI (H2D of queries + context ids, D2H of results inside
the timed call). Prints one JSON object per context count with the biased scan's tensor-core rate; the fp32 context
scan (the round-1 engine) is timed beside it on a 256-query slice."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chronomind_b200 import datasets, engine                      # noqa: E402
from chronomind_b200.engine import ChronoMindB200, default_config  # noqa: E402

NOW = (1_790_000_000, 0)


def main():
    n, dim, nq, k = int(os.environ.get("CM_ROWS", 1_000_000)), 768, 10_000, 10
    gen = datasets.ChunkedCorpus("emb", n, dim, 0x5EED)
    raw = gen.shard(0, 1) * np.random.default_rng(1).uniform(0.5, 4.0, size=(n, 1)).astype(np.float32)   # stored rows are not unit length
    q = gen.queries(nq) * 2.0
    try:                                          # page-locked host buffer, as bench.py's legs use (a serving process would)
        import torch
        q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).pin_memory().numpy()
    except Exception:
        pass
    ts_s, ts_n, dec = datasets.temporal_attrs(n, 1, NOW[0])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for n_ctx in [int(v) for v in os.environ.get("CM_CONTEXTS", "4,1000").split(",")]:
        rng = np.random.default_rng(n_ctx)
        ctx = rng.integers(0, n_ctx, size=n).astype(np.uint32)
        want = rng.integers(0, n_ctx, size=nq).astype(np.uint32)
        st = ChronoMindB200.from_matrix(raw, default_config(dimensions=dim), ts_s, ts_n, dec).set_contexts(raw, ctx).upload(0)
        for _ in range(2):
            st.search_in_context(want, q, k, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
        engine.profile_enable(True)
        engine.profile_read()
        steps = 10
        t0 = time.perf_counter()
        for _ in range(steps):
            ids, sc = st.search_in_context(want, q, k, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
        secs = time.perf_counter() - t0
        kms, kn = engine.profile_read()
        engine.profile_enable(False)
        c = engine.last_counters()
        st.set_exact_engine(1)
        t1 = time.perf_counter()
        i1, s1 = st.search_in_context(want[:256], q[:256], k, now=NOW, temporal_weight=0.3, base_decay_rate=0.1)
        fp32_secs = time.perf_counter() - t1
        same = bool(np.array_equal(i1, ids[:256]) and np.array_equal(s1.view(np.uint32), sc[:256].view(np.uint32)))
        tflops = 2.0 * nq * n * dim / (kms / max(kn, 1) * 1e-3) / 1e12 if kms else None
        print(json.dumps({"workload": "search_in_context %dx%d stored rows, %d contexts, %d-query batch, k=%d, w=0.3" % (n, dim, n_ctx, nq, k),
                          "e2e_qps": nq * steps / secs, "ms_per_batch": 1000 * secs / steps, "engine": c["exact_engine"],
                          "fallback_queries": c["fallback_queries"], "rescored_per_query": c["rescored"] / nq,
                          "scan_kernel_ms": kms / max(kn, 1), "scan_tflops": tflops,
                          "scan_frac_of_burst_bf16": (tflops / peaks["bf16_tflops"]) if tflops and peaks else None,
                          "fp32_context_scan_qps_256_queries": 256 / fp32_secs, "bit_identical_to_fp32_scan_on_256": same}), flush=True)
        st.close()


if __name__ == "__main__":
    main()
