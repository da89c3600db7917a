#!/usr/bin/env python
"""K1 sweep: resident query blocks x ring slots of `scan_tc_kernel` on BASELINE configs[1] (1M x 768 emb-like corpus,
10,000-query batches, k = 10), each setting timed with the library's own CUDA-event bracket around the kernel after
warm-up, every setting returning the same bits. One JSON line per setting."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chronomind_b200 import datasets, engine                      # noqa: E402
from chronomind_b200.engine import ChronoMindB200, default_config  # noqa: E402


def main():
    n, dim, nq, k = int(os.environ.get("CM_ROWS", 1_000_000)), int(os.environ.get("CM_DIM", 768)), 10_000, 10
    gen = datasets.ChunkedCorpus("emb", n, dim, 0x5EED)
    st = ChronoMindB200.from_matrix(gen.shard(0, 1), default_config(dimensions=dim)).upload(0).set_exact_engine(2)
    qs = [gen.queries(nq, stream=i) for i in range(3)]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    settings = [tuple(int(v) for v in s.split(":")) for s in os.environ.get("CM_SWEEP", "0:10,4:9,6:7,8:5,9:4,10:3").split(",")]
    ref = None
    for rnd in range(int(os.environ.get("CM_SWEEP_ROUNDS", 2))):
        for q_res, slots in settings:
            os.environ["CM_SCAN_QRES"], os.environ["CM_SCAN_SLOTS"] = str(q_res), str(slots)
            for i in range(3):
                st.search_exact(qs[i % 3], k)
            engine.profile_enable(True)
            engine.profile_read()
            burst = []
            if os.environ.get("CM_SWEEP_BURST"):      # one launch at a time on a rested GPU: the clock the power cap allows a short kernel
                for i in range(5):
                    time.sleep(1.0)
                    ids, dist = st.search_exact(qs[i % 3], k)
                    burst.append(engine.profile_read()[0])
            for i in range(12):
                ids, dist = st.search_exact(qs[i % 3], k)
            ms, cnt = engine.profile_read()
            engine.profile_enable(False)
            c = engine.last_counters()
            if ref is None:
                ref = (ids.copy(), dist.copy())
            same = bool(np.array_equal(ids, ref[0]) and np.array_equal(dist.view(np.uint32), ref[1].view(np.uint32)))
            tf = 2.0 * nq * n * dim / (ms / cnt * 1e-3) / 1e12
            print(json.dumps({"round": rnd, "q_res": q_res, "slots": slots, "kernel_ms": ms / cnt, "tflops": tf,
                              "frac_burst": tf / peaks["bf16_tflops"], "same_bits": same, "fallback": c["fallback_queries"],
                              "rested_single_launch_ms": burst}), flush=True)


if __name__ == "__main__":
    main()
