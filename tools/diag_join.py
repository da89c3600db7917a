"""Time the consolidate self-join (cm_similar_pairs) on the bench corpus with planted near-duplicates (not a test)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from chronomind_b200 import engine
from chronomind_b200.engine import ChronoMindB200, default_config, last_counters
n = int(os.environ.get("DIAG_ROWS", "1000000"))
raw, _, _ = bench.make_data(n, 16, os.environ.get("DIAG_DATA", "uniform"))
rng = np.random.default_rng(5)
dup = 2000
raw[n - dup:] = raw[:dup] * 2.0 + 0.002 * rng.normal(size=(dup, raw.shape[1])).astype(np.float32)
st = ChronoMindB200.from_matrix(raw, default_config(dimensions=raw.shape[1])).set_contexts(raw).upload(0)
for rep in range(3):
    engine.profile_enable(True); engine.profile_read()
    t0 = time.time()
    i, j, s = st.similar_pairs(0.95)
    dt = time.time() - t0
    kms, kn = engine.profile_read(); engine.profile_enable(False)
    flop = n * (n - 1) / 2 * 2 * raw.shape[1]
    print("rep %d: %d pairs (planted %d) in %.3f s wall; scan kernel %.1f ms = %.0f TFLOP/s (upper-triangle FLOP); counters %s"
          % (rep, len(i), dup, dt, kms, flop / (kms * 1e-3) / 1e12 if kms else 0, last_counters()), flush=True)
