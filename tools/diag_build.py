"""Host vs GPU-assisted graph construction: build time and recall@10 (not a test). DIAG_ROWS / DIAG_DATA env."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from chronomind_b200.engine import ChronoMindB200, default_config, last_counters
n = int(os.environ.get("DIAG_ROWS", "100000")); kind = os.environ.get("DIAG_DATA", "uniform")
data, q, _ = bench.make_data(n, 2000, kind)
cfg = default_config(dimensions=768, ef_construction=100)
def evaluate(st, label, secs):
    truth, _ = st.search_exact(q, 10)
    out = []
    for ef in (64, 128, 200, 400):
        ids, _ = st.search_index(q, ef, 10)
        c = last_counters()
        r = float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ids.tolist(), truth.tolist())]))
        out.append("ef=%d recall %.4f evals/q %.0f" % (ef, r, c["dist_evals"] / len(q)))
    print("%-22s build %.1fs invariants=%d | %s" % (label, secs, st.check_invariants(), " | ".join(out)), flush=True)
t = time.time(); st = ChronoMindB200.from_matrix(data, cfg).build_graph(1, os.cpu_count()); s = time.time() - t
evaluate(st.upload(0), "host efC=100", s)
for C in (32, 64, 127):
    t = time.time(); st = ChronoMindB200.from_matrix(data, cfg).build_graph_gpu(1, 0, os.cpu_count(), C); s = time.time() - t
    evaluate(st.upload(0), "gpu C=%d" % C, s)
