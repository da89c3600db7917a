"""Sweep the HNSW kernel's launch parameters on the bench corpus (not a test): rows per copy round (CM_HNSW_ROWS),
where `visited` lives (CM_HNSW_VMAP: 0 shared-memory hash table, 1 byte map in HBM) and threads per CTA
(CM_HNSW_THREADS), per ef. Results must not change — the last column checks ids against the default configuration."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chronomind_b200.engine import ChronoMindB200, default_config
import bench
n = int(os.environ.get("DIAG_ROWS", "1000000")); nq = 10000
corpus, queries, _ = bench.make_data(n, nq)
st = ChronoMindB200.from_matrix(corpus, default_config(dimensions=768, ef_construction=100))
t0 = time.time(); st.build_graph_gpu(bench.SEED, 0, os.cpu_count()); print("build %.1fs" % (time.time() - t0), flush=True)
st.upload(0)
q = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
efs = [int(e) for e in os.environ.get("DIAG_EFS", "32,64,128,200").split(",")]
for ef in efs:
    for k in ("CM_HNSW_ROWS", "CM_HNSW_VMAP", "CM_HNSW_THREADS"):
        os.environ.pop(k, None)
    st.search_index_cuda(q, ef, 10, ids, d); torch.cuda.synchronize()
    want = ids.clone()
    for vmap in (None, 0, 1):
        for threads in (None, 64, 128, 256):
            for rows in (None, 12, 10, 8, 6):
                for k, v in (("CM_HNSW_ROWS", rows), ("CM_HNSW_VMAP", vmap), ("CM_HNSW_THREADS", threads)):
                    if v is None: os.environ.pop(k, None)
                    else: os.environ[k] = str(v)
                if (vmap is None) != (threads is None) or (vmap is None and rows is not None):
                    continue                      # the all-default line once, then the full grid
                for _ in range(2): st.search_index_cuda(q, ef, 10, ids, d)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4): st.search_index_cuda(q, ef, 10, ids, d)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 4
                print("ef=%d vmap=%s threads=%s rows=%s: %.2f ms/batch  %.0f qps  same=%s"
                      % (ef, vmap, threads, rows, ms, nq / ms * 1e3, bool(torch.equal(ids, want))), flush=True)
