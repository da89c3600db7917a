"""Sweep the HNSW kernel's rows-per-round (CM_HNSW_ROWS) at each ef on the bench corpus (not a test)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chronomind_b200 import datasets, engine
from chronomind_b200.engine import ChronoMindB200, default_config
import bench
n = int(os.environ.get("DIAG_ROWS", "1000000")); nq = 10000
corpus, queries = bench.make_data(n, nq)
st = ChronoMindB200.from_matrix(corpus, default_config(dimensions=768, ef_construction=100))
t0 = time.time(); st.build_graph(bench.SEED, os.cpu_count()); print("build %.1fs" % (time.time() - t0), flush=True)
st.upload(0)
q = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
for ef in (64, 128, 200):
    for rows in (0, 16, 14, 12, 10, 8, 6):
        if rows: os.environ["CM_HNSW_ROWS"] = str(rows)
        else: os.environ.pop("CM_HNSW_ROWS", None)
        for _ in range(2): st.search_index_cuda(q, ef, 10, ids, d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): st.search_index_cuda(q, ef, 10, ids, d)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print("ef=%d rows=%s: %.2f ms/batch  %.0f qps" % (ef, rows or "auto", ms, nq / ms * 1e3), flush=True)
