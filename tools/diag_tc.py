"""Diagnostics for the tensor-core exact scan on a B200 (not a test): counters at the bench size and the NaN row."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import ChronoMindB200, default_config, last_counters

def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)

rng = np.random.default_rng(3)
data = datasets.unit_vectors(rng, 9000, 64)
q = datasets.unit_vectors(rng, 70, 64)
q[13, 5] = np.nan
st = ChronoMindB200.from_matrix(data, default_config(dimensions=64)).upload(0).set_exact_engine(2)
ids, d = st.search_exact(q, 10)
print("nan case counters", last_counters())
wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared")
bad = np.nonzero((bits(d) != bits(wd)).any(axis=1))[0]
print("rows with differing dist bits:", bad)
for r in bad[:3]:
    print(r, ids[r], wi[r]); print([hex(x) for x in bits(d)[r]]); print([hex(x) for x in bits(wd)[r]])

n = int(os.environ.get("DIAG_ROWS", "1000000"))
for kind in ("emb", "uni"):
    t0 = time.time()
    if kind == "emb":
        rng = np.random.default_rng(0x5EED)
        basis = datasets.unit_vectors(rng, 16, 768)
        corpus = datasets.embedding_rows(rng, basis, n)
        queries = datasets.embedding_rows(np.random.default_rng(1), basis, 10000)
    else:
        corpus, queries = datasets.uniform_dataset(n, 10000, 768, 7)
    st = ChronoMindB200.from_matrix(corpus, default_config(dimensions=768)).upload(0).set_exact_engine(2)
    print(kind, "built in %.1fs" % (time.time() - t0), flush=True)
    for rep in range(2):
        t0 = time.time()
        ids, d = st.search_exact(queries, 10)
        dt = time.time() - t0
        c = last_counters()
        print(kind, "rep", rep, "host-api %.1f ms" % (dt * 1e3), c, "rescored/query %.1f" % (c["rescored"] / 10000), flush=True)
    prep = oracle.preprocess_rows(corpus, threads=os.cpu_count())
    wi, wd = oracle.brute_force(prep, queries[:64], 10, mode="prepared", threads=os.cpu_count())
    print(kind, "parity on 64 queries:", np.array_equal(ids[:64], wi), np.array_equal(bits(d[:64]), bits(wd)))
    st.close()
