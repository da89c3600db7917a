"""CPU oracle for the ChronoMind read path — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/libchrono_oracle.so`` (built from
``chrono_oracle.cpp`` by ``oracle/Makefile``) plus the pure-Python snapshot
codec in ``oracle/snapshot_codec.py``. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this package; the product (``chrono-mind_b200/``)
never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libchrono_oracle.so")

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle if the .so is missing or older than its source."""
    src = os.path.join(_HERE, "chrono_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)

    def sig(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args

    vp = C.c_void_p
    sig("orc_cpu_has_avx2_fma", C.c_int, [])
    sig("orc_total_key", C.c_uint32, [C.c_float])
    sig("orc_dot_scalar", C.c_float, [_f32p, _f32p, C.c_uint64])
    sig("orc_dot_avx2", C.c_float, [_f32p, _f32p, C.c_uint64])
    sig("orc_preprocess", None, [_f32p, C.c_uint64, _f32p])
    sig("orc_preprocess_rows", None, [_f32p, C.c_uint64, C.c_uint64, _f32p, C.c_int])
    sig("orc_distance_prepared", C.c_float, [_f32p, C.c_uint64, _f32p, C.c_uint64])
    sig("orc_distance", C.c_float, [_f32p, C.c_uint64, _f32p, C.c_uint64])
    sig("orc_distance_scalar", C.c_float, [_f32p, C.c_uint64, _f32p, C.c_uint64])
    sig("orc_similarity", C.c_float, [_f32p, C.c_uint64, _f32p, C.c_uint64])
    sig("orc_splitmix64", C.c_uint64, [C.c_uint64])
    sig("orc_random_layer_at", C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64])
    sig("orc_combined_score", C.c_float,
        [C.c_float, C.c_uint64, C.c_uint32, C.c_float, C.c_uint64, C.c_uint32, C.c_float, C.c_float])
    sig("orc_hnsw_new", vp, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64])
    sig("orc_hnsw_free", None, [vp])
    sig("orc_hnsw_insert", C.c_int64, [vp, _f32p, C.c_uint64])
    sig("orc_hnsw_insert_matrix", C.c_uint64, [vp, _f32p, C.c_uint64, C.c_uint64])
    sig("orc_hnsw_remove", C.c_int, [vp, C.c_uint32])
    sig("orc_hnsw_len", C.c_uint64, [vp])
    sig("orc_hnsw_slots", C.c_uint64, [vp])
    sig("orc_hnsw_check_invariants", C.c_int, [vp])
    sig("orc_hnsw_entry", C.c_uint64, [vp])
    sig("orc_hnsw_top_layer", C.c_uint32, [vp, C.c_uint32])
    sig("orc_hnsw_neighbors", C.c_uint32, [vp, C.c_uint32, C.c_uint32, _u32p, C.c_uint32])
    sig("orc_hnsw_vector", None, [vp, C.c_uint32, _f32p])
    sig("orc_hnsw_import", None,
        [vp, _f32p, C.c_uint64, C.c_uint64, _u8p, _u32p, C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32, _u32p,
         C.c_uint64, vp])
    sig("orc_hnsw_search", C.c_uint64, [vp, _f32p, C.c_uint64, C.c_uint64, _u32p, _f32p, C.c_uint64, vp])
    sig("orc_hnsw_search_batch", None,
        [vp, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _u32p, _f32p, vp, C.c_int, vp])
    sig("orc_brute_force_batch", None,
        [_f32p, C.c_uint64, C.c_uint64, vp, _f32p, C.c_uint64, C.c_uint64, C.c_int, _u32p, _f32p, C.c_int])
    sig("orc_store_search_batch", C.c_int,
        [vp, C.c_uint64, C.c_uint64, C.c_float, C.c_float, _u64p, _u32p, _f32p, C.c_uint64, C.c_uint32, _f32p,
         C.c_uint64, C.c_uint64, C.c_uint64, _u32p, _f32p, vp, C.c_int])
    sig("orc_rerank_pool", None,
        [_u32p, _f32p, C.c_uint64, C.c_uint64, C.c_float, C.c_float, _u64p, _u32p, _f32p, C.c_uint64, C.c_uint32,
         C.c_uint64, _u32p, _f32p])
    sig("orc_search_in_context", C.c_int,
        [_f32p, C.c_uint64, C.c_uint64, _u32p, C.c_void_p, C.c_float, C.c_float, _u64p, _u32p, _f32p, C.c_uint64,
         C.c_uint32, _f32p, C.c_uint64, C.c_uint64, _u32p, C.c_uint64, _u32p, _f32p, C.c_int])
    sig("orc_similar_pairs", C.c_uint64,
        [_f32p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_float, _u32p, _u32p, _f32p, C.c_uint64])
    sig("orc_decay_sweep", None, [_f32p, _u64p, _u64p, _f32p, C.c_uint64, C.c_float, C.c_uint64])
    sig("orc_sharded_merge", None,
        [_u32p, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _u32p, _f32p])
    if not L.orc_cpu_has_avx2_fma():
        raise RuntimeError("oracle restates the reference's AVX2+FMA path; this CPU has neither")
    _lib = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- metric surface (src/metric.rs) ---------------------------------------
def preprocess(v) -> np.ndarray:
    v = _f32(v).ravel()
    out = np.empty_like(v)
    lib().orc_preprocess(v, v.size, out)
    return out


def preprocess_rows(x, threads=1) -> np.ndarray:
    x = _f32(x)
    out = np.empty_like(x)
    if x.shape[0]:
        lib().orc_preprocess_rows(x, x.shape[0], x.shape[1], out, threads)
    return out


def distance(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_distance(a, a.size, b, b.size))


def distance_scalar(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_distance_scalar(a, a.size, b, b.size))


def similarity(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_similarity(a, a.size, b, b.size))


def distance_prepared(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_distance_prepared(a, a.size, b, b.size))


def dot_avx2(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_dot_avx2(a, b, a.size))


def dot_scalar(a, b) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return float(lib().orc_dot_scalar(a, b, a.size))


def total_key(f: float) -> int:
    return int(lib().orc_total_key(C.c_float(f)))


def splitmix64(x: int) -> int:
    return int(lib().orc_splitmix64(C.c_uint64(x & (2**64 - 1))))


def random_layer(seed: int, ticket: int, max_connections: int = 16) -> int:
    return int(lib().orc_random_layer_at(seed, ticket, max_connections))


def combined_score(dist, ts_secs, ts_nanos, decay_rate, now_secs, now_nanos, w, base_decay_rate) -> float:
    return float(lib().orc_combined_score(dist, ts_secs, ts_nanos, decay_rate, now_secs, now_nanos, w, base_decay_rate))


# ---- index (src/index/lockfree_hnsw.rs) ------------------------------------
class OracleHnsw:
    """`LockFreeHnsw::with_seed(params, CosineDistance, seed)` under single-threaded use."""

    def __init__(self, max_connections=16, ef_construction=200, ef_search=50, seed=0):
        self.params = (max_connections, ef_construction, ef_search)
        self._h = lib().orc_hnsw_new(max_connections, ef_construction, ef_search, seed & (2**64 - 1))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hnsw_free(self._h)
            self._h = None

    def insert(self, v):
        v = _f32(v).ravel()
        r = lib().orc_hnsw_insert(self._h, v, v.size)
        return None if r < 0 else int(r)

    def insert_matrix(self, x) -> int:
        x = _f32(x)
        return int(lib().orc_hnsw_insert_matrix(self._h, x, x.shape[0], x.shape[1]))

    def remove(self, handle: int) -> bool:
        return bool(lib().orc_hnsw_remove(self._h, handle))

    def __len__(self):
        return int(lib().orc_hnsw_len(self._h))

    @property
    def slots(self) -> int:
        return int(lib().orc_hnsw_slots(self._h))

    def check_invariants(self) -> int:
        return int(lib().orc_hnsw_check_invariants(self._h))

    def entry(self):
        w = int(lib().orc_hnsw_entry(self._h))
        return None if w == 2**64 - 1 else (w & 0xFFFFFFFF, w >> 32)

    def top_layer(self, handle) -> int:
        return int(lib().orc_hnsw_top_layer(self._h, handle))

    def neighbors(self, handle, layer):
        out = np.empty(256, dtype=np.uint32)
        n = lib().orc_hnsw_neighbors(self._h, handle, layer, out, out.size)
        return out[:n].copy()

    def vector(self, handle, dim) -> np.ndarray:
        out = np.empty(dim, dtype=np.float32)
        lib().orc_hnsw_vector(self._h, handle, out)
        return out

    def import_graph(self, x_prepared, graph: dict, deleted=None):
        """Adopt a graph read by `read_graph_sidecar` over already-preprocessed rows."""
        x = _f32(x_prepared)
        dptr = None
        if deleted is not None:
            deleted = np.ascontiguousarray(deleted, dtype=np.uint8)
            dptr = deleted.ctypes.data_as(C.c_void_p)
        g = graph
        nbr_up = g["nbr_up"] if g["nbr_up"].size else np.zeros(1, np.uint32)
        deg_up = g["deg_up"] if g["deg_up"].size else np.zeros(1, np.uint32)
        lib().orc_hnsw_import(self._h, x, x.shape[0], x.shape[1], g["top_layer"], g["nbr0"], g["cap0"], g["deg0"],
                              g["upper_base"], nbr_up, g["capU"], deg_up, g["entry"], dptr)

    def search(self, q, ef, counters=None):
        """VectorIndex::search -> list of (handle, distance), ascending."""
        q = _f32(q).ravel()
        cap = max(int(ef), 1)
        ids = np.empty(cap, dtype=np.uint32)
        d = np.empty(cap, dtype=np.float32)
        cptr = counters.ctypes.data_as(C.c_void_p) if counters is not None else None
        n = int(lib().orc_hnsw_search(self._h, q, q.size, ef, ids, d, cap, cptr))
        return ids[:n].copy(), d[:n].copy()

    def search_batch(self, q, ef, k, threads=1):
        """PyO3 Index.batch_query precedent: per row search(q, ef) then take(k)."""
        q = _f32(q)
        nq, dim = q.shape
        ids = np.empty((nq, k), dtype=np.uint32)
        d = np.empty((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        counters = np.zeros(3, dtype=np.uint64)
        lib().orc_hnsw_search_batch(self._h, q, nq, dim, ef, k, ids, d, counts.ctypes.data_as(C.c_void_p), threads,
                                    counters.ctypes.data_as(C.c_void_p))
        return ids, d, counts, counters

    def store_search_batch(self, q, k, dimensions, ef_search, w, base_decay, ts_secs, ts_nanos, decay, now_secs,
                           now_nanos, threads=1):
        """ChronoMind::search (src/store.rs:321-346) for every row of q."""
        q = _f32(q)
        nq, qdim = q.shape
        ids = np.empty((nq, k), dtype=np.uint32)
        sc = np.empty((nq, k), dtype=np.float32)
        d = np.empty((nq, k), dtype=np.float32)
        rc = lib().orc_store_search_batch(
            self._h, dimensions, ef_search, w, base_decay, np.ascontiguousarray(ts_secs, np.uint64),
            np.ascontiguousarray(ts_nanos, np.uint32), _f32(decay), now_secs, now_nanos, q, nq, qdim, k, ids, sc,
            d.ctypes.data_as(C.c_void_p), threads)
        return rc, ids, sc, d


# ---- exact scans -------------------------------------------------------------
def brute_force(x, q, k, mode="prepared", deleted=None, threads=1):
    """mode 'distance': tests/recall_test.rs:70-79; 'prepared': tests/property_test.rs:201-206
    (x must already be preprocess()ed rows)."""
    x, q = _f32(x), _f32(q)
    nq = q.shape[0]
    ids = np.empty((nq, k), dtype=np.uint32)
    d = np.empty((nq, k), dtype=np.float32)
    dptr = None
    if deleted is not None:
        deleted = np.ascontiguousarray(deleted, dtype=np.uint8)
        dptr = deleted.ctypes.data_as(C.c_void_p)
    lib().orc_brute_force_batch(x, x.shape[0], x.shape[1], dptr, q, nq, k, 1 if mode == "prepared" else 0, ids, d,
                                threads)
    return ids, d


def rerank_pool(cand_ids, cand_dists, k, w, base_decay, ts_secs, ts_nanos, decay, now_secs, now_nanos):
    cand_ids = np.ascontiguousarray(cand_ids, np.uint32)
    cand_dists = _f32(cand_dists)
    nq, pool = cand_ids.shape
    ids = np.empty((nq, k), dtype=np.uint32)
    sc = np.empty((nq, k), dtype=np.float32)
    lib().orc_rerank_pool(cand_ids, cand_dists, nq, pool, w, base_decay, np.ascontiguousarray(ts_secs, np.uint64),
                          np.ascontiguousarray(ts_nanos, np.uint32), _f32(decay), now_secs, now_nanos, k, ids, sc)
    return ids, sc


def search_in_context(raw, context_of, q, want_ctx, k, w, base_decay, ts_secs, ts_nanos, decay, now_secs, now_nanos,
                      deleted=None, threads=1):
    """`ChronoMind::search_in_context` (src/store.rs:353-377) per query row: exact scan of the context's live members
    with `metric.distance` on RAW rows + `combined_score`. Returns (rc, ids[nq,k], scores[nq,k])."""
    raw, q = _f32(raw), _f32(q)
    nq = q.shape[0]
    ids = np.empty((nq, k), dtype=np.uint32)
    sc = np.empty((nq, k), dtype=np.float32)
    dptr = None
    if deleted is not None:
        deleted = np.ascontiguousarray(deleted, dtype=np.uint8)
        dptr = deleted.ctypes.data_as(C.c_void_p)
    rc = lib().orc_search_in_context(raw, raw.shape[0], raw.shape[1], np.ascontiguousarray(context_of, np.uint32), dptr,
                                     w, base_decay, np.ascontiguousarray(ts_secs, np.uint64),
                                     np.ascontiguousarray(ts_nanos, np.uint32), _f32(decay), now_secs, now_nanos, q, nq,
                                     q.shape[1], np.ascontiguousarray(want_ctx, np.uint32), k, ids, sc, threads)
    return rc, ids, sc


def similar_pairs(raw, threshold, deleted=None, cap=1 << 20):
    """The pair test of `ChronoMind::consolidate` (src/store.rs:531-534): (i, j, similarity), i < j, ascending (i, j)."""
    raw = _f32(raw)
    oi = np.empty(cap, dtype=np.uint32)
    oj = np.empty(cap, dtype=np.uint32)
    osim = np.empty(cap, dtype=np.float32)
    dptr = None
    if deleted is not None:
        deleted = np.ascontiguousarray(deleted, dtype=np.uint8)
        dptr = deleted.ctypes.data_as(C.c_void_p)
    n = int(lib().orc_similar_pairs(raw, raw.shape[0], raw.shape[1], dptr, threshold, oi, oj, osim, cap))
    assert n <= cap
    return oi[:n], oj[:n], osim[:n]


def decay_sweep(importance, decayed_through_nanos, last_access_nanos, decay_rate, base_rate, now_nanos):
    """One `apply_decay` sweep (src/store.rs:428-458); returns the new (importance, decayed_through_nanos)."""
    imp = _f32(importance).copy()
    dt = np.ascontiguousarray(decayed_through_nanos, np.uint64).copy()
    lib().orc_decay_sweep(imp, dt, np.ascontiguousarray(last_access_nanos, np.uint64), _f32(decay_rate), imp.size,
                          base_rate, now_nanos)
    return imp, dt


def sharded_merge(ids, dists, out_len):
    """ids/dists: [n_shards, nq, len] local handles -> merged global (local*n_shards+shard)."""
    ids = np.ascontiguousarray(ids, np.uint32)
    dists = _f32(dists)
    s, nq, ln = ids.shape
    oi = np.empty((nq, out_len), dtype=np.uint32)
    od = np.empty((nq, out_len), dtype=np.float32)
    lib().orc_sharded_merge(ids, dists, s, nq, ln, out_len, oi, od)
    return oi, od


def read_graph_sidecar(path) -> dict:
    """Parse the engine's graph sidecar ("CMGRAPH1", layout in chrono-mind_b200/csrc/api.cu) with numpy."""
    with open(path, "rb") as f:
        blob = f.read()
    assert blob[:8] == b"CMGRAPH1"
    n, = np.frombuffer(blob, np.uint64, 1, 8)
    cap0, capU = np.frombuffer(blob, np.uint32, 2, 16)
    entry, slots = np.frombuffer(blob, np.uint64, 2, 24)
    n, cap0, capU, entry, slots = int(n), int(cap0), int(capU), int(entry), int(slots)
    o = 40
    out = {"n": n, "cap0": cap0, "capU": capU, "entry": entry}
    for name, dt, cnt in (("top_layer", np.uint8, n), ("deg0", np.uint32, n), ("nbr0", np.uint32, n * cap0),
                          ("upper_base", np.uint32, n), ("deg_up", np.uint32, slots),
                          ("nbr_up", np.uint32, slots * capU)):
        out[name] = np.frombuffer(blob, dt, cnt, o).copy()
        o += cnt * np.dtype(dt).itemsize
    return out
