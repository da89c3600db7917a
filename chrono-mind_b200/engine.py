"""Host-side mirror of the reference's interfaces over the C ABI (`include/chronomind_b200.h`).

* `Index` — the PyO3 `Index` of `bindings/python/src/lib.rs:35-123` (`fit`, `set_ef_search`, `query`,
  `batch_query`, `__len__`), same argument meaning and defaults.
* `ChronoMindB200` — the read side of `ChronoMind` (`src/store.rs`): `search` (`:321-346`) for a batch of
  queries with an explicit `now`, plus `search_exact` (the brute-force ground truth of
  `tests/property_test.rs:201-206`) and `search_index` (`VectorIndex::search`).

Everything is a thin ctypes call into `libchronomind_b200.so`; there is no Python or CPU compute path. If the
library (or a CUDA device) is missing the calls raise `CmError`.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchronomind_b200.so")
NO_ID = 0xFFFFFFFF


class CmError(RuntimeError):
    """A non-zero cm_status; `.status` carries the code (names follow `src/error.rs`)."""

    NAMES = {1: "InvalidDimensions", 2: "InvalidVector", 3: "Config", 4: "Io", 5: "Serialization",
             6: "InvalidSnapshot", 7: "IndexFull", 8: "InvalidImportance", 9: "CapacityExceeded",
             50: "InvalidArgument", 51: "State", 100: "Cuda", 101: "Nccl"}

    def __init__(self, status, message):
        super().__init__("%s: %s" % (self.NAMES.get(status, status), message))
        self.status = status
        self.kind = self.NAMES.get(status, str(status))


class IndexParams(C.Structure):
    _fields_ = [("max_connections", C.c_uint64), ("ef_construction", C.c_uint64), ("ef_search", C.c_uint64)]


class Config(C.Structure):
    _fields_ = [("dimensions", C.c_uint64), ("max_memories", C.c_uint64), ("base_decay_rate", C.c_float),
                ("temporal_weight", C.c_float), ("similarity_threshold", C.c_float),
                ("max_relationships", C.c_uint64), ("index", IndexParams)]


class Counters(C.Structure):
    _fields_ = [("dist_evals", C.c_uint64), ("expansions", C.c_uint64), ("list_entries", C.c_uint64),
                ("rescored", C.c_uint64), ("fallback_queries", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("exact_engine", C.c_uint64), ("tie_replays", C.c_uint64), ("failures", C.c_uint64)]


_lib = None

# name -> (restype, argtypes); every symbol include/chronomind_b200.h declares.
_vp, _u32, _u64, _i, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
ABI = {
    "cm_config_default": (None, [C.POINTER(Config)]),
    "cm_index_from_matrix": (_i, [_vp, _u64, _u32, C.POINTER(Config), C.POINTER(_vp)]),
    "cm_index_from_snapshot": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "cm_index_set_attrs": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cm_index_set_contexts": (_i, [_vp, _vp, _vp]),
    "cm_index_context_id": (_u32, [_vp, C.c_char_p]),
    "cm_index_remove": (_i, [_vp, _u32, C.POINTER(_i)]),
    "cm_index_sync_deleted": (_i, [_vp]),
    "cm_index_build_graph": (_i, [_vp, _u64, _i]),
    "cm_index_build_graph_gpu": (_i, [_vp, _u64, _i, _i, _u32]),
    "cm_index_build_graph_from_candidates": (_i, [_vp, _u64, _i, _u32, _vp, _vp]),
    "cm_index_layer_members": (_u64, [_vp, _u64, _u32, _vp, _u64, C.POINTER(_u32)]),
    "cm_index_set_shard": (_i, [_vp, _u32, _u32]),
    "cm_index_upload": (_i, [_vp, _i]),
    "cm_index_free": (None, [_vp]),
    "cm_index_slots": (_u64, [_vp]),
    "cm_index_len": (_u64, [_vp]),
    "cm_index_dim": (_u32, [_vp]),
    "cm_index_get_config": (_i, [_vp, C.POINTER(Config)]),
    "cm_index_external_id": (C.c_char_p, [_vp, _u32]),
    "cm_index_context_label": (C.c_char_p, [_vp, _u32]),
    "cm_index_importance": (_f, [_vp, _u32]),
    "cm_index_check_invariants": (_i, [_vp]),
    "cm_index_entry": (_u64, [_vp]),
    "cm_index_unreachable": (_u64, [_vp]),
    "cm_index_top_layer": (_u32, [_vp, _u32]),
    "cm_index_neighbors": (_u32, [_vp, _u32, _u32, _vp, _u32]),
    "cm_index_vector": (_i, [_vp, _u32, _vp]),
    "cm_index_save_graph": (_i, [_vp, C.c_char_p]),
    "cm_index_load_graph": (_i, [_vp, C.c_char_p]),
    "cm_search_exact": (_i, [_vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "cm_search_hnsw": (_i, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp]),
    "cm_search_temporal": (_i, [_vp, _vp, _u32, _u32, _u32, _u32, _f, _f, _u64, _u32, _i, _vp, _vp, _vp]),
    "cm_search_in_context": (_i, [_vp, _vp, _u32, _u32, _vp, _u32, _f, _f, _u64, _u32, _vp, _vp]),
    "cm_search_in_context_dev": (_i, [_vp, _vp, _u32, _vp, _u32, _f, _f, _u64, _u32, _vp, _vp, _vp]),
    "cm_search_exact_dev": (_i, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "cm_search_hnsw_dev": (_i, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "cm_search_temporal_dev": (_i, [_vp, _vp, _u32, _u32, _u32, _f, _f, _u64, _u32, _i, _vp, _vp, _vp, _vp]),
    "cm_index_dev_status": (_i, [_vp, C.POINTER(_u64), _i]),
    "cm_index_dev_counters": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), _i]),
    "cm_group_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "cm_group_free": (None, [_vp]),
    "cm_group_from_matrix": (_i, [_vp, _vp, _u64, _u32, C.POINTER(Config)]),
    "cm_group_set_attrs": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cm_group_build_graphs": (_i, [_vp, _u64, _i, _i]),
    "cm_group_upload": (_i, [_vp]),
    "cm_group_size": (_u32, [_vp]),
    "cm_group_slots": (_u64, [_vp]),
    "cm_group_len": (_u64, [_vp]),
    "cm_group_shard": (_vp, [_vp, _u32]),
    "cm_group_search_exact": (_i, [_vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "cm_group_search_hnsw": (_i, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp]),
    "cm_group_search_temporal": (_i, [_vp, _vp, _u32, _u32, _u32, _u32, _f, _f, _u64, _u32, _i, _vp, _vp, _vp]),
    "cm_similar_pairs": (_i, [_vp, _f, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "cm_decay_sweep_dev": (_i, [_vp, _vp, _vp, _vp, _u64, _f, _u64, _i, _vp]),
    "cm_merge_topk_dev": (_i, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _i, _vp]),
    "cm_pack_topk_dev": (_i, [_vp, _vp, _u64, _vp, _i, _vp]),
    "cm_merge_topk_keys_dev": (_i, [_vp, _u32, _u32, _u32, _u32, _vp, _vp, _i, _vp]),
    "cm_rerank_dev": (_i, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _f, _f, _u64, _u32, _vp, _vp, _vp, _i, _vp]),
    "cm_metric_preprocess_dev": (_i, [_vp, _u64, _u32, _vp, _i, _vp]),
    "cm_metric_distance_prepared_dev": (_i, [_vp, _vp, _u64, _u32, _vp, _i, _vp]),
    "cm_metric_name": (C.c_char_p, []),
    "cm_set_exact_engine": (_i, [_vp, _i]),
    "cm_last_counters": (_i, [C.POINTER(Counters)]),
    "cm_profile_enable": (_i, [_i]),
    "cm_profile_read": (_i, [C.POINTER(C.c_float), C.POINTER(_u32)]),
    "cm_launch_count": (_u64, []),
    "cm_last_error": (C.c_char_p, []),
    "cm_version": (C.c_char_p, []),
}


def native():
    """Load libchronomind_b200.so (built in-tree by `__graft_entry__.build()` / `csrc/Makefile`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmError(100, "native library missing: %s (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no fallback path" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in ABI.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _check(status):
    if status != 0:
        raise CmError(status, native().cm_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def last_counters() -> dict:
    c = Counters()
    native().cm_last_counters(C.byref(c))
    return {k: int(getattr(c, k)) for k, _ in Counters._fields_}


def profile_enable(on: bool):
    native().cm_profile_enable(1 if on else 0)


def profile_read():
    """(summed device ms of the bracketed dominant-kernel launches, their count) since the last read."""
    ms, n = C.c_float(0), C.c_uint32(0)
    native().cm_profile_read(C.byref(ms), C.byref(n))
    return float(ms.value), int(n.value)


def launch_count() -> int:
    return int(native().cm_launch_count())


def default_config(**overrides) -> Config:
    c = Config()
    native().cm_config_default(C.byref(c))
    for k, v in overrides.items():
        if k in ("max_connections", "ef_construction", "ef_search"):
            setattr(c.index, k, v)
        else:
            setattr(c, k, v)
    return c


def _now():
    t = time.time_ns()
    return t // 1_000_000_000, t % 1_000_000_000


class ChronoMindB200:
    """Device-resident read replica of a ChronoMind store."""

    def __init__(self, handle, device=None):
        self._h = handle
        self.device = device

    # ---- construction -------------------------------------------------------------------------------------
    @classmethod
    def from_matrix(cls, x, config: Config | None = None, ts_secs=None, ts_nanos=None, decay_rate=None,
                    deleted=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError("x must be [n, dim]")
        h = C.c_void_p()
        cfg = config if config is not None else default_config()
        _check(native().cm_index_from_matrix(_ptr(x), x.shape[0], x.shape[1], C.byref(cfg), C.byref(h)))
        self = cls(h)
        if any(a is not None for a in (ts_secs, ts_nanos, decay_rate, deleted)):
            self.set_attrs(ts_secs, ts_nanos, decay_rate, deleted)
        return self

    @classmethod
    def from_snapshot(cls, path):
        """`load_snapshot` (src/persistence.rs:73-123) minus the index build (see `build_graph`)."""
        h = C.c_void_p()
        _check(native().cm_index_from_snapshot(os.fsencode(path), C.byref(h)))
        return cls(h)

    def set_attrs(self, ts_secs=None, ts_nanos=None, decay_rate=None, deleted=None):
        n = self.slots
        arrs = []
        for a, dt in ((ts_secs, np.uint64), (ts_nanos, np.uint32), (decay_rate, np.float32), (deleted, np.uint8)):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=dt)
                if a.shape != (n,):
                    raise ValueError("attribute arrays must have one entry per handle")
            arrs.append(a)
        _check(native().cm_index_set_attrs(self._h, *[_ptr(a) for a in arrs]))

    def similar_pairs(self, threshold=None, cap: int = 1 << 20):
        """The pair test of `ChronoMind::consolidate` (src/store.rs:531-534) for every pair of live memories:
        (i[], j[], similarity[]) with i < j and similarity > threshold (default: the config's similarity_threshold)."""
        thr = self.config.similarity_threshold if threshold is None else threshold
        while True:
            i = np.empty(cap, dtype=np.uint32)
            j = np.empty(cap, dtype=np.uint32)
            sim = np.empty(cap, dtype=np.float32)
            cnt = C.c_uint64(0)
            _check(native().cm_similar_pairs(self._h, thr, _ptr(i), _ptr(j), _ptr(sim), cap, C.byref(cnt)))
            if cnt.value <= cap:
                return i[:cnt.value], j[:cnt.value], sim[:cnt.value]
            cap = int(cnt.value)

    def set_contexts(self, raw_x, context_ids=None):
        """Store-level contexts: the rows as stored (not preprocessed) and one context id per handle."""
        raw_x = np.ascontiguousarray(raw_x, dtype=np.float32)
        if context_ids is not None:
            context_ids = np.ascontiguousarray(context_ids, dtype=np.uint32)
        if raw_x.shape != (self.slots, self.dim) or (context_ids is not None and context_ids.shape != (self.slots,)):
            raise ValueError("raw_x must be [slots, dim] and context_ids [slots]")
        _check(native().cm_index_set_contexts(self._h, _ptr(raw_x), _ptr(context_ids)))
        return self

    def context_id(self, label: str):
        """Id of a context label of a snapshot-loaded store (None when no memory carries it)."""
        r = int(native().cm_index_context_id(self._h, label.encode("utf-8")))
        return None if r == NO_ID else r

    def remove(self, handle: int) -> bool:
        r = C.c_int(0)
        _check(native().cm_index_remove(self._h, handle, C.byref(r)))
        return bool(r.value)

    def sync_deleted(self):
        """Push the tombstones recorded by `remove` since the last upload to the device replica (bytes, not a re-upload)."""
        _check(native().cm_index_sync_deleted(self._h))
        return self

    def build_graph(self, layer_seed: int = 0x5EED, threads: int = 1):
        _check(native().cm_index_build_graph(self._h, layer_seed & (2 ** 64 - 1), threads))
        return self

    def build_graph_gpu(self, layer_seed: int = 0x5EED, device: int = 0, threads: int = 0, n_candidates: int = 0):
        """GPU-assisted construction: per-layer predecessor kNN, select_diverse and reverse-edge re-selection on the device
        (`cm_index_build_graph_gpu`); the host keeps the reverse-edge CSR and the reachability check."""
        _check(native().cm_index_build_graph_gpu(self._h, layer_seed & (2 ** 64 - 1), device, threads, n_candidates))
        return self

    def layer_members(self, layer_seed: int):
        """Per layer, the handles that drew at least that layer (ascending) for `layer_seed`."""
        nl = C.c_uint32(0)
        native().cm_index_layer_members(self._h, layer_seed & (2 ** 64 - 1), 0, None, 0, C.byref(nl))
        out = []
        for layer in range(nl.value):
            cnt = int(native().cm_index_layer_members(self._h, layer_seed & (2 ** 64 - 1), layer, None, 0, None))
            buf = np.empty(cnt, dtype=np.uint32)
            native().cm_index_layer_members(self._h, layer_seed & (2 ** 64 - 1), layer, _ptr(buf), cnt, None)
            out.append(buf)
        return out

    def build_graph_from_candidates(self, layer_seed: int, candidates, threads: int = 0):
        """Host half of the GPU-assisted construction: `candidates[l]` is a [members_l, width_l] uint32 array of
        candidate handles for the members of layer l (order of `layer_members`)."""
        cands = [np.ascontiguousarray(c, dtype=np.uint32) for c in candidates]
        ptrs = (C.c_void_p * len(cands))(*[c.ctypes.data for c in cands])
        widths = np.array([c.shape[1] if c.ndim == 2 else 0 for c in cands], dtype=np.uint32)
        _check(native().cm_index_build_graph_from_candidates(self._h, layer_seed & (2 ** 64 - 1), threads, len(cands),
                                                             C.cast(ptrs, C.c_void_p), _ptr(widths)))
        return self

    def save_graph(self, path):
        _check(native().cm_index_save_graph(self._h, os.fsencode(path)))

    def load_graph(self, path):
        _check(native().cm_index_load_graph(self._h, os.fsencode(path)))
        return self

    def set_shard(self, rank: int, count: int):
        _check(native().cm_index_set_shard(self._h, rank, count))
        return self

    def upload(self, device: int = 0):
        _check(native().cm_index_upload(self._h, device))
        self.device = device
        return self

    def dev_status(self, reset: bool = True) -> int:
        """Queries the asynchronous (`*_cuda`) searches enqueued so far could not answer like the reference (waits for
        them). 0 in normal operation."""
        f = C.c_uint64(0)
        _check(native().cm_index_dev_status(self._h, C.byref(f), 1 if reset else 0))
        return int(f.value)

    def dev_counters(self, reset: bool = True) -> dict:
        """Work counters of the asynchronous exact searches enqueued so far (waits for them)."""
        r, f = C.c_uint64(0), C.c_uint64(0)
        _check(native().cm_index_dev_counters(self._h, C.byref(r), C.byref(f), 1 if reset else 0))
        return {"rescored": int(r.value), "fallback_queries": int(f.value)}

    def set_exact_engine(self, engine: int):
        _check(native().cm_set_exact_engine(self._h, engine))
        return self

    def close(self):
        if self._h:
            native().cm_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ----------------------------------------------------------------------------------------
    def __len__(self):
        return int(native().cm_index_len(self._h))

    @property
    def slots(self) -> int:
        return int(native().cm_index_slots(self._h))

    @property
    def dim(self) -> int:
        return int(native().cm_index_dim(self._h))

    @property
    def config(self) -> Config:
        c = Config()
        _check(native().cm_index_get_config(self._h, C.byref(c)))
        return c

    def external_id(self, handle: int) -> str:
        return native().cm_index_external_id(self._h, handle).decode("utf-8")

    def context_label(self, handle: int) -> str:
        return native().cm_index_context_label(self._h, handle).decode("utf-8")

    def importance(self, handle: int) -> float:
        return float(native().cm_index_importance(self._h, handle))

    def check_invariants(self) -> int:
        return int(native().cm_index_check_invariants(self._h))

    def unreachable(self) -> int:
        """Layer-0 nodes the entry point cannot reach (0: every node is searchable)."""
        return int(native().cm_index_unreachable(self._h))

    def entry(self):
        w = int(native().cm_index_entry(self._h))
        return None if w == 2 ** 64 - 1 else (w & 0xFFFFFFFF, w >> 32)

    def top_layer(self, handle: int) -> int:
        return int(native().cm_index_top_layer(self._h, handle))

    def neighbors(self, handle: int, layer: int) -> np.ndarray:
        out = np.empty(256, dtype=np.uint32)
        n = native().cm_index_neighbors(self._h, handle, layer, _ptr(out), out.size)
        return out[:n].copy()

    def vector(self, handle: int) -> np.ndarray:
        out = np.empty(self.dim, dtype=np.float32)
        _check(native().cm_index_vector(self._h, handle, _ptr(out)))
        return out

    # ---- search (host buffers) ----------------------------------------------------------------------------------
    @staticmethod
    def _queries(q):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        return q

    def search_exact(self, q, k: int):
        """Brute-force ground truth: (ids[nq,k], dist[nq,k]) ascending (distance, handle)."""
        q = self._queries(q)
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        _check(native().cm_search_exact(self._h, _ptr(q), q.shape[0], q.shape[1], k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def search_index(self, q, ef: int, k: int | None = None):
        """`VectorIndex::search(query, ef)` per row, first k (default ef) results."""
        q = self._queries(q)
        k = ef if k is None else k
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        _check(native().cm_search_hnsw(self._h, _ptr(q), q.shape[0], q.shape[1], ef, k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def search(self, q, k: int, now=None, exact: bool = False, ef_search=None, temporal_weight=None,
               base_decay_rate=None, want_dist=False):
        """`ChronoMind::search(query, k)` per row -> (handles[nq,k], scores[nq,k]) ascending score.
        `now` = (secs, nanos) since the epoch; the reference reads the wall clock (src/store.rs:324)."""
        q = self._queries(q)
        cfg = self.config
        efs = cfg.index.ef_search if ef_search is None else ef_search
        w = cfg.temporal_weight if temporal_weight is None else temporal_weight
        base = cfg.base_decay_rate if base_decay_rate is None else base_decay_rate
        secs, nanos = _now() if now is None else now
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        score = np.empty((q.shape[0], k), dtype=np.float32)
        dist = np.empty((q.shape[0], k), dtype=np.float32) if want_dist else None
        _check(native().cm_search_temporal(self._h, _ptr(q), q.shape[0], q.shape[1], k, efs, w, base, secs, nanos,
                                           1 if exact else 0, _ptr(ids), _ptr(score), _ptr(dist)))
        return (ids, score, dist) if want_dist else (ids, score)

    def search_in_context(self, context, q, k: int, now=None, temporal_weight=None, base_decay_rate=None):
        """`ChronoMind::search_in_context(context, query, k)` per row (src/store.rs:353-377) -> (handles, scores).
        `context`: a label (snapshot-loaded stores), an id, or one id per query."""
        q = self._queries(q)
        cfg = self.config
        w = cfg.temporal_weight if temporal_weight is None else temporal_weight
        base = cfg.base_decay_rate if base_decay_rate is None else base_decay_rate
        secs, nanos = _now() if now is None else now
        if isinstance(context, str):
            cid = self.context_id(context)
            context = NO_ID if cid is None else cid
        want = np.ascontiguousarray(np.broadcast_to(np.asarray(context, dtype=np.uint32), (q.shape[0],)))
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        score = np.empty((q.shape[0], k), dtype=np.float32)
        _check(native().cm_search_in_context(self._h, _ptr(q), q.shape[0], q.shape[1], _ptr(want), k, w, base, secs, nanos,
                                             _ptr(ids), _ptr(score)))
        return ids, score

    # ---- search (torch CUDA tensors, asynchronous on the current stream) ------------------------------------------
    def search_exact_cuda(self, q, k: int, out_ids=None, out_dist=None):
        import torch
        nq = q.shape[0]
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.shape[1] == self.dim
        ids = out_ids if out_ids is not None else torch.empty((nq, k), dtype=torch.int32, device=q.device)
        dist = out_dist if out_dist is not None else torch.empty((nq, k), dtype=torch.float32, device=q.device)
        s = torch.cuda.current_stream(q.device).cuda_stream
        _check(native().cm_search_exact_dev(self._h, q.data_ptr(), nq, k, ids.data_ptr(), dist.data_ptr(), s))
        return ids, dist

    def search_index_cuda(self, q, ef: int, k: int, out_ids=None, out_dist=None):
        import torch
        nq = q.shape[0]
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.shape[1] == self.dim
        ids = out_ids if out_ids is not None else torch.empty((nq, k), dtype=torch.int32, device=q.device)
        dist = out_dist if out_dist is not None else torch.empty((nq, k), dtype=torch.float32, device=q.device)
        s = torch.cuda.current_stream(q.device).cuda_stream
        _check(native().cm_search_hnsw_dev(self._h, q.data_ptr(), nq, ef, k, ids.data_ptr(), dist.data_ptr(), s))
        return ids, dist

    def search_cuda(self, q, k: int, now, exact=False, ef_search=None, temporal_weight=None, base_decay_rate=None):
        import torch
        nq = q.shape[0]
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.shape[1] == self.dim
        cfg = self.config
        efs = cfg.index.ef_search if ef_search is None else ef_search
        w = cfg.temporal_weight if temporal_weight is None else temporal_weight
        base = cfg.base_decay_rate if base_decay_rate is None else base_decay_rate
        ids = torch.empty((nq, k), dtype=torch.int32, device=q.device)
        score = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        s = torch.cuda.current_stream(q.device).cuda_stream
        _check(native().cm_search_temporal_dev(self._h, q.data_ptr(), nq, k, efs, w, base, now[0], now[1],
                                               1 if exact else 0, ids.data_ptr(), score.data_ptr(), dist.data_ptr(), s))
        return ids, score, dist


class ShardedChronoMindB200:
    """One store sharded by vector id over several GPUs of THIS process (`cm_group`): the read side of
    `ShardedRwLockHnsw` (src/index/sharded_rwlock.rs:43-94) — one object, one call per batch, global handles."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(native().cm_group_create(C.cast(devs, C.c_void_p), len(devices), C.byref(h)))
        self._h, self.devices, self.dim = h, list(devices), None

    @classmethod
    def from_matrix(cls, devices, x, config: Config | None = None, ts_secs=None, ts_nanos=None, decay_rate=None, deleted=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        self = cls(devices)
        cfg = config if config is not None else default_config(dimensions=x.shape[1])
        _check(native().cm_group_from_matrix(self._h, _ptr(x), x.shape[0], x.shape[1], C.byref(cfg)))
        self.dim, self._cfg = x.shape[1], cfg
        if any(a is not None for a in (ts_secs, ts_nanos, decay_rate, deleted)):
            arrs = [None if a is None else np.ascontiguousarray(a, dtype=dt)
                    for a, dt in ((ts_secs, np.uint64), (ts_nanos, np.uint32), (decay_rate, np.float32), (deleted, np.uint8))]
            _check(native().cm_group_set_attrs(self._h, *[_ptr(a) for a in arrs]))
        return self

    def build_graphs(self, layer_seed: int = 0x5EED, threads: int = 0, gpu: bool = True):
        _check(native().cm_group_build_graphs(self._h, layer_seed & (2 ** 64 - 1), threads, 1 if gpu else 0))
        return self

    def upload(self):
        _check(native().cm_group_upload(self._h))
        return self

    def shard(self, i: int) -> "ChronoMindB200":
        """Borrowed view of shard i (owned by the group: never closed through this object)."""
        view = ChronoMindB200(C.c_void_p(native().cm_group_shard(self._h, i)), self.devices[i])
        view.close = lambda: None
        return view

    def __len__(self):
        return int(native().cm_group_len(self._h))

    @property
    def slots(self) -> int:
        return int(native().cm_group_slots(self._h))

    def close(self):
        if self._h:
            native().cm_group_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search_exact(self, q, k: int):
        q = ChronoMindB200._queries(q)
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        _check(native().cm_group_search_exact(self._h, _ptr(q), q.shape[0], q.shape[1], k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def search_index(self, q, ef: int, k: int | None = None):
        q = ChronoMindB200._queries(q)
        k = ef if k is None else k
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        _check(native().cm_group_search_hnsw(self._h, _ptr(q), q.shape[0], q.shape[1], ef, k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def search(self, q, k: int, now=None, exact: bool = False, ef_search=None, temporal_weight=None, base_decay_rate=None,
               want_dist=False):
        q = ChronoMindB200._queries(q)
        cfg = self._cfg
        efs = cfg.index.ef_search if ef_search is None else ef_search
        w = cfg.temporal_weight if temporal_weight is None else temporal_weight
        base = cfg.base_decay_rate if base_decay_rate is None else base_decay_rate
        secs, nanos = _now() if now is None else now
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        score = np.empty((q.shape[0], k), dtype=np.float32)
        dist = np.empty((q.shape[0], k), dtype=np.float32) if want_dist else None
        _check(native().cm_group_search_temporal(self._h, _ptr(q), q.shape[0], q.shape[1], k, efs, w, base, secs, nanos,
                                                 1 if exact else 0, _ptr(ids), _ptr(score), _ptr(dist)))
        return (ids, score, dist) if want_dist else (ids, score)


def merge_topk_cuda(ids, dist, out_len: int):
    """Shard merge (src/index/sharded_rwlock.rs:81-94) of allgathered lists: ids/dist [n_shards, nq, len] CUDA
    tensors of LOCAL handles -> ([nq, out_len] global handles, distances)."""
    import torch
    s_, nq, ln = ids.shape
    oi = torch.empty((nq, out_len), dtype=torch.int32, device=ids.device)
    od = torch.empty((nq, out_len), dtype=torch.float32, device=ids.device)
    st = torch.cuda.current_stream(ids.device).cuda_stream
    _check(native().cm_merge_topk_dev(ids.data_ptr(), dist.data_ptr(), s_, nq, ln, out_len, oi.data_ptr(),
                                      od.data_ptr(), ids.device.index or 0, st))
    return oi, od


def pack_topk_cuda(ids, dist, out=None):
    """(distance, local handle) lists -> sortable int64 keys of the same shape (one array to all-gather)."""
    import torch
    keys = out if out is not None else torch.empty(ids.shape, dtype=torch.int64, device=ids.device)
    st = torch.cuda.current_stream(ids.device).cuda_stream
    _check(native().cm_pack_topk_dev(ids.data_ptr(), dist.data_ptr(), ids.numel(), keys.data_ptr(), ids.device.index or 0, st))
    return keys


def merge_topk_keys_cuda(keys, out_len: int):
    """Shard merge of all-gathered packed keys `[n_shards, nq, len]` -> ([nq, out_len] global handles, distances)."""
    import torch
    s_, nq, ln = keys.shape
    oi = torch.empty((nq, out_len), dtype=torch.int32, device=keys.device)
    od = torch.empty((nq, out_len), dtype=torch.float32, device=keys.device)
    st = torch.cuda.current_stream(keys.device).cuda_stream
    _check(native().cm_merge_topk_keys_dev(keys.data_ptr(), s_, nq, ln, out_len, oi.data_ptr(), od.data_ptr(),
                                           keys.device.index or 0, st))
    return oi, od


def rerank_cuda(pool_ids, pool_dist, k: int, ts_secs, ts_nanos, decay_rate, temporal_weight: float,
                base_decay_rate: float, now):
    """The rerank tail of `ChronoMind::search` (src/store.rs:327-344) on a merged pool of GLOBAL handles:
    pool_ids/pool_dist `[nq, pool]` CUDA tensors (ascending (distance, handle), CM_NO_ID padded); ts_secs (int64 bits
    of the u64 seconds), ts_nanos (int32), decay_rate (float32): CUDA tensors indexed by global handle.
    Returns ([nq, k] handles, scores, distances)."""
    import torch
    nq, pool = pool_ids.shape
    dev = pool_ids.device
    oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
    osc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _check(native().cm_rerank_dev(pool_ids.data_ptr(), pool_dist.data_ptr(), nq, pool, k, ts_secs.data_ptr(),
                                  ts_nanos.data_ptr(), decay_rate.data_ptr(), temporal_weight, base_decay_rate,
                                  now[0], now[1], oi.data_ptr(), osc.data_ptr(), od.data_ptr(), dev.index or 0, st))
    return oi, osc, od


def decay_sweep_cuda(importance, decayed_through_nanos, last_access_nanos, decay_rate, base_decay_rate: float,
                     now_nanos: int):
    """One `apply_decay` sweep (src/store.rs:428-458) in place over CUDA tensors: importance f32[n],
    decayed_through_nanos int64[n] (u64 bits), last_access_nanos int64[n], decay_rate f32[n]."""
    import torch
    dev = importance.device
    st = torch.cuda.current_stream(dev).cuda_stream
    _check(native().cm_decay_sweep_dev(importance.data_ptr(), decayed_through_nanos.data_ptr(), last_access_nanos.data_ptr(),
                                       decay_rate.data_ptr(), importance.numel(), base_decay_rate, now_nanos,
                                       dev.index or 0, st))


class Index:
    """PyO3 `Index` (bindings/python/src/lib.rs:35-123) on the GPU engine.

    `Index(dim, max_connections=16, ef_construction=100, ef_search=50, seed=0x5EED)`; `fit(data)` inserts row i
    as handle i (single-threaded build, deterministic, like the reference); `query` / `batch_query` return the
    handles of the n nearest neighbours with `ef = max(ef_search, n)`.
    """

    def __init__(self, dim, max_connections=16, ef_construction=100, ef_search=50, seed=0x5EED, device=0,
                 build_threads=1):
        self.dim = dim
        self.ef_search = ef_search
        self.seed = seed
        self.device = device
        self.build_threads = build_threads
        self._cfg = default_config(dimensions=dim, max_connections=max_connections,
                                   ef_construction=ef_construction, ef_search=ef_search)
        self._store = None

    def fit(self, data):
        self._store = ChronoMindB200.from_matrix(data, self._cfg)
        self._store.build_graph(self.seed, self.build_threads).upload(self.device)

    def set_ef_search(self, ef):
        self.ef_search = ef

    def query(self, v, n):
        return self.batch_query(np.asarray(v, dtype=np.float32)[None, :], n)[0]

    def batch_query(self, queries, n):
        ef = max(self.ef_search, n)
        ids, _ = self._store.search_index(queries, ef, n)
        return [[int(h) for h in row if h != NO_ID] for row in ids]

    def __len__(self):
        return 0 if self._store is None else len(self._store)
