"""Synthetic workloads of the shapes the reference's tests and benches use.

Distribution parity only: the reference draws from rand 0.8 `StdRng` (ChaCha12), which is not
vendored under the reference tree, so byte parity of the datasets is unpinned.
Generators follow `tests/recall_test.rs:38-67,112-131` and `benches/comparison.rs:80-121`
(corpus and queries share ONE basis).
"""
from __future__ import annotations

import numpy as np


def _normalise(v: np.ndarray) -> np.ndarray:
    n = np.sqrt((v.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32, keepdims=True))
    n[n == 0] = 1.0
    return (v / n).astype(np.float32)


def unit_vectors(rng: np.random.Generator, n: int, dim: int) -> np.ndarray:
    """`unit_vector`: dim draws of U(-1,1), L2-normalised (tests/recall_test.rs:38-47)."""
    return _normalise(rng.uniform(-1.0, 1.0, size=(n, dim)).astype(np.float32))


def uniform_dataset(n: int, nq: int, dim: int, seed: int):
    """`uniform_dataset` (tests/recall_test.rs:112-117)."""
    rng = np.random.default_rng(seed)
    return unit_vectors(rng, n, dim), unit_vectors(rng, nq, dim)


def embedding_rows(rng: np.random.Generator, basis: np.ndarray, n: int, chunk: int = 65536) -> np.ndarray:
    """`embedded_vector`: 16 U(-1,1) coefficients over a fixed basis, normalised (:52-67)."""
    out = np.empty((n, basis.shape[1]), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        coeff = rng.uniform(-1.0, 1.0, size=(e - s, basis.shape[0])).astype(np.float32)
        out[s:e] = _normalise(coeff @ basis)
    return out


def embedding_dataset(n: int, nq: int, seed: int, dim: int = 768, intrinsic: int = 16):
    """`embedding_dataset` (tests/recall_test.rs:119-131): basis, then rows, then queries, one stream."""
    rng = np.random.default_rng(seed)
    basis = unit_vectors(rng, intrinsic, dim)
    return embedding_rows(rng, basis, n), embedding_rows(rng, basis, nq)


def temporal_attrs(n: int, seed: int, now_secs: int, span_days: float = 30.0):
    """Timestamps uniform over the last `span_days`; decay_rate 0 (-> base) for half the rows, else one
    of the `examples/sample_vectors.json` rates."""
    rng = np.random.default_rng(seed ^ 0x7E11)
    age = rng.uniform(0.0, span_days * 86400.0, size=n)
    ts = now_secs - age
    ts_secs = np.floor(ts).astype(np.uint64)
    ts_nanos = ((ts - np.floor(ts)) * 1e9).astype(np.uint32)
    rates = np.array([0.0, 0.0, 0.0, 0.0, 0.05, 0.1, 0.15, 0.2, 0.25], dtype=np.float32)
    decay = rates[rng.integers(0, len(rates), size=n)]
    return ts_secs, ts_nanos, decay


class ChunkedCorpus:
    """A deterministic synthetic corpus generated in independent chunks (one RNG stream per 65,536-row chunk), so
    that a rank can materialise only ITS shard of a corpus larger than host memory (10M x 768 = 30.7 GB) and a
    checker can stream the whole corpus without ever holding it. kind: "emb" (`embedded_vector`,
    tests/recall_test.rs:52-67, one shared basis) or "uniform" (`unit_vector`, :38-47)."""

    CHUNK = 65536

    def __init__(self, kind: str, n_rows: int, dim: int, seed: int, intrinsic: int = 16, chunk: int = 65536):
        assert kind in ("emb", "uniform")
        self.CHUNK = int(chunk)
        self.kind, self.n_rows, self.dim, self.seed = kind, int(n_rows), int(dim), int(seed)
        self.basis = unit_vectors(np.random.default_rng([self.seed, 0xBA515]), intrinsic, dim) if kind == "emb" else None

    @property
    def n_chunks(self) -> int:
        return (self.n_rows + self.CHUNK - 1) // self.CHUNK

    def _rows(self, rng, n):
        if self.kind == "emb":
            coeff = rng.uniform(-1.0, 1.0, size=(n, self.basis.shape[0])).astype(np.float32)
            return _normalise(coeff @ self.basis)
        return unit_vectors(rng, n, self.dim)

    def chunk(self, ci: int) -> np.ndarray:
        """Rows [ci * CHUNK, min(n_rows, (ci + 1) * CHUNK))."""
        n = min(self.CHUNK, self.n_rows - ci * self.CHUNK)
        return self._rows(np.random.default_rng([self.seed, 1, ci]), n)

    def shard(self, rank: int = 0, world: int = 1) -> np.ndarray:
        """Rows with global id = rank (mod world), in id order (src/index/sharded_rwlock.rs:70-74)."""
        n_local = (self.n_rows - rank + world - 1) // world
        out = np.empty((n_local, self.dim), dtype=np.float32)
        o = 0
        for ci in range(self.n_chunks):
            start = ci * self.CHUNK
            first = (rank - start) % world                     # first row of this chunk that belongs to `rank`
            rows = self.chunk(ci)[first::world]
            out[o:o + rows.shape[0]] = rows
            o += rows.shape[0]
        assert o == n_local
        return out

    def queries(self, nq: int, stream: int = 0) -> np.ndarray:
        return self._rows(np.random.default_rng([self.seed, 2, stream]), nq)


# ---------------------------------------------------------------------------------------------------------------------
# Synthetic stores as the reference's own snapshot files
# ---------------------------------------------------------------------------------------------------------------------
def write_snapshot(path: str, rows: np.ndarray, ts_secs, ts_nanos, decay_rate, context_ids=None, n_contexts: int = 1,
                   importance=None, config: dict | None = None, chunk: int = 32768) -> dict:
    """Write `rows` (one memory per row, in row order) as a ChronoMind snapshot v2 file — the container of
    `save_snapshot` (src/persistence.rs:41-64: "CHRONO1" | version 2 | crc32(body) LE | body) around the bincode 1.3.3
    encoding of `SnapshotBody{config, memories}` (src/config.rs:11-25,43-71; src/types.rs:13-18,32-49,68-73; SURVEY.md
    App. B) — so that a million-record synthetic store reaches the engine through `cm_index_from_snapshot`, the way a
    ChronoMind deployment's data would. Ids are "m%09d", contexts "ctx%05d" (fixed widths make a record a fixed-size
    numpy struct: the file is streamed out at memory speed), relationships empty, access_count 0, last_access =
    timestamp. Returns {"bytes", "records", "crc32"}. The byte layout is asserted equal to the checker's record-by-record
    writer in tests/test_host_engine.py."""
    import struct
    import zlib
    n, dim = rows.shape
    cfg = {"dimensions": dim, "max_memories": max(n, 100_000), "base_decay_rate": 0.1, "temporal_weight": 0.3,
           "similarity_threshold": 0.95, "max_relationships": 50, "max_connections": 16, "ef_construction": 200,
           "ef_search": 50}
    cfg.update(config or {})
    rec = np.dtype([("idlen", "<u8"), ("id", "S10"), ("dlen", "<u8"), ("data", "<f4", (dim,)), ("ts_s", "<u8"),
                    ("ts_n", "<u4"), ("imp", "<f4"), ("clen", "<u8"), ("ctx", "S8"), ("decay", "<f4"), ("nrel", "<u8"),
                    ("acc", "<u4"), ("la_s", "<u8"), ("la_n", "<u4")])
    assert rec.itemsize == 8 + 10 + 8 + 4 * dim + 12 + 4 + 8 + 8 + 4 + 8 + 4 + 12
    head = struct.pack("<QQfffQQQQ", cfg["dimensions"], cfg["max_memories"], cfg["base_decay_rate"], cfg["temporal_weight"],
                       cfg["similarity_threshold"], cfg["max_relationships"], cfg["max_connections"],
                       cfg["ef_construction"], cfg["ef_search"]) + struct.pack("<Q", n)
    ctx_labels = np.char.mod("ctx%05d", np.arange(max(1, n_contexts))).astype("S8")
    crc = zlib.crc32(head)
    with open(path, "wb") as f:
        f.write(b"CHRONO1" + bytes([2]) + b"\0\0\0\0")          # checksum patched in below
        f.write(head)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            r = np.zeros(e - s, dtype=rec)
            r["idlen"], r["dlen"], r["clen"] = 10, dim, 8
            r["id"] = np.char.mod("m%09d", np.arange(s, e)).astype("S10")
            r["data"] = rows[s:e]
            r["ts_s"], r["ts_n"] = ts_secs[s:e], ts_nanos[s:e]
            r["la_s"], r["la_n"] = ts_secs[s:e], ts_nanos[s:e]
            r["imp"] = 0.5 if importance is None else importance[s:e]
            r["ctx"] = ctx_labels[0] if context_ids is None else ctx_labels[context_ids[s:e]]
            r["decay"] = decay_rate[s:e]
            b = r.tobytes()
            crc = zlib.crc32(b, crc)
            f.write(b)
        total = f.tell()
        f.seek(8)
        f.write(struct.pack("<I", crc & 0xFFFFFFFF))
    return {"bytes": total, "records": n, "crc32": crc & 0xFFFFFFFF}
