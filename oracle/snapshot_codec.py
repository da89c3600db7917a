"""Snapshot wire-format oracle — TEST INFRASTRUCTURE ONLY.

Pure-Python restatement of the ChronoMind snapshot v2 container
(`src/persistence.rs:24-31,41-64,73-114`) and of the bincode 1.3.3 default
encoding (little-endian, fixed-width integers, u64 length prefixes) of
`SnapshotBody{config: Config, memories: Vec<Memory>}`
(`src/config.rs:11-25,43-71`, `src/types.rs:13-18,32-49,68-73`; `SystemTime`
serialises as `{secs_since_epoch: u64, nanos_since_epoch: u32}` per serde).

The reference cannot produce files in this image (no rustc), and its tests
check the format by round trip only (`tests/persistence_test.rs:31-45`), so the
byte layout is "parity unpinned": it is derived from the published bincode
1.3.3 / serde 1.0 encodings, not from golden bytes. CRC is crc32fast 1.5.0 ==
CRC-32/ISO-HDLC == zlib.crc32.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from typing import List

MAGIC = b"CHRONO1"          # src/persistence.rs:24
FORMAT_VERSION = 2          # src/persistence.rs:25


class InvalidSnapshot(Exception):
    pass


class SerializationError(Exception):
    pass


@dataclass
class IndexParams:          # src/config.rs:11-35
    max_connections: int = 16
    ef_construction: int = 200
    ef_search: int = 50


@dataclass
class Config:               # src/config.rs:43-85
    dimensions: int = 768
    max_memories: int = 100_000
    base_decay_rate: float = 0.1
    temporal_weight: float = 0.3
    similarity_threshold: float = 0.95
    max_relationships: int = 50
    index: IndexParams = field(default_factory=IndexParams)


@dataclass
class Memory:               # src/types.rs:13-18,32-49,68-73
    id: str
    data: List[float]
    ts_secs: int = 0
    ts_nanos: int = 0
    importance: float = 0.5
    context: str = ""
    decay_rate: float = 0.0
    relationships: List[str] = field(default_factory=list)
    access_count: int = 0
    last_access_secs: int = 0
    last_access_nanos: int = 0


def _str(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("<Q", len(b)) + b


def encode_body(config: Config, memories: List[Memory]) -> bytes:
    out = [struct.pack("<QQfffQQQQ", config.dimensions, config.max_memories, config.base_decay_rate,
                       config.temporal_weight, config.similarity_threshold, config.max_relationships,
                       config.index.max_connections, config.index.ef_construction, config.index.ef_search)]
    out.append(struct.pack("<Q", len(memories)))
    for m in memories:
        out.append(_str(m.id))
        out.append(struct.pack("<Q", len(m.data)))
        out.append(struct.pack("<%df" % len(m.data), *m.data))
        out.append(struct.pack("<QI", m.ts_secs, m.ts_nanos))
        out.append(struct.pack("<f", m.importance))
        out.append(_str(m.context))
        out.append(struct.pack("<f", m.decay_rate))
        out.append(struct.pack("<Q", len(m.relationships)))
        for r in m.relationships:
            out.append(_str(r))
        out.append(struct.pack("<I", m.access_count))
        out.append(struct.pack("<QI", m.last_access_secs, m.last_access_nanos))
    return b"".join(out)


def encode_snapshot(config: Config, memories: List[Memory]) -> bytes:
    """save_snapshot (src/persistence.rs:41-64) minus the temp-file dance."""
    body = encode_body(config, memories)
    return MAGIC + bytes([FORMAT_VERSION]) + struct.pack("<I", zlib.crc32(body) & 0xFFFFFFFF) + body


def write_snapshot(path, config: Config, memories: List[Memory]) -> None:
    with open(path, "wb") as f:
        f.write(encode_snapshot(config, memories))


class _Reader:
    def __init__(self, b: bytes):
        self.b, self.o = b, 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise SerializationError("unexpected end of body")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def f32(self):
        return struct.unpack("<f", self.take(4))[0]

    def string(self):
        n = self.u64()
        try:
            return self.take(n).decode("utf-8")
        except UnicodeDecodeError as e:
            raise SerializationError(str(e))


def decode_snapshot(blob: bytes):
    """load_snapshot's header checks + bincode::deserialize (src/persistence.rs:73-114).
    Returns (Config, [Memory]); raises InvalidSnapshot / SerializationError like the
    reference's Error::InvalidSnapshot / Error::Serialization."""
    if len(blob) < 7:
        raise InvalidSnapshot("file too short to be a ChronoMind snapshot")
    if blob[:7] != MAGIC:
        raise InvalidSnapshot("bad magic bytes: not a ChronoMind snapshot")
    if len(blob) < 8:
        raise InvalidSnapshot("missing format version")
    if blob[7] != FORMAT_VERSION:
        raise InvalidSnapshot("unsupported format version %d (supported: %d)" % (blob[7], FORMAT_VERSION))
    if len(blob) < 12:
        raise InvalidSnapshot("missing body checksum")
    expected = struct.unpack("<I", blob[8:12])[0]
    body = blob[12:]
    actual = zlib.crc32(body) & 0xFFFFFFFF
    if actual != expected:
        raise InvalidSnapshot("body checksum mismatch (expected %08x, got %08x): the file is corrupt"
                              % (expected, actual))
    r = _Reader(body)
    cfg = Config()
    cfg.dimensions, cfg.max_memories = r.u64(), r.u64()
    cfg.base_decay_rate, cfg.temporal_weight, cfg.similarity_threshold = r.f32(), r.f32(), r.f32()
    cfg.max_relationships = r.u64()
    cfg.index = IndexParams(r.u64(), r.u64(), r.u64())
    n = r.u64()
    mems = []
    for _ in range(n):
        mid = r.string()
        dl = r.u64()
        data = list(struct.unpack("<%df" % dl, r.take(4 * dl)))
        ts_s, ts_n = r.u64(), r.u32()
        imp = r.f32()
        ctx = r.string()
        dec = r.f32()
        rels = [r.string() for _ in range(r.u64())]
        acc = r.u32()
        la_s, la_n = r.u64(), r.u32()
        mems.append(Memory(mid, data, ts_s, ts_n, imp, ctx, dec, rels, acc, la_s, la_n))
    # bincode::deserialize ignores trailing bytes by default (no reject_trailing_bytes).
    return cfg, mems


def read_snapshot(path):
    with open(path, "rb") as f:
        return decode_snapshot(f.read())
