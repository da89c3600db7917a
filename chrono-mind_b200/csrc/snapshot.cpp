// snapshot.cpp — reader for the reference's snapshot container (src/persistence.rs:73-114).
//
//   0   7  "CHRONO1"                       persistence.rs:24
//   7   1  format version == 2             persistence.rs:25
//   8   4  CRC-32 (ISO-HDLC, crc32fast) of the body, little endian
//   12  …  bincode 1.3.3 default encoding of SnapshotBody{config, memories}:
//          little-endian, fixed-width integers, u64 length prefixes, usize as u64,
//          SystemTime as {secs_since_epoch: u64, nanos_since_epoch: u32}.
//
// Error mapping follows load_snapshot: short/bad magic/version/checksum -> InvalidSnapshot, open/read
// failures -> Io, body decode failures -> Serialization; then Config::validate (src/config.rs:97-143) and
// Memory::validate (src/types.rs:93-121) as `ChronoMind::new` / `insert` would apply them.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <unistd.h>

#include "cm_internal.h"

namespace cm {
namespace {

// CRC-32/ISO-HDLC, slicing-by-8.
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};

uint32_t crc32(const uint8_t* p, size_t n) {
  static const CrcTables T;
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t a, b;
    std::memcpy(&a, p, 4);
    std::memcpy(&b, p + 4, 4);
    a ^= c;
    c = T.t[7][a & 0xFF] ^ T.t[6][(a >> 8) & 0xFF] ^ T.t[5][(a >> 16) & 0xFF] ^ T.t[4][a >> 24] ^
        T.t[3][b & 0xFF] ^ T.t[2][(b >> 8) & 0xFF] ^ T.t[1][(b >> 16) & 0xFF] ^ T.t[0][b >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}

// crc(A || B) from crc(A), crc(B) and |B|: multiply crc(A) by x^(8|B|) modulo the CRC polynomial, by squaring a
// GF(2) 32x32 operator (the classic combine construction), then xor crc(B).
uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  for (; vec; vec >>= 1, ++mat)
    if (vec & 1) sum ^= *mat;
  return sum;
}
void gf2_square(uint32_t* sq, const uint32_t* mat) {
  for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]);
}
uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;  // operator for one zero BIT
  for (uint32_t n = 1, row = 1; n < 32; ++n, row <<= 1) odd[n] = row;
  gf2_square(even, odd);  // two bits
  gf2_square(odd, even);  // four bits
  do {                    // apply len2 zero BYTES: first squaring gives the one-byte operator
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}

unsigned host_threads() {
  unsigned t = std::thread::hardware_concurrency();
  return t ? (t > 32 ? 32 : t) : 1;
}

// fn(part, parts) on `parts` threads.
template <class F>
void run_parts(unsigned parts, F fn) {
  if (parts <= 1) {
    fn(0u, 1u);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 1; t < parts; ++t) th.emplace_back(fn, t, parts);
  fn(0u, parts);
  for (auto& x : th) x.join();
}

// CRC of a large buffer: slices in parallel, combined in order.
uint32_t crc32_parallel(const uint8_t* p, size_t n) {
  const unsigned parts = n < (8u << 20) ? 1u : host_threads();
  if (parts == 1) return crc32(p, n);
  std::vector<uint32_t> part_crc(parts);
  std::vector<size_t> part_len(parts);
  const size_t step = (n / parts + 63) & ~(size_t)63;
  run_parts(parts, [&](unsigned t, unsigned) {
    const size_t lo = std::min(n, (size_t)t * step), hi = t + 1 == parts ? n : std::min(n, lo + step);
    part_len[t] = hi - lo;
    part_crc[t] = crc32(p + lo, hi - lo);
  });
  uint32_t c = part_crc[0];
  for (unsigned t = 1; t < parts; ++t) c = crc32_combine(c, part_crc[t], part_len[t]);
  return c;
}

// The snapshot file, mapped read-only (no 3 GB copy for a million 768-d records).
struct MappedFile {
  const uint8_t* p = nullptr;
  size_t n = 0;
  ~MappedFile() {
    if (p && n) munmap(const_cast<uint8_t*>(p), n);
  }
};

struct Cursor {
  const uint8_t* p;
  size_t left;
  bool ok = true;
  bool take(void* dst, size_t n) {
    if (!ok || n > left) return ok = false;
    std::memcpy(dst, p, n);
    p += n;
    left -= n;
    return true;
  }
  uint64_t u64() {
    uint64_t v = 0;
    take(&v, 8);
    return v;
  }
  uint32_t u32() {
    uint32_t v = 0;
    take(&v, 4);
    return v;
  }
  float f32() {
    float v = 0;
    take(&v, 4);
    return v;
  }
  bool str(std::string* s) {
    uint64_t n = u64();
    if (!ok || n > left) return ok = false;
    s->assign((const char*)p, (size_t)n);
    p += n;
    left -= n;
    return true;
  }
};

bool valid_utf8(const std::string& s) {  // bincode rejects non-UTF-8 strings
  const unsigned char* p = (const unsigned char*)s.data();
  size_t n = s.size(), i = 0;
  while (i < n) {
    unsigned char c = p[i];
    size_t len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
    if (!len || i + len > n) return false;
    for (size_t j = 1; j < len; ++j)
      if ((p[i + j] & 0xC0) != 0x80) return false;
    i += len;
  }
  return true;
}

}  // namespace

cm_status validate_config(const cm_config& c) {  // src/config.rs:97-143
  auto bad = [](const char* m) {
    set_error(std::string("invalid configuration: ") + m);
    return CM_ERR_CONFIG;
  };
  if (c.dimensions == 0) return bad("dimensions must be greater than 0");
  if (c.max_memories == 0) return bad("max_memories must be greater than 0");
  if (!std::isfinite(c.base_decay_rate) || c.base_decay_rate <= 0.0f)
    return bad("base_decay_rate must be a positive finite number");
  if (!std::isfinite(c.temporal_weight) || c.temporal_weight < 0.0f || c.temporal_weight > 1.0f)
    return bad("temporal_weight must be within [0.0, 1.0]");
  if (!std::isfinite(c.similarity_threshold) || c.similarity_threshold <= 0.0f || c.similarity_threshold >= 1.0f)
    return bad("similarity_threshold must be within (0.0, 1.0)");
  if (c.max_relationships == 0) return bad("max_relationships must be greater than 0");
  if (c.index.max_connections < 2) return bad("index.max_connections must be at least 2");
  if (c.index.ef_construction < c.index.max_connections)
    return bad("index.ef_construction must be at least index.max_connections");
  if (c.index.ef_search == 0) return bad("index.ef_search must be greater than 0");
  return CM_OK;
}

cm_status read_snapshot(const char* path, SnapshotData* out) {
  // CM_LOAD_TIMING=1: stage times of the reader on stderr
  const bool timing = getenv("CM_LOAD_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[read_snapshot] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };
  MappedFile file;
  {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) {
      set_error(std::string("io error: ") + std::strerror(errno) + ": " + path);
      return CM_ERR_IO;
    }
    struct stat sb;
    if (::fstat(fd, &sb) != 0 || sb.st_size < 0) {
      ::close(fd);
      set_error("io error: cannot size snapshot");
      return CM_ERR_IO;
    }
    file.n = (size_t)sb.st_size;
    if (file.n) {
      void* m = ::mmap(nullptr, file.n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) {
        ::close(fd);
        file.n = 0;
        set_error(std::string("io error: ") + std::strerror(errno) + ": " + path);
        return CM_ERR_IO;
      }
      file.p = static_cast<const uint8_t*>(m);
      ::madvise(m, file.n, MADV_SEQUENTIAL);
    }
    ::close(fd);
  }
  struct Blob {  // what the header checks below read
    const uint8_t* p;
    size_t n;
    size_t size() const { return n; }
    const uint8_t* data() const { return p; }
    uint8_t operator[](size_t i) const { return p[i]; }
  } blob{file.p, file.n};
  auto invalid = [](const std::string& m) {
    set_error("invalid snapshot: " + m);
    return CM_ERR_INVALID_SNAPSHOT;
  };
  if (blob.size() < 7) return invalid("file too short to be a ChronoMind snapshot");
  if (std::memcmp(blob.data(), "CHRONO1", 7) != 0) return invalid("bad magic bytes: not a ChronoMind snapshot");
  if (blob.size() < 8) return invalid("missing format version");
  if (blob[7] != 2)
    return invalid("unsupported format version " + std::to_string((int)blob[7]) + " (supported: 2)");
  if (blob.size() < 12) return invalid("missing body checksum");
  uint32_t expected;
  std::memcpy(&expected, blob.data() + 8, 4);
  uint32_t actual = crc32_parallel(blob.data() + 12, blob.size() - 12);
  lap("map + CRC-32");
  if (actual != expected) {
    char buf[128];
    std::snprintf(buf, sizeof buf, "body checksum mismatch (expected %08x, got %08x): the file is corrupt", expected,
                  actual);
    return invalid(buf);
  }

  Cursor c{blob.data() + 12, blob.size() - 12};
  auto ser = [](const char* m) {
    set_error(std::string("serialization error: ") + m);
    return CM_ERR_SERIALIZATION;
  };
  cm_config& cfg = out->config;
  cfg.dimensions = c.u64();
  cfg.max_memories = c.u64();
  cfg.base_decay_rate = c.f32();
  cfg.temporal_weight = c.f32();
  cfg.similarity_threshold = c.f32();
  cfg.max_relationships = c.u64();
  cfg.index.max_connections = c.u64();
  cfg.index.ef_construction = c.u64();
  cfg.index.ef_search = c.u64();
  uint64_t count = c.u64();
  if (!c.ok) return ser("unexpected end of body (config)");
  // Each memory needs > 60 bytes; reject absurd counts before reserving.
  if (count > c.left / 60 + 1) return ser("memory count exceeds body size");
  out->memories.clear();
  out->memories.reserve((size_t)count);
  out->vectors.clear();
  std::vector<const uint8_t*> payload;  // where each record's floats sit in the file
  payload.reserve((size_t)count);
  uint64_t total_floats = 0;
  for (uint64_t i = 0; i < count; ++i) {
    SnapshotMemory m;
    if (!c.str(&m.id) || !valid_utf8(m.id)) return ser("bad memory id string");
    uint64_t len = c.u64();
    if (!c.ok || len > c.left / 4) return ser("vector length exceeds body size");
    // decode first, judge later: bincode reads the whole body before `load_snapshot` inserts anything, so a
    // truncated tail is a Serialization error even when an earlier record would fail validation
    m.data_offset = total_floats;
    m.data_len = len;
    payload.push_back(c.p);  // copied (and checked for NaN/Inf) in parallel once every record's place is known
    total_floats += len;
    c.p += len * 4;
    c.left -= len * 4;
    m.ts_secs = c.u64();
    m.ts_nanos = c.u32();
    m.importance = c.f32();
    if (!c.str(&m.context) || !valid_utf8(m.context)) return ser("bad context string");
    m.decay_rate = c.f32();
    uint64_t nrel = c.u64();
    if (!c.ok || nrel > c.left / 8) return ser("relationship count exceeds body size");
    m.relationships.resize((size_t)nrel);
    for (auto& r : m.relationships)
      if (!c.str(&r) || !valid_utf8(r)) return ser("bad relationship string");
    m.access_count = c.u32();
    m.last_access_secs = c.u64();
    m.last_access_nanos = c.u32();
    if (!c.ok) return ser("unexpected end of body (memory)");
    if (m.ts_nanos >= 1000000000u || m.last_access_nanos >= 1000000000u)
      return ser("invalid SystemTime nanoseconds");  // serde's SystemTime visitor rejects these
    out->memories.push_back(std::move(m));
  }
  // bincode::deserialize tolerates trailing bytes.
  lap("record headers");

  // The payloads: one allocation, copied by all host threads; the NaN/Inf test `Memory::validate` needs rides along.
  out->vectors.resize((size_t)total_floats);
  {
    float* dst = out->vectors.data();
    std::vector<SnapshotMemory>& mem = out->memories;
    const uint64_t n = mem.size();
    run_parts(n < 4096 ? 1u : host_threads(), [&](unsigned t, unsigned parts) {
      for (uint64_t i = n * t / parts, e = n * (t + 1) / parts; i < e; ++i) {
        float* d = dst + mem[i].data_offset;
        std::memcpy(d, payload[i], mem[i].data_len * 4);
        bool fin = true;
        for (uint64_t j = 0; j < mem[i].data_len; ++j) fin &= std::isfinite(d[j]);
        mem[i].data_finite = fin;
      }
    });
  }
  lap("payload copy + finite check");
  return validate_config(cfg);  // ChronoMind::new (src/store.rs:188) — before any insert
}

// `Memory::validate` (src/types.rs:93-121) for record i, in the reference's order of checks. `load_snapshot` applies it
// through `store.insert` record by record (src/persistence.rs:117-119), interleaved with the capacity check — the
// caller replays that order.
cm_status validate_snapshot_memory(const SnapshotData& snap, uint64_t i) {
  const SnapshotMemory& m = snap.memories[i];
  const cm_config& cfg = snap.config;
  if (m.id.empty()) {
    set_error("invalid vector data: id must not be empty");
    return CM_ERR_INVALID_VECTOR;
  }
  if (m.data_len != cfg.dimensions) {
    char buf[96];
    std::snprintf(buf, sizeof buf, "invalid dimensions: got %llu, expected %llu", (unsigned long long)m.data_len,
                  (unsigned long long)cfg.dimensions);
    set_error(buf);
    return CM_ERR_INVALID_DIMENSIONS;
  }
  if (!m.data_finite) {  // computed by the parallel payload pass of read_snapshot
    set_error("invalid vector data: vector " + m.id + " contains NaN or infinite components");
    return CM_ERR_INVALID_VECTOR;
  }
  if (!std::isfinite(m.importance) || m.importance < 0.0f || m.importance > 1.0f) {
    set_error("invalid importance value: must be within [0.0, 1.0]");
    return CM_ERR_INVALID_IMPORTANCE;
  }
  if (!std::isfinite(m.decay_rate) || m.decay_rate < 0.0f) {
    set_error("invalid vector data: vector " + m.id + " has an invalid decay rate");
    return CM_ERR_INVALID_VECTOR;
  }
  return CM_OK;
}

}  // namespace cm
