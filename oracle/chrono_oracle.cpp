// chrono_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the ChronoMind 0.2.5 read path (and of the index
// build it depends on), following the reference line by line. It exists to
// CHECK the CUDA engine; nothing under chrono-mind_b200/ links, loads or calls
// it. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may use it.
//
// Parity status: PINNED against the known-answer vectors and behavioural
// fixtures of the reference's own tests (see tests/test_oracle_*.py):
//   src/metric.rs:233-287, src/index/lockfree_hnsw.rs:502-573,
//   tests/property_test.rs:146-229, tests/store_test.rs:38-117,
//   tests/recall_test.rs:140-214 (thresholds), tests/persistence_test.rs.
// The Rust reference itself cannot be compiled in this image (no rustc/cargo),
// so there is no oracle/_ref; see DESIGN.md "Oracle".
//
// Build: g++ -O2 -mavx2 -mfma -ffp-contract=off -fno-fast-math -shared -fPIC
// (Rust never contracts a*b+c and never reassociates float sums; the explicit
// FMAs below are the reference's own _mm256_fmadd_ps.)
//
// Every function cites the reference file:line it restates (paths relative to
// the reference checkout).

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

constexpr float F32_EPSILON = 1.1920929e-7f;  // f32::EPSILON

// ---------------------------------------------------------------------------
// src/index/mod.rs:37-52 — TotalF32: f32::total_cmp as an unsigned key.
// total_cmp orders by the sign-magnitude bit pattern: flip all bits of
// negatives, flip only the sign bit of non-negatives.
// ---------------------------------------------------------------------------
inline uint32_t total_key(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct DistId {
  float d;
  uint32_t id;
};
// Tuple ordering of (TotalF32, u32): lockfree_hnsw.rs:149-150.
inline bool tuple_less(const DistId& a, const DistId& b) {
  uint32_t ka = total_key(a.d), kb = total_key(b.d);
  if (ka != kb) return ka < kb;
  return a.id < b.id;
}
struct MaxHeapCmp {  // std::priority_queue top() == greatest
  bool operator()(const DistId& a, const DistId& b) const { return tuple_less(a, b); }
};
struct MinHeapCmp {  // BinaryHeap<Reverse<..>>
  bool operator()(const DistId& a, const DistId& b) const { return tuple_less(b, a); }
};

// ---------------------------------------------------------------------------
// src/metric.rs:119-121 — dot_scalar: sequential sum of products.
// ---------------------------------------------------------------------------
float dot_scalar(const float* a, const float* b, size_t n) {
  float s = 0.0f;
  for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// src/metric.rs:95-102 / 137-141 — the reference's horizontal sum order:
// lo128+hi128, then +movehl, then lane0 + lane1.
inline float hsum256(__m256 v) {
  __m128 lo = _mm256_castps256_ps128(v);
  __m128 hi = _mm256_extractf128_ps(v, 1);
  __m128 sum128 = _mm_add_ps(lo, hi);
  __m128 sum64 = _mm_add_ps(sum128, _mm_movehl_ps(sum128, sum128));
  __m128 sum32 = _mm_add_ss(sum64, _mm_shuffle_ps(sum64, sum64, 1));
  return _mm_cvtss_f32(sum32);
}

// src/metric.rs:129-147 — dot_avx2.
float dot_avx2(const float* a, const float* b, size_t n) {
  __m256 acc = _mm256_setzero_ps();
  size_t chunks = n / 8 * 8;
  for (size_t i = 0; i < chunks; i += 8) {
    __m256 va = _mm256_loadu_ps(a + i);
    __m256 vb = _mm256_loadu_ps(b + i);
    acc = _mm256_fmadd_ps(va, vb, acc);
  }
  float s = hsum256(acc);
  for (size_t i = chunks; i < n; ++i) s += a[i] * b[i];
  return s;
}

// src/metric.rs:152-161 — dot(): AVX2+FMA when the CPU has it. Every host this
// runs on has both (probed at load by orc_cpu_has_avx2_fma()).
inline float dot(const float* a, const float* b, size_t n) { return dot_avx2(a, b, n); }

// src/metric.rs:61-71 — dot_and_norms_scalar.
void dot_and_norms_scalar(const float* a, const float* b, size_t n, float* d, float* na, float* nb) {
  float dot_ = 0.0f, norm_a = 0.0f, norm_b = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    float x = a[i], y = b[i];
    dot_ += x * y;
    norm_a += x * x;
    norm_b += y * y;
  }
  *d = dot_;
  *na = norm_a;
  *nb = norm_b;
}

// src/metric.rs:78-116 — dot_and_norms_avx2.
void dot_and_norms_avx2(const float* a, const float* b, size_t n, float* d, float* na, float* nb) {
  __m256 vd = _mm256_setzero_ps(), vna = _mm256_setzero_ps(), vnb = _mm256_setzero_ps();
  size_t chunks = n / 8 * 8;
  for (size_t i = 0; i < chunks; i += 8) {
    __m256 va = _mm256_loadu_ps(a + i);
    __m256 vb = _mm256_loadu_ps(b + i);
    vd = _mm256_fmadd_ps(va, vb, vd);
    vna = _mm256_fmadd_ps(va, va, vna);
    vnb = _mm256_fmadd_ps(vb, vb, vnb);
  }
  float dot_s = hsum256(vd), na_s = hsum256(vna), nb_s = hsum256(vnb);
  for (size_t i = chunks; i < n; ++i) {
    float x = a[i], y = b[i];
    dot_s += x * y;
    na_s += x * x;
    nb_s += y * y;
  }
  *d = dot_s;
  *na = na_s;
  *nb = nb_s;
}

// f32::clamp: NaN passes through.
inline float clampf(float x, float lo, float hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}

// src/metric.rs:163-186 — cosine(): returns false for "None".
bool cosine(const float* a, size_t na_len, const float* b, size_t nb_len, bool simd, float* out) {
  if (na_len == 0 || na_len != nb_len) return false;
  float d, na, nb;
  if (simd)
    dot_and_norms_avx2(a, b, na_len, &d, &na, &nb);
  else
    dot_and_norms_scalar(a, b, na_len, &d, &na, &nb);
  float denom = std::sqrt(na * nb);
  if (denom <= F32_EPSILON) return false;
  *out = clampf(d / denom, -1.0f, 1.0f);
  return true;
}

// src/metric.rs:190-195 — distance().
float metric_distance(const float* a, size_t na, const float* b, size_t nb) {
  float c;
  if (cosine(a, na, b, nb, true, &c)) return 1.0f - c;
  return 2.0f;
}

// src/metric.rs:208-214 — preprocess(): sequential Σx², sqrt, true division.
void preprocess(const float* v, size_t n, float* out) {
  float s = 0.0f;
  for (size_t i = 0; i < n; ++i) s += v[i] * v[i];
  float norm = std::sqrt(s);
  if (norm <= F32_EPSILON) {
    if (out != v) std::memcpy(out, v, n * sizeof(float));
    return;
  }
  for (size_t i = 0; i < n; ++i) out[i] = v[i] / norm;
}

// src/metric.rs:219-224 — distance_prepared().
inline float distance_prepared(const float* a, size_t na, const float* b, size_t nb) {
  if (na == 0 || na != nb) return 2.0f;
  return clampf(1.0f - dot(a, b, na), 0.0f, 2.0f);
}

// ---------------------------------------------------------------------------
// src/index/lockfree_hnsw.rs:71-77 — splitmix64.
// ---------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr size_t MAX_LAYER = 31;                         // lockfree_hnsw.rs:43
constexpr uint64_t EMPTY_ENTRY = ~0ull;                  // lockfree_hnsw.rs:46
constexpr size_t ARENA_CAPACITY = 4096ull * 4096ull;     // arena.rs:47-50,101

struct Counters {
  uint64_t dist_evals = 0;     // distance_prepared calls
  uint64_t expansions = 0;     // nodes whose neighbor list was walked
  uint64_t list_entries = 0;   // neighbor ids read (Σ deg over expansions)
};

struct Node {  // lockfree_hnsw.rs:48-54
  std::vector<float> vector;
  size_t top_layer;
  bool deleted;
  std::vector<std::vector<uint32_t>> layers;
};

struct Hnsw {
  size_t max_connections, ef_construction, ef_search;  // config.rs:11-25
  double ml;                                           // lockfree_hnsw.rs:104
  std::deque<Node> nodes;
  uint64_t entry = EMPTY_ENTRY;
  size_t live = 0;
  uint64_t layer_ticket = 0;
  uint64_t seed = 0;

  Hnsw(size_t m, size_t efc, size_t efs, uint64_t s)
      : max_connections(m), ef_construction(efc), ef_search(efs),
        ml(1.0 / std::log((double)m)), seed(s) {}

  size_t cap(size_t layer) const {  // lockfree_hnsw.rs:117-123
    return layer == 0 ? max_connections * 2 : max_connections;
  }

  size_t random_layer() {  // lockfree_hnsw.rs:126-132
    uint64_t ticket = layer_ticket++;
    uint64_t bits = splitmix64(seed ^ ticket);
    double u = 1.0 - (double)(bits >> 11) / (double)(1ull << 53);
    double v = -std::log(u) * ml;
    // Rust `as usize` saturates; v is finite and >= 0 here.
    size_t l = (size_t)v;
    return std::min(l, MAX_LAYER);
  }

  const Node* node(uint32_t id) const { return id < nodes.size() ? &nodes[id] : nullptr; }

  // lockfree_hnsw.rs:140-194 — search_layer.
  std::vector<DistId> search_layer(const float* q, size_t qlen, const uint32_t* eps, size_t n_eps,
                                   size_t ef, size_t layer, Counters* c) const {
    std::unordered_set<uint32_t> visited;
    std::priority_queue<DistId, std::vector<DistId>, MinHeapCmp> frontier;
    std::priority_queue<DistId, std::vector<DistId>, MaxHeapCmp> best;

    for (size_t i = 0; i < n_eps; ++i) {
      uint32_t ep = eps[i];
      const Node* nd = node(ep);
      if (!nd) continue;
      if (layer > nd->top_layer || !visited.insert(ep).second) continue;
      float d = distance_prepared(nd->vector.data(), nd->vector.size(), q, qlen);
      if (c) c->dist_evals++;
      frontier.push({d, ep});
      best.push({d, ep});
    }
    while (best.size() > ef) best.pop();

    while (!frontier.empty()) {
      DistId cur = frontier.top();
      frontier.pop();
      if (!best.empty()) {
        float worst = best.top().d;
        // `dist > worst` on TotalF32: total order on the distance only.
        if (best.size() >= ef && total_key(cur.d) > total_key(worst)) break;
      }
      const Node* nd = node(cur.id);
      if (!nd) continue;
      const std::vector<uint32_t>& nbrs = nd->layers[layer];
      if (c) {
        c->expansions++;
        c->list_entries += nbrs.size();
      }
      for (uint32_t nb : nbrs) {
        if (!visited.insert(nb).second) continue;
        const Node* nn = node(nb);
        if (!nn) continue;
        float d = distance_prepared(nn->vector.data(), nn->vector.size(), q, qlen);
        if (c) c->dist_evals++;
        bool admit = best.size() < ef || total_key(d) < total_key(best.top().d);
        if (admit) {
          frontier.push({d, nb});
          best.push({d, nb});
          if (best.size() > ef) best.pop();
        }
      }
    }

    std::vector<DistId> out;  // into_sorted_vec(): ascending
    out.reserve(best.size());
    while (!best.empty()) {
      out.push_back(best.top());
      best.pop();
    }
    std::reverse(out.begin(), out.end());
    return out;
  }

  // lockfree_hnsw.rs:196-201 — descend_layer.
  uint32_t descend_layer(const float* q, size_t qlen, uint32_t entry_id, size_t layer, Counters* c) const {
    std::vector<DistId> r = search_layer(q, qlen, &entry_id, 1, 1, layer, c);
    return r.empty() ? entry_id : r[0].id;
  }

  // lockfree_hnsw.rs:205-240 — select_diverse (Alg. 4 + keepPrunedConnections).
  std::vector<uint32_t> select_diverse(const std::vector<DistId>& candidates, size_t m) const {
    std::vector<DistId> selected;
    std::vector<uint32_t> rejected;
    for (const DistId& cand : candidates) {
      if (selected.size() >= m) break;
      const Node* cn = node(cand.id);
      if (!cn) continue;
      bool diverse = true;
      for (const DistId& ch : selected) {
        const Node* hn = node(ch.id);
        if (!hn) continue;
        float dc = distance_prepared(cn->vector.data(), cn->vector.size(), hn->vector.data(), hn->vector.size());
        // `dist_to_query.0 < dist_to_chosen`: a raw f32 compare (:225), not TotalF32.
        if (!(cand.d < dc)) {
          diverse = false;
          break;
        }
      }
      if (diverse)
        selected.push_back(cand);
      else
        rejected.push_back(cand.id);
    }
    std::vector<uint32_t> result;
    for (const DistId& s : selected) result.push_back(s.id);
    for (uint32_t r : rejected) {
      if (result.size() >= m) break;
      result.push_back(r);
    }
    return result;
  }

  // lockfree_hnsw.rs:246-280 — add_backlink.
  void add_backlink(uint32_t from, uint32_t to, size_t layer) {
    if (from >= nodes.size()) return;
    Node& fn = nodes[from];
    size_t cap_ = cap(layer);
    std::vector<uint32_t>& current = fn.layers[layer];
    if (std::find(current.begin(), current.end(), to) != current.end()) return;
    if (current.size() < cap_) {
      current.push_back(to);
      return;
    }
    std::vector<DistId> with_d;
    with_d.reserve(current.size() + 1);
    auto add = [&](uint32_t id) {
      const Node* n = node(id);
      if (!n) return;
      with_d.push_back({distance_prepared(n->vector.data(), n->vector.size(), fn.vector.data(), fn.vector.size()), id});
    };
    for (uint32_t id : current) add(id);
    add(to);
    std::sort(with_d.begin(), with_d.end(), tuple_less);
    current = select_diverse(with_d, cap_);
  }

  // lockfree_hnsw.rs:375-449 — insert (single-threaded restatement; CAS
  // loops collapse to plain stores).
  int64_t insert(const float* v, size_t len) {
    std::vector<float> vec(len);
    preprocess(v, len, vec.data());
    size_t top_layer = random_layer();
    if (nodes.size() >= ARENA_CAPACITY) return -1;  // arena.rs:114-118
    uint32_t id = (uint32_t)nodes.size();
    nodes.push_back(Node{vec, top_layer, false, std::vector<std::vector<uint32_t>>(top_layer + 1)});

    if (entry == EMPTY_ENTRY) {
      entry = ((uint64_t)top_layer << 32) | id;  // pack_entry, :56-59
      live++;
      return id;
    }
    uint32_t entry_id = (uint32_t)entry;
    size_t entry_top = (size_t)(entry >> 32);

    uint32_t ep = entry_id;
    for (size_t layer = entry_top; layer > top_layer; --layer)  // (top_layer+1..=entry_top).rev()
      ep = descend_layer(vec.data(), len, ep, layer, nullptr);

    std::vector<uint32_t> entry_points{ep};
    std::vector<std::pair<size_t, std::vector<uint32_t>>> selected_per_layer;
    size_t start = std::min(top_layer, entry_top);
    for (size_t layer = start + 1; layer-- > 0;) {
      std::vector<DistId> cands = search_layer(vec.data(), len, entry_points.data(), entry_points.size(),
                                               ef_construction, layer, nullptr);
      selected_per_layer.emplace_back(layer, select_diverse(cands, max_connections));
      entry_points.clear();
      for (const DistId& c : cands) entry_points.push_back(c.id);
    }

    for (auto& ls : selected_per_layer) nodes[id].layers[ls.first] = ls.second;
    for (auto& ls : selected_per_layer)
      for (uint32_t nb : ls.second) add_backlink(nb, id, ls.first);

    if (top_layer > entry_top) entry = ((uint64_t)top_layer << 32) | id;  // raise_entry :351-371
    live++;
    return id;
  }

  // lockfree_hnsw.rs:451-467 — remove.
  bool remove(uint32_t id) {
    if (id >= nodes.size()) return false;
    if (nodes[id].deleted) return false;
    nodes[id].deleted = true;
    live--;
    return true;
  }

  // lockfree_hnsw.rs:469-495 — search.
  std::vector<DistId> search(const float* query, size_t len, size_t ef, Counters* c) const {
    std::vector<DistId> out;
    if (ef == 0) return out;
    if (entry == EMPTY_ENTRY) return out;
    uint32_t entry_id = (uint32_t)entry;
    size_t entry_top = (size_t)(entry >> 32);
    std::vector<float> q(len);
    preprocess(query, len, q.data());
    uint32_t ep = entry_id;
    for (size_t layer = entry_top; layer >= 1; --layer) ep = descend_layer(q.data(), len, ep, layer, c);
    std::vector<DistId> r = search_layer(q.data(), len, &ep, 1, ef, 0, c);
    for (const DistId& x : r) {
      const Node* n = node(x.id);
      if (n && !n->deleted) out.push_back(x);
    }
    return out;
  }

  // lockfree_hnsw.rs:295-347 — check_invariants. Returns 0 when OK, else a
  // code: 1 entry missing, 2 entry layer, 3 cap, 4 self-loop, 5 dangling,
  // 6 duplicate, 7 link not in layer.
  int check_invariants() const {
    uint32_t total = (uint32_t)nodes.size();
    if (entry != EMPTY_ENTRY) {
      uint32_t eid = (uint32_t)entry;
      size_t etop = (size_t)(entry >> 32);
      if (eid >= total) return 1;
      if (nodes[eid].top_layer < etop) return 2;
    }
    for (uint32_t id = 0; id < total; ++id) {
      const Node& n = nodes[id];
      for (size_t layer = 0; layer < n.layers.size(); ++layer) {
        const auto& ids = n.layers[layer];
        if (ids.size() > cap(layer)) return 3;
        std::unordered_set<uint32_t> seen;
        for (uint32_t x : ids) {
          if (x == id) return 4;
          if (x >= total) return 5;
          if (!seen.insert(x).second) return 6;
          if (nodes[x].top_layer < layer) return 7;
        }
      }
    }
    return 0;
  }
};

// src/store.rs:395-415 — combined_score. Duration::as_secs_f32 is
// (secs as f32) + (nanos as f32) / 1e9f32 (Rust std, core/src/time.rs).
float combined_score(float distance, uint64_t ts_secs, uint32_t ts_nanos, float decay_rate,
                     uint64_t now_secs, uint32_t now_nanos, float w, float base_decay_rate) {
  uint64_t dsecs = 0;
  uint32_t dnanos = 0;
  bool later = (now_secs > ts_secs) || (now_secs == ts_secs && now_nanos >= ts_nanos);
  if (later) {  // duration_since(..).unwrap_or_default()
    dsecs = now_secs - ts_secs;
    if (now_nanos >= ts_nanos) {
      dnanos = now_nanos - ts_nanos;
    } else {
      dsecs -= 1;
      dnanos = now_nanos + 1000000000u - ts_nanos;
    }
  }
  float age_secs = (float)dsecs + (float)dnanos / 1.0e9f;
  float age_hours = age_secs / 3600.0f;
  float rate = decay_rate > 0.0f ? decay_rate : base_decay_rate;
  float temporal_relevance = std::exp(-rate * age_hours);  // f32::exp -> expf
  return (1.0f - w) * (distance / 2.0f) + w * (1.0f - temporal_relevance);
}

struct ScoredHandle {
  uint32_t handle;
  float score;
};

template <class F>
void parallel_for(size_t n, int threads, F f) {  // benches/comparison.rs:125-139 timed_parallel
  if (threads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  size_t chunk = (n + threads - 1) / threads;
  std::vector<std::thread> ts;
  for (size_t s = 0; s < n; s += chunk) {
    size_t e = std::min(n, s + chunk);
    ts.emplace_back([=] {
      for (size_t i = s; i < e; ++i) f(i);
    });
  }
  for (auto& t : ts) t.join();
}

}  // namespace

// ===========================================================================
// C surface (ctypes). All pointers are host pointers owned by the caller.
// ===========================================================================
extern "C" {

int orc_cpu_has_avx2_fma() { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }

uint32_t orc_total_key(float f) { return total_key(f); }
float orc_dot_scalar(const float* a, const float* b, uint64_t n) { return dot_scalar(a, b, n); }
float orc_dot_avx2(const float* a, const float* b, uint64_t n) { return dot_avx2(a, b, n); }
void orc_preprocess(const float* v, uint64_t n, float* out) { preprocess(v, n, out); }
// preprocess() for every row of a row-major matrix (rows are independent; `threads` splits them statically).
void orc_preprocess_rows(const float* x, uint64_t n, uint64_t dim, float* out, int threads) {
  parallel_for(n, threads, [&](size_t i) { preprocess(x + i * dim, dim, out + i * dim); });
}
float orc_distance_prepared(const float* a, uint64_t na, const float* b, uint64_t nb) {
  return distance_prepared(a, na, b, nb);
}
float orc_distance(const float* a, uint64_t na, const float* b, uint64_t nb) { return metric_distance(a, na, b, nb); }
// src/metric.rs:197-199 — similarity().
float orc_similarity(const float* a, uint64_t na, const float* b, uint64_t nb) {
  float c;
  return cosine(a, na, b, nb, true, &c) ? c : 0.0f;
}
// distance() through the scalar kernel (src/metric.rs:61-71) — the "expected"
// side of the reference's simd_and_scalar_paths_agree test.
float orc_distance_scalar(const float* a, uint64_t na, const float* b, uint64_t nb) {
  float c;
  return cosine(a, na, b, nb, false, &c) ? 1.0f - c : 2.0f;
}
uint64_t orc_splitmix64(uint64_t x) { return splitmix64(x); }

float orc_combined_score(float distance, uint64_t ts_secs, uint32_t ts_nanos, float decay_rate, uint64_t now_secs,
                         uint32_t now_nanos, float w, float base_decay_rate) {
  return combined_score(distance, ts_secs, ts_nanos, decay_rate, now_secs, now_nanos, w, base_decay_rate);
}

// ---- index ---------------------------------------------------------------
void* orc_hnsw_new(uint64_t max_connections, uint64_t ef_construction, uint64_t ef_search, uint64_t seed) {
  return new Hnsw(max_connections, ef_construction, ef_search, seed);
}
void orc_hnsw_free(void* h) { delete (Hnsw*)h; }
int64_t orc_hnsw_insert(void* h, const float* v, uint64_t len) { return ((Hnsw*)h)->insert(v, len); }
// Row-major bulk insert; returns number inserted.
uint64_t orc_hnsw_insert_matrix(void* h, const float* x, uint64_t n, uint64_t dim) {
  Hnsw* g = (Hnsw*)h;
  uint64_t done = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (g->insert(x + i * dim, dim) < 0) break;
    ++done;
  }
  return done;
}
int orc_hnsw_remove(void* h, uint32_t id) { return ((Hnsw*)h)->remove(id) ? 1 : 0; }
uint64_t orc_hnsw_len(void* h) { return ((Hnsw*)h)->live; }
uint64_t orc_hnsw_slots(void* h) { return ((Hnsw*)h)->nodes.size(); }
int orc_hnsw_check_invariants(void* h) { return ((Hnsw*)h)->check_invariants(); }
uint64_t orc_hnsw_entry(void* h) { return ((Hnsw*)h)->entry; }
uint64_t orc_random_layer_at(uint64_t seed, uint64_t ticket, uint64_t max_connections) {
  Hnsw g(max_connections, 1, 1, seed);
  g.layer_ticket = ticket;
  return g.random_layer();
}
uint32_t orc_hnsw_top_layer(void* h, uint32_t id) { return (uint32_t)((Hnsw*)h)->nodes[id].top_layer; }
// Copy the neighbor list of (id, layer) into out (capacity cap); returns its length.
uint32_t orc_hnsw_neighbors(void* h, uint32_t id, uint32_t layer, uint32_t* out, uint32_t cap) {
  Hnsw* g = (Hnsw*)h;
  if (id >= g->nodes.size() || layer >= g->nodes[id].layers.size()) return 0;
  const auto& l = g->nodes[id].layers[layer];
  for (size_t i = 0; i < l.size() && i < cap; ++i) out[i] = l[i];
  return (uint32_t)l.size();
}
void orc_hnsw_vector(void* h, uint32_t id, float* out) {
  Hnsw* g = (Hnsw*)h;
  std::memcpy(out, g->nodes[id].vector.data(), g->nodes[id].vector.size() * 4);
}

// Adopt an already-built graph (flat layout written by the engine's graph sidecar) so that the oracle's
// search_layer/search run over exactly the graph the GPU searches ("same snapshot graph" parity):
// rows are ALREADY preprocess()ed; nbr0[n][cap0]/deg0[n]; upper lists of node v (top_layer t >= 1) at slots
// upper_base[v] .. upper_base[v]+t-1 of nbr_up[slot][capU]/deg_up[slot]. Replaces the index content.
void orc_hnsw_import(void* h, const float* x, uint64_t n, uint64_t dim, const uint8_t* top_layer, const uint32_t* nbr0,
                     uint32_t cap0, const uint32_t* deg0, const uint32_t* upper_base, const uint32_t* nbr_up,
                     uint32_t capU, const uint32_t* deg_up, uint64_t entry, const uint8_t* deleted) {
  Hnsw* g = (Hnsw*)h;
  g->nodes.clear();
  g->live = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Node nd;
    nd.vector.assign(x + i * dim, x + (i + 1) * dim);
    nd.top_layer = top_layer[i];
    nd.deleted = deleted && deleted[i];
    nd.layers.resize(nd.top_layer + 1);
    nd.layers[0].assign(nbr0 + i * cap0, nbr0 + i * cap0 + deg0[i]);
    for (size_t l = 1; l <= nd.top_layer; ++l) {
      uint64_t slot = (uint64_t)upper_base[i] + l - 1;
      nd.layers[l].assign(nbr_up + slot * capU, nbr_up + slot * capU + deg_up[slot]);
    }
    if (!nd.deleted) g->live++;
    g->nodes.push_back(std::move(nd));
  }
  g->entry = entry;
  g->layer_ticket = n;
}

// VectorIndex::search (src/index/mod.rs:70-72). Writes up to `cap` pairs;
// returns the number of results. counters (optional): [dist_evals, expansions, list_entries].
uint64_t orc_hnsw_search(void* h, const float* q, uint64_t len, uint64_t ef, uint32_t* ids, float* dists,
                         uint64_t cap, uint64_t* counters) {
  Counters c;
  std::vector<DistId> r = ((Hnsw*)h)->search(q, len, ef, &c);
  uint64_t n = std::min<uint64_t>(r.size(), cap);
  for (uint64_t i = 0; i < n; ++i) {
    ids[i] = r[i].id;
    dists[i] = r[i].d;
  }
  if (counters) {
    counters[0] += c.dist_evals;
    counters[1] += c.expansions;
    counters[2] += c.list_entries;
  }
  return r.size();
}

// Batched form of the PyO3 Index.batch_query loop (bindings/python/src/lib.rs:92-117):
// search(q, ef), take(k). Rows padded with id=UINT32_MAX, dist=+inf. `threads`
// splits queries statically (benches/comparison.rs:125-139).
void orc_hnsw_search_batch(void* h, const float* q, uint64_t nq, uint64_t dim, uint64_t ef, uint64_t k,
                           uint32_t* ids, float* dists, uint32_t* counts, int threads, uint64_t* counters) {
  Hnsw* g = (Hnsw*)h;
  std::atomic<uint64_t> ev{0}, ex{0}, le{0};
  parallel_for(nq, threads, [&](size_t i) {
    Counters c;
    std::vector<DistId> r = g->search(q + i * dim, dim, ef, &c);
    uint64_t n = std::min<uint64_t>(r.size(), k);
    for (uint64_t j = 0; j < k; ++j) {
      ids[i * k + j] = j < n ? r[j].id : 0xFFFFFFFFu;
      dists[i * k + j] = j < n ? r[j].d : std::numeric_limits<float>::infinity();
    }
    if (counts) counts[i] = (uint32_t)n;
    ev += c.dist_evals;
    ex += c.expansions;
    le += c.list_entries;
  });
  if (counters) {
    counters[0] += ev;
    counters[1] += ex;
    counters[2] += le;
  }
}

// ---- exact scans -----------------------------------------------------------
// mode 0: tests/recall_test.rs:70-79 — metric.distance(v, q) on the rows as
//         given, stable sort by distance, take k.
// mode 1: tests/property_test.rs:201-206 — distance_prepared(preprocess(v),
//         preprocess(q)); `x` must already hold preprocess()ed rows, the query
//         is preprocessed here. Stable sort == ascending (distance, index).
// `deleted` (optional, n bytes) drops tombstoned rows (property_test.rs:190-195).
void orc_brute_force_batch(const float* x, uint64_t n, uint64_t dim, const uint8_t* deleted, const float* q,
                           uint64_t nq, uint64_t k, int mode, uint32_t* ids, float* dists, int threads) {
  parallel_for(nq, threads, [&](size_t qi) {
    std::vector<float> qn(dim);
    const float* qq = q + qi * dim;
    if (mode == 1) {
      preprocess(qq, dim, qn.data());
      qq = qn.data();
    }
    std::vector<DistId> scored;
    scored.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
      if (deleted && deleted[i]) continue;
      float d = mode == 1 ? distance_prepared(x + i * dim, dim, qq, dim) : metric_distance(x + i * dim, dim, qq, dim);
      scored.push_back({d, (uint32_t)i});
    }
    uint64_t kk = std::min<uint64_t>(k, scored.size());
    // Stable sort by distance only == partial sort by (distance, index) since
    // indexes are pushed in ascending order.
    std::partial_sort(scored.begin(), scored.begin() + kk, scored.end(), tuple_less);
    for (uint64_t j = 0; j < k; ++j) {
      ids[qi * k + j] = j < kk ? scored[j].id : 0xFFFFFFFFu;
      dists[qi * k + j] = j < kk ? scored[j].d : std::numeric_limits<float>::infinity();
    }
  });
}

// ---- store-level search -----------------------------------------------------
// src/store.rs:321-346 (ChronoMind::search) over an index whose handle i maps
// 1:1 to memory i (true for a loaded snapshot without duplicate ids), or
// src/store.rs:353-377-style exact candidate pool when `exact` != 0 (pool =
// exact top-ef by distance_prepared; used to check the GPU exact+rerank path).
// Returns 0, or 1 = InvalidDimensions, 2 = InvalidVector (store.rs:379-392).
int orc_store_search_batch(void* h, uint64_t dimensions, uint64_t ef_search, float temporal_weight,
                           float base_decay_rate, const uint64_t* ts_secs, const uint32_t* ts_nanos,
                           const float* decay_rate, uint64_t now_secs, uint32_t now_nanos, const float* q,
                           uint64_t nq, uint64_t qdim, uint64_t k, uint32_t* ids, float* scores, float* dists,
                           int threads) {
  Hnsw* g = (Hnsw*)h;
  if (qdim != dimensions) return 1;
  for (uint64_t i = 0; i < nq * qdim; ++i)
    if (!std::isfinite(q[i])) return 2;
  uint64_t ef = std::max<uint64_t>(ef_search, k * 3);  // OVERSAMPLE, store.rs:40,323
  parallel_for(nq, threads, [&](size_t qi) {
    std::vector<DistId> r = g->search(q + qi * qdim, qdim, ef, nullptr);
    struct S {
      uint32_t handle;
      float dist;
      float score;
    };
    std::vector<S> scored;
    scored.reserve(r.size());
    for (const DistId& c : r)
      scored.push_back({c.id, c.d,
                        combined_score(c.d, ts_secs[c.id], ts_nanos[c.id], decay_rate[c.id], now_secs, now_nanos,
                                       temporal_weight, base_decay_rate)});
    std::stable_sort(scored.begin(), scored.end(),
                     [](const S& a, const S& b) { return total_key(a.score) < total_key(b.score); });
    for (uint64_t j = 0; j < k; ++j) {
      bool ok = j < scored.size();
      ids[qi * k + j] = ok ? scored[j].handle : 0xFFFFFFFFu;
      scores[qi * k + j] = ok ? scored[j].score : std::numeric_limits<float>::infinity();
      if (dists) dists[qi * k + j] = ok ? scored[j].dist : std::numeric_limits<float>::infinity();
    }
  });
  return 0;
}

// Rerank an already-computed candidate pool (ascending (dist,id)) exactly as
// store.rs:327-344 does: score, stable sort by score, truncate k.
void orc_rerank_pool(const uint32_t* cand_ids, const float* cand_dists, uint64_t nq, uint64_t pool,
                     float temporal_weight, float base_decay_rate, const uint64_t* ts_secs,
                     const uint32_t* ts_nanos, const float* decay_rate, uint64_t now_secs, uint32_t now_nanos,
                     uint64_t k, uint32_t* ids, float* scores) {
  for (uint64_t qi = 0; qi < nq; ++qi) {
    std::vector<ScoredHandle> scored;
    for (uint64_t j = 0; j < pool; ++j) {
      uint32_t id = cand_ids[qi * pool + j];
      if (id == 0xFFFFFFFFu) continue;
      scored.push_back({id, combined_score(cand_dists[qi * pool + j], ts_secs[id], ts_nanos[id], decay_rate[id],
                                           now_secs, now_nanos, temporal_weight, base_decay_rate)});
    }
    std::stable_sort(scored.begin(), scored.end(), [](const ScoredHandle& a, const ScoredHandle& b) {
      return total_key(a.score) < total_key(b.score);
    });
    for (uint64_t j = 0; j < k; ++j) {
      bool ok = j < scored.size();
      ids[qi * k + j] = ok ? scored[j].handle : 0xFFFFFFFFu;
      scores[qi * k + j] = ok ? scored[j].score : std::numeric_limits<float>::infinity();
    }
  }
}

// src/store.rs:353-377 — ChronoMind::search_in_context: validate; for every LIVE memory whose context matches:
// distance = metric.distance(raw stored data, raw query) (the full cosine with norms, NOT distance_prepared),
// score = combined_score; sort by score (total_cmp); truncate k. The reference iterates a concurrent hash map, so
// the order among EQUAL scores is unspecified there; here (and on the device) ties resolve toward the lower handle.
// raw: [n][dim] un-preprocessed rows; context_of: [n] context id per handle; deleted: optional tombstones (handles
// replaced or removed are not in `by_id`). want_ctx: [nq] context id per query.
// Returns 0, or 1 = InvalidDimensions, 2 = InvalidVector.
int orc_search_in_context(const float* raw, uint64_t n, uint64_t dim, const uint32_t* context_of, const uint8_t* deleted,
                          float temporal_weight, float base_decay_rate, const uint64_t* ts_secs, const uint32_t* ts_nanos,
                          const float* decay_rate, uint64_t now_secs, uint32_t now_nanos, const float* q, uint64_t nq,
                          uint64_t qdim, const uint32_t* want_ctx, uint64_t k, uint32_t* ids, float* scores, int threads) {
  if (qdim != dim) return 1;
  for (uint64_t i = 0; i < nq * qdim; ++i)
    if (!std::isfinite(q[i])) return 2;
  parallel_for(nq, threads, [&](size_t qi) {
    std::vector<ScoredHandle> scored;
    for (uint64_t r = 0; r < n; ++r) {
      if (context_of[r] != want_ctx[qi] || (deleted && deleted[r])) continue;
      float d = metric_distance(raw + r * dim, dim, q + qi * qdim, qdim);
      scored.push_back({(uint32_t)r, combined_score(d, ts_secs[r], ts_nanos[r], decay_rate[r], now_secs, now_nanos,
                                                    temporal_weight, base_decay_rate)});
    }
    std::stable_sort(scored.begin(), scored.end(), [](const ScoredHandle& a, const ScoredHandle& b) {
      return total_key(a.score) < total_key(b.score);
    });
    for (uint64_t j = 0; j < k; ++j) {
      bool ok = j < scored.size();
      ids[qi * k + j] = ok ? scored[j].handle : 0xFFFFFFFFu;
      scores[qi * k + j] = ok ? scored[j].score : std::numeric_limits<float>::infinity();
    }
  });
  return 0;
}

// src/store.rs:516-568 — the pair test of ChronoMind::consolidate: for every pair i < j of live memories,
// similarity = metric.similarity(a.data, b.data) (cosine on the raw rows, 0 when undefined); the pair qualifies
// when `similarity > similarity_threshold` (:531-534). Only the test is restated: the greedy keeper/absorbed
// bookkeeping that follows is write-path state. Returns the number of qualifying pairs; writes at most `cap`.
uint64_t orc_similar_pairs(const float* raw, uint64_t n, uint64_t dim, const uint8_t* deleted, float threshold,
                           uint32_t* out_i, uint32_t* out_j, float* out_sim, uint64_t cap) {
  uint64_t found = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (deleted && deleted[i]) continue;
    for (uint64_t j = i + 1; j < n; ++j) {
      if (deleted && deleted[j]) continue;
      float c;
      float sim = cosine(raw + i * dim, dim, raw + j * dim, dim, true, &c) ? c : 0.0f;
      if (sim <= threshold) continue;
      if (found < cap) {
        out_i[found] = (uint32_t)i;
        out_j[found] = (uint32_t)j;
        out_sim[found] = sim;
      }
      ++found;
    }
  }
  return found;
}

// src/store.rs:428-458 — one apply_decay sweep over plain arrays (the per-memory CAS gate only matters for
// concurrent sweeps): from = max(previous sweep, last access); skip when now <= from; claim; hours = (now - from)
// as f32 / 1e9 / 3600; importance = clamp(importance * exp(-rate * hours), 0, 1) (:114-128).
void orc_decay_sweep(float* importance, uint64_t* decayed_through, const uint64_t* last_access, const float* decay_rate,
                     uint64_t n, float base_rate, uint64_t now_nanos) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t from = std::max(decayed_through[i], last_access[i]);
    if (now_nanos <= from) continue;
    decayed_through[i] = now_nanos;
    float hours = (float)(now_nanos - from) / 1e9f / 3600.0f;
    float rate = decay_rate[i] > 0.0f ? decay_rate[i] : base_rate;
    float v = importance[i] * std::exp(-rate * hours);
    importance[i] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  }
}

// src/index/sharded_rwlock.rs:81-94 — merge per-shard lists: concat, sort by
// (dist, global id), truncate. lists: [n_shards][nq][len] local ids/dists
// (UINT32_MAX = padding); global id = local * n_shards + shard (the
// generalisation of join_handle :60-66, which is (local << 4) | shard for 16).
void orc_sharded_merge(const uint32_t* ids, const float* dists, uint64_t n_shards, uint64_t nq, uint64_t len,
                       uint64_t out_len, uint32_t* out_ids, float* out_dists) {
  for (uint64_t qi = 0; qi < nq; ++qi) {
    std::vector<DistId> merged;
    for (uint64_t s = 0; s < n_shards; ++s)
      for (uint64_t j = 0; j < len; ++j) {
        uint64_t o = (s * nq + qi) * len + j;
        if (ids[o] == 0xFFFFFFFFu) continue;
        merged.push_back({dists[o], (uint32_t)(ids[o] * n_shards + s)});
      }
    std::sort(merged.begin(), merged.end(), tuple_less);
    for (uint64_t j = 0; j < out_len; ++j) {
      bool ok = j < merged.size();
      out_ids[qi * out_len + j] = ok ? merged[j].id : 0xFFFFFFFFu;
      out_dists[qi * out_len + j] = ok ? merged[j].d : std::numeric_limits<float>::infinity();
    }
  }
}

}  // extern "C"
