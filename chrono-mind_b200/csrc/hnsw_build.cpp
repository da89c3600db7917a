// hnsw_build.cpp — host-side HNSW construction for the device index.
//
// The reference snapshot carries no graph (src/persistence.rs:4-6,114-119: "The index is rebuilt on load"),
// so the engine has to construct one before it has anything to flatten. This builder follows
// `LockFreeHnsw::insert` (src/index/lockfree_hnsw.rs:375-449) — layer draw (:126-132), greedy descent
// (:196-201), per-layer `search_layer(ef_construction)` (:140-194), Algorithm-4 `select_diverse` with
// keepPruned backfill (:205-240), `add_backlink` with overflow re-selection (:246-280), `raise_entry`
// (:351-371) — but works directly in the flat padded-CSR arrays the GPU consumes (cm::Graph), with epoch
// visited arrays and reusable heaps instead of per-call HashSet/BinaryHeap allocations, four distance
// chains in flight per query, and per-node spinlocks instead of COW lists + epoch GC when threads > 1.
//
// threads == 1 gives the reference's deterministic single-threaded graph (ticket == handle); threads > 1
// mirrors its concurrent build: correct, but linkage depends on interleaving (:21-26).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

#include "cm_internal.h"
#include "metric_host.h"

namespace cm {
namespace {

inline uint64_t splitmix64(uint64_t x) {  // src/index/lockfree_hnsw.rs:71-77
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

inline uint32_t draw_layer(uint64_t seed, uint64_t ticket, double ml) {  // :126-132
  uint64_t bits = splitmix64(seed ^ ticket);
  double u = 1.0 - (double)(bits >> 11) / 9007199254740992.0;
  double v = -std::log(u) * ml;
  uint64_t l = (uint64_t)v;
  return (uint32_t)std::min<uint64_t>(l, kMaxLayer);
}

// (TotalF32, u32) tuples packed so that u64 order == tuple order.
inline uint64_t pack(float d, uint32_t id) { return ((uint64_t)total_key(d) << 32) | id; }
inline uint32_t key_id(uint64_t k) { return (uint32_t)k; }
inline uint32_t key_dist(uint64_t k) { return (uint32_t)(k >> 32); }
inline float key_float(uint64_t k) {
  uint32_t t = key_dist(k);
  uint32_t b = (t & 0x80000000u) ? (t & 0x7FFFFFFFu) : ~t;
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}

struct Builder;

struct Scratch {  // per-thread, reused across inserts
  std::vector<uint32_t> visited;
  uint32_t epoch = 0;
  std::vector<uint64_t> frontier;  // min-heap (std::greater)
  std::vector<uint64_t> best;      // max-heap (std::less)
  std::vector<uint64_t> cands;     // search_layer output, ascending
  std::vector<uint32_t> entry_points;
  std::vector<uint32_t> nbuf;      // neighbor-list copy + unvisited ids
  std::vector<float> dbuf;
  std::vector<uint64_t> sel;
  std::vector<uint32_t> rej, result;
  std::vector<std::vector<uint32_t>> selected_per_layer;
  std::vector<uint64_t> with_d;

  void next_epoch(uint64_t n) {
    if (visited.size() != n) visited.assign(n, 0), epoch = 0;
    if (++epoch == 0) {
      std::fill(visited.begin(), visited.end(), 0);
      epoch = 1;
    }
  }
  bool visit(uint32_t id) {  // HashSet::insert: true when newly inserted
    if (visited[id] == epoch) return false;
    visited[id] = epoch;
    return true;
  }
};

struct Builder {
  const float* x;
  uint64_t n;
  uint32_t dim;
  uint32_t M, efC, cap0, capU;
  Graph* g;
  bool mt;
  std::unique_ptr<std::atomic<uint8_t>[]> locks;
  std::atomic<uint64_t> entry{kEmptyEntry};

  const float* row(uint32_t id) const { return x + (uint64_t)id * dim; }
  uint32_t cap(uint32_t layer) const { return layer == 0 ? cap0 : capU; }  // :117-123

  void lock(uint32_t id) {
    if (!mt) return;
    while (locks[id].exchange(1, std::memory_order_acquire)) {
      while (locks[id].load(std::memory_order_relaxed)) _mm_pause();
    }
  }
  void unlock(uint32_t id) {
    if (mt) locks[id].store(0, std::memory_order_release);
  }

  uint32_t* list_ptr(uint32_t v, uint32_t layer, uint32_t** deg) {
    if (layer == 0) {
      *deg = &g->deg0[v];
      return &g->nbr0[(uint64_t)v * cap0];
    }
    uint32_t slot = g->upper_base[v] + layer - 1;
    *deg = &g->deg_up[slot];
    return &g->nbr_up[(uint64_t)slot * capU];
  }

  // Snapshot a neighbor list (NeighborList::load, src/index/neighbors.rs:48-57) into s.nbuf[0..deg).
  uint32_t load_list(uint32_t v, uint32_t layer, Scratch& s) {
    uint32_t* deg;
    lock(v);
    uint32_t* l = list_ptr(v, layer, &deg);
    uint32_t d = *deg;
    if (s.nbuf.size() < 2 * (size_t)cap0 + 8) s.nbuf.resize(2 * (size_t)cap0 + 8);
    std::memcpy(s.nbuf.data(), l, d * sizeof(uint32_t));
    unlock(v);
    return d;
  }

  // src/index/lockfree_hnsw.rs:140-194. Result in s.cands (ascending (dist,id)).
  void search_layer(const float* q, const uint32_t* eps, size_t n_eps, size_t ef, uint32_t layer, Scratch& s) {
    s.next_epoch(n);
    s.frontier.clear();
    s.best.clear();
    auto fpush = [&](uint64_t k) {
      s.frontier.push_back(k);
      std::push_heap(s.frontier.begin(), s.frontier.end(), std::greater<uint64_t>());
    };
    auto bpush = [&](uint64_t k) {
      s.best.push_back(k);
      std::push_heap(s.best.begin(), s.best.end());
    };
    auto bpop = [&]() {
      std::pop_heap(s.best.begin(), s.best.end());
      s.best.pop_back();
    };
    for (size_t i = 0; i < n_eps; ++i) {
      uint32_t ep = eps[i];
      if (layer > g->top_layer[ep] || !s.visit(ep)) continue;
      uint64_t k = pack(dist_prepared(row(ep), q, dim), ep);
      fpush(k);
      bpush(k);
    }
    while (s.best.size() > ef) bpop();

    while (!s.frontier.empty()) {
      std::pop_heap(s.frontier.begin(), s.frontier.end(), std::greater<uint64_t>());
      uint64_t cur = s.frontier.back();
      s.frontier.pop_back();
      if (!s.best.empty() && s.best.size() >= ef && key_dist(cur) > key_dist(s.best.front())) break;
      uint32_t deg = load_list(key_id(cur), layer, s);
      // Keep the not-yet-visited neighbors, in list order (visited.insert happens before the distance).
      uint32_t m = 0;
      uint32_t* fresh = s.nbuf.data() + cap0 + 4;
      for (uint32_t j = 0; j < deg; ++j) {
        uint32_t nb = s.nbuf[j];
        if (s.visit(nb)) {
          fresh[m++] = nb;
          _mm_prefetch((const char*)row(nb), _MM_HINT_T0);
        }
      }
      if (s.dbuf.size() < m + 4) s.dbuf.resize(m + 4);
      uint32_t j = 0;
      for (; j + 4 <= m; j += 4) {
        const float* rows[4] = {row(fresh[j]), row(fresh[j + 1]), row(fresh[j + 2]), row(fresh[j + 3])};
        dist_prepared_x4(q, rows, dim, &s.dbuf[j]);
      }
      for (; j < m; ++j) s.dbuf[j] = dist_prepared(row(fresh[j]), q, dim);
      // Admission is sequential in list order against the evolving worst (strict, distance only).
      for (j = 0; j < m; ++j) {
        uint64_t k = pack(s.dbuf[j], fresh[j]);
        bool admit = s.best.size() < ef || key_dist(k) < key_dist(s.best.front());
        if (admit) {
          fpush(k);
          bpush(k);
          if (s.best.size() > ef) bpop();
        }
      }
    }
    s.cands.assign(s.best.begin(), s.best.end());
    std::sort(s.cands.begin(), s.cands.end());
  }

  uint32_t descend(const float* q, uint32_t ep, uint32_t layer, Scratch& s) {  // :196-201
    search_layer(q, &ep, 1, 1, layer, s);
    return s.cands.empty() ? ep : key_id(s.cands[0]);
  }

  // :205-240. cands ascending; result in s.result.
  void select_diverse(const uint64_t* cands, size_t nc, size_t m, Scratch& s) {
    s.sel.clear();
    s.rej.clear();
    for (size_t i = 0; i < nc; ++i) {
      if (s.sel.size() >= m) break;
      uint32_t cand = key_id(cands[i]);
      float dq = key_float(cands[i]);
      bool diverse = true;
      for (uint64_t ch : s.sel) {
        float dc = dist_prepared(row(cand), row(key_id(ch)), dim);
        if (!(dq < dc)) {  // raw f32 `<` (:225)
          diverse = false;
          break;
        }
      }
      if (diverse)
        s.sel.push_back(cands[i]);
      else
        s.rej.push_back(cand);
    }
    s.result.clear();
    for (uint64_t k : s.sel) s.result.push_back(key_id(k));
    for (uint32_t r : s.rej) {
      if (s.result.size() >= m) break;
      s.result.push_back(r);
    }
  }

  void add_backlink(uint32_t from, uint32_t to, uint32_t layer, Scratch& s) {  // :246-280
    uint32_t* deg;
    lock(from);
    uint32_t* l = list_ptr(from, layer, &deg);
    uint32_t d = *deg, c = cap(layer);
    for (uint32_t j = 0; j < d; ++j)
      if (l[j] == to) {
        unlock(from);
        return;
      }
    if (d < c) {
      l[d] = to;
      *deg = d + 1;
      unlock(from);
      return;
    }
    s.with_d.clear();
    for (uint32_t j = 0; j < d; ++j) s.with_d.push_back(pack(dist_prepared(row(l[j]), row(from), dim), l[j]));
    s.with_d.push_back(pack(dist_prepared(row(to), row(from), dim), to));
    std::sort(s.with_d.begin(), s.with_d.end());
    select_diverse(s.with_d.data(), s.with_d.size(), c, s);
    std::memcpy(l, s.result.data(), s.result.size() * sizeof(uint32_t));
    *deg = (uint32_t)s.result.size();
    unlock(from);
  }

  void raise_entry(uint32_t id, uint32_t top) {  // :351-371
    uint64_t cur = entry.load(std::memory_order_acquire);
    for (;;) {
      bool replace = cur == kEmptyEntry || top > (uint32_t)(cur >> 32);
      if (!replace) return;
      if (entry.compare_exchange_weak(cur, ((uint64_t)top << 32) | id, std::memory_order_acq_rel,
                                      std::memory_order_acquire))
        return;
    }
  }

  void insert(uint32_t id, Scratch& s) {  // :375-449
    const float* q = row(id);
    uint32_t top = g->top_layer[id];
    uint64_t e = entry.load(std::memory_order_acquire);
    if (e == kEmptyEntry) {
      uint64_t expected = kEmptyEntry;
      if (entry.compare_exchange_strong(expected, ((uint64_t)top << 32) | id, std::memory_order_acq_rel,
                                        std::memory_order_acquire))
        return;
      e = expected;
    }
    uint32_t entry_id = (uint32_t)e, entry_top = (uint32_t)(e >> 32);
    uint32_t ep = entry_id;
    for (uint32_t layer = entry_top; layer > top; --layer) ep = descend(q, ep, layer, s);

    s.entry_points.assign(1, ep);
    uint32_t start = std::min(top, entry_top);
    if (s.selected_per_layer.size() < (size_t)start + 1) s.selected_per_layer.resize(start + 1);
    for (uint32_t layer = start + 1; layer-- > 0;) {
      search_layer(q, s.entry_points.data(), s.entry_points.size(), efC, layer, s);
      select_diverse(s.cands.data(), s.cands.size(), M, s);
      s.selected_per_layer[layer] = s.result;
      s.entry_points.clear();
      for (uint64_t k : s.cands) s.entry_points.push_back(key_id(k));
    }
    lock(id);
    for (uint32_t layer = 0; layer <= start; ++layer) {
      uint32_t* deg;
      uint32_t* l = list_ptr(id, layer, &deg);
      const auto& sel = s.selected_per_layer[layer];
      std::memcpy(l, sel.data(), sel.size() * sizeof(uint32_t));
      *deg = (uint32_t)sel.size();
    }
    unlock(id);
    for (uint32_t layer = start + 1; layer-- > 0;) {
      // A private copy: add_backlink reuses s.result.
      std::vector<uint32_t>& sel = s.selected_per_layer[layer];
      for (size_t j = 0; j < sel.size(); ++j) add_backlink(sel[j], id, layer, s);
    }
    raise_entry(id, top);
  }
};

}  // namespace

// Layer draw (ticket == handle) and the empty flat layout.
static void init_layout(uint64_t n, const cm_index_params& p, uint64_t seed, Graph* g) {
  g->n = n;
  g->cap0 = (uint32_t)(p.max_connections * 2);
  g->capU = (uint32_t)p.max_connections;
  g->top_layer.assign(n, 0);
  g->upper_base.assign(n, CM_NO_ID);
  double ml = 1.0 / std::log((double)p.max_connections);  // :104
  uint64_t slots = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t t = draw_layer(seed, i, ml);  // ticket == handle under in-order insertion
    g->top_layer[i] = (uint8_t)t;
    if (t > 0) {
      g->upper_base[i] = (uint32_t)slots;
      slots += t;
    }
  }
  g->nbr0.assign(n * g->cap0, CM_NO_ID);
  g->deg0.assign(n, 0);
  g->nbr_up.assign(slots * g->capU, CM_NO_ID);
  g->deg_up.assign(slots, 0);
  g->entry = kEmptyEntry;
}

void build_graph(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed, int threads,
                 Graph* g) {
  init_layout(n, p, seed, g);

  Builder b;
  b.x = x;
  b.n = n;
  b.dim = dim;
  b.M = (uint32_t)p.max_connections;
  b.efC = (uint32_t)p.ef_construction;
  b.cap0 = g->cap0;
  b.capU = g->capU;
  b.g = g;
  if (threads < 1) threads = 1;
  b.mt = threads > 1;
  if (b.mt) {
    b.locks.reset(new std::atomic<uint8_t>[n]);
    for (uint64_t i = 0; i < n; ++i) b.locks[i].store(0, std::memory_order_relaxed);
  }
  if (n > 0) {
    // The first insert (and a short serial prefix that seeds the upper layers) runs on this thread.
    uint64_t serial = b.mt ? std::min<uint64_t>(n, 256) : n;
    {
      Scratch s;
      for (uint64_t i = 0; i < serial; ++i) b.insert((uint32_t)i, s);
    }
    if (serial < n) {
      std::atomic<uint64_t> next{serial};
      std::vector<std::thread> ts;
      for (int t = 0; t < threads; ++t)
        ts.emplace_back([&] {
          Scratch s;
          for (;;) {
            uint64_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) break;
            b.insert((uint32_t)i, s);
          }
        });
      for (auto& t : ts) t.join();
    }
  }
  g->entry = b.entry.load();
  g->built = true;
}

// ---------------------------------------------------------------------------------------------------------
// GPU-assisted construction. The expensive part of `LockFreeHnsw::insert` is the
// `search_layer(ef_construction)` that finds each new node's candidate neighbors (:419-432); here the candidates of
// EVERY node of a layer come from one exact all-pairs kNN over the layer's members on the tensor cores (the K1 scan
// with the corpus as its own query set, bf16-ranked — see cm_index_build_graph_gpu in api.cu). The host then does
// what the reference does with candidates: exact distances, `select_diverse(cands, M)` (:205-240) for the node's
// own list, and `add_backlink`'s rule (:246-280) for the reverse edges, batched per target. Layers and the entry
// point are the reference's (:126-132, :351-371). The graph is not the one sequential insertion would build (no
// GPU/CPU build can be, the reference's own concurrent build is order dependent, :21-26) but satisfies
// `check_invariants` and is searched by the CPU oracle and the device alike.
// ---------------------------------------------------------------------------------------------------------
template <class F>
static void parallel_range(uint64_t n, int threads, F f) {
  if (threads <= 1 || n < 1024) {
    Scratch s;
    for (uint64_t i = 0; i < n; ++i) f(i, s);
    return;
  }
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; ++t)
    ts.emplace_back([&] {
      Scratch s;
      for (;;) {
        uint64_t b0 = next.fetch_add(256, std::memory_order_relaxed);
        if (b0 >= n) break;
        for (uint64_t i = b0; i < std::min(n, b0 + 256); ++i) f(i, s);
      }
    });
  for (auto& t : ts) t.join();
}

// seen[v] = 1 for every layer-0 node a search can arrive at: from the entry point along the edges of the top layer,
// then, layer by layer, from everything reached above along the edges of the layer below (`LockFreeHnsw::search`
// descends exactly like that, src/index/lockfree_hnsw.rs:481-486). Returns the number of nodes reached.
static uint64_t mark_reachable(const Graph& g, std::vector<uint8_t>& seen) {
  seen.assign(g.n, 0);
  if (g.entry == kEmptyEntry || g.n == 0) return 0;
  std::vector<uint32_t> reached{(uint32_t)g.entry}, stack;
  seen[(uint32_t)g.entry] = 1;
  for (uint32_t layer = (uint32_t)(g.entry >> 32) + 1; layer-- > 0;) {
    stack = reached;  // everything reached above is a starting point on this layer
    while (!stack.empty()) {
      const uint32_t v = stack.back();
      stack.pop_back();
      uint32_t deg;
      const uint32_t* l = g.list(v, layer, &deg);
      for (uint32_t j = 0; j < deg; ++j)
        if (!seen[l[j]]) seen[l[j]] = 1, reached.push_back(l[j]), stack.push_back(l[j]);
    }
  }
  return reached.size();
}

// A kNN-derived graph has no connectivity guarantee: incremental insertion links every new node to something the entry
// point already reaches and creates long-range edges while lists are still short, but all-pairs candidates only ever
// name a node's nearest neighbors, so on clustered data layer 0 can fall apart into components the beam search cannot
// cross. After the build every layer-0 node must be reachable the way a search moves (mark_reachable): each unreached
// component is bridged by the reference's own means — `search_layer(ef_construction)` from the entry finds the
// nearest REACHED nodes to an orphan u, `select_diverse` picks up to four of them, and each pick v gets the edge
// v -> u (appended, or in place of v's farthest neighbor when the list is full; a node orphaned by that is picked up
// by the next round). u -> v is added when u has room. Sequential and deterministic; a no-op on graphs that are already
// connected (one walk, ~0.1 s per million nodes).
static uint64_t repair_layer0_reachability(Builder& b) {
  constexpr size_t kBridges = 4;  // back-links forced per bridged node
  Graph* g = b.g;
  const uint64_t n = b.n;
  if (g->entry == kEmptyEntry || n == 0) return 0;
  const uint32_t entry_id = (uint32_t)g->entry, entry_top = (uint32_t)(g->entry >> 32);
  std::vector<uint8_t> seen;
  std::vector<uint32_t> stack;
  Scratch s;
  uint64_t bridged = 0;
  auto walk_from = [&](uint32_t start) {
    if (seen[start]) return;
    seen[start] = 1;
    stack.push_back(start);
    while (!stack.empty()) {
      const uint32_t v = stack.back();
      stack.pop_back();
      const uint32_t* l = &g->nbr0[(uint64_t)v * b.cap0];
      for (uint32_t j = 0; j < g->deg0[v]; ++j)
        if (!seen[l[j]]) seen[l[j]] = 1, stack.push_back(l[j]);
    }
  };
  for (int round = 0; round < 8; ++round) {
    if (mark_reachable(*g, seen) == n) break;
    uint64_t fixed_this_round = 0;
    for (uint64_t u = 0; u < n; ++u) {
      if (seen[u]) continue;
      // nearest reached nodes to u: the search itself only ever visits nodes the entry reaches
      const float* q = b.row((uint32_t)u);
      uint32_t ep = entry_id;
      for (uint32_t layer = entry_top; layer >= 1; --layer) ep = b.descend(q, ep, layer, s);
      b.search_layer(q, &ep, 1, b.efC, 0, s);
      // what `insert` would link u to: a diverse pick of the nearest reached nodes (:419-432) ...
      s.with_d.clear();
      for (uint64_t k : s.cands)
        if (key_id(k) != u && seen[key_id(k)]) s.with_d.push_back(k);
      if (s.with_d.empty()) continue;
      b.select_diverse(s.with_d.data(), s.with_d.size(), kBridges, s);
      const std::vector<uint32_t> picks = s.result;
      uint32_t* lu = &g->nbr0[u * b.cap0];
      uint32_t& du = g->deg0[u];
      for (uint32_t v : picks) {
        // ... and the back-link v -> u that makes u reachable, which `add_backlink` (:246-280) may drop when v's list
        // is full: here it must stay, so it takes the place of v's farthest neighbor instead
        uint32_t* lv = &g->nbr0[(uint64_t)v * b.cap0];
        uint32_t& dv = g->deg0[v];
        if (std::find(lv, lv + dv, (uint32_t)u) == lv + dv) {
          if (dv < b.cap0) {
            lv[dv++] = (uint32_t)u;
          } else {
            uint32_t far = 0;
            uint64_t far_key = 0;
            for (uint32_t j = 0; j < dv; ++j) {
              const uint64_t k = pack(dist_prepared(b.row(lv[j]), b.row(v), b.dim), lv[j]);
              if (k >= far_key) far_key = k, far = j;
            }
            lv[far] = (uint32_t)u;
          }
        }
        if (du < b.cap0 && std::find(lu, lu + du, v) == lu + du) lu[du++] = v;
      }
      walk_from((uint32_t)u);  // u's whole component is reached now
      ++fixed_this_round;
    }
    bridged += fixed_this_round;
    if (fixed_this_round == 0) break;
  }
  return bridged;
}

void build_graph_from_candidates(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed,
                                 int threads, const std::vector<LayerCandidates>& layers, Graph* g) {
  init_layout(n, p, seed, g);
  Builder b;
  b.x = x;
  b.n = n;
  b.dim = dim;
  b.M = (uint32_t)p.max_connections;
  b.efC = (uint32_t)p.ef_construction;
  b.cap0 = g->cap0;
  b.capU = g->capU;
  b.g = g;
  if (threads < 1) threads = 1;
  b.mt = threads > 1;
  if (b.mt) {
    b.locks.reset(new std::atomic<uint8_t>[n]);
    for (uint64_t i = 0; i < n; ++i) b.locks[i].store(0, std::memory_order_relaxed);
  }
  for (uint32_t layer = 0; layer < layers.size(); ++layer) {
    const LayerCandidates& lc = layers[layer];
    const uint64_t nm = lc.members.size();
    std::vector<uint32_t> own(nm * b.M, CM_NO_ID);
    std::vector<uint32_t> own_deg(nm, 0);  // max_connections has no upper bound in Config::validate
    // phase A: every member's own list from its candidates (no shared writes)
    parallel_range(nm, threads, [&](uint64_t i, Scratch& s) {
      const uint32_t u = lc.members[i];
      s.with_d.clear();
      for (uint32_t j = 0; j < lc.width; ++j) {
        uint32_t v = lc.ids[i * lc.width + j];
        if (v == CM_NO_ID || v == u || v >= n || g->top_layer[v] < layer) continue;  // only members of this layer link here
        s.with_d.push_back(pack(dist_prepared(b.row(v), b.row(u), dim), v));
      }
      std::sort(s.with_d.begin(), s.with_d.end());
      s.with_d.erase(std::unique(s.with_d.begin(), s.with_d.end()), s.with_d.end());
      b.select_diverse(s.with_d.data(), s.with_d.size(), b.M, s);
      own_deg[i] = (uint32_t)s.result.size();
      std::memcpy(&own[i * b.M], s.result.data(), s.result.size() * sizeof(uint32_t));
      uint32_t* deg;
      uint32_t* l = b.list_ptr(u, layer, &deg);
      std::memcpy(l, s.result.data(), s.result.size() * sizeof(uint32_t));
      *deg = (uint32_t)s.result.size();
    });
    // phase B: reverse edges, batched per target. `add_backlink` (:246-280) appends u to v's list and, when the
    // list overflows, re-selects it from (list + u) sorted by distance to v; replaying that one edge at a time
    // re-selects a popular node many times over. All edges into v are known here, so v's list is selected ONCE
    // from own(v) + every u with v in own(u): same selection rule, same caps, no locks (each node writes only its
    // own list), and a result that does not depend on thread interleaving.
    std::vector<uint32_t> slot_of(n, CM_NO_ID);  // handle -> position in lc.members
    for (uint64_t i = 0; i < nm; ++i) slot_of[lc.members[i]] = (uint32_t)i;
    std::vector<uint32_t> in_start(nm + 1, 0);
    for (uint64_t i = 0; i < nm; ++i)
      for (uint32_t j = 0; j < own_deg[i]; ++j) ++in_start[slot_of[own[i * b.M + j]] + 1];
    for (uint64_t i = 0; i < nm; ++i) in_start[i + 1] += in_start[i];
    std::vector<uint32_t> incoming(in_start[nm]);
    {
      std::vector<uint32_t> fill(in_start.begin(), in_start.end() - 1);
      for (uint64_t i = 0; i < nm; ++i)  // ascending source handle within each target
        for (uint32_t j = 0; j < own_deg[i]; ++j) incoming[fill[slot_of[own[i * b.M + j]]]++] = lc.members[i];
    }
    const uint32_t cap = b.cap(layer);
    parallel_range(nm, threads, [&](uint64_t i, Scratch& s) {
      const uint32_t v = lc.members[i];
      uint32_t* deg;
      uint32_t* l = b.list_ptr(v, layer, &deg);
      // own list first, then the sources not already in it (what appending in handle order would leave)
      s.nbuf.clear();
      for (uint32_t j = 0; j < own_deg[i]; ++j) s.nbuf.push_back(own[i * b.M + j]);
      for (uint32_t e = in_start[i]; e < in_start[i + 1]; ++e) {
        const uint32_t u = incoming[e];
        if (std::find(s.nbuf.begin(), s.nbuf.begin() + own_deg[i], u) == s.nbuf.begin() + own_deg[i]) s.nbuf.push_back(u);
      }
      if (s.nbuf.size() <= cap) {
        std::memcpy(l, s.nbuf.data(), s.nbuf.size() * sizeof(uint32_t));
        *deg = (uint32_t)s.nbuf.size();
        return;
      }
      s.with_d.clear();
      for (uint32_t u : s.nbuf) s.with_d.push_back(pack(dist_prepared(b.row(u), b.row(v), dim), u));
      std::sort(s.with_d.begin(), s.with_d.end());
      b.select_diverse(s.with_d.data(), s.with_d.size(), cap, s);
      std::memcpy(l, s.result.data(), s.result.size() * sizeof(uint32_t));
      *deg = (uint32_t)s.result.size();
    });
  }
  // raise_entry under in-order insertion (:351-371): the first handle that reaches the highest layer
  for (uint64_t i = 0; i < n; ++i)
    if (g->entry == kEmptyEntry || g->top_layer[i] > (uint32_t)(g->entry >> 32)) g->entry = ((uint64_t)g->top_layer[i] << 32) | i;
  b.entry.store(g->entry);
  b.mt = false;  // the repair below is sequential (and deterministic)
  repair_layer0_reachability(b);
  g->built = true;
}

// Reverse-edge lists by target (member-local indices): for every member, the members that chose it, in ascending
// source order — what `add_backlink` sees when nodes are inserted in handle order.
void build_incoming_csr(const uint32_t* own, const uint32_t* own_deg, uint64_t nm, uint32_t width,
                        std::vector<uint32_t>* in_start, std::vector<uint32_t>* incoming) {
  in_start->assign(nm + 1, 0);
  for (uint64_t i = 0; i < nm; ++i)
    for (uint32_t j = 0; j < own_deg[i]; ++j) ++(*in_start)[own[i * width + j] + 1];
  for (uint64_t i = 0; i < nm; ++i) (*in_start)[i + 1] += (*in_start)[i];
  incoming->resize((*in_start)[nm]);
  std::vector<uint32_t> fill(in_start->begin(), in_start->end() - 1);
  for (uint64_t i = 0; i < nm; ++i)
    for (uint32_t j = 0; j < own_deg[i]; ++j) (*incoming)[fill[own[i * width + j]]++] = (uint32_t)i;
}

// The graph from per-layer neighbor lists selected on the device (build_select.cu): write them into the flat layout,
// select the few items the kernel left to the host (hubs with more incoming edges than it sorts) with the same code the
// host path uses, then entry point and reachability as in build_graph_from_candidates.
void build_graph_from_layer_lists(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed,
                                  int threads, const std::vector<LayerLists>& layers, Graph* g) {
  init_layout(n, p, seed, g);
  Builder b;
  b.x = x;
  b.n = n;
  b.dim = dim;
  b.M = (uint32_t)p.max_connections;
  b.efC = (uint32_t)p.ef_construction;
  b.cap0 = g->cap0;
  b.capU = g->capU;
  b.g = g;
  b.mt = false;
  for (uint32_t layer = 0; layer < layers.size(); ++layer) {
    const LayerLists& ll = layers[layer];
    const uint64_t nm = ll.members.size();
    if (ll.deg.empty()) continue;
    parallel_range(nm, threads, [&](uint64_t i, Scratch&) {
      uint32_t* deg;
      uint32_t* l = b.list_ptr(ll.members[i], layer, &deg);
      const uint32_t d = ll.deg[i];
      for (uint32_t j = 0; j < d; ++j) l[j] = ll.members[ll.lists[i * ll.cap + j]];
      *deg = d;
    });
    const uint32_t cap = b.cap(layer);
    parallel_range(ll.too_big.size(), threads, [&](uint64_t t, Scratch& s) {
      const uint32_t i = ll.too_big[t];
      const uint32_t v = ll.members[i];
      s.nbuf.clear();
      const uint32_t od = ll.own_deg[i];
      for (uint32_t j = 0; j < od; ++j) s.nbuf.push_back(ll.members[ll.own[(uint64_t)i * ll.own_width + j]]);
      for (uint32_t e = ll.in_start[i]; e < ll.in_start[i + 1]; ++e) {
        const uint32_t u = ll.members[ll.incoming[e]];
        if (std::find(s.nbuf.begin(), s.nbuf.begin() + od, u) == s.nbuf.begin() + od) s.nbuf.push_back(u);
      }
      s.with_d.clear();
      for (uint32_t u : s.nbuf) s.with_d.push_back(pack(dist_prepared(b.row(u), b.row(v), dim), u));
      std::sort(s.with_d.begin(), s.with_d.end());
      b.select_diverse(s.with_d.data(), s.with_d.size(), cap, s);
      uint32_t* deg;
      uint32_t* l = b.list_ptr(v, layer, &deg);
      std::memcpy(l, s.result.data(), s.result.size() * sizeof(uint32_t));
      *deg = (uint32_t)s.result.size();
    });
  }
  for (uint64_t i = 0; i < n; ++i)
    if (g->entry == kEmptyEntry || g->top_layer[i] > (uint32_t)(g->entry >> 32)) g->entry = ((uint64_t)g->top_layer[i] << 32) | i;
  b.entry.store(g->entry);
  repair_layer0_reachability(b);
  g->built = true;
}

// Layer-0 nodes a search from the entry point cannot arrive at (see mark_reachable). 0 for an empty graph.
uint64_t count_unreachable(const Graph& g) {
  std::vector<uint8_t> seen;
  return g.n - (g.entry == kEmptyEntry ? g.n : mark_reachable(g, seen));
}

// Members of layer `layer` (ascending handles) for the layer draw of (n, params, seed).
void layer_members(uint64_t n, const cm_index_params& p, uint64_t seed, std::vector<std::vector<uint32_t>>* out) {
  double ml = 1.0 / std::log((double)p.max_connections);
  out->clear();
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t t = draw_layer(seed, i, ml);
    if (out->size() < (size_t)t + 1) out->resize(t + 1);
    for (uint32_t l = 0; l <= t; ++l) (*out)[l].push_back((uint32_t)i);
  }
}

// src/index/lockfree_hnsw.rs:295-347 on the flat layout. 0 == sound.
int check_graph_invariants(const Graph& g) {
  uint64_t total = g.n;
  if (g.entry != kEmptyEntry) {
    uint32_t eid = (uint32_t)g.entry, etop = (uint32_t)(g.entry >> 32);
    if (eid >= total) return 1;
    if (g.top_layer[eid] < etop) return 2;
  }
  std::vector<uint32_t> seen;
  for (uint64_t id = 0; id < total; ++id) {
    for (uint32_t layer = 0; layer <= g.top_layer[id]; ++layer) {
      uint32_t deg;
      const uint32_t* l = g.list((uint32_t)id, layer, &deg);
      if (deg > (layer == 0 ? g.cap0 : g.capU)) return 3;
      seen.assign(l, l + deg);
      std::sort(seen.begin(), seen.end());
      if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) return 6;
      for (uint32_t j = 0; j < deg; ++j) {
        if (l[j] == id) return 4;
        if (l[j] >= total) return 5;
        if (g.top_layer[l[j]] < layer) return 7;
      }
    }
  }
  return 0;
}

float dist_prepared_host(const float* a, const float* b, uint32_t n) { return dist_prepared(a, b, n); }

void preprocess_row(const float* v, uint32_t n, float* out) {  // src/metric.rs:208-214
  float s = 0.0f;
  for (uint32_t i = 0; i < n; ++i) s += v[i] * v[i];
  float norm = std::sqrt(s);
  if (norm <= kF32Epsilon) {
    if (out != v) std::memcpy(out, v, (size_t)n * sizeof(float));
    return;
  }
  for (uint32_t i = 0; i < n; ++i) out[i] = v[i] / norm;
}

}  // namespace cm
