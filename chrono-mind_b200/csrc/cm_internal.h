// cm_internal.h — shared declarations between the host runtime (index, snapshot, graph builder) and the
// CUDA side. Not part of the ABI.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/chronomind_b200.h"

namespace cm {

constexpr uint64_t kArenaCapacity = 4096ull * 4096ull;  // src/index/arena.rs:47-50,101
constexpr uint32_t kMaxLayer = 31;                      // src/index/lockfree_hnsw.rs:43
constexpr uint64_t kEmptyEntry = ~0ull;                 // src/index/lockfree_hnsw.rs:46
constexpr float kF32Epsilon = 1.1920929e-7f;

// f32::total_cmp as an unsigned key (src/index/mod.rs:37-52).
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t total_key(float f) {
#ifdef __CUDA_ARCH__
  // An opaque bit cast: with __float_as_uint nvcc 12.9 may keep the value in the float domain and set the sign
  // bit with `FADD -|x|, -0`, which canonicalises a NaN (payload and sign lost => wrong key for NaN distances).
  uint32_t b;
  asm("mov.b32 %0, %1;" : "=r"(b) : "f"(f));
#else
  uint32_t b;
  std::memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

void set_error(const std::string& msg);  // thread-local, read by cm_last_error()

// ---------------------------------------------------------------------------------------------------------
// Host-side HNSW graph in the flat layout the device consumes.
//   layer 0 : nbr0[n][cap0] + deg0[n]                  (cap0 = 2M)
//   layer>0 : only nodes with top_layer >= 1 own lists. upper_base[v] is the slot of v's layer-1 list in
//             nbr_up[slot][capU] / deg_up[slot] (capU = M); layer l lives at slot upper_base[v] + l - 1.
// ---------------------------------------------------------------------------------------------------------
struct Graph {
  uint32_t cap0 = 0, capU = 0;
  uint64_t n = 0;
  std::vector<uint8_t> top_layer;    // [n]
  std::vector<uint32_t> nbr0;        // [n * cap0]
  std::vector<uint32_t> deg0;        // [n]
  std::vector<uint32_t> upper_base;  // [n], CM_NO_ID when top_layer == 0
  std::vector<uint32_t> nbr_up;      // [slots * capU]
  std::vector<uint32_t> deg_up;      // [slots]
  uint64_t entry = kEmptyEntry;      // (top_layer << 32) | id
  bool built = false;

  const uint32_t* list(uint32_t v, uint32_t layer, uint32_t* deg) const {
    if (layer == 0) {
      *deg = deg0[v];
      return &nbr0[(uint64_t)v * cap0];
    }
    uint32_t slot = upper_base[v] + layer - 1;
    *deg = deg_up[slot];
    return &nbr_up[(uint64_t)slot * capU];
  }
};

// Build `g` over the preprocessed rows x[n][dim] following LockFreeHnsw::insert. See hnsw_build.cpp.
void build_graph(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed, int threads,
                 Graph* g);
int check_graph_invariants(const Graph& g);
uint64_t count_unreachable(const Graph& g);  // layer-0 nodes the entry point cannot reach

// GPU-assisted construction (hnsw_build.cpp): per layer, the members (ascending handles) and `width` candidate
// neighbors per member (global handles, CM_NO_ID padded; may contain the member itself).
struct LayerCandidates {
  std::vector<uint32_t> members;
  std::vector<uint32_t> ids;  // [members.size()][width]
  uint32_t width = 0;
};
void layer_members(uint64_t n, const cm_index_params& p, uint64_t seed, std::vector<std::vector<uint32_t>>* out);
void build_graph_from_candidates(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed,
                                 int threads, const std::vector<LayerCandidates>& layers, Graph* g);

// Per-layer neighbor lists selected on the device (member-local indices), see build_select.cu / hnsw_build.cpp.
struct LayerLists {
  std::vector<uint32_t> members;
  uint32_t cap = 0, own_width = 0;
  std::vector<uint32_t> lists, deg;    // [members][cap], [members] — final lists (empty: the layer has no edges)
  std::vector<uint32_t> own, own_deg;  // [members][own_width], [members] — each member's own pick (phase A)
  std::vector<uint32_t> in_start, incoming;
  std::vector<uint32_t> too_big;       // items the kernel left to the host
};
void build_incoming_csr(const uint32_t* own, const uint32_t* own_deg, uint64_t nm, uint32_t width,
                        std::vector<uint32_t>* in_start, std::vector<uint32_t>* incoming);
void build_graph_from_layer_lists(const float* x, uint64_t n, uint32_t dim, const cm_index_params& p, uint64_t seed,
                                  int threads, const std::vector<LayerLists>& layers, Graph* g);

// Host metric kernels (same arithmetic as src/metric.rs; see metric_host.h).
void preprocess_row(const float* v, uint32_t n, float* out);
float dist_prepared_host(const float* a, const float* b, uint32_t n);  // distance_prepared, AVX2+FMA lane order

// Row matrices (gigabytes at the BASELINE sizes) are always filled right after they are sized — by all host threads —
// so `resize` must not value-initialise them first: a single-threaded zero fill of 3 GB was two thirds of the snapshot
// load time. `construct(p)` without arguments default-initialises (no store); everything else is std::allocator.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAllocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U>
  void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void*>(p)) U;
  }
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
using RowMatrix = std::vector<float, DefaultInitAllocator<float>>;

// ---------------------------------------------------------------------------------------------------------
// Snapshot records (src/types.rs) as parsed by snapshot.cpp.
// ---------------------------------------------------------------------------------------------------------
struct SnapshotMemory {
  std::string id;
  uint64_t data_offset;  // into SnapshotData::vectors (floats)
  uint64_t data_len;     // floats in the record (== config.dimensions once validated)
  bool data_finite = true;  // no NaN/Inf among them
  uint64_t ts_secs;
  uint32_t ts_nanos;
  float importance;
  std::string context;
  float decay_rate;
  std::vector<std::string> relationships;
  uint32_t access_count;
  uint64_t last_access_secs;
  uint32_t last_access_nanos;
};
struct SnapshotData {
  cm_config config;
  std::vector<SnapshotMemory> memories;
  RowMatrix vectors;  // the records' data back to back, file order (data_offset / data_len)
};
// Returns CM_OK or the status load_snapshot would map to; message via set_error().
cm_status read_snapshot(const char* path, SnapshotData* out);  // decode + Config::validate only
cm_status validate_snapshot_memory(const SnapshotData& snap, uint64_t i);  // Memory::validate of record i
cm_status validate_config(const cm_config& c);

// ---------------------------------------------------------------------------------------------------------
// Device-resident index (filled by cm_index_upload; consumed by the kernels).
// ---------------------------------------------------------------------------------------------------------
struct DeviceIndex {
  int device = -1;
  uint64_t n = 0;
  uint32_t dim = 0;
  uint32_t dim_pad = 0;       // fp32 row stride in floats (multiple of 8 -> 32-byte aligned rows)
  float* x32 = nullptr;       // [n][dim_pad] preprocessed rows, zero padded
  void* xbf16 = nullptr;      // [n_pad][k_pad] bf16 copy for the tensor-core scan
  uint64_t n_pad = 0;         // rows padded to the scan's tile height
  uint32_t k_pad = 0;         // columns padded to a multiple of 64
  uint8_t* deleted = nullptr; // [n]
  // store-level context scan (`search_in_context`): raw (un-preprocessed) rows, their squared norms, context ids
  float* xraw = nullptr;      // [n][dim_pad], only when the index carries contexts
  float* norm2_raw = nullptr; // [n]
  uint32_t* ctx = nullptr;    // [n_pad] (padding = CM_NO_ID)
  // the same rows SORTED BY CONTEXT (position p holds handle perm[p]): a query only has to scan the tiles of its context
  void* xbf16_ctx = nullptr;     // [n_pad][k_pad] bf16, NaN-masked like xbf16
  uint32_t* ctx_sorted = nullptr;  // [n_pad] context id per position (padding = CM_NO_ID)
  uint32_t* perm = nullptr;      // [n_pad] position -> handle
  uint32_t* inv_perm = nullptr;  // [n] handle -> position
  uint32_t n_contexts = 0;       // distinct context ids (the sorted copy exists from 2 up)
  uint64_t* ts_secs = nullptr;
  uint32_t* ts_nanos = nullptr;
  float* decay = nullptr;
  // graph
  bool has_graph = false;
  uint32_t cap0 = 0, capU = 0;
  uint32_t* nbr0 = nullptr;
  uint32_t* deg0 = nullptr;
  uint32_t* upper_base = nullptr;
  uint32_t* nbr_up = nullptr;
  uint32_t* deg_up = nullptr;
  uint32_t entry_id = 0, entry_top = 0;
  bool has_entry = false;
  uint64_t live = 0;
  int sm_count = 0;
  // Device status words for the asynchronous (_dev) entry points, which cannot report through a return value:
  // [0] HNSW queries that could not be answered like the reference (visited-table exhaustion, oversized tie group),
  // [1] exact-scan queries left unresolved (more flagged queries than the in-stream fp32 fallback absorbs),
  // [2..3] rows re-scored by K2 (u64) and [4] queries routed to the fp32 fallback since the last reset — the work
  // counters a host-buffer call reports through cm_last_counters, for callers that never synchronise.
  uint32_t* status = nullptr;
};

}  // namespace cm

// The opaque ABI type.
struct cm_index {
  cm_config config{};
  uint64_t n = 0;
  uint32_t dim = 0;
  cm::RowMatrix x;                    // [n][dim] preprocessed rows (host copy: graph build + upload source)
  std::vector<uint8_t> deleted;     // [n]
  std::vector<uint32_t> pending_removed;  // tombstoned since the last upload / cm_index_sync_deleted
  std::vector<uint64_t> ts_secs;    // [n]
  std::vector<uint32_t> ts_nanos;   // [n]
  std::vector<float> decay;         // [n]
  std::vector<std::string> ext_ids; // snapshot-loaded only
  cm::RowMatrix raw;                  // [n][dim] rows as stored (`StoredMemory.data`), only with contexts
  std::vector<uint32_t> ctx;        // [n] context id per handle (empty: no contexts)
  std::vector<uint32_t> ctx_sorted; // context ids in (context, handle) order: the row ranges of the context-sorted device copy
  std::vector<std::string> ctx_labels;  // snapshot-loaded: id -> `MemoryAttributes.context`
  std::vector<float> importance;    // snapshot-loaded: `MemoryAttributes.importance`
  uint64_t live = 0;
  bool all_finite = true;           // no NaN/Inf component in any stored row (gate of the tensor-core scan)
  cm::Graph graph;
  uint32_t shard_rank = 0, shard_count = 1;
  int exact_engine = 0;
  bool uploaded = false;
  cm::DeviceIndex dev;
  void* workspace_pool = nullptr;   // cm::WorkspacePool*, owned (see api.cu)
};
