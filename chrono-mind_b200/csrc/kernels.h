// kernels.h — launchers of the CUDA kernels (defined in the .cu files of this directory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cm_internal.h"

namespace cm {

void count_launch();  // bumps the process-wide and the thread-local launch counters

// K0 — CosineDistance::preprocess per row (src/metric.rs:208-214). out may alias x when ld_in == ld_out.
// Columns [dim, ld_out) of out are zeroed. bad_row (optional, device u32 preset to UINT32_MAX) receives the
// smallest row index holding a non-finite component (validate_query, src/store.rs:386-390).
cudaError_t launch_preprocess_rows(const float* x, uint64_t n, uint32_t dim, uint32_t ld_in, float* out,
                                   uint32_t ld_out, uint32_t* bad_row, cudaStream_t s);

// K0 fused with the bf16 cast for query batches: out [n][ld_out] fp32 prepared rows (columns [dim, ld_out) zeroed) and
// out16 [n_pad16][k_pad] bf16 (zero padded rows and columns) in one pass. Needs dim % 4 == 0, 16-byte aligned x and
// rows short enough for 16 of them in shared memory (prepare_queries_fused_supported); x must not alias out.
bool prepare_queries_fused_supported(const float* x, uint32_t dim);
// src_rows (optional, device [n]): output row i is built from x row src_rows[i] (bad_row reports the x row).
cudaError_t launch_prepare_queries(const float* x, uint32_t n, uint32_t dim, float* out, uint32_t ld_out, void* out16,
                                   uint32_t n_pad16, uint32_t k_pad, uint32_t* bad_row, cudaStream_t s,
                                   const uint32_t* src_rows = nullptr);

// fp32 -> bf16 copy of prepared rows into the tensor-core scan layout [n_pad][k_pad] (zero padded).
// mask_rows (corpus copies): tombstoned rows (deleted[row] != 0) and padding rows score NaN in the scan.
// row_map (optional, device [n]): output row i is x row row_map[i] (a permuted copy; `deleted` is indexed by x row).
cudaError_t launch_to_bf16(const float* x, uint64_t n, uint32_t dim, uint32_t ld_in, void* out, uint64_t n_pad,
                           uint32_t k_pad, const uint8_t* deleted, bool mask_rows, cudaStream_t s,
                           const uint32_t* row_map = nullptr);

// Tombstone `m` handles on the device replica (deleted byte + NaN mask of the bf16 row).
// (xbf16_ctx / inv_perm: the context-sorted bf16 copy and handle -> position in it, when the index carries contexts)
cudaError_t launch_apply_tombstones(const uint32_t* handles, uint32_t m, uint8_t* deleted, void* xbf16, uint32_t k_pad,
                                    void* xbf16_ctx, const uint32_t* inv_perm, cudaStream_t s);

// distance_prepared for n row pairs (src/metric.rs:219-224).
cudaError_t launch_pair_distance(const float* a, const float* b, uint64_t n, uint32_t dim, uint32_t ld, float* out,
                                 cudaStream_t s);

// dot_avx2(a[i], b[i]) for n row pairs (a == b: the squared norm as dot_and_norms_avx2 accumulates it).
cudaError_t launch_pair_dot(const float* a, const float* b, uint64_t n, uint32_t dim, uint32_t ld, float* out,
                            cudaStream_t s);

// Epilogue of the fp32 scan for `ChronoMind::search_in_context` (src/store.rs:353-377): x and q are RAW rows; the
// kernel writes combined_score(distance(x[row], q[qi])) for the live members of context want[qi], +inf elsewhere.
struct ContextEpilogue {
  int enabled = 0;
  const float* na = nullptr;       // [n]  squared norms of the raw rows
  const float* nb = nullptr;       // [nq] squared norms of the raw queries
  const uint32_t* ctx = nullptr;   // [n]  context id per handle
  const uint32_t* want = nullptr;  // [nq] context id per query
  const uint8_t* deleted = nullptr;
  const unsigned long long* ts_secs = nullptr;
  const uint32_t* ts_nanos = nullptr;
  const float* decay = nullptr;
  float w = 0.0f, base = 0.0f;
  unsigned long long now_secs = 0;
  uint32_t now_nanos = 0;
};

// fp32 exact scan, CUDA cores, bit-exact to dot_avx2: D[qi][row] = distance_prepared(x[row], q[qi]).
// active_count (optional, device): query i of this launch is computed only when active_base + i < *active_count.
cudaError_t launch_exact_dist(const float* q, uint32_t nq, uint32_t ldq, const float* x, uint64_t n, uint32_t dim,
                              uint32_t ldx, float* D, uint64_t ldD, const uint32_t* active_count, uint32_t active_base,
                              cudaStream_t s, const ContextEpilogue* ce = nullptr);
cudaError_t launch_fill(float* p, uint64_t n, float v, cudaStream_t s);
// tensor-core scan bookkeeping: fold one query chunk's {rescored, flagged} into the call totals
// (`status`: optional device word that receives the queries beyond `flag_cap`, which stay unresolved)
cudaError_t launch_fold_scan_counters(const void* chunk, void* total, uint32_t flag_cap, uint32_t* status, cudaStream_t s);

// k smallest (distance, handle) per row of D, tombstones dropped.
// qmap/active_count (optional, device): row b of D belongs to output row qmap[b]; rows with
// active_base + b >= *active_count are skipped.
cudaError_t launch_select_from_dist(const float* D, uint64_t ldD, uint32_t nq, uint64_t n, const uint8_t* deleted,
                                    uint32_t k, uint32_t* ids, float* dist, uint32_t out_ld, const uint32_t* qmap,
                                    const uint32_t* active_count, uint32_t active_base, cudaStream_t s,
                                    bool skip_inf = false);

// consolidate's pair test (src/store.rs:531-534) on candidate pairs (a << 32 | b) of raw rows: sim[p] = similarity.
cudaError_t launch_pair_similarity(const float* x, uint32_t ld, uint32_t dim, const float* norm2,
                                   const unsigned long long* pairs, uint32_t n_pairs, float* sim, cudaStream_t s);
// apply_decay sweep (src/store.rs:428-458) over caller-owned state arrays.
cudaError_t launch_decay_sweep(float* importance, unsigned long long* decayed_through, const unsigned long long* last_access,
                               const float* decay_rate, uint64_t n, float base_rate, unsigned long long now_nanos,
                               cudaStream_t s);

// K5 — shard merge (src/index/sharded_rwlock.rs:81-94).
// Input either as (ids, dist) arrays or as `packed` keys (launch_pack_topk), [n_shards][nq][len] local handles.
cudaError_t launch_merge_topk(const uint32_t* ids, const float* dist, const unsigned long long* packed, uint32_t n_shards,
                              uint32_t nq, uint32_t len, uint32_t out_len, uint32_t* out_ids, float* out_dist,
                              cudaStream_t s);
cudaError_t launch_pack_topk(const uint32_t* ids, const float* dist, uint64_t n, unsigned long long* out, cudaStream_t s);

// K4 — rerank tail of ChronoMind::search (src/store.rs:327-344, 395-415).
cudaError_t launch_rerank(const uint32_t* pool_ids, const float* pool_dist, uint32_t nq, uint32_t pool, uint32_t k,
                          const uint64_t* ts_secs, const uint32_t* ts_nanos, const float* decay, float w, float base,
                          uint64_t now_secs, uint32_t now_nanos, uint32_t* out_ids, float* out_score,
                          float* out_dist, cudaStream_t s);

// K3 — batched HNSW search (src/index/lockfree_hnsw.rs:140-194, 469-495). q: prepared queries [nq][ldq].
// Writes the tombstone-filtered candidate pool, ascending (distance, handle): pool_ids/pool_dist [nq][ef]
// padded with CM_NO_ID / +inf.
struct HnswLaunch {
  const float* q;
  uint32_t nq, ldq, ef;
  uint32_t degenerate;       // query length != dim: all distances are 2.0
  uint32_t* pool_ids;
  float* pool_dist;
  unsigned long long* counters;  // device, 8 x u64: dist_evals, expansions, list_entries, gtable restarts, failures
  uint32_t* work_counter;    // device u32, zeroed by the launcher
  uint32_t table_entries;    // shared-memory visited table entries (u32 each)
  uint32_t* gtable;          // per-CTA overflow tables in global memory, grid << ghash_bits entries
  uint32_t ghash_bits;
  unsigned grid;             // from hnsw_grid_size()
  uint32_t rows;             // rows staged per copy round, from hnsw_grid_size()
  uint32_t* status = nullptr;  // optional device word bumped once per query that could not be answered exactly
  // visited set in HBM (large ef): one slab of vmap_stride bytes (>= rows, multiple of 16) per CTA of `grid`, zeroed when
  // allocated, plus the per-CTA epoch words (zeroed with it). table_entries must be 0 then.
  unsigned char* vmap = nullptr;
  uint64_t vmap_stride = 0;
  uint32_t* vmap_epoch = nullptr;
};
cudaError_t hnsw_grid_size(const DeviceIndex& ix, uint32_t ef, uint32_t table_entries, unsigned* grid, uint32_t* rows);
cudaError_t launch_hnsw_search(const DeviceIndex& ix, const HnswLaunch& a, cudaStream_t s);

// In-stream fp32 fallback of the tensor-core scan: one cooperative launch re-runs the *flag_count queries listed in
// flag_list through the bit-exact scan + select and overwrites their rows of ids/dist [nq][k]; an immediate return
// when nothing is flagged. qc: scratch of fallback_exact_query_bytes(ldq), D: scratch of fallback_exact_dist_bytes(n).
size_t fallback_exact_dist_bytes(uint64_t n);
size_t fallback_exact_query_bytes(uint32_t ldq);
cudaError_t launch_fallback_exact(const float* qn, uint32_t ldq, const uint32_t* flag_list, const uint32_t* flag_count,
                                  float* qc, const float* x, uint64_t n, uint32_t dim, uint32_t ldx, float* D,
                                  const uint8_t* deleted, uint32_t k, uint32_t* ids, float* dist, int sm_count,
                                  cudaStream_t s);

// `select_diverse` (src/index/lockfree_hnsw.rs:205-240) for a batch of (center row, candidate list) items on the device,
// bit-identical to the host code of hnsw_build.cpp (see build_select.cu). Item i's center is row i of `ix.x32`; its
// candidates are either cand[i][0..width) (CM_NO_ID padded; phase A: a node's own list, m = M) or own[i][0..own_deg[i])
// followed by incoming[in_start[i]..in_start[i+1]) minus duplicates (phase B: the reverse edges, m = cap(layer); items
// that fit are written out unselected, like add_backlink's append branch). Items with more than 1024 candidates are
// listed in too_big ([0] = count) for the host.
struct SelectDiverseLaunch {
  uint32_t n_items = 0, m = 0;
  const uint32_t* cand = nullptr;
  uint32_t width = 0;
  const uint32_t* own = nullptr;
  const uint32_t* own_deg = nullptr;
  uint32_t own_width = 0;
  const uint32_t* in_start = nullptr;
  const uint32_t* incoming = nullptr;
  uint32_t* out = nullptr;       // [n_items][m]
  uint32_t* out_deg = nullptr;   // [n_items]
  uint32_t* too_big = nullptr;   // [1 + too_big_cap], zeroed by the caller
  uint32_t too_big_cap = 0;
  uint32_t* work_counter = nullptr;
};
bool select_diverse_supported(uint32_t dim_pad, uint32_t m);
cudaError_t launch_select_diverse(const DeviceIndex& ix, const SelectDiverseLaunch& l, cudaStream_t s);

// Copy the rows named by a device list into a compact matrix.
cudaError_t launch_gather_rows(const float* src, uint32_t ld, const uint32_t* list, const uint32_t* count,
                               uint32_t cap, float* dst, cudaStream_t s);

}  // namespace cm
