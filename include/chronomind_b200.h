/* chronomind_b200.h — C ABI of the B200-native batched search engine for ChronoMind's read path.
 *
 * This header is the drop-in boundary. Every entry point names the reference interface it replaces
 * (paths relative to the JtPerez-Acle/chrono-mind 0.2.5 checkout). A Rust `extern "C"` block, a cgo
 * preamble or a ctypes stub binds these 1:1. No CUDA or torch types appear in
 * the signatures: device buffers travel as `void*`, streams as `void*` (a `cudaStream_t`).
 *
 * Conventions kept from the reference
 *   - handles are dense u32 in insertion order: handle i == row i (src/index/mod.rs:8-10);
 *   - results are ascending by (distance, handle) — the (TotalF32, u32) tuple order of
 *     src/index/lockfree_hnsw.rs:149-150 — or by score for the temporal search (src/store.rs:338);
 *   - fewer than `k` results is legal (small index, tombstones): rows are padded with
 *     id = CM_NO_ID, dist/score = +inf;
 *   - index-level search never fails on degenerate vectors; store-level search returns typed
 *     errors (src/store.rs:379-392) — here: a status code plus cm_last_error();
 *   - all cm_search_* calls are re-entrant on a shared, uploaded index (mirrors `&self` + Send + Sync).
 *
 * There is no CPU fallback anywhere behind this ABI: searches on an index that was not uploaded to a
 * CUDA device fail with CM_ERR_CUDA.
 */
#ifndef CHRONOMIND_B200_H
#define CHRONOMIND_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cm_index cm_index; /* opaque */
typedef struct cm_group cm_group; /* opaque: one index sharded over several GPUs of this process */

#define CM_NO_ID 0xFFFFFFFFu

/* Mirrors `enum Error` (src/error.rs:9-60) for the variants the read path can raise. */
typedef enum {
  CM_OK = 0,
  CM_ERR_INVALID_DIMENSIONS = 1, /* Error::InvalidDimensions */
  CM_ERR_INVALID_VECTOR = 2,     /* Error::InvalidVector */
  CM_ERR_CONFIG = 3,             /* Error::Config */
  CM_ERR_IO = 4,                 /* Error::Io */
  CM_ERR_SERIALIZATION = 5,      /* Error::Serialization */
  CM_ERR_INVALID_SNAPSHOT = 6,   /* Error::InvalidSnapshot */
  CM_ERR_INDEX_FULL = 7,         /* Error::IndexFull */
  CM_ERR_INVALID_IMPORTANCE = 8, /* Error::InvalidImportance */
  CM_ERR_CAPACITY_EXCEEDED = 9,  /* Error::CapacityExceeded */
  CM_ERR_INVALID_ARGUMENT = 50,  /* ABI misuse (null pointer, k == 0, k > CM_MAX_K ...) */
  CM_ERR_STATE = 51,             /* call order (search before upload, HNSW search without a graph) */
  CM_ERR_CUDA = 100,
  CM_ERR_NCCL = 101
} cm_status;

/* `IndexParams` (src/config.rs:11-35). */
typedef struct {
  uint64_t max_connections; /* M; layer 0 allows 2M */
  uint64_t ef_construction;
  uint64_t ef_search;
} cm_index_params;

/* `Config` (src/config.rs:43-85); what a snapshot body carries. */
typedef struct {
  uint64_t dimensions;
  uint64_t max_memories;
  float base_decay_rate;
  float temporal_weight;
  float similarity_threshold;
  uint64_t max_relationships;
  cm_index_params index;
} cm_config;

/* Per-call work counters of the last search on this thread (algorithmic bytes). */
typedef struct {
  uint64_t dist_evals;      /* distance_prepared evaluations (HNSW) or nq*N (exact) */
  uint64_t expansions;      /* neighbor lists walked */
  uint64_t list_entries;    /* neighbor ids read */
  uint64_t rescored;        /* fp32 rescore candidates (tensor-core scan); host-buffer calls only */
  uint64_t fallback_queries; /* queries re-run through the fp32 scan because the bf16 pool could not prove exactness
                                (host-buffer calls only: the _dev variants never synchronise to read it back) */
  uint64_t kernel_launches; /* kernels launched by this call */
  uint64_t exact_engine;    /* exact search: 1 = fp32 CUDA-core scan, 2 = tcgen05 scan + fp32 rescore; 0 otherwise */
  uint64_t tie_replays;     /* HNSW: expansions replayed sequentially because bit-identical distances straddled the cut
                               at ef (src/index/lockfree_hnsw.rs:165-189 admits on distance only, strictly) */
  uint64_t failures;        /* queries the engine could NOT answer like the reference (visited-table exhaustion, a tie
                               group larger than the kernel parks, unresolved exact rows). Host-buffer calls turn a
                               non-zero count into CM_ERR_CAPACITY_EXCEEDED; the _dev variants add it to the index's
                               device status word (cm_index_dev_status) */
} cm_counters;

#define CM_MAX_K 1024u  /* largest k / ef the select kernels accept */

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* `Config::default()` (src/config.rs:73-85). */
void cm_config_default(cm_config* out);

/* Replaces `LockFreeHnsw::with_seed` + a `VectorIndex::insert` loop (src/index/lockfree_hnsw.rs:101-113,
 * 375-449; PyO3 `Index.fit`, bindings/python/src/lib.rs:64-72) for the READ side: rows are copied and
 * `CosineDistance::preprocess`ed (src/metric.rs:208-214) on the host; handle i == row i. No graph is built
 * until cm_index_build_graph. Fails with CM_ERR_INDEX_FULL beyond 16,777,216 rows (src/index/arena.rs:47-50). */
cm_status cm_index_from_matrix(const float* x, uint64_t n, uint32_t dim, const cm_config* cfg, cm_index** out);

/* Replaces `load_snapshot` (src/persistence.rs:73-123) up to, not including, index construction: header,
 * CRC-32, bincode body, `Config::validate` (src/config.rs:97-143), `Memory::validate` (src/types.rs:93-121),
 * then replays `ChronoMind::insert` handle assignment in file order (a repeated id tombstones the earlier
 * handle, src/store.rs:252-262). */
cm_status cm_index_from_snapshot(const char* path, cm_index** out);

/* Temporal attributes per handle (`MemoryAttributes.timestamp`, `.decay_rate`, src/types.rs:32-49) and
 * tombstones (`Node.deleted`, src/index/lockfree_hnsw.rs:51). Any pointer may be NULL (left unchanged).
 * Must precede cm_index_upload (or be followed by another upload). */
cm_status cm_index_set_attrs(cm_index* idx, const uint64_t* ts_secs, const uint32_t* ts_nanos,
                             const float* decay_rate, const uint8_t* deleted);

/* Store-level contexts (`MemoryAttributes.context`, src/types.rs:32-49) for cm_search_in_context: `raw_x` [n][dim]
 * are the rows AS STORED (`StoredMemory.data`, not preprocessed — `search_in_context` calls `metric.distance` on
 * them), `context_ids` [n] one id per handle. Snapshot-loaded indexes carry both already (ids in order of first
 * appearance, see cm_index_context_id). `context_ids` may be NULL (one context). Must precede cm_index_upload.
 * With two or more distinct ids the upload also keeps a copy of the bf16 rows ORDERED BY CONTEXT (2 bytes per element
 * of device memory): a batch of context searches then only scans the row ranges of the contexts it asks for. */
cm_status cm_index_set_contexts(cm_index* idx, const float* raw_x, const uint32_t* context_ids);
/* Context id of a label of a snapshot-loaded index; CM_NO_ID when no memory carries it. */
uint32_t cm_index_context_id(const cm_index* idx, const char* label);

/* `VectorIndex::remove` (src/index/lockfree_hnsw.rs:451-467): tombstone one handle. *removed = 1 when this
 * call flipped it. O(1) on the host; the device replica learns of it at cm_index_sync_deleted (or the next upload). */
cm_status cm_index_remove(cm_index* idx, uint32_t handle, int* removed);
/* Push the tombstones recorded since the last upload / sync to the device: a few bytes per removed handle (the
 * `deleted` flag the HNSW final-list filter reads, src/index/lockfree_hnsw.rs:488-492, and a NaN mask on the row's
 * bf16 copy so the tensor-core scans stop emitting it) instead of re-uploading the index. Tombstoned nodes keep
 * routing, as in the reference. Call between batches: searches already enqueued may or may not see the removal. */
cm_status cm_index_sync_deleted(cm_index* idx);

/* Host HNSW construction following `LockFreeHnsw::insert` (src/index/lockfree_hnsw.rs:375-449) with
 * `with_seed(layer_seed)`. threads == 1 reproduces the reference's single-threaded (deterministic) graph;
 * threads > 1 inserts concurrently like the reference's concurrent build (:21-26), so linkage varies run
 * to run. Needed before cm_search_hnsw / cm_search_temporal(exact = 0). */
cm_status cm_index_build_graph(cm_index* idx, uint64_t layer_seed, int threads);

/* GPU-assisted construction. `LockFreeHnsw::insert` links node i to a diverse pick of its nearest
 * ALREADY INSERTED nodes (`search_layer(ef_construction)` over the graph so far, src/index/lockfree_hnsw.rs:419-432).
 * Here those candidates come from an exact PREDECESSOR kNN of each layer's members on CUDA device `device`: member j
 * against members 0..j-1 (tensor-core scan, lower triangle, `n_candidates` nearest per node, 0 = min(ef_construction,
 * 64), at most 127); the host then applies the reference's `select_diverse` (:205-240) and `add_backlink` (:246-280)
 * with `threads` threads. Same layer draw and entry point as cm_index_build_graph; the linkage differs (as it does
 * between any two concurrent builds of the reference, :21-26) and satisfies cm_index_check_invariants. Components the
 * entry point cannot reach are bridged afterwards (see cm_index_unreachable). */
cm_status cm_index_build_graph_gpu(cm_index* idx, uint64_t layer_seed, int device, int threads, uint32_t n_candidates);

/* The host half of the GPU-assisted construction on caller-supplied candidates (any kNN source): `cand_ids[l]` holds
 * `widths[l]` candidate handles (CM_NO_ID padded; the member itself is skipped) for each member of layer l, members in
 * ascending handle order as cm_index_layer_members reports them for the same seed. Candidates that are not members
 * of the layer are ignored. */
cm_status cm_index_build_graph_from_candidates(cm_index* idx, uint64_t layer_seed, int threads, uint32_t n_layers,
                                               const uint32_t* const* cand_ids, const uint32_t* widths);
/* Members of `layer` under the layer draw (`random_layer`, src/index/lockfree_hnsw.rs:126-132, ticket == handle) of
 * this index for `layer_seed`: returns the count, writes at most `cap` handles; *n_layers (optional) = layers drawn. */
uint64_t cm_index_layer_members(const cm_index* idx, uint64_t layer_seed, uint32_t layer, uint32_t* out, uint64_t cap,
                                uint32_t* n_layers);

/* Sharding by vector id (generalises `join_handle`/`split`, src/index/sharded_rwlock.rs:54-66):
 * this index holds the rows whose global handle = local * shard_count + shard_rank. Only affects the ids the
 * merge kernel emits. Default (0, 1). */
cm_status cm_index_set_shard(cm_index* idx, uint32_t shard_rank, uint32_t shard_count);

/* Flatten to the device layout (packed fp32 + bf16 vector matrices, padded CSR adjacency, attribute
 * arrays) on CUDA device `device`. The index is immutable on the device until the next upload. */
cm_status cm_index_upload(cm_index* idx, int device);

void cm_index_free(cm_index* idx);

/* ---- introspection ---------------------------------------------------------------------------- */
uint64_t cm_index_slots(const cm_index* idx);    /* handles ever assigned (tombstones included) */
uint64_t cm_index_len(const cm_index* idx);      /* `VectorIndex::len`: live handles */
uint32_t cm_index_dim(const cm_index* idx);
cm_status cm_index_get_config(const cm_index* idx, cm_config* out);
/* External id of a handle (snapshot-loaded indexes; "" otherwise). Pointer valid until cm_index_free. */
const char* cm_index_external_id(const cm_index* idx, uint32_t handle);
/* `MemoryAttributes.context` / `.importance` of a handle (snapshot-loaded indexes; "" / 0 otherwise) — what the
 * reference CLI prints per result (src/main.rs:213-222). Pointer valid until cm_index_free. */
const char* cm_index_context_label(const cm_index* idx, uint32_t handle);
float cm_index_importance(const cm_index* idx, uint32_t handle);
/* `check_invariants` (src/index/lockfree_hnsw.rs:295-347): 0 when the built graph is sound. */
int cm_index_check_invariants(const cm_index* idx);
/* Layer-0 nodes a search cannot arrive at from the entry point — down through the upper layers, then along layer-0
 * edges, as `LockFreeHnsw::search` moves (src/index/lockfree_hnsw.rs:481-486); 0 when every node is searchable. The reference's
 * incremental insert links each new node into what the entry already reaches; the candidate-based builders bridge
 * unreached components after linking and this reports what is left. */
uint64_t cm_index_unreachable(const cm_index* idx);
/* Graph readout for tests / sidecar files: entry word ((top_layer << 32) | id, ~0 when empty),
 * per-node top layer, and one neighbor list. */
uint64_t cm_index_entry(const cm_index* idx);
uint32_t cm_index_top_layer(const cm_index* idx, uint32_t handle);
uint32_t cm_index_neighbors(const cm_index* idx, uint32_t handle, uint32_t layer, uint32_t* out, uint32_t cap);
/* Copy the preprocessed row of a handle (dim floats). */
cm_status cm_index_vector(const cm_index* idx, uint32_t handle, float* out);

/* Sidecar graph file next to a snapshot (the reference lists "Graph serialization" as
 * roadmap, README.md:318). Save after cm_index_build_graph; load instead of building. */
cm_status cm_index_save_graph(const cm_index* idx, const char* path);
cm_status cm_index_load_graph(cm_index* idx, const char* path);

/* ---- search: host buffers (copies inside the call) ----------------------------------------------- */

/* Exact brute-force kNN — the differential-oracle arithmetic of tests/property_test.rs:201-206:
 * distance_prepared(preprocess(v), preprocess(q)), ascending (distance, handle), tombstones dropped.
 * `ids`/`dist` are [nq][k]. qdim != dim follows the index-level convention (every distance 2.0,
 * src/metric.rs:220-222). */
cm_status cm_search_exact(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                          uint32_t* ids, float* dist);

/* `VectorIndex::search(query, ef)` then `take(k)` for each row — PyO3 `Index.batch_query`
 * (bindings/python/src/lib.rs:92-117) over `LockFreeHnsw::search` (src/index/lockfree_hnsw.rs:469-495). */
cm_status cm_search_hnsw(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t ef,
                         uint32_t k, uint32_t* ids, float* dist);

/* `ChronoMind::search(query, k)` for each row (src/store.rs:321-346): validate (`:379-392`, the whole call
 * fails with the first bad row's status), ef = max(ef_search, 3k), candidate pool from the HNSW search
 * (exact == 0) or from the exact scan (exact != 0), `combined_score` (`:395-415`) with the caller's `now`
 * (the reference reads the wall clock, `:324`), stable sort by score, truncate k.
 * `score`, `dist` are [nq][k]; `dist` may be NULL. */
cm_status cm_search_temporal(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                             uint32_t ef_search, float temporal_weight, float base_decay_rate,
                             uint64_t now_secs, uint32_t now_nanos, int exact, uint32_t* ids, float* score,
                             float* dist);

/* `ChronoMind::search_in_context(context, query, k)` for each row (src/store.rs:353-377): validate, then an EXACT
 * scan of the live members of the row's context with `metric.distance` on the raw stored rows (full cosine with
 * both norms, src/metric.rs:163-195), `combined_score`, sort by score, truncate k. `context_ids` [nq] (one per
 * query; CM_NO_ID matches nothing). The reference iterates a hash map, so its order among EQUAL scores is
 * unspecified; here ties resolve toward the lower handle. */
cm_status cm_search_in_context(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim,
                               const uint32_t* context_ids, uint32_t k, float temporal_weight, float base_decay_rate,
                               uint64_t now_secs, uint32_t now_nanos, uint32_t* ids, float* score);

/* ---- search: device buffers on the index's device, asynchronous on `stream` ------------------------ */
/* Same contracts; q/ids/dist/score are device pointers; qdim must equal dim; no validation of query
 * values (the caller owns them). Work is enqueued on `stream` (a cudaStream_t, NULL = default stream). */
cm_status cm_search_exact_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t k, void* ids, void* dist,
                              void* stream);
cm_status cm_search_hnsw_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t ef, uint32_t k, void* ids,
                             void* dist, void* stream);
cm_status cm_search_temporal_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t k, uint32_t ef_search,
                                 float temporal_weight, float base_decay_rate, uint64_t now_secs,
                                 uint32_t now_nanos, int exact, void* ids, void* score, void* dist, void* stream);
cm_status cm_search_in_context_dev(const cm_index* idx, const void* q, uint32_t nq, const void* context_ids, uint32_t k,
                                   float temporal_weight, float base_decay_rate, uint64_t now_secs, uint32_t now_nanos,
                                   void* ids, void* score, void* stream);

/* The asynchronous variants cannot report a query they could not answer like the reference through their return
 * value (HNSW: visited-table exhaustion or a tie group larger than the kernel parks; exact: more queries flagged for
 * the fp32 fallback than the in-stream path absorbs — those rows are left CM_NO_ID / NaN, never a plausible wrong
 * answer). They count such queries in a status word on the index's device instead. This call waits for the work
 * enqueued so far on that device, returns the count in *failures and optionally clears it. The host-buffer variants
 * need no such call: they redo the batch or return CM_ERR_CAPACITY_EXCEEDED. */
cm_status cm_index_dev_status(const cm_index* idx, uint64_t* failures, int reset);
/* Work counters of the asynchronous exact searches since the last reset (what cm_last_counters reports for a
 * host-buffer call): rows re-scored by the fp32 rescore and queries routed to the fp32 fallback. Waits like the above. */
cm_status cm_index_dev_counters(const cm_index* idx, uint64_t* rescored, uint64_t* fallback_queries, int reset);

/* Merge step of `ShardedRwLockHnsw::search` (src/index/sharded_rwlock.rs:81-94): concatenate per-shard lists,
 * sort by (distance, global handle), truncate. ids/dist: device [n_shards][nq][len] LOCAL handles as produced
 * by each shard's search (the layout an allgather leaves); global = local * n_shards + shard.
 * out_ids/out_dist: device [nq][out_len]. `device` selects the GPU. */
cm_status cm_merge_topk_dev(const void* ids, const void* dist, uint32_t n_shards, uint32_t nq, uint32_t len,
                            uint32_t out_len, void* out_ids, void* out_dist, int device, void* stream);

/* The same merge over ONE exchanged array per step: cm_pack_topk_dev turns `count` (distance, local handle) pairs into
 * sortable 64-bit keys ((TotalF32 distance << 32) | handle; all-ones for CM_NO_ID padding) and
 * cm_merge_topk_keys_dev consumes the all-gathered keys [n_shards][nq][len]. One collective instead of two. */
cm_status cm_pack_topk_dev(const void* ids, const void* dist, uint64_t count, void* out_keys, int device, void* stream);
cm_status cm_merge_topk_keys_dev(const void* keys, uint32_t n_shards, uint32_t nq, uint32_t len, uint32_t out_len,
                                 void* out_ids, void* out_dist, int device, void* stream);

/* The rerank tail of `ChronoMind::search` (src/store.rs:327-344) on an already merged pool of GLOBAL handles
 * (multi-GPU: merge first, then rerank). Attribute arrays are device pointers indexed by global
 * handle. pool_ids/pool_dist: [nq][pool]; out: [nq][k]. */
cm_status cm_rerank_dev(const void* pool_ids, const void* pool_dist, uint32_t nq, uint32_t pool, uint32_t k,
                        const void* ts_secs, const void* ts_nanos, const void* decay_rate, float temporal_weight,
                        float base_decay_rate, uint64_t now_secs, uint32_t now_nanos, void* out_ids,
                        void* out_score, void* out_dist, int device, void* stream);

/* ---- sharded search: ONE object over G GPUs of this process ------------------------------------------------------
 * Replaces `ShardedRwLockHnsw` (src/index/sharded_rwlock.rs:43-94): rows are dealt out round-robin (shard = id mod G,
 * local = id div G, :54-74), every shard is an independent index (its own HNSW sub-graph) resident on its own GPU,
 * and a search is ONE call: each device copies 1/G of the batch over its own PCIe link and the slices are
 * all-gathered over NVLink, every shard searches the whole batch, the per-shard lists are all-gathered (NCCL) and
 * merged on the device by (distance, global handle) (:81-94); the store-level search reranks the MERGED pool
 * (src/store.rs:323-344). Results carry GLOBAL handles. NCCL is loaded at cm_group_create (n_devices > 1); a missing or
 * failing NCCL is CM_ERR_NCCL. One batch at a time per group (calls on one group serialise). */
cm_status cm_group_create(const int* devices, int n_devices, cm_group** out);
void cm_group_free(cm_group* g);
/* `insert` loop of the sharded index: row i becomes global handle i (shard i mod G). Replaces any previous contents. */
cm_status cm_group_from_matrix(cm_group* g, const float* x, uint64_t n, uint32_t dim, const cm_config* cfg);
/* Attribute arrays indexed by GLOBAL handle (any pointer may be NULL); before cm_group_upload. */
cm_status cm_group_set_attrs(cm_group* g, const uint64_t* ts_secs, const uint32_t* ts_nanos, const float* decay_rate,
                             const uint8_t* deleted);
/* One HNSW sub-graph per shard (seed + shard), all shards concurrently; gpu_assisted != 0 uses each shard's own GPU
 * (cm_index_build_graph_gpu), else the host replay of `LockFreeHnsw::insert` with threads / G threads per shard. */
cm_status cm_group_build_graphs(cm_group* g, uint64_t layer_seed, int threads, int gpu_assisted);
cm_status cm_group_upload(cm_group* g);
uint32_t cm_group_size(const cm_group* g);
uint64_t cm_group_slots(const cm_group* g);
uint64_t cm_group_len(const cm_group* g);
cm_index* cm_group_shard(const cm_group* g, uint32_t shard); /* borrowed: introspection, sidecars */
/* Same contracts as cm_search_exact / cm_search_hnsw / cm_search_temporal, host buffers, GLOBAL handles. */
cm_status cm_group_search_exact(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t k, uint32_t* ids,
                                float* dist);
cm_status cm_group_search_hnsw(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t ef, uint32_t k,
                               uint32_t* ids, float* dist);
cm_status cm_group_search_temporal(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                                   uint32_t ef_search, float temporal_weight, float base_decay_rate, uint64_t now_secs,
                                   uint32_t now_nanos, int exact, uint32_t* ids, float* score, float* dist);

/* ---- maintenance passes over the same device-resident arrays ------------------------------- */
/* The all-pairs test of `ChronoMind::consolidate` (src/store.rs:516-568: `metric.similarity(a.data, b.data) >
 * similarity_threshold` for every pair of live memories) as a threshold self-join on the tensor cores: candidate pairs
 * from the bf16 scan of the corpus against itself, decided by the exact similarity on the rows as stored (needs
 * cm_index_set_contexts / a snapshot). Pairs i < j ascending (i, j); *count = pairs found, at most `cap` written (any
 * out pointer may be NULL). The greedy keeper/absorbed bookkeeping stays with the store (write path). */
cm_status cm_similar_pairs(const cm_index* idx, float threshold, uint32_t* out_i, uint32_t* out_j, float* out_sim,
                           uint64_t cap, uint64_t* count);
/* One `apply_decay` sweep (src/store.rs:428-458) over caller-owned device arrays: importance f32[n] (in/out),
 * decayed_through_nanos u64[n] (in/out), last_access_nanos u64[n], decay_rate f32[n]. */
cm_status cm_decay_sweep_dev(void* importance, void* decayed_through_nanos, const void* last_access_nanos,
                             const void* decay_rate, uint64_t n, float base_decay_rate, uint64_t now_nanos, int device,
                             void* stream);

/* ---- metric surface (`DistanceMetric`, src/metric.rs:13-48) on device rows ---------------------------- */
/* `preprocess` for each row of a device matrix [n][dim] -> out [n][dim] (may alias). */
cm_status cm_metric_preprocess_dev(const void* x, uint64_t n, uint32_t dim, void* out, int device, void* stream);
/* `distance_prepared(a[i], b[i])` for n row pairs of prepared device matrices -> out[n]. */
cm_status cm_metric_distance_prepared_dev(const void* a, const void* b, uint64_t n, uint32_t dim, void* out,
                                          int device, void* stream);
/* "cosine" */
const char* cm_metric_name(void);

/* ---- tuning / diagnostics ------------------------------------------------------------------------------ */
/* Exact-scan engine: 0 = auto (tensor-core scan for corpora of 8192 rows or more), 1 = force the fp32
 * CUDA-core scan, 2 = force the tcgen05 scan + fp32 rescore. */
cm_status cm_set_exact_engine(cm_index* idx, int engine);
cm_status cm_last_counters(cm_counters* out); /* thread-local */
/* Roofline instrumentation: when enabled, the search calls bracket their dominant kernel (the scan for the exact
 * path, the traversal for the HNSW path) with CUDA events on the launching stream. cm_profile_read waits for
 * the recorded events, returns their summed device time and count since the last read, and clears them. */
cm_status cm_profile_enable(int on);
cm_status cm_profile_read(float* kernel_ms, uint32_t* kernel_launches);
uint64_t cm_launch_count(void);               /* process-wide count of kernels launched by this library */
const char* cm_last_error(void);              /* thread-local message of the last failing call */
const char* cm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CHRONOMIND_B200_H */
