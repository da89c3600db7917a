// api.cu — the C ABI (include/chronomind_b200.h): index lifecycle on the host, flattening to the device
// layout, and the search entry points that enqueue the kernels. No CPU search path exists here.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string_view>
#include <thread>
#include <unordered_map>

#include "cm_internal.h"
#include "kernels.h"
#include "scan_tc.h"

namespace cm {

// ---- errors / counters -----------------------------------------------------------------------------------
static thread_local std::string tl_error;
static thread_local cm_counters tl_counters;
static thread_local bool tl_used_tc = false;  // the last exact search on this thread ran the tensor-core scan
static thread_local uint32_t tl_scan_worst_chunk = 0;  // most queries one chunk of that search sent to the fallback
static thread_local bool tl_scan_bounded = false;       // ... and that fallback was the bounded (kFlagCap) one
static thread_local bool tl_force_fp32 = false;        // redo of a call whose tensor-core pool could not be completed
static std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { tl_error = msg; }
void count_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  tl_counters.kernel_launches++;
}

static cm_status fail(cm_status st, const std::string& msg) {
  set_error(msg);
  return st;
}
static cm_status cuda_fail(cudaError_t e, const char* what) {
  return fail(CM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CM_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---- roofline instrumentation (cm_profile_*) ---------------------------------------------------------------
static std::atomic<int> g_profile{0};
static std::mutex g_profile_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events;
struct ProfileScope {  // brackets one dominant-kernel launch with events on its stream
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  explicit ProfileScope(cudaStream_t stream, bool enabled = true) : s(stream) {
    if (!enabled || !g_profile.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
      a = b = nullptr;
      return;
    }
    cudaEventRecord(a, s);
  }
  ~ProfileScope() {
    if (!a) return;
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> g(g_profile_mu);
    g_profile_events.emplace_back(a, b);
  }
};

// ---- device buffers ----------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Scratch for one search call. A workspace is leased for the duration of the HOST call; the kernels a `_dev`
// variant enqueued may still be running when it returns, so the lease records a completion event on the stream it
// used. The next user of the workspace either runs on that same stream (ordered behind the earlier work anyway) or
// makes its stream wait on the event first — two calls never touch the same scratch concurrently on the device,
// whatever mix of host threads and streams the caller uses.
struct Workspace {
  DevBuf qraw, qn, D, pool_ids, pool_dist, out_ids, out_a, out_b, small, gtable, qnorm, qctx;
  DevBuf qbf16, tc_state, cand, qflag;  // tensor-core scan: bf16 queries, tau/count arrays, candidates, fallback queries
  DevBuf bias;                          // search_in_context: per-row temporal bias of the call
  DevBuf ctxsort;                       // search_in_context: query order / grouped context ids / per-pair tile ranges
  DevBuf vmap, vmap_epoch;              // HNSW at large ef: visited byte maps in HBM (one slab per CTA) and their epochs
  cudaStream_t stream = nullptr;  // host-buffer variants
  // Exact scan: everything after K1 (rescore, fallback, counters) runs on a LOW-priority side stream, forked from and
  // joined back into the call's stream with events. For one caller stream nothing changes; a caller that keeps two
  // batches in flight on two streams gets the tail of batch i scheduled UNDER the scan of batch i + 1 (the scan's blocks
  // win the SMs first, a rescore CTA fits in the shared memory the scan leaves free). The default stream priority is
  // already the lowest one, so it is the CALLER's streams that must sit a notch higher (bench.py creates them so; the
  // host-buffer variants' own streams are created so).
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t ev_scan = nullptr, ev_tail = nullptr;
  cudaEvent_t done = nullptr;     // completion of the last call that used this workspace
  cudaStream_t last_stream = nullptr;
  bool pending = false;           // `done` was recorded and may not have completed yet
  unsigned hnsw_grid = 0;
  void release() {
    for (DevBuf* b : {&qraw, &qn, &D, &pool_ids, &pool_dist, &out_ids, &out_a, &out_b, &small, &gtable, &qnorm, &qctx, &qbf16,
                      &tc_state, &cand, &qflag, &bias, &ctxsort, &vmap, &vmap_epoch})
      b->release();
    if (stream) cudaStreamDestroy(stream);
    if (tail_stream) cudaStreamDestroy(tail_stream);
    for (cudaEvent_t* e : {&done, &ev_scan, &ev_tail}) {
      if (*e) cudaEventDestroy(*e);
      *e = nullptr;
    }
    stream = nullptr;
    tail_stream = nullptr;
  }
};

struct WorkspacePool {
  std::mutex mu;
  std::vector<Workspace*> free_list;
  std::vector<Workspace*> all;
  // Prefer a workspace whose last call ran on `want` (or has nothing in flight): no cross-stream wait needed.
  Workspace* acquire(cudaStream_t want, bool want_known) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = free_list.size(); i-- > 0;) {
      Workspace* w = free_list[i];
      if (!w->pending || (want_known && w->last_stream == want)) {
        free_list.erase(free_list.begin() + (long)i);
        return w;
      }
    }
    for (size_t i = free_list.size(); i-- > 0;) {  // its last call has finished by now
      Workspace* w = free_list[i];
      if (w->done && cudaEventQuery(w->done) == cudaSuccess) {
        w->pending = false;
        free_list.erase(free_list.begin() + (long)i);
        return w;
      }
    }
    if (!free_list.empty() && all.size() >= 4) {  // bounded pool: reuse a busy one, the lease makes the stream wait
      Workspace* w = free_list.back();
      free_list.pop_back();
      return w;
    }
    Workspace* w = new Workspace();
    all.push_back(w);
    return w;
  }
  void give_back(Workspace* w) {
    std::lock_guard<std::mutex> g(mu);
    free_list.push_back(w);
  }
  // Is the caller keeping more than one batch in flight? True when another workspace is leased right now (a second
  // host thread inside a call) or still has device work pending (an asynchronous call on another stream).
  bool overlapped(const Workspace* mine) {
    std::lock_guard<std::mutex> g(mu);
    for (Workspace* w : all) {
      if (w == mine) continue;
      const bool is_free = std::find(free_list.begin(), free_list.end(), w) != free_list.end();
      if (!is_free) return true;
      if (w->pending && w->done && cudaEventQuery(w->done) == cudaErrorNotReady) return true;
    }
    return false;
  }
  ~WorkspacePool() {
    for (Workspace* w : all) {
      w->release();
      delete w;
    }
  }
};

// Lease for a call that enqueues on the caller's `stream` (`_dev` variants) or on the workspace's own stream
// (host-buffer variants: host = true; they synchronise before returning, so nothing stays pending).
struct WorkspaceLease {
  WorkspacePool* pool;
  Workspace* ws;
  cudaStream_t s = nullptr;
  bool host;
  cudaError_t err = cudaSuccess;
  WorkspaceLease(const cm_index* idx, cudaStream_t stream, bool host_call)
      : pool((WorkspacePool*)idx->workspace_pool), ws(pool->acquire(stream, !host_call)), host(host_call) {
    if (host) {
      if (!ws->stream) {  // one notch above the default priority: the low-priority tail streams (default priority is the
        int lo = 0, hi = 0;  // lowest there is) then yield to the scans of other in-flight calls
        err = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (err == cudaSuccess) err = cudaStreamCreateWithPriority(&ws->stream, cudaStreamNonBlocking, hi < lo ? lo - 1 : lo);
      }
      s = ws->stream;
    } else {
      s = stream;
    }
    if (err == cudaSuccess && ws->pending && ws->last_stream != s) err = cudaStreamWaitEvent(s, ws->done, 0);
  }
  ~WorkspaceLease() {
    if (host) {
      ws->pending = false;  // the call synchronised its stream (or failed before enqueuing anything that matters)
      if (err != cudaSuccess || cudaStreamQuery(s) != cudaSuccess) cudaStreamSynchronize(s);
    } else {
      if (!ws->done && cudaEventCreateWithFlags(&ws->done, cudaEventDisableTiming) != cudaSuccess) ws->done = nullptr;
      if (ws->done && cudaEventRecord(ws->done, s) == cudaSuccess) {
        ws->pending = true;
        ws->last_stream = s;
      } else {
        cudaStreamSynchronize(s);  // no event: fall back to finishing the work before anyone reuses the scratch
        ws->pending = false;
      }
    }
    pool->give_back(ws);
  }
};

constexpr uint32_t kGHashBits = 17;             // per-CTA overflow visited table: 128 Ki entries
constexpr size_t kDistWorkspaceBytes = 1ull << 30;  // fp32 scan: distance block [QB][n] floats
constexpr uint32_t kFlagCap = 512;             // queries per call the in-stream fp32 fallback can absorb
constexpr uint32_t kFlagChunk = 64;            // ... in fp32 scans of this many queries (empty chunks exit at once)
// ws->small layout (bytes): 0 HNSW counters (8 x u64) | per query chunk: 64 rescored u64, 72 flag_count u32 |
// per call: 80 rescored total u64, 88 flagged total u32, 92 most flagged in one chunk u32 |
// 128 work_counter / pair_count u32 | 192 bad_row u32 | 256 flag_list (kFlagCap x u32)
constexpr uint32_t kTcQueryChunk = 32768;      // queries per tensor-core scan (bounds the candidate buffer: 1 GB at k <= 16)
constexpr size_t kSmallBytes = 256 + kFlagCap * 4;

static void free_device(DeviceIndex& d) {
  if (d.device < 0) return;
  cudaSetDevice(d.device);
  for (void* p : {(void*)d.x32, d.xbf16, (void*)d.deleted, (void*)d.xraw, (void*)d.norm2_raw, (void*)d.ctx, d.xbf16_ctx,
                  (void*)d.ctx_sorted, (void*)d.perm, (void*)d.inv_perm,
                  (void*)d.ts_secs, (void*)d.ts_nanos, (void*)d.decay,
                  (void*)d.nbr0, (void*)d.deg0, (void*)d.upper_base, (void*)d.nbr_up, (void*)d.deg_up, (void*)d.status})
    if (p) cudaFree(p);
  d = DeviceIndex();
}

template <class F>
static void parallel_rows(uint64_t n, F f) {
  unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 4096 || hw == 1) {
    for (uint64_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> ts;
  uint64_t chunk = (n + hw - 1) / hw;
  for (uint64_t s = 0; s < n; s += chunk) {
    uint64_t e = std::min(n, s + chunk);
    ts.emplace_back([=] {
      for (uint64_t i = s; i < e; ++i) f(i);
    });
  }
  for (auto& t : ts) t.join();
}

static bool row_is_finite(const float* v, uint32_t dim) {
  float acc = 0.0f;
  for (uint32_t e = 0; e < dim; ++e) acc += v[e] * 0.0f;  // NaN/Inf poison the sum; vectorises
  return acc == 0.0f;
}

// Entries of the shared-memory visited table: a search visits ~17-19 nodes per unit of ef on the bench corpus
// (dist_evals / ef), the table restarts on the global one past 3/4 full. Not rounded to a power of two: this table
// is the largest shared-memory consumer at large ef and decides how many queries an SM holds.
static uint32_t visited_entries_for(uint32_t ef) {
  return std::min<uint32_t>(std::max<uint32_t>(2048u, 30u * ef), 32768u);
}

// ---- search plumbing (device pointers, async on `s`) --------------------------------------------------------
struct PreparedQueries {
  const float* qn = nullptr;
  uint32_t ld = 0;
  bool degenerate = false;
  bool bf16_ready = false;  // ws->qbf16 already holds the scan-layout bf16 copy of ALL nq rows (fused K0)
};

// Would an exact search of k neighbors on this index take the tensor-core scan? (one predicate for the fused query
// preparation and for exact_topk)
static bool exact_uses_tc(const cm_index* idx, bool degenerate, uint32_t k);

static cm_status prepare_queries(const cm_index* idx, Workspace* ws, const float* q_dev, uint32_t nq, uint32_t qdim,
                                 uint32_t* bad_row, cudaStream_t s, PreparedQueries* out, bool want_bf16 = false) {
  const DeviceIndex& d = idx->dev;
  out->degenerate = qdim != d.dim;
  out->ld = d.dim_pad;
  if (out->degenerate) {  // distances are all 2.0; the query values are never read
    out->qn = nullptr;
    return CM_OK;
  }
  CM_CUDA(ws->qn.ensure((size_t)std::max(nq, 1u) * d.dim_pad * sizeof(float)));
  out->qn = ws->qn.as<float>();
  if (want_bf16 && nq && prepare_queries_fused_supported(q_dev, d.dim)) {
    // one pass: normalise + bf16 copy in the scan layout (every 32,768-query chunk is a whole number of tile pairs)
    const uint32_t nq_pad = scan_tc_pad_queries(nq);
    CM_CUDA(ws->qbf16.ensure((size_t)nq_pad * d.k_pad * 2));
    CM_CUDA(launch_prepare_queries(q_dev, nq, d.dim, ws->qn.as<float>(), d.dim_pad, ws->qbf16.p, nq_pad, d.k_pad, bad_row, s));
    out->bf16_ready = true;
    return CM_OK;
  }
  CM_CUDA(launch_preprocess_rows(q_dev, nq, d.dim, d.dim, ws->qn.as<float>(), d.dim_pad, bad_row, s));
  return CM_OK;
}

// fp32 CUDA-core scan in query blocks: D[QB][n] then select. qmap/active_count route the fallback of the
// tensor-core scan (only the first *active_count rows of q are live; results land in row qmap[i]).
static cm_status exact_topk_fp32(const cm_index* idx, Workspace* ws, const float* qn, uint32_t ldq, bool degenerate,
                                 uint32_t nq, uint32_t k, uint32_t* ids, float* dist, const uint32_t* qmap,
                                 const uint32_t* active_count, cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  uint64_t ldD = (d.n + 3) & ~3ull;
  uint64_t qb = std::max<uint64_t>(16, (kDistWorkspaceBytes / (ldD * 4)) & ~15ull);
  qb = std::min<uint64_t>(qb, ((uint64_t)nq + 15) & ~15ull);
  if (qmap) qb = std::min<uint64_t>(qb, kFlagChunk);  // fallback: small chunks, the ones past *active_count exit at once
  CM_CUDA(ws->D.ensure(qb * ldD * sizeof(float)));
  float* D = ws->D.as<float>();
  for (uint32_t q0 = 0; q0 < nq; q0 += (uint32_t)qb) {
    uint32_t nb = std::min<uint32_t>((uint32_t)qb, nq - q0);
    if (degenerate) {
      CM_CUDA(launch_fill(D, (uint64_t)nb * ldD, 2.0f, s));
    } else {
      ProfileScope prof(qmap ? nullptr : s, qmap == nullptr);
      CM_CUDA(launch_exact_dist(qn + (size_t)q0 * ldq, nb, ldq, d.x32, d.n, d.dim, d.dim_pad, D, ldD, active_count, q0, s));
    }
    CM_CUDA(launch_select_from_dist(D, ldD, nb, d.n, d.deleted, k, qmap ? ids : ids + (size_t)q0 * k,
                                    qmap ? dist : dist + (size_t)q0 * k, k, qmap ? qmap + q0 : nullptr, active_count, q0, s));
  }
  return CM_OK;
}

// tcgen05 bf16 scan + exact fp32 rescore (+ in-stream fp32 fallback for flagged queries). See scan_tc.cu.
static cm_status exact_topk_tc(const cm_index* idx, Workspace* ws, const PreparedQueries& pq, uint32_t q_off, uint32_t nq,
                               uint32_t k, uint32_t* ids, float* dist, cudaStream_t s, bool approx = false) {
  const DeviceIndex& d = idx->dev;
  const uint32_t nq_pad = scan_tc_pad_queries(nq);
  tl_used_tc = true;
  if (!pq.bf16_ready) CM_CUDA(ws->qbf16.ensure((size_t)nq_pad * d.k_pad * 2));
  CM_CUDA(ws->tc_state.ensure((size_t)nq_pad * 8));
  const uint32_t cap = scan_tc_cand_cap(k);
  CM_CUDA(ws->cand.ensure((size_t)nq_pad * cap * 8));
  CM_CUDA(ws->small.ensure(kSmallBytes));
  // Flagged queries (candidate overflow / too few survivors) are re-run through the exact fp32 scan in the same
  // stream. Normal shapes: ONE cooperative launch that handles any number of them (returns at once when there are
  // none). Rows too long for its 16-query tile: the chunked fp32 engine, bounded by kFlagCap per call.
  const bool one_launch = (size_t)(d.dim & ~7u) * 16 * sizeof(float) <= 200 * 1024;
  tl_scan_bounded = !one_launch;
  const size_t qc_bytes = (fallback_exact_query_bytes(d.dim_pad) + 255) & ~(size_t)255;
  if (one_launch)
    CM_CUDA(ws->qflag.ensure(qc_bytes + (size_t)nq_pad * 4));
  else
    CM_CUDA(ws->qflag.ensure((size_t)kFlagCap * d.dim_pad * 4));
  uint32_t* tau_key = ws->tc_state.as<uint32_t>();
  uint32_t* cand_cnt = tau_key + nq_pad;
  uint8_t* small = ws->small.as<uint8_t>();
  // CM_SCAN_WARM=1 (measurement only: valid only when the SAME batch is repeated): keep the previous launch's thresholds,
  // i.e. time the scan without its cold start
  static const bool warm_diag = getenv("CM_SCAN_WARM") != nullptr;
  if (warm_diag)
    CM_CUDA(cudaMemsetAsync(cand_cnt, 0, (size_t)nq_pad * 4, s));
  else
    CM_CUDA(cudaMemsetAsync(ws->tc_state.p, 0, (size_t)nq_pad * 8, s));
  if (q_off) CM_CUDA(cudaMemsetAsync(small + 64, 0, 16, s));  // this chunk's counters (the first chunk's: zeroed with the totals)
  const void* q_bf16 = ws->qbf16.p;
  if (pq.bf16_ready)
    q_bf16 = static_cast<const uint8_t*>(ws->qbf16.p) + (size_t)q_off * d.k_pad * 2;  // this chunk of the fused copy
  else
    CM_CUDA(launch_to_bf16(pq.qn, nq, d.dim, pq.ld, ws->qbf16.p, nq_pad, d.k_pad, nullptr, false, s));
  ScanTcLaunch sl;
  // A caller with a second batch in flight gets a scan that leaves room on every SM for one rescore CTA (the tail of
  // the other batch then runs UNDER this scan); a single batch in flight gets all of the shared memory for resident
  // query blocks.
  sl.leave_room = ((WorkspacePool*)idx->workspace_pool)->overlapped(ws);
  sl.q_bf16 = q_bf16;
  sl.nq = nq;
  sl.k = k;
  sl.tau_key = tau_key;
  sl.cand_cnt = cand_cnt;
  sl.cand = ws->cand.as<unsigned long long>();
  sl.cap = cap;
  {
    ProfileScope prof(s);
    CM_CUDA(launch_scan_tc(d, sl, s));
  }
  // fork: with a second batch in flight the tail of this one goes to the workspace's low-priority stream (it then runs
  // under the other batch's scan); a lone batch keeps its tail on the caller's stream — two event hops less per step
  cudaStream_t caller = s;
  const bool fork_tail = sl.leave_room;
  if (fork_tail && !ws->tail_stream) {
    int lo = 0, hi = 0;
    CM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = lowest priority (numerically greatest)
    CM_CUDA(cudaStreamCreateWithPriority(&ws->tail_stream, cudaStreamNonBlocking, lo));
    CM_CUDA(cudaEventCreateWithFlags(&ws->ev_scan, cudaEventDisableTiming));
    CM_CUDA(cudaEventCreateWithFlags(&ws->ev_tail, cudaEventDisableTiming));
  }
  if (fork_tail) {
    CM_CUDA(cudaEventRecord(ws->ev_scan, caller));
    s = ws->tail_stream;
    CM_CUDA(cudaStreamWaitEvent(s, ws->ev_scan, 0));
  }
  RescoreLaunch rl;
  rl.qn = pq.qn;
  rl.ldq = pq.ld;
  rl.nq = nq;
  rl.k = k;
  rl.tau_key = tau_key;
  rl.cand_cnt = cand_cnt;
  rl.cand = sl.cand;
  rl.cap = cap;
  rl.ids = ids;
  rl.dist = dist;
  rl.rescored_total = reinterpret_cast<unsigned long long*>(small + 64);
  rl.flag_count = reinterpret_cast<uint32_t*>(small + 72);
  rl.flag_list = one_launch ? reinterpret_cast<uint32_t*>(ws->qflag.as<uint8_t>() + qc_bytes) : reinterpret_cast<uint32_t*>(small + 256);
  rl.flag_cap = one_launch ? nq : kFlagCap;
  rl.approx = approx;
  CM_CUDA(launch_rescore(d, rl, s));
  if (one_launch) {
    CM_CUDA(ws->D.ensure(fallback_exact_dist_bytes(d.n)));
    CM_CUDA(launch_fallback_exact(pq.qn, pq.ld, rl.flag_list, rl.flag_count, ws->qflag.as<float>(), d.x32, d.n, d.dim, d.dim_pad,
                                  ws->D.as<float>(), d.deleted, k, ids, dist, d.sm_count, s));
  } else {
    CM_CUDA(launch_gather_rows(pq.qn, pq.ld, rl.flag_list, rl.flag_count, kFlagCap, ws->qflag.as<float>(), s));
    cm_status st = exact_topk_fp32(idx, ws, ws->qflag.as<float>(), pq.ld, false, kFlagCap, k, ids, dist, rl.flag_list,
                                   rl.flag_count, s);
    if (st != CM_OK) return st;
  }
  // chunk counters -> call totals; queries beyond what the fallback absorbed are counted in the status word
  CM_CUDA(launch_fold_scan_counters(small + 64, small + 80, one_launch ? 0xFFFFFFFFu : kFlagCap, d.status + 1, s));
  // join: whatever the caller enqueues next on its stream sees this batch's results
  if (fork_tail) {
    CM_CUDA(cudaEventRecord(ws->ev_tail, s));
    CM_CUDA(cudaStreamWaitEvent(caller, ws->ev_tail, 0));
  }
  return CM_OK;
}

static bool exact_uses_tc(const cm_index* idx, bool degenerate, uint32_t k) {
  const DeviceIndex& d = idx->dev;
  // Rows with NaN/Inf components order by f32::total_cmp in the reference; the threshold select cannot
  // reproduce that, so such an index always takes the fp32 scan.
  const bool tc_ok = !degenerate && idx->all_finite && d.n > 0 && scan_tc_supported(d, k);
  if (tl_force_fp32) return false;
  // auto: the tensor-core scan wins from one query up once the corpus is a few thousand rows (1M rows, 1 query:
  // ~0.7 ms against ~3 ms for the fp32 scan + select); tiny corpora stay on the fp32 engine.
  return idx->exact_engine == 2 ? tc_ok : (idx->exact_engine == 0 && tc_ok && d.n >= 8192);
}

// Exact top-`k` by (distance, handle) into ids/dist [nq][k] (device).
static cm_status exact_topk(const cm_index* idx, Workspace* ws, const PreparedQueries& pq, uint32_t nq, uint32_t k,
                            uint32_t* ids, float* dist, cudaStream_t s, bool approx = false) {
  const DeviceIndex& d = idx->dev;
  if (nq == 0) return CM_OK;
  if (d.n == 0) {
    CM_CUDA(cudaMemsetAsync(ids, 0xFF, (size_t)nq * k * 4, s));
    CM_CUDA(launch_fill(dist, (uint64_t)nq * k, INFINITY, s));
    return CM_OK;
  }
  tl_counters.dist_evals += (uint64_t)nq * d.n;
  bool tc_ok = !pq.degenerate && idx->all_finite && scan_tc_supported(d, k);
  const bool use_tc = exact_uses_tc(idx, pq.degenerate, k);
  if (!tl_force_fp32 && idx->exact_engine == 2 && !tc_ok && !pq.degenerate)
    return fail(CM_ERR_INVALID_ARGUMENT, "tensor-core scan needs k <= 128 and a corpus without NaN/Inf components");
  tl_counters.exact_engine = use_tc ? 2 : 1;
  if (use_tc) {
    CM_CUDA(ws->small.ensure(kSmallBytes));
    CM_CUDA(cudaMemsetAsync(ws->small.as<uint8_t>() + 64, 0, 32, s));  // the first chunk's counters and the call totals
    for (uint32_t q0 = 0; q0 < nq; q0 += kTcQueryChunk) {  // bounded candidate buffer for very large batches
      const uint32_t nb = std::min(kTcQueryChunk, nq - q0);
      PreparedQueries part = pq;
      part.qn = pq.qn + (size_t)q0 * pq.ld;
      cm_status st = exact_topk_tc(idx, ws, part, q0, nb, k, ids + (size_t)q0 * k, dist + (size_t)q0 * k, s, approx);
      if (st != CM_OK) return st;
    }
    return CM_OK;
  }
  return exact_topk_fp32(idx, ws, pq.qn, pq.ld, pq.degenerate, nq, k, ids, dist, nullptr, nullptr, s);
}

// HNSW candidate pool [nq][ef] (tombstones filtered, ascending) into ws->pool_*.
static cm_status hnsw_pool(const cm_index* idx, Workspace* ws, const PreparedQueries& pq, uint32_t nq, uint32_t ef,
                           cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  if (!d.has_graph) return fail(CM_ERR_STATE, "HNSW search needs cm_index_build_graph / cm_index_load_graph before upload");
  CM_CUDA(ws->pool_ids.ensure((size_t)std::max(nq, 1u) * ef * 4));
  CM_CUDA(ws->pool_dist.ensure((size_t)std::max(nq, 1u) * ef * 4));
  if (nq == 0) return CM_OK;
  if (!d.has_entry) {  // empty index: search returns [] (:474-476)
    CM_CUDA(cudaMemsetAsync(ws->pool_ids.p, 0xFF, (size_t)nq * ef * 4, s));
    CM_CUDA(launch_fill(ws->pool_dist.as<float>(), (uint64_t)nq * ef, INFINITY, s));
    return CM_OK;
  }
  // visited set: shared-memory hash table for small ef, one byte per node per CTA in HBM from ef = 48 up, where the
  // table (30 * ef entries) would be what limits the queries in flight per SM (measured cross-over between ef = 32
  // and 64, hnsw_search.cu). CM_HNSW_VMAP=0/1 forces either.
  bool use_vmap = ef >= 48;
  if (const char* env = getenv("CM_HNSW_VMAP")) use_vmap = atoi(env) != 0;
  uint32_t hb = use_vmap ? 0 : visited_entries_for(ef);
  unsigned grid = 0;
  uint32_t rows = 0;
  cudaError_t ge = hnsw_grid_size(d, ef, hb, &grid, &rows);
  if (ge != cudaSuccess) return fail(CM_ERR_INVALID_ARGUMENT, "ef / dim too large for the HNSW kernel's shared memory");
  CM_CUDA(ws->small.ensure(kSmallBytes));
  const uint64_t vstride = (d.n + 15) & ~15ull;
  if (use_vmap) {
    const size_t need = (size_t)grid * vstride;
    if (need > ws->vmap.cap || (size_t)grid * 4 > ws->vmap_epoch.cap) {  // fresh slabs start at epoch 0, all zero
      CM_CUDA(ws->vmap.ensure(need));
      CM_CUDA(ws->vmap_epoch.ensure((size_t)grid * 4));
      CM_CUDA(cudaMemsetAsync(ws->vmap.p, 0, ws->vmap.cap, s));
      CM_CUDA(cudaMemsetAsync(ws->vmap_epoch.p, 0, ws->vmap_epoch.cap, s));
    }
  } else {
    CM_CUDA(ws->gtable.ensure(((size_t)grid << kGHashBits) * 4));
  }
  CM_CUDA(cudaMemsetAsync(ws->small.p, 0, 160, s));  // counters + work counter (bad_row at 192 is the caller's)
  HnswLaunch l;
  l.q = pq.qn;
  l.nq = nq;
  l.ldq = pq.ld;
  l.ef = ef;
  l.degenerate = pq.degenerate ? 1 : 0;
  l.pool_ids = ws->pool_ids.as<uint32_t>();
  l.pool_dist = ws->pool_dist.as<float>();
  l.counters = ws->small.as<unsigned long long>();          // 8 x u64 at offset 0
  l.work_counter = ws->small.as<uint32_t>() + 32;           // offset 128
  l.table_entries = hb;
  l.gtable = ws->gtable.as<uint32_t>();
  l.ghash_bits = kGHashBits;
  l.grid = grid;
  l.rows = rows;
  l.status = d.status;
  if (use_vmap) {
    l.vmap = ws->vmap.as<unsigned char>();
    l.vmap_stride = vstride;
    l.vmap_epoch = ws->vmap_epoch.as<uint32_t>();
  }
  {
    ProfileScope prof(s);
    CM_CUDA(launch_hnsw_search(d, l, s));
  }
  return CM_OK;
}

static cm_status check_search_args(const cm_index* idx, const void* q, uint32_t nq, uint32_t k, const void* ids,
                                   const void* out) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  if (!idx->uploaded) return fail(CM_ERR_CUDA, "index is not resident on a CUDA device (call cm_index_upload); there is no CPU path");
  if (nq && (!q || !ids || !out)) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  if (k == 0 || k > CM_MAX_K) return fail(CM_ERR_INVALID_ARGUMENT, "k/ef must be in [1, CM_MAX_K]");
  return CM_OK;
}

static cm_status set_device(const cm_index* idx) {
  CM_CUDA(cudaSetDevice(idx->dev.device));
  return CM_OK;
}

static cm_status strided_copy(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                              cudaStream_t s) {
  if (rows == 0 || width == 0) return CM_OK;
  CM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice, s));
  return CM_OK;
}

static cm_status search_exact_dev_impl(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq, uint32_t qdim,
                                       uint32_t k, uint32_t* ids, float* dist, cudaStream_t s, bool approx = false) {
  PreparedQueries pq;
  cm_status st = prepare_queries(idx, ws, q, nq, qdim, nullptr, s, &pq, exact_uses_tc(idx, qdim != idx->dev.dim, k));
  if (st != CM_OK) return st;
  return exact_topk(idx, ws, pq, nq, k, ids, dist, s, approx);
}

static cm_status search_hnsw_dev_impl(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq, uint32_t qdim,
                                      uint32_t ef, uint32_t k, uint32_t* ids, float* dist, cudaStream_t s) {
  PreparedQueries pq;
  cm_status st = prepare_queries(idx, ws, q, nq, qdim, nullptr, s, &pq);
  if (st != CM_OK) return st;
  uint32_t pool = std::max(ef, k);  // PyO3 Index.query: ef = ef_search.max(n)
  st = hnsw_pool(idx, ws, pq, nq, pool, s);
  if (st != CM_OK) return st;
  st = strided_copy(ids, (size_t)k * 4, ws->pool_ids.p, (size_t)pool * 4, (size_t)k * 4, nq, s);
  if (st != CM_OK) return st;
  return strided_copy(dist, (size_t)k * 4, ws->pool_dist.p, (size_t)pool * 4, (size_t)k * 4, nq, s);
}

static cm_status search_temporal_dev_impl(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq,
                                          uint32_t qdim, uint32_t k, uint32_t ef_search, float w, float base,
                                          uint64_t now_secs, uint32_t now_nanos, int exact, uint32_t* bad_row,
                                          uint32_t* ids, float* score, float* dist, cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  uint64_t ef64 = std::max<uint64_t>(ef_search, (uint64_t)k * 3);  // OVERSAMPLE (src/store.rs:40,323)
  if (ef64 > CM_MAX_K) return fail(CM_ERR_INVALID_ARGUMENT, "max(ef_search, 3k) exceeds CM_MAX_K");
  uint32_t ef = (uint32_t)ef64;
  PreparedQueries pq;
  cm_status st = prepare_queries(idx, ws, q, nq, qdim, bad_row, s, &pq, exact && exact_uses_tc(idx, qdim != d.dim, ef));
  if (st != CM_OK) return st;
  if (exact) {
    CM_CUDA(ws->pool_ids.ensure((size_t)std::max(nq, 1u) * ef * 4));
    CM_CUDA(ws->pool_dist.ensure((size_t)std::max(nq, 1u) * ef * 4));
    st = exact_topk(idx, ws, pq, nq, ef, ws->pool_ids.as<uint32_t>(), ws->pool_dist.as<float>(), s);
  } else {
    st = hnsw_pool(idx, ws, pq, nq, ef, s);
  }
  if (st != CM_OK) return st;
  CM_CUDA(launch_rerank(ws->pool_ids.as<uint32_t>(), ws->pool_dist.as<float>(), nq, ef, k, d.ts_secs, d.ts_nanos,
                        d.decay, w, base, now_secs, now_nanos, ids, score, dist, s));
  return CM_OK;
}

// ChronoMind::search_in_context on the tensor cores (scan_tc.cu): biased threshold scan of the bf16 rows restricted to
// each query's context, then the reference's exact arithmetic on the stored rows of the few dozen candidates. Flagged
// queries (candidate overflow) are counted; the host-buffer caller redoes the call on the fp32 scan.
static cm_status search_in_context_tc(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq, const uint32_t* want,
                                      uint32_t k, float w, float base, uint64_t now_secs, uint32_t now_nanos,
                                      uint32_t* bad_row, uint32_t* ids, float* score, cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  tl_used_tc = true;
  tl_scan_bounded = true;  // no in-stream fallback here: any flagged query means "redo"
  tl_counters.exact_engine = 2;
  // Context-sorted layout (two or more contexts): the scan runs over the copy of the rows ordered by context, the queries
  // are grouped by context (device-side sort per chunk), and every query-tile pair only visits the tiles of its contexts.
  const bool sorted = d.xbf16_ctx != nullptr;
  const uint32_t chunk = sorted ? scan_tc_context_chunk() : kTcQueryChunk;
  const uint32_t nq_pad_all = scan_tc_pad_queries(nq);
  uint32_t *order = nullptr, *qctx_sorted = nullptr, *pair_lo = nullptr, *pair_hi = nullptr;
  if (sorted) {
    const uint32_t n_pairs_all = ((nq + chunk - 1) / chunk) * (chunk / 256);
    CM_CUDA(ws->ctxsort.ensure(((size_t)nq * 2 + (size_t)n_pairs_all * 2) * 4));
    order = ws->ctxsort.as<uint32_t>();
    qctx_sorted = order + nq;
    pair_lo = qctx_sorted + nq;
    pair_hi = pair_lo + n_pairs_all;
    CM_CUDA(launch_group_queries_by_context(want, nq, order, qctx_sorted, s));
    CM_CUDA(launch_context_pair_ranges(d, qctx_sorted, nq, pair_lo, pair_hi, s));
  }
  CM_CUDA(ws->qn.ensure((size_t)nq * d.dim_pad * sizeof(float)));
  CM_CUDA(ws->qbf16.ensure((size_t)nq_pad_all * d.k_pad * 2));
  if (prepare_queries_fused_supported(q, d.dim)) {
    CM_CUDA(launch_prepare_queries(q, nq, d.dim, ws->qn.as<float>(), d.dim_pad, ws->qbf16.p, nq_pad_all, d.k_pad, bad_row, s,
                                   order));
  } else {
    CM_CUDA(launch_preprocess_rows(q, nq, d.dim, d.dim, ws->qn.as<float>(), d.dim_pad, bad_row, s));
    CM_CUDA(launch_to_bf16(ws->qn.as<float>(), nq, d.dim, d.dim_pad, ws->qbf16.p, nq_pad_all, d.k_pad, nullptr, false, s, order));
  }
  CM_CUDA(ws->qnorm.ensure((size_t)nq * 4));
  CM_CUDA(launch_pair_dot(q, q, nq, d.dim, d.dim, ws->qnorm.as<float>(), s));
  CM_CUDA(ws->bias.ensure(d.n_pad * sizeof(float)));
  CM_CUDA(launch_context_bias(d, w, base, now_secs, now_nanos, ws->bias.as<float>(), s, sorted ? d.perm : nullptr));
  const uint32_t cap = scan_tc_cand_cap(k);
  CM_CUDA(ws->small.ensure(kSmallBytes));
  uint8_t* small = ws->small.as<uint8_t>();
  CM_CUDA(cudaMemsetAsync(small + 80, 0, 16, s));  // call totals
  for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
    const uint32_t nb = std::min(chunk, nq - q0);
    const uint32_t nq_pad = scan_tc_pad_queries(nb);
    CM_CUDA(ws->tc_state.ensure((size_t)nq_pad * 8));
    CM_CUDA(ws->cand.ensure((size_t)nq_pad * cap * 8));
    CM_CUDA(ws->qflag.ensure((size_t)nq_pad * 4));
    uint32_t* tau_key = ws->tc_state.as<uint32_t>();
    uint32_t* cand_cnt = tau_key + nq_pad;
    CM_CUDA(cudaMemsetAsync(ws->tc_state.p, 0, (size_t)nq_pad * 8, s));
    CM_CUDA(cudaMemsetAsync(small + 64, 0, 16, s));
    ScanTcLaunch sl;
    sl.q_bf16 = static_cast<const uint8_t*>(ws->qbf16.p) + (size_t)q0 * d.k_pad * 2;
    sl.nq = nb;
    sl.k = k;
    sl.tau_key = tau_key;
    sl.cand_cnt = cand_cnt;
    sl.cand = ws->cand.as<unsigned long long>();
    sl.cap = cap;
    sl.bias = ws->bias.as<float>();
    sl.row_ctx = sorted ? d.ctx_sorted : d.ctx;
    sl.q_ctx = sorted ? qctx_sorted + q0 : want + q0;
    if (sorted) {
      sl.x_bf16 = d.xbf16_ctx;
      sl.pair_lo = pair_lo + q0 / 256;
      sl.pair_hi = pair_hi + q0 / 256;
      // expected range of a pair: the contexts its 256 grouped queries span (if the batch spreads evenly over the
      // contexts) plus one for the straddle, as a fraction of the corpus — only sizes the corpus splits
      const uint64_t C = d.n_contexts;
      const uint64_t span = std::min<uint64_t>(C, (256ull * C + nb - 1) / nb + 1);
      sl.pair_tiles_avg = (uint32_t)std::max<uint64_t>(1, (d.n_pad / 256) * span / C);
    }
    {
      ProfileScope prof(s);
      CM_CUDA(launch_scan_tc(d, sl, s));
    }
    ContextRescoreLaunch rl;
    // (sorted: the scan's query i is the caller's row order[q0 + i]; inputs and outputs are addressed through it)
    rl.q_raw = sorted ? q : q + (size_t)q0 * d.dim;
    rl.ldq = d.dim;
    rl.nq = nb;
    rl.k = k;
    rl.qnorm2 = sorted ? ws->qnorm.as<float>() : ws->qnorm.as<float>() + q0;
    rl.w = w;
    rl.base = base;
    rl.now_secs = now_secs;
    rl.now_nanos = now_nanos;
    rl.tau_key = tau_key;
    rl.cand_cnt = cand_cnt;
    rl.cand = sl.cand;
    rl.cap = cap;
    rl.ids = sorted ? ids : ids + (size_t)q0 * k;
    rl.score = sorted ? score : score + (size_t)q0 * k;
    rl.rescored_total = reinterpret_cast<unsigned long long*>(small + 64);
    rl.flag_count = reinterpret_cast<uint32_t*>(small + 72);
    rl.flag_list = ws->qflag.as<uint32_t>();
    rl.flag_cap = nq_pad;
    if (sorted) {
      rl.perm = d.perm;
      rl.q_order = order + q0;
    }
    CM_CUDA(launch_rescore_context(d, rl, s));
    // every flagged query is unresolved here; the host-buffer caller (bad_row != nullptr) reads the count and redoes
    // the call, the asynchronous one reports through the status word
    CM_CUDA(launch_fold_scan_counters(small + 64, small + 80, 0, bad_row ? nullptr : d.status + 1, s));
  }
  return CM_OK;
}

// ChronoMind::search_in_context (src/store.rs:353-377): fp32 scan of the RAW rows with the context epilogue, then
// the k smallest (score, handle). q: raw queries [nq][dim] (device), want: [nq] context ids (device).
static cm_status search_in_context_dev_impl(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq,
                                            const uint32_t* want, uint32_t k, float w, float base, uint64_t now_secs,
                                            uint32_t now_nanos, uint32_t* bad_row, uint32_t* ids, float* score,
                                            cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  if (!d.ctx && d.n) return fail(CM_ERR_STATE, "this index carries no contexts (cm_index_set_contexts / a snapshot, then upload)");
  if (nq == 0) return CM_OK;
  if (d.n == 0) {
    CM_CUDA(cudaMemsetAsync(ids, 0xFF, (size_t)nq * k * 4, s));
    CM_CUDA(launch_fill(score, (uint64_t)nq * k, INFINITY, s));
    return CM_OK;
  }
  tl_counters.dist_evals += (uint64_t)nq * d.n;
  // The tensor-core path needs 0 <= w < 1 (the bias scales with 1 / (1 - w)) and an index the scan accepts.
  if (w >= 0.0f && w <= 0.999f && exact_uses_tc(idx, false, k))
    return search_in_context_tc(idx, ws, q, nq, want, k, w, base, now_secs, now_nanos, bad_row, ids, score, s);
  tl_counters.exact_engine = 1;
  if (bad_row) {  // validate_query's finiteness check (the normalised copy is scratch)
    CM_CUDA(ws->qn.ensure((size_t)nq * d.dim_pad * sizeof(float)));
    CM_CUDA(launch_preprocess_rows(q, nq, d.dim, d.dim, ws->qn.as<float>(), d.dim_pad, bad_row, s));
  }
  CM_CUDA(ws->qnorm.ensure((size_t)nq * 4));
  CM_CUDA(launch_pair_dot(q, q, nq, d.dim, d.dim, ws->qnorm.as<float>(), s));
  uint64_t ldD = (d.n + 3) & ~3ull;
  uint64_t qb = std::max<uint64_t>(16, (kDistWorkspaceBytes / (ldD * 4)) & ~15ull);
  qb = std::min<uint64_t>(qb, ((uint64_t)nq + 15) & ~15ull);
  CM_CUDA(ws->D.ensure(qb * ldD * sizeof(float)));
  float* D = ws->D.as<float>();
  ContextEpilogue ce;
  ce.enabled = 1;
  ce.na = d.norm2_raw;
  ce.ctx = d.ctx;
  ce.deleted = d.deleted;
  ce.ts_secs = reinterpret_cast<const unsigned long long*>(d.ts_secs);
  ce.ts_nanos = d.ts_nanos;
  ce.decay = d.decay;
  ce.w = w;
  ce.base = base;
  ce.now_secs = now_secs;
  ce.now_nanos = now_nanos;
  for (uint32_t q0 = 0; q0 < nq; q0 += (uint32_t)qb) {
    uint32_t nb = std::min<uint32_t>((uint32_t)qb, nq - q0);
    ce.nb = ws->qnorm.as<float>() + q0;
    ce.want = want + q0;
    CM_CUDA(launch_exact_dist(q + (size_t)q0 * d.dim, nb, d.dim, d.xraw, d.n, d.dim, d.dim_pad, D, ldD, nullptr, 0, s, &ce));
    CM_CUDA(launch_select_from_dist(D, ldD, nb, d.n, nullptr, k, ids + (size_t)q0 * k, score + (size_t)q0 * k, k, nullptr,
                                    nullptr, 0, s, /*skip_inf=*/true));
  }
  return CM_OK;
}

static void fetch_scan_counters(Workspace* ws) {
  if (!ws->small.p) return;
  unsigned long long r = 0;
  uint32_t f = 0;
  uint32_t worst = 0;
  if (cudaMemcpy(&r, ws->small.as<uint8_t>() + 80, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  if (cudaMemcpy(&f, ws->small.as<uint8_t>() + 88, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  if (cudaMemcpy(&worst, ws->small.as<uint8_t>() + 92, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  tl_counters.rescored += r;
  tl_counters.fallback_queries += f;
  tl_scan_worst_chunk = worst;
}

static void fetch_hnsw_counters(Workspace* ws) {
  if (!ws->small.p) return;
  unsigned long long c[8];
  if (cudaMemcpy(c, ws->small.p, sizeof c, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  tl_counters.dist_evals += c[0];
  tl_counters.expansions += c[1];
  tl_counters.list_entries += c[2];
  tl_counters.fallback_queries += c[3];
  tl_counters.tie_replays += c[5];
  tl_counters.failures += c[4] + c[6];
}

// Host-buffer HNSW calls: a query the kernel could not answer like the reference is an error, never a quiet CM_OK.
static cm_status hnsw_failures_to_status() {
  if (tl_counters.failures == 0) return CM_OK;
  return fail(CM_ERR_CAPACITY_EXCEEDED,
              std::to_string(tl_counters.failures) + " quer(ies) exceeded the traversal's visited table or parked more than " +
                  "a kernel's worth of tied candidates; results for them are not the reference's");
}

}  // namespace cm

using namespace cm;

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

const char* cm_last_error(void) { return tl_error.c_str(); }
const char* cm_version(void) { return "chronomind_b200 0.1.0 (sm_100a)"; }
const char* cm_metric_name(void) { return "cosine"; }  // src/metric.rs:201-203
uint64_t cm_launch_count(void) { return g_launches.load(); }
cm_status cm_last_counters(cm_counters* out) {
  if (!out) return fail(CM_ERR_INVALID_ARGUMENT, "null out");
  *out = tl_counters;
  return CM_OK;
}

cm_status cm_profile_enable(int on) {
  g_profile.store(on ? 1 : 0);
  return CM_OK;
}
cm_status cm_profile_read(float* kernel_ms, uint32_t* kernel_launches) {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
  {
    std::lock_guard<std::mutex> g(g_profile_mu);
    evs.swap(g_profile_events);
  }
  float total = 0.0f;
  for (auto& e : evs) {
    float ms = 0.0f;
    if (cudaEventSynchronize(e.second) == cudaSuccess && cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess)
      total += ms;
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  if (kernel_ms) *kernel_ms = total;
  if (kernel_launches) *kernel_launches = (uint32_t)evs.size();
  return CM_OK;
}

void cm_config_default(cm_config* c) {  // src/config.rs:73-85, 27-35
  if (!c) return;
  c->dimensions = 768;
  c->max_memories = 100000;
  c->base_decay_rate = 0.1f;
  c->temporal_weight = 0.3f;
  c->similarity_threshold = 0.95f;
  c->max_relationships = 50;
  c->index.max_connections = 16;
  c->index.ef_construction = 200;
  c->index.ef_search = 50;
}

static cm_index* new_index(const cm_config& cfg, uint64_t n, uint32_t dim) {
  cm_index* idx = new cm_index();
  idx->config = cfg;
  idx->n = n;
  idx->dim = dim;
  idx->deleted.assign(n, 0);
  idx->ts_secs.assign(n, 0);
  idx->ts_nanos.assign(n, 0);
  idx->decay.assign(n, 0.0f);
  idx->live = n;
  idx->workspace_pool = new WorkspacePool();
  return idx;
}

cm_status cm_index_from_matrix(const float* x, uint64_t n, uint32_t dim, const cm_config* cfg, cm_index** out) {
  if (!out) return fail(CM_ERR_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  if (n && !x) return fail(CM_ERR_INVALID_ARGUMENT, "null matrix");
  cm_config c;
  if (cfg)
    c = *cfg;
  else
    cm_config_default(&c);
  c.dimensions = dim;
  cm_status st = validate_config(c);
  if (st != CM_OK) return st;
  if (n > kArenaCapacity)
    return fail(CM_ERR_INDEX_FULL, "index storage exhausted (16777216 slots, including tombstones); compact via snapshot reload");
  cm_index* idx = new_index(c, n, dim);
  idx->x.resize(n * (uint64_t)dim);
  float* dst = idx->x.data();
  std::atomic<bool> finite{true};
  std::atomic<bool>* fin = &finite;
  parallel_rows(n, [=](uint64_t i) {
    preprocess_row(x + i * dim, dim, dst + i * dim);
    if (!row_is_finite(dst + i * dim, dim)) fin->store(false, std::memory_order_relaxed);
  });
  idx->all_finite = finite.load();
  *out = idx;
  return CM_OK;
}

cm_status cm_index_from_snapshot(const char* path, cm_index** out) {
  if (!out || !path) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  const bool timing = getenv("CM_LOAD_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[index_from_snapshot] %-22s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };
  SnapshotData snap;
  cm_status st = read_snapshot(path, &snap);
  if (st != CM_OK) return st;
  lap("read_snapshot");
  uint64_t n = snap.memories.size();
  uint32_t dim = (uint32_t)snap.config.dimensions;
  // Replay `store.insert` record by record in file order (src/persistence.rs:117-119, src/store.rs:226-264):
  // validate, then the capacity check (distinct ids), then the handle bookkeeping — a repeated id publishes the new
  // handle and tombstones the one it replaces. The first failing record decides the error, as in the reference.
  std::vector<uint8_t> deleted(n, 0);
  uint64_t live = n;
  {
    // keys are views of the decoded records' own strings (snap.memories is not resized any more): no copy per record
    std::unordered_map<std::string_view, uint32_t> by_id;
    by_id.reserve(n * 2);
    for (uint64_t i = 0; i < n; ++i) {
      if ((st = validate_snapshot_memory(snap, i)) != CM_OK) return st;
      const SnapshotMemory& m = snap.memories[i];
      const std::string_view key(m.id);
      auto it = by_id.find(key);
      if (it != by_id.end()) {
        deleted[it->second] = 1;
        live--;
        it->second = (uint32_t)i;
      } else {
        if (by_id.size() >= snap.config.max_memories)
          return fail(CM_ERR_CAPACITY_EXCEEDED, "store is at capacity (" + std::to_string(snap.config.max_memories) + " memories)");
        by_id.emplace(key, (uint32_t)i);
      }
      if (i + 1 > kArenaCapacity)  // `VectorIndex::insert` -> None (src/index/arena.rs:114-118)
        return fail(CM_ERR_INDEX_FULL, "index storage exhausted (16777216 slots, including tombstones); compact via snapshot reload");
    }
  }
  lap("validate + id replay");
  cm_index* idx = new_index(snap.config, n, dim);
  idx->deleted = std::move(deleted);
  idx->live = live;
  idx->x.resize(n * (uint64_t)dim);
  lap("allocate rows");
  const float* src = snap.vectors.data();  // every record has exactly `dim` floats now: offsets are i * dim
  float* dst = idx->x.data();
  std::atomic<bool> finite{true};
  std::atomic<bool>* fin = &finite;
  parallel_rows(n, [=](uint64_t i) {
    preprocess_row(src + i * dim, dim, dst + i * dim);
    if (!row_is_finite(dst + i * dim, dim)) fin->store(false, std::memory_order_relaxed);
  });
  idx->all_finite = finite.load();
  lap("preprocess rows");
  idx->ext_ids.resize(n);
  // contexts in order of first appearance; the raw rows stay for search_in_context (metric.distance on stored data)
  idx->ctx.resize(n);
  idx->importance.resize(n);
  {
    std::unordered_map<std::string, uint32_t> ctx_of;
    for (uint64_t i = 0; i < n; ++i) {
      SnapshotMemory& m = snap.memories[i];
      idx->importance[i] = m.importance;
      idx->ext_ids[i] = std::move(m.id);  // the snapshot records are dropped right after this loop
      idx->ts_secs[i] = m.ts_secs;
      idx->ts_nanos[i] = m.ts_nanos;
      idx->decay[i] = m.decay_rate;
      auto it = ctx_of.find(m.context);
      if (it == ctx_of.end()) {
        it = ctx_of.emplace(m.context, (uint32_t)idx->ctx_labels.size()).first;
        idx->ctx_labels.push_back(m.context);
      }
      idx->ctx[i] = it->second;
    }
  }
  idx->raw = std::move(snap.vectors);
  lap("attributes + contexts");
  *out = idx;
  return CM_OK;
}

cm_status cm_index_set_contexts(cm_index* idx, const float* raw_x, const uint32_t* context_ids) {
  if (!idx || (idx->n && !raw_x)) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  idx->raw.assign(raw_x, raw_x + idx->n * (uint64_t)idx->dim);
  if (context_ids)
    idx->ctx.assign(context_ids, context_ids + idx->n);
  else
    idx->ctx.assign(idx->n, 0);  // one context
  return CM_OK;
}

const char* cm_index_context_label(const cm_index* idx, uint32_t handle) {
  if (!idx || handle >= idx->ctx.size() || idx->ctx[handle] >= idx->ctx_labels.size()) return "";
  return idx->ctx_labels[idx->ctx[handle]].c_str();
}

float cm_index_importance(const cm_index* idx, uint32_t handle) {
  return (idx && handle < idx->importance.size()) ? idx->importance[handle] : 0.0f;
}

uint32_t cm_index_context_id(const cm_index* idx, const char* label) {
  if (!idx || !label) return CM_NO_ID;
  for (size_t i = 0; i < idx->ctx_labels.size(); ++i)
    if (idx->ctx_labels[i] == label) return (uint32_t)i;
  return CM_NO_ID;
}

cm_status cm_index_set_attrs(cm_index* idx, const uint64_t* ts_secs, const uint32_t* ts_nanos, const float* decay_rate,
                             const uint8_t* deleted) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  uint64_t n = idx->n;
  if (decay_rate)
    for (uint64_t i = 0; i < n; ++i)
      if (!std::isfinite(decay_rate[i]) || decay_rate[i] < 0.0f)  // Memory::validate (src/types.rs:112-118)
        return fail(CM_ERR_INVALID_VECTOR, "invalid vector data: vector has an invalid decay rate");
  if (ts_nanos)
    for (uint64_t i = 0; i < n; ++i)
      if (ts_nanos[i] >= 1000000000u) return fail(CM_ERR_INVALID_ARGUMENT, "timestamp nanoseconds out of range");
  if (ts_secs) std::copy(ts_secs, ts_secs + n, idx->ts_secs.begin());
  if (ts_nanos) std::copy(ts_nanos, ts_nanos + n, idx->ts_nanos.begin());
  if (decay_rate) std::copy(decay_rate, decay_rate + n, idx->decay.begin());
  if (deleted) {
    uint64_t live = 0;
    for (uint64_t i = 0; i < n; ++i) {
      idx->deleted[i] = deleted[i] ? 1 : 0;
      live += deleted[i] ? 0 : 1;
    }
    idx->live = live;
  }
  return CM_OK;
}

cm_status cm_index_remove(cm_index* idx, uint32_t handle, int* removed) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  int r = 0;
  if (handle < idx->n && !idx->deleted[handle]) {
    idx->deleted[handle] = 1;
    idx->live--;
    if (idx->uploaded) idx->pending_removed.push_back(handle);  // reaches the device at cm_index_sync_deleted / upload
    r = 1;
  }
  if (removed) *removed = r;
  return CM_OK;
}

cm_status cm_index_sync_deleted(cm_index* idx) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  if (!idx->uploaded) return fail(CM_ERR_CUDA, "index is not resident on a CUDA device (call cm_index_upload); there is no CPU path");
  if (idx->pending_removed.empty()) return CM_OK;
  cm_status st = set_device(idx);
  if (st != CM_OK) return st;
  DeviceIndex& d = idx->dev;
  const size_t m = idx->pending_removed.size();
  uint32_t* handles = nullptr;
  CM_CUDA(cudaMalloc(&handles, m * 4));
  cudaError_t e = cudaMemcpy(handles, idx->pending_removed.data(), m * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_apply_tombstones(handles, (uint32_t)m, d.deleted, d.xbf16, d.k_pad, d.xbf16_ctx, d.inv_perm, nullptr);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(handles);
  if (e != cudaSuccess) return cuda_fail(e, "cm_index_sync_deleted");
  d.live = idx->live;
  idx->pending_removed.clear();
  return CM_OK;
}

cm_status cm_index_build_graph(cm_index* idx, uint64_t layer_seed, int threads) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  build_graph(idx->x.data(), idx->n, idx->dim, idx->config.index, layer_seed, threads, &idx->graph);
  return CM_OK;
}

cm_status cm_index_set_shard(cm_index* idx, uint32_t shard_rank, uint32_t shard_count) {
  if (!idx || shard_count == 0 || shard_rank >= shard_count) return fail(CM_ERR_INVALID_ARGUMENT, "bad shard");
  idx->shard_rank = shard_rank;
  idx->shard_count = shard_count;
  return CM_OK;
}

cm_status cm_index_upload(cm_index* idx, int device) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(CM_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(CM_ERR_INVALID_ARGUMENT, "device ordinal out of range");
  free_device(idx->dev);
  idx->uploaded = false;
  idx->pending_removed.clear();  // the upload below carries every tombstone
  CM_CUDA(cudaSetDevice(device));
  DeviceIndex& d = idx->dev;
  d.device = device;
  d.n = idx->n;
  d.dim = idx->dim;
  d.dim_pad = (idx->dim + 7) & ~7u;
  d.live = idx->live;
  cudaDeviceProp prop;
  CM_CUDA(cudaGetDeviceProperties(&prop, device));
  d.sm_count = prop.multiProcessorCount;
  uint64_t n = idx->n;
  size_t rows = std::max<uint64_t>(n, 1);
  CM_CUDA(cudaMalloc(&d.x32, rows * d.dim_pad * sizeof(float)));
  if (n) {
    if (d.dim_pad != d.dim) CM_CUDA(cudaMemset(d.x32, 0, rows * d.dim_pad * sizeof(float)));
    CM_CUDA(cudaMemcpy2D(d.x32, (size_t)d.dim_pad * 4, idx->x.data(), (size_t)d.dim * 4, (size_t)d.dim * 4, n,
                         cudaMemcpyHostToDevice));
  }
  CM_CUDA(cudaMalloc(&d.status, 8 * sizeof(uint32_t)));
  CM_CUDA(cudaMemset(d.status, 0, 8 * sizeof(uint32_t)));
  CM_CUDA(cudaMalloc(&d.deleted, rows));
  CM_CUDA(cudaMalloc(&d.ts_secs, rows * 8));
  CM_CUDA(cudaMalloc(&d.ts_nanos, rows * 4));
  CM_CUDA(cudaMalloc(&d.decay, rows * 4));
  if (n) {
    CM_CUDA(cudaMemcpy(d.deleted, idx->deleted.data(), n, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(d.ts_secs, idx->ts_secs.data(), n * 8, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(d.ts_nanos, idx->ts_nanos.data(), n * 4, cudaMemcpyHostToDevice));
    CM_CUDA(cudaMemcpy(d.decay, idx->decay.data(), n * 4, cudaMemcpyHostToDevice));
  }
  if (!idx->ctx.empty() && n) {  // store-level context scan: raw rows, their squared norms, context ids
    CM_CUDA(cudaMalloc(&d.xraw, rows * d.dim_pad * sizeof(float)));
    if (d.dim_pad != d.dim) CM_CUDA(cudaMemset(d.xraw, 0, rows * d.dim_pad * sizeof(float)));
    CM_CUDA(cudaMemcpy2D(d.xraw, (size_t)d.dim_pad * 4, idx->raw.data(), (size_t)d.dim * 4, (size_t)d.dim * 4, n,
                         cudaMemcpyHostToDevice));
    CM_CUDA(cudaMalloc(&d.norm2_raw, rows * 4));
    CM_CUDA(launch_pair_dot(d.xraw, d.xraw, n, d.dim, d.dim_pad, d.norm2_raw, nullptr));
    // padded to the scan's tile height (the biased scan stages whole tiles of context ids); padding = CM_NO_ID
    CM_CUDA(cudaMalloc(&d.ctx, scan_tc_pad_rows(rows) * 4));
    CM_CUDA(cudaMemset(d.ctx, 0xFF, scan_tc_pad_rows(rows) * 4));
    CM_CUDA(cudaMemcpy(d.ctx, idx->ctx.data(), n * 4, cudaMemcpyHostToDevice));
  }
  // bf16 copy for the tensor-core scan, padded to whole tiles; tombstoned and padding rows are NaN-masked.
  d.n_pad = scan_tc_pad_rows(std::max<uint64_t>(n, 1));
  d.k_pad = scan_tc_pad_k(d.dim);
  CM_CUDA(cudaMalloc(&d.xbf16, d.n_pad * d.k_pad * 2));
  CM_CUDA(launch_to_bf16(d.x32, n, d.dim, d.dim_pad, d.xbf16, d.n_pad, d.k_pad, n ? d.deleted : nullptr, true, nullptr));
  d.n_contexts = 0;
  if (!idx->ctx.empty() && n) {
    // Context-sorted copy for `search_in_context`: rows ordered by (context, handle), so that a context is a contiguous
    // row range and a batch of queries grouped by context only scans the tiles that hold its contexts.
    std::vector<uint32_t> perm(n), inv(n);
    for (uint64_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
    const std::vector<uint32_t>& cx = idx->ctx;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return cx[a] < cx[b]; });
    idx->ctx_sorted.resize(n);
    for (uint64_t p = 0; p < n; ++p) {
      idx->ctx_sorted[p] = cx[perm[p]];
      inv[perm[p]] = (uint32_t)p;
      if (p == 0 || idx->ctx_sorted[p] != idx->ctx_sorted[p - 1]) ++d.n_contexts;
    }
    // (one context: the plain layout already is the sorted one; CM_CONTEXT_UNSORTED=1 keeps the mask-only scan, for A/B runs)
    if (d.n_contexts >= 2 && !getenv("CM_CONTEXT_UNSORTED")) {
      CM_CUDA(cudaMalloc(&d.perm, d.n_pad * 4));
      CM_CUDA(cudaMalloc(&d.inv_perm, rows * 4));
      CM_CUDA(cudaMalloc(&d.ctx_sorted, d.n_pad * 4));
      CM_CUDA(cudaMemset(d.perm, 0, d.n_pad * 4));
      CM_CUDA(cudaMemset(d.ctx_sorted, 0xFF, d.n_pad * 4));
      CM_CUDA(cudaMemcpy(d.perm, perm.data(), n * 4, cudaMemcpyHostToDevice));
      CM_CUDA(cudaMemcpy(d.inv_perm, inv.data(), n * 4, cudaMemcpyHostToDevice));
      CM_CUDA(cudaMemcpy(d.ctx_sorted, idx->ctx_sorted.data(), n * 4, cudaMemcpyHostToDevice));
      CM_CUDA(cudaMalloc(&d.xbf16_ctx, d.n_pad * d.k_pad * 2));
      CM_CUDA(launch_to_bf16(d.x32, n, d.dim, d.dim_pad, d.xbf16_ctx, d.n_pad, d.k_pad, d.deleted, true, nullptr, d.perm));
    }
  }
  const Graph& g = idx->graph;
  d.has_graph = g.built;
  if (g.built) {
    d.cap0 = g.cap0;
    d.capU = g.capU;
    d.has_entry = g.entry != kEmptyEntry;
    d.entry_id = (uint32_t)g.entry;
    d.entry_top = d.has_entry ? (uint32_t)(g.entry >> 32) : 0;
    auto up = [&](uint32_t** dst, const std::vector<uint32_t>& v) -> cudaError_t {
      cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(v.size(), 1) * 4);
      if (e != cudaSuccess) return e;
      if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
      return e;
    };
    CM_CUDA(up(&d.nbr0, g.nbr0));
    CM_CUDA(up(&d.deg0, g.deg0));
    CM_CUDA(up(&d.upper_base, g.upper_base));
    CM_CUDA(up(&d.nbr_up, g.nbr_up));
    CM_CUDA(up(&d.deg_up, g.deg_up));
  }
  CM_CUDA(cudaDeviceSynchronize());
  idx->uploaded = true;
  return CM_OK;
}

void cm_index_free(cm_index* idx) {
  if (!idx) return;
  if (idx->dev.device >= 0) cudaSetDevice(idx->dev.device);
  delete (WorkspacePool*)idx->workspace_pool;
  free_device(idx->dev);
  delete idx;
}

uint64_t cm_index_slots(const cm_index* idx) { return idx ? idx->n : 0; }
uint64_t cm_index_len(const cm_index* idx) { return idx ? idx->live : 0; }
uint32_t cm_index_dim(const cm_index* idx) { return idx ? idx->dim : 0; }
cm_status cm_index_get_config(const cm_index* idx, cm_config* out) {
  if (!idx || !out) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  *out = idx->config;
  return CM_OK;
}
const char* cm_index_external_id(const cm_index* idx, uint32_t handle) {
  if (!idx || handle >= idx->ext_ids.size()) return "";
  return idx->ext_ids[handle].c_str();
}
int cm_index_check_invariants(const cm_index* idx) {
  if (!idx || !idx->graph.built) return -1;
  return check_graph_invariants(idx->graph);
}
uint64_t cm_index_unreachable(const cm_index* idx) {
  if (!idx || !idx->graph.built) return 0;
  return count_unreachable(idx->graph);
}
uint64_t cm_index_entry(const cm_index* idx) { return idx && idx->graph.built ? idx->graph.entry : kEmptyEntry; }
uint32_t cm_index_top_layer(const cm_index* idx, uint32_t handle) {
  if (!idx || !idx->graph.built || handle >= idx->n) return 0;
  return idx->graph.top_layer[handle];
}
uint32_t cm_index_neighbors(const cm_index* idx, uint32_t handle, uint32_t layer, uint32_t* out, uint32_t cap) {
  if (!idx || !idx->graph.built || handle >= idx->n || layer > idx->graph.top_layer[handle]) return 0;
  uint32_t deg;
  const uint32_t* l = idx->graph.list(handle, layer, &deg);
  for (uint32_t i = 0; i < deg && i < cap; ++i) out[i] = l[i];
  return deg;
}
cm_status cm_index_vector(const cm_index* idx, uint32_t handle, float* out) {
  if (!idx || !out || handle >= idx->n) return fail(CM_ERR_INVALID_ARGUMENT, "bad handle");
  std::memcpy(out, idx->x.data() + (uint64_t)handle * idx->dim, (size_t)idx->dim * 4);
  return CM_OK;
}

// ---- graph sidecar ------------------------------------------------------------------------------------------
// "CMGRAPH1" | n u64 | cap0 u32 | capU u32 | entry u64 | slots u64 | top_layer[n] u8 | deg0[n] | nbr0[n*cap0] |
// upper_base[n] | deg_up[slots] | nbr_up[slots*capU]   (all little endian)
cm_status cm_index_save_graph(const cm_index* idx, const char* path) {
  if (!idx || !path) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  const Graph& g = idx->graph;
  if (!g.built) return fail(CM_ERR_STATE, "no graph to save");
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(CM_ERR_IO, std::string("io error: cannot open ") + path);
  uint64_t slots = g.deg_up.size();
  bool ok = std::fwrite("CMGRAPH1", 1, 8, f) == 8;
  auto w = [&](const void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
  w(&g.n, 8);
  w(&g.cap0, 4);
  w(&g.capU, 4);
  w(&g.entry, 8);
  w(&slots, 8);
  w(g.top_layer.data(), g.top_layer.size());
  w(g.deg0.data(), g.deg0.size() * 4);
  w(g.nbr0.data(), g.nbr0.size() * 4);
  w(g.upper_base.data(), g.upper_base.size() * 4);
  w(g.deg_up.data(), g.deg_up.size() * 4);
  w(g.nbr_up.data(), g.nbr_up.size() * 4);
  ok = (std::fclose(f) == 0) && ok;
  return ok ? CM_OK : fail(CM_ERR_IO, "io error: short write");
}

cm_status cm_index_load_graph(cm_index* idx, const char* path) {
  if (!idx || !path) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(CM_ERR_IO, std::string("io error: cannot open ") + path);
  Graph g;
  char magic[8];
  uint64_t slots = 0;
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "CMGRAPH1", 8) == 0;
  auto r = [&](void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fread(p, 1, bytes, f) == bytes); };
  r(&g.n, 8);
  r(&g.cap0, 4);
  r(&g.capU, 4);
  r(&g.entry, 8);
  r(&slots, 8);
  if (!ok || g.n != idx->n || g.cap0 != idx->config.index.max_connections * 2 ||
      g.capU != idx->config.index.max_connections || slots > g.n * 32) {
    std::fclose(f);
    return fail(CM_ERR_INVALID_SNAPSHOT, "invalid snapshot: graph sidecar does not match this index");
  }
  g.top_layer.resize(g.n);
  g.deg0.resize(g.n);
  g.nbr0.resize(g.n * g.cap0);
  g.upper_base.resize(g.n);
  g.deg_up.resize(slots);
  g.nbr_up.resize(slots * g.capU);
  r(g.top_layer.data(), g.top_layer.size());
  r(g.deg0.data(), g.deg0.size() * 4);
  r(g.nbr0.data(), g.nbr0.size() * 4);
  r(g.upper_base.data(), g.upper_base.size() * 4);
  r(g.deg_up.data(), g.deg_up.size() * 4);
  r(g.nbr_up.data(), g.nbr_up.size() * 4);
  std::fclose(f);
  if (!ok) return fail(CM_ERR_INVALID_SNAPSHOT, "invalid snapshot: truncated graph sidecar");
  g.built = true;
  for (uint64_t i = 0; i < g.n; ++i) {  // bounds before trusting it on the device
    uint32_t t = g.top_layer[i];
    if (t > kMaxLayer || (t > 0 && (uint64_t)g.upper_base[i] + t > slots))
      return fail(CM_ERR_INVALID_SNAPSHOT, "invalid snapshot: corrupt graph sidecar");
  }
  if (check_graph_invariants(g) != 0) return fail(CM_ERR_INVALID_SNAPSHOT, "invalid snapshot: graph sidecar fails invariants");
  idx->graph = std::move(g);
  return CM_OK;
}

// ---- search: device buffers -----------------------------------------------------------------------------------
cm_status cm_search_exact_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t k, void* ids, void* dist,
                              void* stream) {
  cm_status st = check_search_args(idx, q, nq, k, ids, dist);
  if (st != CM_OK) return st;
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, (cudaStream_t)stream, false);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  return search_exact_dev_impl(idx, lease.ws, (const float*)q, nq, idx->dim, k, (uint32_t*)ids, (float*)dist,
                               (cudaStream_t)stream);
}

cm_status cm_search_hnsw_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t ef, uint32_t k, void* ids,
                             void* dist, void* stream) {
  cm_status st = check_search_args(idx, q, nq, k, ids, dist);
  if (st != CM_OK) return st;
  if (ef > CM_MAX_K) return fail(CM_ERR_INVALID_ARGUMENT, "ef exceeds CM_MAX_K");
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, (cudaStream_t)stream, false);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  return search_hnsw_dev_impl(idx, lease.ws, (const float*)q, nq, idx->dim, ef, k, (uint32_t*)ids, (float*)dist,
                              (cudaStream_t)stream);
}

cm_status cm_search_temporal_dev(const cm_index* idx, const void* q, uint32_t nq, uint32_t k, uint32_t ef_search,
                                 float w, float base, uint64_t now_secs, uint32_t now_nanos, int exact, void* ids,
                                 void* score, void* dist, void* stream) {
  cm_status st = check_search_args(idx, q, nq, k, ids, score);
  if (st != CM_OK) return st;
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, (cudaStream_t)stream, false);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  return search_temporal_dev_impl(idx, lease.ws, (const float*)q, nq, idx->dim, k, ef_search, w, base, now_secs,
                                  now_nanos, exact, nullptr, (uint32_t*)ids, (float*)score, (float*)dist,
                                  (cudaStream_t)stream);
}

cm_status cm_search_in_context_dev(const cm_index* idx, const void* q, uint32_t nq, const void* context_ids, uint32_t k,
                                   float w, float base, uint64_t now_secs, uint32_t now_nanos, void* ids, void* score,
                                   void* stream) {
  cm_status st = check_search_args(idx, q, nq, k, ids, score);
  if (st != CM_OK) return st;
  if (nq && !context_ids) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, (cudaStream_t)stream, false);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  return search_in_context_dev_impl(idx, lease.ws, (const float*)q, nq, (const uint32_t*)context_ids, k, w, base, now_secs,
                                    now_nanos, nullptr, (uint32_t*)ids, (float*)score, (cudaStream_t)stream);
}

cm_status cm_merge_topk_dev(const void* ids, const void* dist, uint32_t n_shards, uint32_t nq, uint32_t len,
                            uint32_t out_len, void* out_ids, void* out_dist, int device, void* stream) {
  if (!ids || !dist || !out_ids || !out_dist || n_shards == 0 || len == 0 || out_len == 0)
    return fail(CM_ERR_INVALID_ARGUMENT, "bad merge arguments");
  if ((uint64_t)n_shards * len > 16384) return fail(CM_ERR_INVALID_ARGUMENT, "n_shards * len exceeds 16384");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_merge_topk((const uint32_t*)ids, (const float*)dist, nullptr, n_shards, nq, len, out_len, (uint32_t*)out_ids,
                            (float*)out_dist, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_pack_topk_dev(const void* ids, const void* dist, uint64_t count, void* out_keys, int device, void* stream) {
  if (count && (!ids || !dist || !out_keys)) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_pack_topk((const uint32_t*)ids, (const float*)dist, count, (unsigned long long*)out_keys, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_merge_topk_keys_dev(const void* keys, uint32_t n_shards, uint32_t nq, uint32_t len, uint32_t out_len,
                                 void* out_ids, void* out_dist, int device, void* stream) {
  if (!keys || !out_ids || !out_dist || n_shards == 0 || len == 0 || out_len == 0)
    return fail(CM_ERR_INVALID_ARGUMENT, "bad merge arguments");
  if ((uint64_t)n_shards * len > 16384) return fail(CM_ERR_INVALID_ARGUMENT, "n_shards * len exceeds 16384");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_merge_topk(nullptr, nullptr, (const unsigned long long*)keys, n_shards, nq, len, out_len, (uint32_t*)out_ids,
                            (float*)out_dist, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_rerank_dev(const void* pool_ids, const void* pool_dist, uint32_t nq, uint32_t pool, uint32_t k,
                        const void* ts_secs, const void* ts_nanos, const void* decay_rate, float w, float base,
                        uint64_t now_secs, uint32_t now_nanos, void* out_ids, void* out_score, void* out_dist,
                        int device, void* stream) {
  if (!pool_ids || !pool_dist || !ts_secs || !ts_nanos || !decay_rate || !out_ids || !out_score || pool == 0 ||
      k == 0 || pool > 8192)
    return fail(CM_ERR_INVALID_ARGUMENT, "bad rerank arguments");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_rerank((const uint32_t*)pool_ids, (const float*)pool_dist, nq, pool, k, (const uint64_t*)ts_secs,
                        (const uint32_t*)ts_nanos, (const float*)decay_rate, w, base, now_secs, now_nanos,
                        (uint32_t*)out_ids, (float*)out_score, (float*)out_dist, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_metric_preprocess_dev(const void* x, uint64_t n, uint32_t dim, void* out, int device, void* stream) {
  if (n && (!x || !out)) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_preprocess_rows((const float*)x, n, dim, dim, (float*)out, dim, nullptr, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_metric_distance_prepared_dev(const void* a, const void* b, uint64_t n, uint32_t dim, void* out, int device,
                                          void* stream) {
  if (n && (!a || !b || !out)) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_pair_distance((const float*)a, (const float*)b, n, dim, dim, (float*)out, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_index_dev_status(const cm_index* idx, uint64_t* failures, int reset) {
  if (!idx || !failures) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  *failures = 0;
  if (!idx->uploaded) return fail(CM_ERR_CUDA, "index is not resident on a CUDA device (call cm_index_upload); there is no CPU path");
  cm_status st = set_device(idx);
  if (st != CM_OK) return st;
  uint32_t w[8];
  CM_CUDA(cudaMemcpy(w, idx->dev.status, sizeof w, cudaMemcpyDeviceToHost));  // waits for the work enqueued so far
  *failures = (uint64_t)w[0] + w[1];
  if (reset) CM_CUDA(cudaMemset(idx->dev.status, 0, 2 * sizeof(uint32_t)));  // the work counters have their own reset
  return CM_OK;
}

cm_status cm_index_dev_counters(const cm_index* idx, uint64_t* rescored, uint64_t* fallback_queries, int reset) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  if (!idx->uploaded) return fail(CM_ERR_CUDA, "index is not resident on a CUDA device (call cm_index_upload); there is no CPU path");
  cm_status st = set_device(idx);
  if (st != CM_OK) return st;
  uint32_t w[8];
  CM_CUDA(cudaMemcpy(w, idx->dev.status, sizeof w, cudaMemcpyDeviceToHost));  // waits for the work enqueued so far
  if (rescored) *rescored = ((uint64_t)w[3] << 32) | w[2];
  if (fallback_queries) *fallback_queries = w[4];
  if (reset) CM_CUDA(cudaMemset(idx->dev.status + 2, 0, 3 * sizeof(uint32_t)));
  return CM_OK;
}

cm_status cm_set_exact_engine(cm_index* idx, int engine) {
  if (!idx || engine < 0 || engine > 2) return fail(CM_ERR_INVALID_ARGUMENT, "engine must be 0, 1 or 2");
  idx->exact_engine = engine;
  return CM_OK;
}

// ---- search: host buffers -----------------------------------------------------------------------------------------
// Stage q on the device, run the device path on the workspace's stream, copy results back, synchronise.
struct HostCall {
  Workspace* ws;
  cudaStream_t s;
  float* q_dev = nullptr;
  uint32_t* ids_dev = nullptr;
  float* a_dev = nullptr;
  float* b_dev = nullptr;
};

static cm_status host_begin(const cm_index* idx, Workspace* ws, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                            HostCall* hc) {
  hc->ws = ws;
  hc->s = ws->stream;
  size_t qbytes = (size_t)nq * qdim * sizeof(float);
  CM_CUDA(ws->qraw.ensure(std::max<size_t>(qbytes, 16)));
  CM_CUDA(ws->out_ids.ensure(std::max<size_t>((size_t)nq * k * 4, 16)));
  CM_CUDA(ws->out_a.ensure(std::max<size_t>((size_t)nq * k * 4, 16)));
  CM_CUDA(ws->out_b.ensure(std::max<size_t>((size_t)nq * k * 4, 16)));
  hc->q_dev = ws->qraw.as<float>();
  hc->ids_dev = ws->out_ids.as<uint32_t>();
  hc->a_dev = ws->out_a.as<float>();
  hc->b_dev = ws->out_b.as<float>();
  if (qbytes) CM_CUDA(cudaMemcpyAsync(hc->q_dev, q, qbytes, cudaMemcpyHostToDevice, hc->s));
  (void)idx;
  return CM_OK;
}

static cm_status search_exact_host(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                                   uint32_t* ids, float* dist, bool approx) {
  cm_status st = check_search_args(idx, q, nq, k, ids, dist);
  if (st != CM_OK) return st;
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  HostCall hc;
  if ((st = host_begin(idx, lease.ws, q, nq, qdim, k, &hc)) != CM_OK) return st;
  tl_used_tc = false;
  st = search_exact_dev_impl(idx, lease.ws, hc.q_dev, nq, qdim, k, hc.ids_dev, hc.a_dev, hc.s, approx);
  if (st != CM_OK) return st;
  size_t ob = (size_t)nq * k * 4;
  if (ob) {
    CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
    CM_CUDA(cudaMemcpyAsync(dist, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
  }
  CM_CUDA(cudaStreamSynchronize(hc.s));
  if (tl_used_tc) {
    fetch_scan_counters(lease.ws);
    if (tl_scan_bounded && tl_scan_worst_chunk > kFlagCap) {
      // More queries needed the fp32 fallback than the in-stream path absorbs: redo the batch on the fp32 engine.
      PreparedQueries pq;
      pq.qn = lease.ws->qn.as<float>();
      pq.ld = idx->dev.dim_pad;
      st = exact_topk_fp32(idx, lease.ws, pq.qn, pq.ld, false, nq, k, hc.ids_dev, hc.a_dev, nullptr, nullptr, hc.s);
      if (st != CM_OK) return st;
      CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
      CM_CUDA(cudaMemcpyAsync(dist, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
      CM_CUDA(cudaStreamSynchronize(hc.s));
    }
  }
  return CM_OK;
}

cm_status cm_search_exact(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t k, uint32_t* ids,
                          float* dist) {
  return search_exact_host(idx, q, nq, qdim, k, ids, dist, false);
}

// One batch of the predecessor kNN behind the GPU-assisted construction: rows [q0, q0 + nb) of the resident index
// against the rows BEFORE each of them (tensor-core scan in predecessor mode, bf16-ranked top-k, no exact rescore:
// the host re-measures every candidate). ids_dev/dist_dev: [nb][k] device scratch; flagged (overflowed) queries come
// back as CM_NO_ID rows and are redone by the caller.
static cm_status pred_knn_batch(const cm_index* idx, Workspace* ws, uint32_t q0, uint32_t nb, uint32_t k, uint32_t* ids_dev,
                                float* dist_dev, cudaStream_t s) {
  const DeviceIndex& d = idx->dev;
  const uint32_t nq_pad = scan_tc_pad_queries(nb);
  const uint32_t cap = scan_tc_cand_cap(k);
  CM_CUDA(ws->tc_state.ensure((size_t)nq_pad * 8));
  CM_CUDA(ws->cand.ensure((size_t)nq_pad * cap * 8));
  CM_CUDA(ws->small.ensure(kSmallBytes));
  CM_CUDA(ws->qflag.ensure((size_t)nq_pad * 4));
  uint32_t* tau_key = ws->tc_state.as<uint32_t>();
  uint32_t* cand_cnt = tau_key + nq_pad;
  uint8_t* small = ws->small.as<uint8_t>();
  CM_CUDA(cudaMemsetAsync(ws->tc_state.p, 0, (size_t)nq_pad * 8, s));
  CM_CUDA(cudaMemsetAsync(small + 64, 0, 32, s));
  ScanTcLaunch sl;
  sl.q_bf16 = static_cast<const uint8_t*>(d.xbf16) + (size_t)q0 * d.k_pad * 2;  // the corpus rows ARE the queries
  sl.nq = nb;
  sl.k = k;
  sl.tau_key = tau_key;
  sl.cand_cnt = cand_cnt;
  sl.cand = ws->cand.as<unsigned long long>();
  sl.cap = cap;
  sl.pred = true;
  sl.q_base = q0;
  {
    ProfileScope prof(s);
    CM_CUDA(launch_scan_tc(d, sl, s));
  }
  RescoreLaunch rl;
  rl.qn = nullptr;  // approx mode never reads the fp32 queries
  rl.ldq = d.dim_pad;
  rl.nq = nb;
  rl.k = k;
  rl.tau_key = tau_key;
  rl.cand_cnt = cand_cnt;
  rl.cand = sl.cand;
  rl.cap = cap;
  rl.ids = ids_dev;
  rl.dist = dist_dev;
  rl.rescored_total = reinterpret_cast<unsigned long long*>(small + 64);
  rl.flag_count = reinterpret_cast<uint32_t*>(small + 72);
  rl.flag_list = ws->qflag.as<uint32_t>();
  rl.flag_cap = nq_pad;
  rl.approx = true;
  rl.pred = true;
  rl.pred_base = q0;
  CM_CUDA(launch_rescore(d, rl, s));
  return CM_OK;
}

// GPU-assisted graph construction (host half in hnsw_build.cpp). `LockFreeHnsw::insert` links node
// i to a diverse pick of its nearest ALREADY INSERTED nodes (src/index/lockfree_hnsw.rs:419-432) — that is what makes
// the graph navigable: early nodes link far and wide because nothing nearer exists yet. Per layer, the candidates of
// every member therefore come from a PREDECESSOR kNN on the device: member j against members 0..j-1 only (tensor-core
// scan of the layer's members against themselves, lower triangle, bf16-ranked, in batches of queries) — half the
// work of an all-pairs join, and on clustered data the difference between a searchable graph and islands. The host
// then selects and links like the reference (hnsw_build.cpp).
// RAII for the build's temporaries on the device.
struct DevTmp {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  ~DevTmp() {
    if (p) cudaFree(p);
  }
  uint32_t* u32() const { return reinterpret_cast<uint32_t*>(p); }
};

// One layer of the GPU-assisted construction, entirely on the device once the members' rows are resident in `tmp`:
// predecessor kNN -> each member's own list (select_diverse, phase A) -> reverse-edge lists by target (phase B). Only
// the reverse-edge CSR is put together on the host in between (a counting sort over nm * M ids).
static cm_status build_layer_on_device(cm_index* tmp, uint32_t k, uint32_t M, uint32_t cap, LayerLists* ll,
                                       std::vector<uint32_t>* redo_rows, bool timing) {
  const uint64_t nm = tmp->n;
  const DeviceIndex& d = tmp->dev;
  const uint32_t dim = tmp->dim;
  auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now_s();
  auto lap = [&](const char* what) {
    if (timing && nm > 100000) std::fprintf(stderr, "[build_graph_gpu]   %-26s %.3f s\n", what, now_s() - t0);
    t0 = now_s();
  };
  WorkspaceLease lease(tmp, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  Workspace* ws = lease.ws;
  cudaStream_t s = lease.s;
  DevTmp cand, own, own_deg, lists, deg, in_start_d, incoming_d, misc;
  CM_CUDA(cand.alloc(nm * k * 4));
  CM_CUDA(misc.alloc((2 + 4096) * 4));  // [0] work counter, [1..] too_big (count + ids)
  const uint32_t batch = 16384;          // a multiple of the tile height: batches start on tile boundaries
  CM_CUDA(ws->out_a.ensure((size_t)batch * k * 4));
  for (uint64_t q0 = 0; q0 < nm; q0 += batch) {
    const uint32_t nb = (uint32_t)std::min<uint64_t>(batch, nm - q0);
    cm_status st = pred_knn_batch(tmp, ws, (uint32_t)q0, nb, k, cand.u32() + q0 * k, ws->out_a.as<float>(), s);
    if (st != CM_OK) return st;
  }
  // members whose candidate buffer overflowed came back empty: exact predecessor scan on the host for those (rare)
  std::vector<uint32_t> first(nm);
  CM_CUDA(cudaMemcpy2DAsync(first.data(), 4, cand.p, (size_t)k * 4, 4, nm, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaStreamSynchronize(s));
  lap("predecessor kNN");
  redo_rows->clear();
  for (uint64_t i = 1; i < nm; ++i)
    if (first[i] == CM_NO_ID) redo_rows->push_back((uint32_t)i);
  if (!redo_rows->empty()) {
    const float* X = tmp->x.data();
    std::vector<uint32_t> fixed(redo_rows->size() * (size_t)k, CM_NO_ID);
    parallel_rows(redo_rows->size(), [&](uint64_t r) {
      const uint32_t i = (*redo_rows)[r];
      std::vector<std::pair<float, uint32_t>> best;
      best.reserve(i);
      for (uint32_t p = 0; p < i; ++p) best.emplace_back(dist_prepared_host(X + (uint64_t)p * dim, X + (uint64_t)i * dim, dim), p);
      const size_t kk = std::min<size_t>(k, best.size());
      std::partial_sort(best.begin(), best.begin() + (long)kk, best.end());
      for (size_t c = 0; c < kk; ++c) fixed[r * k + c] = best[c].second;
    });
    for (size_t r = 0; r < redo_rows->size(); ++r)
      CM_CUDA(cudaMemcpyAsync(cand.u32() + (uint64_t)(*redo_rows)[r] * k, &fixed[r * k], (size_t)k * 4, cudaMemcpyHostToDevice, s));
    CM_CUDA(cudaStreamSynchronize(s));
  }
  // phase A: every member's own list
  CM_CUDA(own.alloc(nm * M * 4));
  CM_CUDA(own_deg.alloc(nm * 4));
  CM_CUDA(cudaMemsetAsync(misc.p, 0, (2 + 4096) * 4, s));
  SelectDiverseLaunch a;
  a.n_items = (uint32_t)nm;
  a.m = M;
  a.cand = cand.u32();
  a.width = k;
  a.out = own.u32();
  a.out_deg = own_deg.u32();
  a.work_counter = misc.u32();
  a.too_big = misc.u32() + 1;
  a.too_big_cap = 4095;
  CM_CUDA(launch_select_diverse(d, a, s));
  ll->own_width = M;
  ll->own.resize(nm * M);
  ll->own_deg.resize(nm);
  CM_CUDA(cudaMemcpyAsync(ll->own.data(), own.p, nm * M * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaMemcpyAsync(ll->own_deg.data(), own_deg.p, nm * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaStreamSynchronize(s));
  lap("own lists (device select)");
  build_incoming_csr(ll->own.data(), ll->own_deg.data(), nm, M, &ll->in_start, &ll->incoming);
  lap("reverse-edge CSR (host)");
  // phase B: the reverse edges, selected once per target
  CM_CUDA(in_start_d.alloc((nm + 1) * 4));
  CM_CUDA(incoming_d.alloc(ll->incoming.size() * 4));
  CM_CUDA(lists.alloc(nm * cap * 4));
  CM_CUDA(deg.alloc(nm * 4));
  CM_CUDA(cudaMemcpyAsync(in_start_d.p, ll->in_start.data(), (nm + 1) * 4, cudaMemcpyHostToDevice, s));
  if (!ll->incoming.empty())
    CM_CUDA(cudaMemcpyAsync(incoming_d.p, ll->incoming.data(), ll->incoming.size() * 4, cudaMemcpyHostToDevice, s));
  CM_CUDA(cudaMemsetAsync(misc.p, 0, (2 + 4096) * 4, s));
  SelectDiverseLaunch b;
  b.n_items = (uint32_t)nm;
  b.m = cap;
  b.own = own.u32();
  b.own_deg = own_deg.u32();
  b.own_width = M;
  b.in_start = in_start_d.u32();
  b.incoming = incoming_d.u32();
  b.out = lists.u32();
  b.out_deg = deg.u32();
  b.work_counter = misc.u32();
  b.too_big = misc.u32() + 1;
  b.too_big_cap = 4095;
  CM_CUDA(launch_select_diverse(d, b, s));
  ll->cap = cap;
  ll->lists.resize(nm * cap);
  ll->deg.resize(nm);
  std::vector<uint32_t> big(1 + 4095);
  CM_CUDA(cudaMemcpyAsync(ll->lists.data(), lists.p, nm * cap * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaMemcpyAsync(ll->deg.data(), deg.p, nm * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaMemcpyAsync(big.data(), misc.u32() + 1, big.size() * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaStreamSynchronize(s));
  if (big[0] > 4095) return fail(CM_ERR_CAPACITY_EXCEEDED, "more oversized reverse-edge lists than the build tracks");
  ll->too_big.assign(big.begin() + 1, big.begin() + 1 + big[0]);
  std::sort(ll->too_big.begin(), ll->too_big.end());
  lap("reverse edges (device select)");
  return CM_OK;
}

cm_status cm_index_build_graph_gpu(cm_index* idx, uint64_t layer_seed, int device, int threads, uint32_t n_candidates) {
  if (!idx) return fail(CM_ERR_INVALID_ARGUMENT, "null index");
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const uint64_t n = idx->n;
  const uint32_t dim = idx->dim;
  const uint32_t M = (uint32_t)idx->config.index.max_connections;
  uint32_t C = n_candidates ? n_candidates : (uint32_t)std::min<uint64_t>(idx->config.index.ef_construction, 64);
  C = std::min(std::max(C, M + 1), 127u);
  const bool timing = getenv("CM_BUILD_TIMING") != nullptr;  // stage times on stderr (diagnostics)
  const char* sel_env = getenv("CM_BUILD_SELECT");          // "host": select and link on the host (equivalence tests)
  const uint32_t dim_pad = (dim + 7) & ~7u;
  const bool device_select = !(sel_env && std::string(sel_env) == "host") && select_diverse_supported(dim_pad, 2 * M) && M >= 2;
  auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_stage = now_s();
  auto lap = [&](const char* what) {
    if (timing) std::fprintf(stderr, "[build_graph_gpu] %-28s %.3f s\n", what, now_s() - t_stage);
    t_stage = now_s();
  };
  std::vector<std::vector<uint32_t>> members;
  layer_members(n, idx->config.index, layer_seed, &members);
  lap("layer draw");
  std::vector<LayerCandidates> layers(members.size());
  std::vector<LayerLists> lists(members.size());
  for (size_t layer = 0; layer < members.size(); ++layer) {
    LayerCandidates& lc = layers[layer];
    lc.members.swap(members[layer]);
    lists[layer].members = lc.members;
    const uint64_t nm = lc.members.size();
    if (nm <= 1) continue;
    const uint32_t k = (uint32_t)std::min<uint64_t>(C, nm - 1);
    lc.width = k;
    cm_index* tmp = new_index(idx->config, nm, dim);
    tmp->all_finite = idx->all_finite;
    if (layer == 0) {
      tmp->x.swap(idx->x);  // layer 0 == every row: borrow the matrix instead of copying it
    } else {
      tmp->x.resize(nm * (uint64_t)dim);
      for (uint64_t i = 0; i < nm; ++i)
        std::memcpy(&tmp->x[i * dim], &idx->x[(uint64_t)lc.members[i] * dim], (size_t)dim * 4);
    }
    cm_status st = cm_index_upload(tmp, device);
    if (layer == 0) lap("layer 0 upload");
    std::vector<uint32_t> redo;  // members whose candidate buffer overflowed on the device
    if (st == CM_OK && !(tmp->all_finite && scan_tc_supported(tmp->dev, k)))
      st = fail(CM_ERR_STATE, "the GPU-assisted build needs the tensor-core scan (finite rows, at most 127 candidates)");
    if (st == CM_OK && device_select) {
      st = build_layer_on_device(tmp, k, M, layer == 0 ? 2 * M : M, &lists[layer], &redo, timing);
      if (layer == 0) lap("layer 0 kNN + select + link");
    } else if (st == CM_OK) {
      lc.ids.assign(nm * k, CM_NO_ID);
      WorkspaceLease lease(tmp, nullptr, true);
      if (lease.err != cudaSuccess) st = cuda_fail(lease.err, "workspace lease");
      const uint32_t batch = 16384;
      Workspace* ws = lease.ws;
      cudaError_t ce = cudaSuccess;
      if (st == CM_OK && ((ce = ws->out_ids.ensure((size_t)batch * k * 4)) != cudaSuccess || (ce = ws->out_a.ensure((size_t)batch * k * 4)) != cudaSuccess))
        st = cuda_fail(ce, "scratch");
      for (uint64_t q0 = 0; st == CM_OK && q0 < nm; q0 += batch) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(batch, nm - q0);
        st = pred_knn_batch(tmp, ws, (uint32_t)q0, nb, k, ws->out_ids.as<uint32_t>(), ws->out_a.as<float>(), lease.s);
        if (st != CM_OK) break;
        ce = cudaMemcpyAsync(&lc.ids[q0 * k], ws->out_ids.p, (size_t)nb * k * 4, cudaMemcpyDeviceToHost, lease.s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(lease.s);
        if (ce != cudaSuccess) st = cuda_fail(ce, "predecessor kNN batch");
      }
      if (st == CM_OK)
        for (uint64_t i = 1; i < nm; ++i)  // member 0 has no predecessor; anything else without candidates was flagged
          if (lc.ids[i * k] == CM_NO_ID) redo.push_back((uint32_t)i);
      if (layer == 0) lap("layer 0 predecessor kNN");
      if (st == CM_OK && !redo.empty()) {  // rare: exact predecessor scan on the host for the flagged members
        const float* X = tmp->x.data();
        parallel_rows(redo.size(), [&](uint64_t r) {
          const uint32_t i = redo[r];
          std::vector<std::pair<float, uint32_t>> best;
          best.reserve(i);
          for (uint32_t p = 0; p < i; ++p) best.emplace_back(dist_prepared_host(X + (uint64_t)p * dim, X + (uint64_t)i * dim, dim), p);
          const size_t kk = std::min<size_t>(k, best.size());
          std::partial_sort(best.begin(), best.begin() + (long)kk, best.end());
          for (size_t c = 0; c < kk; ++c) lc.ids[(uint64_t)i * k + c] = best[c].second;
        });
      }
    }
    if (layer == 0) tmp->x.swap(idx->x);
    cm_index_free(tmp);
    if (st != CM_OK) return st;
    if (!device_select && layer > 0)  // subset-local -> global handles
      for (uint32_t& v : lc.ids)
        if (v != CM_NO_ID) v = lc.members[v];
  }
  lap("upper layers (all)");
  if (device_select)
    build_graph_from_layer_lists(idx->x.data(), n, dim, idx->config.index, layer_seed, threads, lists, &idx->graph);
  else
    build_graph_from_candidates(idx->x.data(), n, dim, idx->config.index, layer_seed, threads, layers, &idx->graph);
  lap(device_select ? "write lists + reachability" : "host select + link");
  return CM_OK;
}

// consolidate's all-pairs pass (src/store.rs:516-568, the `similarity > similarity_threshold` test of every pair)
// as a threshold self-join: the tensor-core scan with the corpus as its own query set emits the pairs whose bf16
// score can exceed the threshold, the exact `metric.similarity` on the raw rows decides. Pairs (i < j, both live),
// ascending (i, j). *count receives the number found; at most `cap` are written.
// The host half of the GPU-assisted construction on caller-supplied candidates (any kNN source): layer `l` holds
// `widths[l]` candidate handles for each member of that layer, members in ascending handle order (the layer draw of
// (n, params, seed): cm_index_layer_members). Used by cm_index_build_graph_gpu's tests without a device.
cm_status cm_index_build_graph_from_candidates(cm_index* idx, uint64_t layer_seed, int threads, uint32_t n_layers,
                                               const uint32_t* const* cand_ids, const uint32_t* widths) {
  if (!idx || (n_layers && (!cand_ids || !widths))) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::vector<uint32_t>> members;
  layer_members(idx->n, idx->config.index, layer_seed, &members);
  if (members.size() != n_layers) return fail(CM_ERR_INVALID_ARGUMENT, "layer count does not match the layer draw of this seed");
  std::vector<LayerCandidates> layers(members.size());
  for (size_t l = 0; l < members.size(); ++l) {
    layers[l].members.swap(members[l]);
    layers[l].width = widths[l];
    const size_t cnt = layers[l].members.size() * (size_t)widths[l];
    if (cnt && !cand_ids[l]) return fail(CM_ERR_INVALID_ARGUMENT, "null candidate array");
    layers[l].ids.assign(cand_ids[l], cand_ids[l] + cnt);
    for (uint32_t v : layers[l].ids)
      if (v != CM_NO_ID && v >= idx->n) return fail(CM_ERR_INVALID_ARGUMENT, "candidate handle out of range");
  }
  build_graph_from_candidates(idx->x.data(), idx->n, idx->dim, idx->config.index, layer_seed, threads, layers, &idx->graph);
  return CM_OK;
}

// Members (ascending handles) of layer `layer` under the layer draw of (slots, params, seed); returns the count and
// writes at most `cap` handles. *n_layers (optional) receives the number of layers.
uint64_t cm_index_layer_members(const cm_index* idx, uint64_t layer_seed, uint32_t layer, uint32_t* out, uint64_t cap,
                                uint32_t* n_layers) {
  if (!idx) return 0;
  std::vector<std::vector<uint32_t>> members;
  layer_members(idx->n, idx->config.index, layer_seed, &members);
  if (n_layers) *n_layers = (uint32_t)members.size();
  if (layer >= members.size()) return 0;
  for (uint64_t i = 0; i < members[layer].size() && i < cap; ++i) out[i] = members[layer][i];
  return members[layer].size();
}

cm_status cm_similar_pairs(const cm_index* idx, float threshold, uint32_t* out_i, uint32_t* out_j, float* out_sim,
                           uint64_t cap, uint64_t* count) {
  if (!idx || !count) return fail(CM_ERR_INVALID_ARGUMENT, "null argument");
  *count = 0;
  if (!idx->uploaded) return fail(CM_ERR_CUDA, "index is not resident on a CUDA device (call cm_index_upload); there is no CPU path");
  const DeviceIndex& d = idx->dev;
  if (d.n == 0) return CM_OK;
  if (!d.xraw) return fail(CM_ERR_STATE, "the pair test runs on the rows as stored: cm_index_set_contexts / a snapshot, then upload");
  if (!idx->all_finite || !scan_tc_supported(d, 1)) return fail(CM_ERR_STATE, "the self-join needs the tensor-core scan (finite rows)");
  cm_status st = set_device(idx);
  if (st != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  Workspace* ws = lease.ws;
  cudaStream_t s = lease.s;
  CM_CUDA(ws->small.ensure(kSmallBytes));
  uint32_t* pair_count = ws->small.as<uint32_t>() + 32;  // offset 128
  uint64_t pcap = std::max<uint64_t>(1u << 20, 4 * cap);
  uint32_t found = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CM_CUDA(ws->cand.ensure(pcap * 8));
    CM_CUDA(cudaMemsetAsync(pair_count, 0, 4, s));
    ScanTcLaunch sl;
    sl.q_bf16 = d.xbf16;
    sl.nq = (uint32_t)d.n;
    sl.k = 1;
    sl.tau_key = nullptr;
    sl.cand_cnt = nullptr;
    sl.cand = nullptr;
    sl.cap = 16;
    sl.join = true;
    sl.join_tau = threshold - scan_tc_eps(d.dim) - 1e-5f;   // bf16 score >= cos - eps; cos of the unit rows ~ similarity
    sl.pairs = ws->cand.as<unsigned long long>();
    sl.pair_count = pair_count;
    sl.pair_cap = (uint32_t)std::min<uint64_t>(pcap, 0xFFFFFFFFu);
    {
      ProfileScope prof(s);
      CM_CUDA(launch_scan_tc(d, sl, s));
    }
    CM_CUDA(cudaMemcpyAsync(&found, pair_count, 4, cudaMemcpyDeviceToHost, s));
    CM_CUDA(cudaStreamSynchronize(s));
    if (found <= pcap) break;
    if (attempt == 1) return fail(CM_ERR_CAPACITY_EXCEEDED, "more candidate pairs than the join buffer holds");
    pcap = found;  // the count is exact: retry once with room for all of them
  }
  tl_counters.dist_evals += d.n * (d.n - 1) / 2;
  tl_counters.rescored += found;
  if (found == 0) return CM_OK;
  CM_CUDA(ws->D.ensure((size_t)found * 4));
  CM_CUDA(launch_pair_similarity(d.xraw, d.dim_pad, d.dim, d.norm2_raw, ws->cand.as<unsigned long long>(), found,
                                 ws->D.as<float>(), s));
  std::vector<unsigned long long> pairs(found);
  std::vector<float> sims(found);
  CM_CUDA(cudaMemcpyAsync(pairs.data(), ws->cand.p, (size_t)found * 8, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaMemcpyAsync(sims.data(), ws->D.p, (size_t)found * 4, cudaMemcpyDeviceToHost, s));
  CM_CUDA(cudaStreamSynchronize(s));
  std::vector<std::pair<unsigned long long, float>> keep;
  for (uint32_t i = 0; i < found; ++i)
    if (sims[i] > threshold) keep.emplace_back(pairs[i], sims[i]);   // `similarity <= threshold => continue`
  std::sort(keep.begin(), keep.end());
  *count = keep.size();
  for (uint64_t i = 0; i < keep.size() && i < cap; ++i) {
    if (out_i) out_i[i] = (uint32_t)(keep[i].first >> 32);
    if (out_j) out_j[i] = (uint32_t)keep[i].first;
    if (out_sim) out_sim[i] = keep[i].second;
  }
  return CM_OK;
}

cm_status cm_decay_sweep_dev(void* importance, void* decayed_through_nanos, const void* last_access_nanos,
                             const void* decay_rate, uint64_t n, float base_decay_rate, uint64_t now_nanos, int device,
                             void* stream) {
  if (n && (!importance || !decayed_through_nanos || !last_access_nanos || !decay_rate))
    return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  CM_CUDA(cudaSetDevice(device));
  CM_CUDA(launch_decay_sweep((float*)importance, (unsigned long long*)decayed_through_nanos,
                             (const unsigned long long*)last_access_nanos, (const float*)decay_rate, n, base_decay_rate,
                             now_nanos, (cudaStream_t)stream));
  return CM_OK;
}

cm_status cm_search_hnsw(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t ef, uint32_t k,
                         uint32_t* ids, float* dist) {
  cm_status st = check_search_args(idx, q, nq, k, ids, dist);
  if (st != CM_OK) return st;
  if (ef > CM_MAX_K) return fail(CM_ERR_INVALID_ARGUMENT, "ef exceeds CM_MAX_K");
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  HostCall hc;
  if ((st = host_begin(idx, lease.ws, q, nq, qdim, k, &hc)) != CM_OK) return st;
  if (ef == 0) {  // search(.., 0) == [] (:470-472)
    for (size_t i = 0; i < (size_t)nq * k; ++i) {
      ids[i] = CM_NO_ID;
      dist[i] = INFINITY;
    }
    return CM_OK;
  }
  st = search_hnsw_dev_impl(idx, lease.ws, hc.q_dev, nq, qdim, ef, k, hc.ids_dev, hc.a_dev, hc.s);
  if (st != CM_OK) return st;
  size_t ob = (size_t)nq * k * 4;
  if (ob) {
    CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
    CM_CUDA(cudaMemcpyAsync(dist, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
  }
  CM_CUDA(cudaStreamSynchronize(hc.s));
  fetch_hnsw_counters(lease.ws);
  return hnsw_failures_to_status();
}

cm_status cm_search_temporal(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                             uint32_t ef_search, float w, float base, uint64_t now_secs, uint32_t now_nanos, int exact,
                             uint32_t* ids, float* score, float* dist) {
  cm_status st = check_search_args(idx, q, nq, k, ids, score);
  if (st != CM_OK) return st;
  // validate_query (src/store.rs:379-392): dimensions first, then finiteness (checked on the device below).
  if (qdim != idx->dim) {
    char buf[96];
    std::snprintf(buf, sizeof buf, "invalid dimensions: got %u, expected %u", qdim, idx->dim);
    return fail(CM_ERR_INVALID_DIMENSIONS, buf);
  }
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  HostCall hc;
  if ((st = host_begin(idx, lease.ws, q, nq, qdim, k, &hc)) != CM_OK) return st;
  CM_CUDA(lease.ws->small.ensure(kSmallBytes));
  uint32_t* bad_row = lease.ws->small.as<uint32_t>() + 48;  // offset 192
  CM_CUDA(cudaMemsetAsync(bad_row, 0xFF, 4, hc.s));
  tl_used_tc = false;
  st = search_temporal_dev_impl(idx, lease.ws, hc.q_dev, nq, qdim, k, ef_search, w, base, now_secs, now_nanos, exact,
                                bad_row, hc.ids_dev, hc.a_dev, hc.b_dev, hc.s);
  if (st != CM_OK) return st;
  uint32_t bad = 0xFFFFFFFFu;
  CM_CUDA(cudaMemcpyAsync(&bad, bad_row, 4, cudaMemcpyDeviceToHost, hc.s));
  size_t ob = (size_t)nq * k * 4;
  if (ob) {
    CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
    CM_CUDA(cudaMemcpyAsync(score, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
    if (dist) CM_CUDA(cudaMemcpyAsync(dist, hc.b_dev, ob, cudaMemcpyDeviceToHost, hc.s));
  }
  CM_CUDA(cudaStreamSynchronize(hc.s));
  if (!exact) fetch_hnsw_counters(lease.ws);
  if (bad != 0xFFFFFFFFu)
    return fail(CM_ERR_INVALID_VECTOR,
                "invalid vector data: query contains NaN or infinite components (row " + std::to_string(bad) + ")");
  if (exact && tl_used_tc) {
    fetch_scan_counters(lease.ws);
    if (tl_scan_bounded && tl_scan_worst_chunk > kFlagCap) {
      // More pool queries needed the fp32 fallback than the in-stream path absorbs: redo the call on the fp32 engine.
      tl_counters = cm_counters();
      tl_force_fp32 = true;
      st = search_temporal_dev_impl(idx, lease.ws, hc.q_dev, nq, qdim, k, ef_search, w, base, now_secs, now_nanos, exact,
                                    nullptr, hc.ids_dev, hc.a_dev, hc.b_dev, hc.s);
      tl_force_fp32 = false;
      if (st != CM_OK) return st;
      if (ob) {
        CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
        CM_CUDA(cudaMemcpyAsync(score, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
        if (dist) CM_CUDA(cudaMemcpyAsync(dist, hc.b_dev, ob, cudaMemcpyDeviceToHost, hc.s));
      }
      CM_CUDA(cudaStreamSynchronize(hc.s));
    }
    return CM_OK;
  }
  return exact ? CM_OK : hnsw_failures_to_status();
}

cm_status cm_search_in_context(const cm_index* idx, const float* q, uint32_t nq, uint32_t qdim,
                               const uint32_t* context_ids, uint32_t k, float w, float base, uint64_t now_secs,
                               uint32_t now_nanos, uint32_t* ids, float* score) {
  cm_status st = check_search_args(idx, q, nq, k, ids, score);
  if (st != CM_OK) return st;
  if (nq && !context_ids) return fail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  if (qdim != idx->dim) {  // validate_query (src/store.rs:379-392): dimensions first
    char buf[96];
    std::snprintf(buf, sizeof buf, "invalid dimensions: got %u, expected %u", qdim, idx->dim);
    return fail(CM_ERR_INVALID_DIMENSIONS, buf);
  }
  if ((st = set_device(idx)) != CM_OK) return st;
  tl_counters = cm_counters();
  WorkspaceLease lease(idx, nullptr, true);
  if (lease.err != cudaSuccess) return cuda_fail(lease.err, "workspace lease");
  HostCall hc;
  if ((st = host_begin(idx, lease.ws, q, nq, qdim, k, &hc)) != CM_OK) return st;
  CM_CUDA(lease.ws->small.ensure(kSmallBytes));
  CM_CUDA(lease.ws->qctx.ensure(std::max<size_t>((size_t)nq * 4, 16)));
  if (nq) CM_CUDA(cudaMemcpyAsync(lease.ws->qctx.p, context_ids, (size_t)nq * 4, cudaMemcpyHostToDevice, hc.s));
  uint32_t* bad_row = lease.ws->small.as<uint32_t>() + 48;  // offset 192
  CM_CUDA(cudaMemsetAsync(bad_row, 0xFF, 4, hc.s));
  tl_used_tc = false;
  st = search_in_context_dev_impl(idx, lease.ws, hc.q_dev, nq, lease.ws->qctx.as<uint32_t>(), k, w, base, now_secs,
                                  now_nanos, bad_row, hc.ids_dev, hc.a_dev, hc.s);
  if (st != CM_OK) return st;
  uint32_t bad = 0xFFFFFFFFu;
  CM_CUDA(cudaMemcpyAsync(&bad, bad_row, 4, cudaMemcpyDeviceToHost, hc.s));
  size_t ob = (size_t)nq * k * 4;
  if (ob) {
    CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
    CM_CUDA(cudaMemcpyAsync(score, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
  }
  CM_CUDA(cudaStreamSynchronize(hc.s));
  if (bad != 0xFFFFFFFFu)
    return fail(CM_ERR_INVALID_VECTOR,
                "invalid vector data: query contains NaN or infinite components (row " + std::to_string(bad) + ")");
  if (tl_used_tc) {
    fetch_scan_counters(lease.ws);
    if (tl_counters.fallback_queries > 0) {
      // Some query's candidate buffer overflowed on the tensor-core path: redo the call on the fp32 context scan.
      const uint64_t flagged = tl_counters.fallback_queries;
      tl_counters = cm_counters();
      tl_counters.fallback_queries = flagged;
      tl_force_fp32 = true;
      st = search_in_context_dev_impl(idx, lease.ws, hc.q_dev, nq, lease.ws->qctx.as<uint32_t>(), k, w, base, now_secs,
                                      now_nanos, nullptr, hc.ids_dev, hc.a_dev, hc.s);
      tl_force_fp32 = false;
      if (st != CM_OK) return st;
      if (ob) {
        CM_CUDA(cudaMemcpyAsync(ids, hc.ids_dev, ob, cudaMemcpyDeviceToHost, hc.s));
        CM_CUDA(cudaMemcpyAsync(score, hc.a_dev, ob, cudaMemcpyDeviceToHost, hc.s));
      }
      CM_CUDA(cudaStreamSynchronize(hc.s));
    }
  }
  return CM_OK;
}

}  // extern "C"
