// group.cu — the sharded read path behind the C ABI: ONE object, ONE call, G GPUs of one process.
//
// Replaces `ShardedRwLockHnsw` (src/index/sharded_rwlock.rs:43-94): `new` makes the shards, `insert` deals rows out
// round-robin (:70-74), `search` scatters the query to every shard, concatenates the per-shard lists, sorts by
// (distance, id) and truncates (:81-94). Here a shard is a cm_index resident on its own GPU and the serial loop over
// shards becomes:
//   1. every device copies 1/G of the query batch over ITS PCIe link and the slices are all-gathered over NVLink
//      (ncclAllGather) — the node's aggregate host bandwidth instead of G full copies;
//   2. every device searches its shard for the whole batch (cm_search_*_dev on its own stream and host thread);
//   3. the per-shard lists, packed to one u64 per entry, are all-gathered (ncclAllGather) and merged on the device
//      by (distance, global handle) — `global = local * G + shard`, the generalisation of `join_handle` (:54-58);
//   4. store-level search: the temporal rerank runs on the MERGED pool (merge-then-rerank keeps
//      `ChronoMind::search` semantics, src/store.rs:323-344), attributes indexed by global handle.
// NCCL is loaded with dlopen at cm_group_create: an engine that never shards has no NCCL dependency, and a missing
// or failing NCCL surfaces as CM_ERR_NCCL, not as a load error of the whole library.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: every entry point is resolved at run time

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <thread>

#include "cm_internal.h"

namespace cm {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool load(std::string* err) {
    if (handle) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      *err = std::string("cannot load NCCL: ") + dlerror();
      return false;
    }
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllGather = (decltype(AllGather))sym("ncclAllGather");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
    if (!CommInitAll || !CommDestroy || !AllGather || !GetErrorString) {
      *err = "NCCL library lacks ncclCommInitAll / ncclAllGather";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

// One host thread per device (the CUDA and NCCL calls of a device always come from the same thread).
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<cm_status(std::string*)> job;
  bool has_job = false, done = false, stop = false;
  cm_status result = CM_OK;
  std::string err;
  void loop() {
    for (;;) {
      std::function<cm_status(std::string*)> j;
      {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return has_job || stop; });
        if (stop) return;
        j = job;
        has_job = false;
      }
      std::string e;
      cm_status r = j(&e);
      {
        std::lock_guard<std::mutex> l(mu);
        result = r;
        err = e;
        done = true;
      }
      cv.notify_all();
    }
  }
};

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t grow(size_t bytes) {
    bytes = std::max<size_t>(bytes, 16);
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct DeviceSide {
  int device = -1;
  cm_index* shard = nullptr;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  // scratch, grown on demand: query slice / gathered queries, this shard's lists, packed keys, gathered keys
  Buf q_slice, q_all, ids, dist, keys, keys_all;
  // merged pool and reranked result (first device only)
  Buf m_ids, m_dist, r_ids, r_score, r_dist;
};

}  // namespace
}  // namespace cm

using namespace cm;

struct cm_group {
  int G = 0;
  uint64_t n = 0;
  uint32_t dim = 0;
  bool uploaded = false;
  std::vector<DeviceSide> dev;
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex call_mu;  // one batch at a time per group (each device runs one stream)
  // attributes by GLOBAL handle on device 0, for the rerank of the merged pool
  std::vector<uint64_t> ts_secs;
  std::vector<uint32_t> ts_nanos;
  std::vector<float> decay;
  void *d_ts_secs = nullptr, *d_ts_nanos = nullptr, *d_decay = nullptr;
  cm_config config{};
};

namespace {

cm_status gfail(cm_status st, const std::string& msg) {
  cm::set_error(msg);
  return st;
}

#define G_CUDA(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      *err = std::string(#call) + ": " + cudaGetErrorString(e__);                                      \
      return CM_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)
#define G_NCCL(call)                                                                                   \
  do {                                                                                                 \
    ncclResult_t r__ = (call);                                                                         \
    if (r__ != ncclSuccess) {                                                                          \
      *err = std::string(#call) + ": " + g_nccl.GetErrorString(r__);                                   \
      return CM_ERR_NCCL;                                                                              \
    }                                                                                                  \
  } while (0)
#define G_CM(call)                                                                                     \
  do {                                                                                                 \
    cm_status s__ = (call);                                                                            \
    if (s__ != CM_OK) {                                                                                \
      *err = cm_last_error();                                                                          \
      return s__;                                                                                      \
    }                                                                                                  \
  } while (0)

// Run `f(rank, err)` on every device's thread and wait; the first failing rank decides the status.
cm_status run_all(cm_group* g, const std::function<cm_status(int, std::string*)>& f) {
  for (int r = 0; r < g->G; ++r) {
    Worker& w = *g->workers[r];
    std::lock_guard<std::mutex> l(w.mu);
    w.job = [&f, r](std::string* e) { return f(r, e); };
    w.has_job = true;
    w.done = false;
    w.cv.notify_all();
  }
  cm_status st = CM_OK;
  std::string msg;
  for (int r = 0; r < g->G; ++r) {
    Worker& w = *g->workers[r];
    std::unique_lock<std::mutex> l(w.mu);
    w.cv.wait(l, [&] { return w.done; });
    if (st == CM_OK && w.result != CM_OK) {
      st = w.result;
      msg = "shard " + std::to_string(r) + ": " + w.err;
    }
  }
  if (st != CM_OK) cm::set_error(msg);
  return st;
}

enum class Kind { Exact, Hnsw, Temporal };

struct SearchCall {
  Kind kind;
  const float* q;
  uint32_t nq, k, ef, pool;  // pool: entries per shard list that are exchanged (k, or max(ef_search, 3k) for Temporal)
  int exact;
  float w, base;
  uint64_t now_secs;
  uint32_t now_nanos;
  uint32_t* ids;
  float* a;  // dist (Exact/Hnsw) or score (Temporal)
  float* b;  // Temporal: dist (optional)
  bool validate = false;                       // store-level call: reject NaN/Inf queries
  std::atomic<uint32_t>* bad_row = nullptr;   // smallest offending row over all slices (UINT32_MAX: none)
};

cm_status search_rank(cm_group* g, const SearchCall& c, int r, std::string* err) {
  DeviceSide& d = g->dev[r];
  const int G = g->G;
  const uint32_t dim = g->dim;
  G_CUDA(cudaSetDevice(d.device));
  cudaStream_t s = d.stream;
  // 1. this device's slice of the batch over its own PCIe link, then the all-gather over NVLink
  const uint32_t per = (c.nq + G - 1) / G;
  const uint32_t lo = std::min<uint64_t>(c.nq, (uint64_t)r * per), hi = std::min<uint64_t>(c.nq, (uint64_t)lo + per);
  const size_t slice_bytes = (size_t)per * dim * 4;
  G_CUDA(d.q_slice.grow(slice_bytes));
  G_CUDA(d.q_all.grow(slice_bytes * G));
  if (c.validate) {  // validate_query's finiteness check (src/store.rs:386-390) on this device's slice, all slices at once
    for (uint32_t row = lo; row < hi; ++row) {
      const float* v = c.q + (size_t)row * dim;
      float acc = 0.0f;
      for (uint32_t e = 0; e < dim; ++e) acc += v[e] * 0.0f;  // NaN/Inf poison the sum; vectorises
      if (!(acc == 0.0f)) {
        uint32_t seen = c.bad_row->load(std::memory_order_relaxed);
        while (row < seen && !c.bad_row->compare_exchange_weak(seen, row, std::memory_order_relaxed)) {
        }
        break;
      }
    }
  }
  if (hi > lo) G_CUDA(cudaMemcpyAsync(d.q_slice.p, c.q + (size_t)lo * dim, (size_t)(hi - lo) * dim * 4, cudaMemcpyHostToDevice, s));
  if (G > 1) G_NCCL(g_nccl.AllGather(d.q_slice.p, d.q_all.p, (size_t)per * dim, ncclFloat32, d.comm, s));
  const void* q_dev = G > 1 ? d.q_all.p : d.q_slice.p;
  // 2. this shard's search for the whole batch
  const size_t list_bytes = (size_t)c.nq * c.pool * 4;
  G_CUDA(d.ids.grow(list_bytes));
  G_CUDA(d.dist.grow(list_bytes));
  G_CUDA(d.keys.grow(list_bytes * 2));
  switch (c.kind) {
    case Kind::Exact:
      G_CM(cm_search_exact_dev(d.shard, q_dev, c.nq, c.pool, d.ids.p, d.dist.p, s));
      break;
    case Kind::Hnsw:
      G_CM(cm_search_hnsw_dev(d.shard, q_dev, c.nq, c.ef, c.pool, d.ids.p, d.dist.p, s));
      break;
    case Kind::Temporal:  // the shard's candidate pool: exact top-pool or `VectorIndex::search(q, pool)`
      if (c.exact)
        G_CM(cm_search_exact_dev(d.shard, q_dev, c.nq, c.pool, d.ids.p, d.dist.p, s));
      else
        G_CM(cm_search_hnsw_dev(d.shard, q_dev, c.nq, c.pool, c.pool, d.ids.p, d.dist.p, s));
      break;
  }
  // 3. one packed array per shard over NVLink, merged by (distance, global handle)
  const uint64_t entries = (uint64_t)c.nq * c.pool;
  G_CM(cm_pack_topk_dev(d.ids.p, d.dist.p, entries, d.keys.p, d.device, s));
  G_CUDA(d.keys_all.grow(entries * 8 * G));
  if (G > 1)
    G_NCCL(g_nccl.AllGather(d.keys.p, d.keys_all.p, entries, ncclUint64, d.comm, s));
  else
    G_CUDA(cudaMemcpyAsync(d.keys_all.p, d.keys.p, entries * 8, cudaMemcpyDeviceToDevice, s));
  if (r == 0) {  // the caller reads ONE result: merge (and rerank) on the first device
    const uint32_t out_len = c.kind == Kind::Temporal ? c.pool : c.k;
    G_CUDA(d.m_ids.grow((size_t)c.nq * out_len * 4));
    G_CUDA(d.m_dist.grow((size_t)c.nq * out_len * 4));
    G_CM(cm_merge_topk_keys_dev(d.keys_all.p, (uint32_t)G, c.nq, c.pool, out_len, d.m_ids.p, d.m_dist.p, d.device, s));
    const size_t ob = (size_t)c.nq * c.k * 4;
    if (c.kind == Kind::Temporal) {
      G_CUDA(d.r_ids.grow(ob));
      G_CUDA(d.r_score.grow(ob));
      G_CUDA(d.r_dist.grow(ob));
      G_CM(cm_rerank_dev(d.m_ids.p, d.m_dist.p, c.nq, c.pool, c.k, g->d_ts_secs, g->d_ts_nanos, g->d_decay, c.w, c.base,
                         c.now_secs, c.now_nanos, d.r_ids.p, d.r_score.p, d.r_dist.p, d.device, s));
      G_CUDA(cudaMemcpyAsync(c.ids, d.r_ids.p, ob, cudaMemcpyDeviceToHost, s));
      G_CUDA(cudaMemcpyAsync(c.a, d.r_score.p, ob, cudaMemcpyDeviceToHost, s));
      if (c.b) G_CUDA(cudaMemcpyAsync(c.b, d.r_dist.p, ob, cudaMemcpyDeviceToHost, s));
    } else {
      G_CUDA(cudaMemcpyAsync(c.ids, d.m_ids.p, ob, cudaMemcpyDeviceToHost, s));
      G_CUDA(cudaMemcpyAsync(c.a, d.m_dist.p, ob, cudaMemcpyDeviceToHost, s));
    }
  }
  G_CUDA(cudaStreamSynchronize(s));
  // a query a shard could not answer like the reference must not pass silently (the _dev calls cannot say so)
  uint64_t failures = 0;
  G_CM(cm_index_dev_status(d.shard, &failures, 1));
  if (failures) {
    *err = std::to_string(failures) + " quer(ies) could not be answered like the reference on this shard";
    return CM_ERR_CAPACITY_EXCEEDED;
  }
  return CM_OK;
}

cm_status group_search(const cm_group* cg, SearchCall c) {
  cm_group* g = const_cast<cm_group*>(cg);  // scratch buffers are per-call state guarded by call_mu
  if (!g) return gfail(CM_ERR_INVALID_ARGUMENT, "null group");
  if (!g->uploaded) return gfail(CM_ERR_CUDA, "group is not resident on its CUDA devices (call cm_group_upload); there is no CPU path");
  if (c.nq && (!c.q || !c.ids || !c.a)) return gfail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  if (c.k == 0 || c.k > CM_MAX_K || c.pool == 0 || c.pool > CM_MAX_K) return gfail(CM_ERR_INVALID_ARGUMENT, "k/ef must be in [1, CM_MAX_K]");
  if ((uint64_t)g->G * c.pool > 16384) return gfail(CM_ERR_INVALID_ARGUMENT, "shards * list length exceeds the merge kernel's 16384");
  if (c.nq == 0) return CM_OK;
  std::lock_guard<std::mutex> lock(g->call_mu);
  return run_all(g, [g, &c](int r, std::string* err) { return search_rank(g, c, r, err); });
}

}  // namespace

extern "C" {

cm_status cm_group_create(const int* devices, int n_devices, cm_group** out) {
  if (!out) return gfail(CM_ERR_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  if (!devices || n_devices <= 0 || n_devices > 64) return gfail(CM_ERR_INVALID_ARGUMENT, "bad device list");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) return gfail(CM_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
  for (int i = 0; i < n_devices; ++i) {
    if (devices[i] < 0 || devices[i] >= count) return gfail(CM_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return gfail(CM_ERR_INVALID_ARGUMENT, "a device appears twice in the list");
  }
  std::unique_ptr<cm_group> g(new cm_group());
  g->G = n_devices;
  g->dev.resize(n_devices);
  std::vector<ncclComm_t> comms(n_devices, nullptr);
  if (n_devices > 1) {
    std::lock_guard<std::mutex> l(g_nccl_mu);
    std::string err;
    if (!g_nccl.load(&err)) return gfail(CM_ERR_NCCL, err);
    ncclResult_t r = g_nccl.CommInitAll(comms.data(), n_devices, devices);
    if (r != ncclSuccess) return gfail(CM_ERR_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
  }
  for (int i = 0; i < n_devices; ++i) {
    g->dev[i].device = devices[i];
    g->dev[i].comm = comms[i];
    if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&g->dev[i].stream, cudaStreamNonBlocking) != cudaSuccess) {
      cm_group_free(g.release());
      return gfail(CM_ERR_CUDA, "cannot create a stream on device " + std::to_string(devices[i]));
    }
    g->workers.emplace_back(new Worker());
    Worker* w = g->workers.back().get();
    w->th = std::thread([w] { w->loop(); });
  }
  *out = g.release();
  return CM_OK;
}

void cm_group_free(cm_group* g) {
  if (!g) return;
  for (auto& w : g->workers) {
    {
      std::lock_guard<std::mutex> l(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
  }
  for (DeviceSide& d : g->dev) {
    if (d.device < 0) continue;
    cudaSetDevice(d.device);
    if (d.stream) cudaStreamSynchronize(d.stream);
    for (Buf* b : {&d.q_slice, &d.q_all, &d.ids, &d.dist, &d.keys, &d.keys_all, &d.m_ids, &d.m_dist, &d.r_ids, &d.r_score, &d.r_dist})
      b->release();
    if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d.comm);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.shard) cm_index_free(d.shard);
  }
  if (!g->dev.empty() && g->dev[0].device >= 0) {
    cudaSetDevice(g->dev[0].device);
    for (void* p : {g->d_ts_secs, g->d_ts_nanos, g->d_decay})
      if (p) cudaFree(p);
  }
  delete g;
}

uint32_t cm_group_size(const cm_group* g) { return g ? (uint32_t)g->G : 0; }
cm_index* cm_group_shard(const cm_group* g, uint32_t shard) { return (g && shard < (uint32_t)g->G) ? g->dev[shard].shard : nullptr; }
uint64_t cm_group_slots(const cm_group* g) { return g ? g->n : 0; }
uint64_t cm_group_len(const cm_group* g) {
  uint64_t live = 0;
  if (g)
    for (const DeviceSide& d : g->dev) live += d.shard ? cm_index_len(d.shard) : 0;
  return live;
}

cm_status cm_group_from_matrix(cm_group* g, const float* x, uint64_t n, uint32_t dim, const cm_config* cfg) {
  if (!g || (n && !x)) return gfail(CM_ERR_INVALID_ARGUMENT, "null argument");
  if (n > kArenaCapacity * (uint64_t)g->G)
    return gfail(CM_ERR_INDEX_FULL, "index storage exhausted (16777216 slots per shard, including tombstones)");
  const int G = g->G;
  for (DeviceSide& d : g->dev)
    if (d.shard) cm_index_free(d.shard), d.shard = nullptr;
  g->uploaded = false;
  std::vector<float> rows;
  for (int r = 0; r < G; ++r) {  // round-robin deal (:70-74): shard r holds global rows r, r + G, r + 2G, ...
    const uint64_t cnt = n > (uint64_t)r ? (n - r + G - 1) / G : 0;
    rows.resize(std::max<uint64_t>(cnt, 1) * dim);
    for (uint64_t i = 0; i < cnt; ++i) std::memcpy(&rows[i * dim], x + (i * G + r) * (uint64_t)dim, (size_t)dim * 4);
    cm_status st = cm_index_from_matrix(rows.data(), cnt, dim, cfg, &g->dev[r].shard);
    if (st != CM_OK) return st;
    if ((st = cm_index_set_shard(g->dev[r].shard, (uint32_t)r, (uint32_t)G)) != CM_OK) return st;
  }
  g->n = n;
  g->dim = dim;
  cm_index_get_config(g->dev[0].shard, &g->config);
  g->ts_secs.assign(n, 0);
  g->ts_nanos.assign(n, 0);
  g->decay.assign(n, 0.0f);
  return CM_OK;
}

cm_status cm_group_set_attrs(cm_group* g, const uint64_t* ts_secs, const uint32_t* ts_nanos, const float* decay_rate,
                             const uint8_t* deleted) {
  if (!g || !g->dev[0].shard) return gfail(CM_ERR_STATE, "cm_group_from_matrix first");
  const int G = g->G;
  const uint64_t n = g->n;
  std::vector<uint64_t> s;
  std::vector<uint32_t> ns;
  std::vector<float> dr;
  std::vector<uint8_t> del;
  for (int r = 0; r < G; ++r) {
    const uint64_t cnt = n > (uint64_t)r ? (n - r + G - 1) / G : 0;
    s.resize(cnt), ns.resize(cnt), dr.resize(cnt), del.resize(cnt);
    for (uint64_t i = 0; i < cnt; ++i) {
      const uint64_t gi = i * G + r;
      if (ts_secs) s[i] = ts_secs[gi];
      if (ts_nanos) ns[i] = ts_nanos[gi];
      if (decay_rate) dr[i] = decay_rate[gi];
      if (deleted) del[i] = deleted[gi];
    }
    cm_status st = cm_index_set_attrs(g->dev[r].shard, ts_secs ? s.data() : nullptr, ts_nanos ? ns.data() : nullptr,
                                      decay_rate ? dr.data() : nullptr, deleted ? del.data() : nullptr);
    if (st != CM_OK) return st;
  }
  if (ts_secs) g->ts_secs.assign(ts_secs, ts_secs + n);
  if (ts_nanos) g->ts_nanos.assign(ts_nanos, ts_nanos + n);
  if (decay_rate) g->decay.assign(decay_rate, decay_rate + n);
  return CM_OK;
}

cm_status cm_group_build_graphs(cm_group* g, uint64_t layer_seed, int threads, int gpu_assisted) {
  if (!g || !g->dev[0].shard) return gfail(CM_ERR_STATE, "cm_group_from_matrix first");
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int per = std::max(1, threads / g->G);
  // one sub-graph per shard, built independently like the 16 RwLockHnsw shards (:43-52), all shards at once
  return run_all(g, [g, layer_seed, per, gpu_assisted](int r, std::string* err) -> cm_status {
    DeviceSide& d = g->dev[r];
    if (gpu_assisted)
      G_CM(cm_index_build_graph_gpu(d.shard, layer_seed + (uint64_t)r, d.device, per, 0));
    else
      G_CM(cm_index_build_graph(d.shard, layer_seed + (uint64_t)r, per));
    return CM_OK;
  });
}

cm_status cm_group_upload(cm_group* g) {
  if (!g || !g->dev[0].shard) return gfail(CM_ERR_STATE, "cm_group_from_matrix first");
  cm_status st = run_all(g, [g](int r, std::string* err) -> cm_status {
    G_CM(cm_index_upload(g->dev[r].shard, g->dev[r].device));
    return CM_OK;
  });
  if (st != CM_OK) return st;
  // global attribute arrays for the rerank of the merged pool, on the merging device
  if (cudaSetDevice(g->dev[0].device) != cudaSuccess) return gfail(CM_ERR_CUDA, "cudaSetDevice");
  for (void** p : {&g->d_ts_secs, &g->d_ts_nanos, &g->d_decay})
    if (*p) cudaFree(*p), *p = nullptr;
  const size_t n1 = std::max<uint64_t>(g->n, 1);
  if (cudaMalloc(&g->d_ts_secs, n1 * 8) != cudaSuccess || cudaMalloc(&g->d_ts_nanos, n1 * 4) != cudaSuccess ||
      cudaMalloc(&g->d_decay, n1 * 4) != cudaSuccess)
    return gfail(CM_ERR_CUDA, "out of device memory for the attribute arrays");
  if (g->n) {
    cudaMemcpy(g->d_ts_secs, g->ts_secs.data(), g->n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_ts_nanos, g->ts_nanos.data(), g->n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_decay, g->decay.data(), g->n * 4, cudaMemcpyHostToDevice);
  }
  g->uploaded = true;
  return CM_OK;
}

cm_status cm_group_search_exact(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t k, uint32_t* ids,
                                float* dist) {
  if (g && qdim != g->dim) return gfail(CM_ERR_INVALID_DIMENSIONS, "invalid dimensions: got " + std::to_string(qdim) + ", expected " + std::to_string(g->dim));
  SearchCall c{Kind::Exact, q, nq, k, 0, k, 1, 0.f, 0.f, 0, 0, ids, dist, nullptr};
  return group_search(g, c);
}

cm_status cm_group_search_hnsw(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t ef, uint32_t k,
                               uint32_t* ids, float* dist) {
  if (g && qdim != g->dim) return gfail(CM_ERR_INVALID_DIMENSIONS, "invalid dimensions: got " + std::to_string(qdim) + ", expected " + std::to_string(g->dim));
  if (ef == 0 || ef > CM_MAX_K) return gfail(CM_ERR_INVALID_ARGUMENT, "ef must be in [1, CM_MAX_K]");
  SearchCall c{Kind::Hnsw, q, nq, k, ef, k, 0, 0.f, 0.f, 0, 0, ids, dist, nullptr};
  return group_search(g, c);
}

cm_status cm_group_search_temporal(const cm_group* g, const float* q, uint32_t nq, uint32_t qdim, uint32_t k,
                                   uint32_t ef_search, float w, float base, uint64_t now_secs, uint32_t now_nanos, int exact,
                                   uint32_t* ids, float* score, float* dist) {
  if (!g) return gfail(CM_ERR_INVALID_ARGUMENT, "null group");
  // validate_query (src/store.rs:379-392): dimensions first, then finiteness — the first bad row decides
  if (qdim != g->dim) return gfail(CM_ERR_INVALID_DIMENSIONS, "invalid dimensions: got " + std::to_string(qdim) + ", expected " + std::to_string(g->dim));
  if (nq && !q) return gfail(CM_ERR_INVALID_ARGUMENT, "null buffer");
  const uint64_t pool = std::max<uint64_t>(ef_search, (uint64_t)k * 3);  // OVERSAMPLE (src/store.rs:40,323)
  if (pool > CM_MAX_K) return gfail(CM_ERR_INVALID_ARGUMENT, "max(ef_search, 3k) exceeds CM_MAX_K");
  SearchCall c{Kind::Temporal, q, nq, k, (uint32_t)pool, (uint32_t)pool, exact, w, base, now_secs, now_nanos, ids, score, dist};
  std::atomic<uint32_t> bad_row{0xFFFFFFFFu};
  c.validate = true;
  c.bad_row = &bad_row;
  cm_status st = group_search(g, c);
  if (bad_row.load() != 0xFFFFFFFFu)  // the whole call fails with the first bad row, as `search` would for that query
    return gfail(CM_ERR_INVALID_VECTOR, "invalid vector data: query contains NaN or infinite components (row " +
                                            std::to_string(bad_row.load()) + ")");
  return st;
}

}  // extern "C"
