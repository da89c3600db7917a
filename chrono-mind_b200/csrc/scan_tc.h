// scan_tc.h — launch interface of the tcgen05 scan (K1) and the fp32 rescore (K2). See scan_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cm_internal.h"

namespace cm {

uint64_t scan_tc_pad_rows(uint64_t n);        // corpus rows padded to the X tile height (256)
uint32_t scan_tc_pad_queries(uint32_t nq);    // queries padded to the Q tile-pair height (256)
uint32_t scan_tc_cand_cap(uint32_t k);        // candidate slots per query
float scan_tc_eps(uint32_t dim);              // bound on |bf16 tensor-core score - fp32 dot| for unit vectors of `dim` elements
uint32_t scan_tc_pad_k(uint32_t dim);         // columns padded to the K block (64)
bool scan_tc_supported(const DeviceIndex& ix, uint32_t k);

struct ScanTcLaunch {
  const void* q_bf16;        // [pad_queries(nq)][k_pad] bf16, zero padded
  uint32_t nq, k;
  uint32_t* tau_key;         // [pad_queries(nq)] zeroed
  uint32_t* cand_cnt;        // [pad_queries(nq)] zeroed
  unsigned long long* cand;  // [pad_queries(nq)][cap]
  uint32_t cap;
  // threshold self-join (q_bf16 == the corpus copy, nq == n): pairs (query < row) whose bf16 score >= join_tau
  bool join = false;
  float join_tau = 0.0f;
  unsigned long long* pairs = nullptr;
  uint32_t* pair_count = nullptr;  // zeroed by the caller
  uint32_t pair_cap = 0;
  // predecessor mode (graph construction): q_bf16 = rows [q_base, q_base + nq) of the corpus copy itself (q_base a
  // multiple of 256); row r competes for query i only when r < q_base + i
  bool pred = false;
  uint32_t q_base = 0;
  // biased mode (`search_in_context`): ranking key = score + bias[row], restricted to rows with row_ctx[row] == q_ctx[q]
  const float* bias = nullptr;        // [pad_rows(n)]
  const uint32_t* row_ctx = nullptr;  // [pad_rows(n)]
  const uint32_t* q_ctx = nullptr;    // [nq]
  // context-sorted layout: scan `x_bf16` (a permuted copy of the corpus, same shape) instead of the index's own copy,
  // and give every query-tile pair its own tile range [pair_lo[m], pair_hi[m]) (device arrays, one entry per pair)
  const void* x_bf16 = nullptr;
  const uint32_t* pair_lo = nullptr;
  const uint32_t* pair_hi = nullptr;
  uint32_t pair_tiles_avg = 0;        // mean range length (sizes the splits)
  // keep ~40 KB of every SM's shared memory free: a rescore CTA of ANOTHER batch in flight can then co-reside with the scan
  bool leave_room = false;
};
cudaError_t launch_scan_tc(const DeviceIndex& ix, const ScanTcLaunch& l, cudaStream_t s);

struct RescoreLaunch {
  const float* qn;           // prepared fp32 queries [nq][ldq]
  uint32_t ldq, nq, k;
  const uint32_t* tau_key;
  const uint32_t* cand_cnt;
  const unsigned long long* cand;
  uint32_t cap;
  uint32_t* ids;             // [nq][k]
  float* dist;               // [nq][k]
  uint32_t* flag_count;      // zeroed; queries that must be re-run through the fp32 scan
  uint32_t* flag_list;       // [flag_cap]
  uint32_t flag_cap;
  unsigned long long* rescored_total;
  bool approx = false;       // top-k by bf16 score, no exact rescore (graph-construction candidates)
  bool pred = false;         // predecessor mode: query i has pred_base + i rows to choose from
  uint32_t pred_base = 0;
};
cudaError_t launch_rescore(const DeviceIndex& ix, const RescoreLaunch& l, cudaStream_t s);

// `search_in_context` on the tensor cores (see scan_tc.cu): the per-row bias of one call, and the exact decision on the
// stored rows.
cudaError_t launch_context_bias(const DeviceIndex& ix, float w, float base, uint64_t now_secs, uint32_t now_nanos, float* bias,
                                cudaStream_t s, const uint32_t* perm = nullptr);
// Context-sorted layout: queries grouped by context in chunks of scan_tc_context_chunk() (order[i] = caller row of scan
// query i, qctx[i] = its context id), and per query-tile pair the corpus tile range that holds the pair's contexts
// (pairs numbered chunk by chunk: chunk c owns pairs [c * chunk / 256, (c + 1) * chunk / 256)).
uint32_t scan_tc_context_chunk();
cudaError_t launch_group_queries_by_context(const uint32_t* want, uint32_t nq, uint32_t* order, uint32_t* qctx, cudaStream_t s);
cudaError_t launch_context_pair_ranges(const DeviceIndex& ix, const uint32_t* qctx, uint32_t nq, uint32_t* pair_lo,
                                       uint32_t* pair_hi, cudaStream_t s);
struct ContextRescoreLaunch {
  const float* q_raw;        // queries as given [nq][ldq]
  uint32_t ldq, nq, k;
  const float* qnorm2;       // [nq]
  float w, base;
  uint64_t now_secs;
  uint32_t now_nanos;
  const uint32_t* tau_key;
  const uint32_t* cand_cnt;
  const unsigned long long* cand;
  uint32_t cap;
  uint32_t* ids;             // [nq][k]
  float* score;              // [nq][k]
  uint32_t* flag_count;
  uint32_t* flag_list;
  uint32_t flag_cap;
  unsigned long long* rescored_total;
  const uint32_t* perm = nullptr;     // candidate row (scan order) -> handle
  const uint32_t* q_order = nullptr;  // scan query index -> caller's query row (q_raw, qnorm2, ids, score are indexed by it)
};
cudaError_t launch_rescore_context(const DeviceIndex& ix, const ContextRescoreLaunch& l, cudaStream_t s);

}  // namespace cm
