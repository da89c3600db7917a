// build_select.cu — the host half of the GPU-assisted graph construction, moved to the device.
//
// `LockFreeHnsw::insert` spends what is left after its `search_layer` in two places: `select_diverse` over the new
// node's candidates (src/index/lockfree_hnsw.rs:205-240, Algorithm 4 + keepPruned backfill) and `add_backlink`'s
// re-selection of a full neighbor list (:246-280). Both are the same procedure over a CENTER row and a list of
// candidate rows:
//     order the candidates by (distance to the center, id);
//     walk them in that order: a candidate is admitted iff it is closer to the center than to every candidate admitted
//     so far (raw f32 `<`, :225) — stop once m are admitted; fill up with the rejected ones in order.
// k_select_diverse runs one such item per CTA with the reference's exact fp32 arithmetic (the same quad chain the
// traversal and the rescore use), so the lists it produces are BIT-FOR-BIT the lists the host code in hnsw_build.cpp
// produces from the same candidates — which is how it is tested (build both ways, compare every list).
//
// Memory shape: rows are gathered by the copy engine (one cp.async.bulk per row). The center row and the rows admitted
// so far stay resident in shared memory (at most m = 32 of them: 99 KB at 768 floats); the candidates are streamed
// through a 16-row window twice — once to measure their distance to the center, once, in sorted order, to be judged —
// and an admitted row is copied from the window into the resident set. One CTA (256 threads = 64 quads) per SM-half.
#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"
#include "row_gather.cuh"

namespace cm {

constexpr uint32_t SD_THREADS = 256;
constexpr uint32_t SD_WINDOW = 16;      // candidate rows staged per round
constexpr uint32_t SD_MAX_CANDS = 1024;  // candidates per item the kernel sorts (larger items go to the host)

struct SelectArgs {
  const float* x;            // rows [n][ldx]
  uint32_t ldx, dim;
  uint32_t n_items, m;       // m = list capacity (M for a node's own list, cap(layer) for a reverse-edge list)
  // item i: center row `center[i]` (or i when center == nullptr); candidates either
  //   (a) fixed width: cand[i * width .. +width), CM_NO_ID padded                            (phase A), or
  //   (b) `own[i * own_width .. +own_deg[i])` followed by `incoming[in_start[i] .. in_start[i+1])` minus the entries
  //       already in `own` — what appending every source to the target's list would leave      (phase B)
  const uint32_t* cand;
  uint32_t width;
  const uint32_t* own;
  const uint32_t* own_deg;
  uint32_t own_width;
  const uint32_t* in_start;
  const uint32_t* incoming;
  uint32_t* out;             // [n_items][m], CM_NO_ID padded
  uint32_t* out_deg;         // [n_items]
  uint32_t* too_big;         // items with more than SD_MAX_CANDS candidates: listed for the host (count at [0], ids from [1])
  uint32_t too_big_cap;
  uint32_t* work_counter;
};

__global__ void __launch_bounds__(SD_THREADS) k_select_diverse(SelectArgs a) {
  extern __shared__ __align__(128) unsigned char sd_smem[];
  const uint32_t row_stride = a.ldx * 4 + RG_ROW_PAD;
  unsigned char* center_row = sd_smem;                                   // 1 row
  unsigned char* sel_rows = center_row + row_stride;                     // m rows (admitted so far)
  unsigned char* window = sel_rows + (size_t)a.m * row_stride;           // SD_WINDOW rows
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(window + (size_t)SD_WINDOW * row_stride);  // [SD_MAX_CANDS]
  uint32_t* ids = reinterpret_cast<uint32_t*>(keys + SD_MAX_CANDS);      // [SD_MAX_CANDS] candidate ids, then sorted ids
  uint32_t* rej = ids + SD_MAX_CANDS;                                    // [SD_MAX_CANDS] rejected ids in order
  __shared__ unsigned long long mbar;
  __shared__ uint32_t s_item, s_cnt, s_nsel, s_nrej, s_fail;
  __shared__ uint32_t s_sel_ids[64];
  const uint32_t tid = threadIdx.x, quad = tid >> 2, c4 = tid & 3;
  if (tid == 0) rg_mbar_init(&mbar);
  uint32_t phase = 0;
  __syncthreads();

  for (;;) {
    if (tid == 0) s_item = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= a.n_items) break;
    const uint32_t center = item;  // items are the layer's members in order; rows are indexed the same way
    // ---- gather the candidate ids -------------------------------------------------------------------------------
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    uint32_t total = 0;
    if (a.cand) {
      for (uint32_t j = tid; j < a.width; j += SD_THREADS) {
        const uint32_t v = a.cand[(uint64_t)item * a.width + j];
        if (v != CM_NO_ID_DEV && v != center) ids[atomicAdd(&s_cnt, 1u)] = v;  // order restored by the sort below
      }
      __syncthreads();
      total = s_cnt;
    } else {
      const uint32_t od = a.own_deg[item];
      const uint32_t lo = a.in_start[item], hi = a.in_start[item + 1];
      total = od + (hi - lo);  // upper bound; duplicates of `own` entries are dropped below
      if (total <= SD_MAX_CANDS) {
        // own list first, then the sources not already in it, in ascending source order: a single thread keeps the
        // order (what appending in handle order would leave; the no-overflow case is written out in exactly this order)
        if (tid == 0) {
          uint32_t c = 0;
          for (uint32_t j = 0; j < od; ++j) ids[c++] = a.own[(uint64_t)item * a.own_width + j];
          for (uint32_t e = lo; e < hi; ++e) {
            const uint32_t u = a.incoming[e];
            bool dup = false;
            for (uint32_t j = 0; j < od; ++j) dup |= ids[j] == u;
            if (!dup) ids[c++] = u;
          }
          s_cnt = c;
        }
        __syncthreads();
        total = s_cnt;
      }
    }
    if (total > SD_MAX_CANDS) {  // a hub with more incoming edges than the kernel sorts: the host selects this one
      if (tid == 0) {
        const uint32_t slot = atomicAdd(&a.too_big[0], 1u);
        if (slot < a.too_big_cap) a.too_big[1 + slot] = item;
        a.out_deg[item] = 0;
      }
      __syncthreads();
      continue;
    }
    if (!a.cand && total <= a.m) {  // add_backlink's first branch: there is room, every edge is kept (:255-259)
      for (uint32_t j = tid; j < a.m; j += SD_THREADS) a.out[(uint64_t)item * a.m + j] = j < total ? ids[j] : CM_NO_ID_DEV;
      if (tid == 0) a.out_deg[item] = total;
      __syncthreads();
      continue;
    }
    // ---- distances to the center: candidates streamed through the window -------------------------------------------
    if (tid == 0) s_sel_ids[0] = center;
    __syncthreads();
    rg_stage_rows(a.x, a.ldx, &mbar, center_row, row_stride, s_sel_ids, 1, phase);
    const float* cq = reinterpret_cast<const float*>(center_row);
    for (uint32_t r0 = 0; r0 < total; r0 += SD_WINDOW) {
      const uint32_t cnt = min(SD_WINDOW, total - r0);
      rg_stage_rows(a.x, a.ldx, &mbar, window, row_stride, ids + r0, cnt, phase);
      const bool active = quad < cnt;
      const float d = rg_quad_distance(cq, window + (size_t)(quad < SD_WINDOW ? quad : 0) * row_stride, a.dim, c4, active);
      if (active && c4 == 0) keys[r0 + quad] = pack_key(d, ids[r0 + quad]);
      __syncthreads();  // window free for the next round
    }
    const uint32_t P = next_pow2(total < 32 ? 32 : total);
    for (uint32_t i = total + tid; i < P; i += SD_THREADS) keys[i] = kKeyMax;
    block_bitonic_sort(keys, P);  // ascending (distance, id) — `sort_unstable` on (TotalF32, u32) tuples (:268-270)
    for (uint32_t i = tid; i < total; i += SD_THREADS) ids[i] = (uint32_t)keys[i];
    if (tid == 0) s_nsel = 0, s_nrej = 0;
    __syncthreads();
    // ---- select_diverse (:205-240): sorted candidates through the window, admitted rows kept resident --------------
    bool full = false;
    for (uint32_t r0 = 0; r0 < total && !full; r0 += SD_WINDOW) {
      const uint32_t cnt = min(SD_WINDOW, total - r0);
      rg_stage_rows(a.x, a.ldx, &mbar, window, row_stride, ids + r0, cnt, phase);
      for (uint32_t w = 0; w < cnt; ++w) {
        const uint32_t nsel = s_nsel;  // uniform: written before the last barrier
        if (nsel >= a.m) {
          full = true;
          break;
        }
        const float dq = key_to_float((uint32_t)(keys[r0 + w] >> 32));
        if (tid == 0) s_fail = 0;
        __syncthreads();
        // one quad per admitted row: distance between the candidate and that row
        const bool active = quad < nsel;
        const float dc = rg_quad_distance(reinterpret_cast<const float*>(window + (size_t)w * row_stride),
                                          sel_rows + (size_t)(active ? quad : 0) * row_stride, a.dim, c4, active);
        if (active && c4 == 0 && !(dq < dc)) s_fail = 1;  // raw f32 `<` (:225); NaN compares false => not diverse
        __syncthreads();
        if (s_fail) {
          if (tid == 0) rej[s_nrej++] = ids[r0 + w];
        } else {
          // admitted: its row joins the resident set
          const uint4* src = reinterpret_cast<const uint4*>(window + (size_t)w * row_stride);
          uint4* dst = reinterpret_cast<uint4*>(sel_rows + (size_t)nsel * row_stride);
          for (uint32_t i = tid; i < (a.ldx * 4) / 16; i += SD_THREADS) dst[i] = src[i];
          if (tid == 0) {
            s_sel_ids[nsel] = ids[r0 + w];
            s_nsel = nsel + 1;
          }
        }
        __syncthreads();
      }
      __syncthreads();  // window free for the next round
    }
    // ---- result: the admitted ones, then the rejected ones in order until m (:233-239) ------------------------------
    const uint32_t nsel = s_nsel, nrej = s_nrej;
    const uint32_t deg = min(a.m, nsel + nrej);
    for (uint32_t j = tid; j < a.m; j += SD_THREADS) {
      uint32_t v = CM_NO_ID_DEV;
      if (j < nsel) v = s_sel_ids[j];
      else if (j < deg) v = rej[j - nsel];
      a.out[(uint64_t)item * a.m + j] = v;
    }
    if (tid == 0) a.out_deg[item] = deg;
    __syncthreads();
  }
}

size_t select_diverse_smem_bytes(uint32_t dim_pad, uint32_t m) {
  const size_t row_stride = (size_t)dim_pad * 4 + RG_ROW_PAD;
  return (1 + (size_t)m + SD_WINDOW) * row_stride + (size_t)SD_MAX_CANDS * (8 + 4 + 4) + 128;
}

bool select_diverse_supported(uint32_t dim_pad, uint32_t m) {
  return m >= 1 && m <= 64 && dim_pad % 4 == 0 && select_diverse_smem_bytes(dim_pad, m) <= 220 * 1024;
}

cudaError_t launch_select_diverse(const DeviceIndex& ix, const SelectDiverseLaunch& l, cudaStream_t s) {
  if (l.n_items == 0) return cudaSuccess;
  SelectArgs a;
  a.x = ix.x32;
  a.ldx = ix.dim_pad;
  a.dim = ix.dim;
  a.n_items = l.n_items;
  a.m = l.m;
  a.cand = l.cand;
  a.width = l.width;
  a.own = l.own;
  a.own_deg = l.own_deg;
  a.own_width = l.own_width;
  a.in_start = l.in_start;
  a.incoming = l.incoming;
  a.out = l.out;
  a.out_deg = l.out_deg;
  a.too_big = l.too_big;
  a.too_big_cap = l.too_big_cap;
  a.work_counter = l.work_counter;
  const size_t smem = select_diverse_smem_bytes(ix.dim_pad, l.m);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_select_diverse, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_select_diverse, SD_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  e = cudaMemsetAsync(l.work_counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ix.sm_count * per_sm, l.n_items);
  k_select_diverse<<<grid, SD_THREADS, smem, s>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cm
