// kernels_basic.cu — CUDA-core kernels of the read path: preprocess (K0), the bit-exact fp32 scan, top-k
// select, shard merge (K5), temporal rerank (K4) and the metric surface.
//
// All of this is HBM/ALU work on CUDA cores (no GEMM shape is forced onto it); the tensor-core scan lives
// in scan_tc.cu and the graph traversal in hnsw_search.cu.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"

namespace cm {

// ---------------------------------------------------------------------------------------------------------
// K0 preprocess — CosineDistance::preprocess (src/metric.rs:208-214).
// One CTA (8 warps) owns 32 rows. Tiles of 32 rows x 32 columns are staged through shared memory by all 8 warps
// (coalesced 128-byte reads, 4 rows per warp in flight at once), then lane r of warp 0 walks ITS row of the
// tile sequentially: s += x*x with a separately rounded multiply and add, exactly like the reference's scalar
// `map(|x| x * x).sum()`. The division pass is again spread over the 8 warps.
// ---------------------------------------------------------------------------------------------------------
constexpr int PP_THREADS = 256;

__global__ void __launch_bounds__(PP_THREADS) k_preprocess_rows(const float* __restrict__ x, uint64_t n, uint32_t dim,
                                                                uint32_t ld_in, float* __restrict__ out,
                                                                uint32_t ld_out, uint32_t* bad_row) {
  __shared__ float tile[2][32][33];
  __shared__ float norms[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t row0 = (uint64_t)blockIdx.x * 32;
  float s = 0.0f;
  auto load_tile = [&](uint32_t c0, int buf) {
#pragma unroll
    for (uint32_t rr = 0; rr < 4; ++rr) {
      const uint32_t r = warp * 4 + rr;
      const uint64_t row = row0 + r;
      const uint32_t c = c0 + lane;
      float v = (row < n && c < dim) ? x[row * ld_in + c] : 0.0f;
      // validate_query (src/store.rs:386-390): remember the first row with a NaN/Inf component.
      if (bad_row && !isfinite(v)) atomicMin(bad_row, (uint32_t)row);
      tile[buf][r][lane] = v;
    }
  };
  int buf = 0;
  if (dim) load_tile(0, 0);
  __syncthreads();
  for (uint32_t c0 = 0; c0 < dim; c0 += 32) {
    if (c0 + 32 < dim) load_tile(c0 + 32, buf ^ 1);  // next tile in flight while warp 0 walks this one
    if (warp == 0) {
      const uint32_t lim = min(32u, dim - c0);
      for (uint32_t c = 0; c < lim; ++c) {
        const float v = tile[buf][lane][c];
        s = __fadd_rn(s, __fmul_rn(v, v));
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  if (warp == 0) norms[lane] = __fsqrt_rn(s);
  __syncthreads();
#pragma unroll
  for (uint32_t rr = 0; rr < 4; ++rr) {
    const uint32_t r = warp * 4 + rr;
    const uint64_t row = row0 + r;
    if (row >= n) continue;
    const float nr = norms[r];
    const bool degenerate = nr <= 1.1920929e-7f;  // f32::EPSILON: left unchanged
    for (uint32_t c = lane; c < ld_out; c += 32) {
      float v = 0.0f;
      if (c < dim) {
        v = x[row * ld_in + c];
        if (!degenerate) v = __fdiv_rn(v, nr);
      }
      out[row * ld_out + c] = v;
    }
  }
}

cudaError_t launch_preprocess_rows(const float* x, uint64_t n, uint32_t dim, uint32_t ld_in, float* out,
                                   uint32_t ld_out, uint32_t* bad_row, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  // In-place use is only safe when the strides match: a CTA reads all of its 32 rows (first loop) before any of
  // them is written (second loop, after a barrier), and no other CTA touches those rows.
  k_preprocess_rows<<<(unsigned)((n + 31) / 32), PP_THREADS, 0, s>>>(x, n, dim, ld_in, out, ld_out, bad_row);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K0 for query batches, fused with the bf16 cast of the tensor-core scan: one pass over the queries instead of two
// kernels (preprocess, then cast) that each stream them through HBM, and no per-tile barrier chain.
// CTA = PQ_ROWS rows. Each row (dim * 4 bytes, dim % 4 == 0) is landed whole in shared memory by ONE bulk copy
// (cp.async.bulk, all of them in flight at once); lane r then walks row r sequentially with LDS.128 — the reference's
// scalar `map(|x| x * x).sum()` order, separately rounded multiply and add — and all threads write the
// normalised fp32 row and its bf16 copy (zero padded to the scan's [nq_pad][k_pad] layout). The row stride of
// dim * 4 + 16 bytes puts the 8 lanes of a quarter-warp on 8 different 16-byte bank groups.
// ---------------------------------------------------------------------------------------------------------
// 8 rows and 128 threads per CTA: 24.7 KB of shared memory at dim 768 -> 9 CTAs per SM, so a 10,000-query batch (1,250
// CTAs) is ONE wave; with 16 rows / 256 threads (4 per SM, 625 CTAs) a second, nearly empty wave doubled the kernel.
constexpr uint32_t PQ_ROWS = 8;
constexpr int PQ_THREADS = 128;

__global__ void __launch_bounds__(PQ_THREADS) k_prepare_queries(const float* __restrict__ x, uint32_t n, uint32_t dim,
                                                                float* __restrict__ out, uint32_t ld_out,
                                                                __nv_bfloat16* __restrict__ out16, uint32_t n_pad16,
                                                                uint32_t k_pad, uint32_t* bad_row,
                                                                const uint32_t* __restrict__ src_rows) {
  extern __shared__ __align__(128) unsigned char pq_smem[];  // PQ_ROWS x row_stride
  __shared__ unsigned long long mbar;
  __shared__ float norms[PQ_ROWS];
  const uint32_t tid = threadIdx.x;
  const uint32_t row_bytes = dim * 4, row_stride = row_bytes + 16;
  const uint32_t row0 = blockIdx.x * PQ_ROWS;
  const uint32_t cnt = row0 < n ? min(PQ_ROWS, n - row0) : 0;  // blocks past n only zero the bf16 padding rows
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cnt) {
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * row_bytes) : "memory");
    if (tid < cnt) {
      // out row i = x row src_rows[i] when a gather order is given (queries grouped by context)
      const float* src = x + (uint64_t)(src_rows ? src_rows[row0 + tid] : row0 + tid) * dim;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"((uint32_t)__cvta_generic_to_shared(pq_smem + (size_t)tid * row_stride)), "l"(src), "r"(row_bytes), "r"(bar)
          : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(0u)
          : "memory");
      if (spin > (1u << 24)) __trap();
    }
    if (tid < cnt) {  // sequential sum of squares, element order (src/metric.rs:209)
      const float4* r4 = reinterpret_cast<const float4*>(pq_smem + (size_t)tid * row_stride);
      float s = 0.0f;
      for (uint32_t t = 0; t < dim / 4; ++t) {
        const float4 v = r4[t];
        s = __fadd_rn(s, __fmul_rn(v.x, v.x));
        s = __fadd_rn(s, __fmul_rn(v.y, v.y));
        s = __fadd_rn(s, __fmul_rn(v.z, v.z));
        s = __fadd_rn(s, __fmul_rn(v.w, v.w));
      }
      norms[tid] = __fsqrt_rn(s);
    }
  }
  __syncthreads();
  // write phase: thread t handles element pairs (bf16x2 stores); all threads sweep the rows one after the other
  const uint32_t pairs16 = k_pad / 2;
  for (uint32_t r = 0; r < PQ_ROWS; ++r) {
    const uint32_t row = row0 + r;
    if (row >= n_pad16 && row >= n) break;
    const bool live = r < cnt;
    const float nr = live ? norms[r] : 0.0f;
    const bool degenerate = nr <= 1.1920929e-7f;  // f32::EPSILON: left unchanged
    const float* rs = reinterpret_cast<const float*>(pq_smem + (size_t)r * row_stride);
    for (uint32_t p = tid; p < pairs16; p += PQ_THREADS) {
      const uint32_t c = 2 * p;
      float a = 0.0f, b = 0.0f;
      if (live) {
        if (c < dim) a = rs[c];
        if (c + 1 < dim) b = rs[c + 1];
        // validate_query (src/store.rs:386-390): remember the first row with a NaN/Inf component
        if (bad_row && (!isfinite(a) || !isfinite(b))) atomicMin(bad_row, src_rows ? src_rows[row] : row);
        if (!degenerate) {
          if (c < dim) a = __fdiv_rn(a, nr);
          if (c + 1 < dim) b = __fdiv_rn(b, nr);
        }
        if (c < ld_out) out[(uint64_t)row * ld_out + c] = a;
        if (c + 1 < ld_out) out[(uint64_t)row * ld_out + c + 1] = b;
      }
      if (row < n_pad16) reinterpret_cast<__nv_bfloat162*>(out16)[(uint64_t)row * pairs16 + p] = __floats2bfloat162_rn(a, b);
    }
  }
}

bool prepare_queries_fused_supported(const float* x, uint32_t dim) {
  return dim % 4 == 0 && dim >= 4 && (size_t)PQ_ROWS * (dim * 4 + 16) <= 200 * 1024 && ((uintptr_t)x & 15) == 0;
}

cudaError_t launch_prepare_queries(const float* x, uint32_t n, uint32_t dim, float* out, uint32_t ld_out, void* out16,
                                   uint32_t n_pad16, uint32_t k_pad, uint32_t* bad_row, cudaStream_t s,
                                   const uint32_t* src_rows) {
  const uint32_t rows = n_pad16 > n ? n_pad16 : n;
  if (rows == 0) return cudaSuccess;
  const size_t smem = (size_t)PQ_ROWS * (dim * 4 + 16);
  cudaError_t e = cudaFuncSetAttribute(k_prepare_queries, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  k_prepare_queries<<<(rows + PQ_ROWS - 1) / PQ_ROWS, PQ_THREADS, smem, s>>>(x, n, dim, out, ld_out, (__nv_bfloat16*)out16,
                                                                             n_pad16, k_pad, bad_row, src_rows);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fp32 -> bf16 (round to nearest even) into the scan layout. With mask_rows set (corpus copies), tombstoned
// rows and the padding rows past n get a NaN in column 0: their tensor-core score is NaN, which never passes
// a threshold (`s >= tau` is false) and is dropped by fmaxf — the scan needs no per-row liveness lookup.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_to_bf16(const float* __restrict__ x, uint64_t n, uint32_t dim, uint32_t ld_in,
                          __nv_bfloat16* __restrict__ out, uint64_t n_pad, uint32_t k_pad,
                          const uint8_t* __restrict__ deleted, int mask_rows, const uint32_t* __restrict__ row_map) {
  uint64_t total = n_pad * (uint64_t)(k_pad / 2);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t row = i / (k_pad / 2);
    uint32_t c = (uint32_t)(i % (k_pad / 2)) * 2;
    float a = 0.0f, b = 0.0f;
    const uint64_t src = (row_map && row < n) ? row_map[row] : row;  // out row `row` = x row row_map[row] (permuted copies)
    if (row < n) {
      if (c < dim) a = x[src * ld_in + c];
      if (c + 1 < dim) b = x[src * ld_in + c + 1];
    }
    if (mask_rows && c == 0 && (row >= n || (deleted && deleted[src]))) a = __int_as_float(0x7FC00000);
    reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(a, b);
  }
}

cudaError_t launch_to_bf16(const float* x, uint64_t n, uint32_t dim, uint32_t ld_in, void* out, uint64_t n_pad,
                           uint32_t k_pad, const uint8_t* deleted, bool mask_rows, cudaStream_t s, const uint32_t* row_map) {
  if (n_pad == 0) return cudaSuccess;
  k_to_bf16<<<148 * 8, 256, 0, s>>>(x, n, dim, ld_in, (__nv_bfloat16*)out, n_pad, k_pad, deleted, mask_rows ? 1 : 0, row_map);
  count_launch();
  return cudaGetLastError();
}

// `VectorIndex::remove` on the device replica (src/index/lockfree_hnsw.rs:451-467 flips `Node.deleted`): set the
// tombstone byte the HNSW final-list filter and the fp32 scans read, and NaN-mask column 0 of the bf16 row so the
// tensor-core scans never emit it again. The node keeps routing traffic, as in the reference.
__global__ void k_apply_tombstones(const uint32_t* __restrict__ handles, uint32_t m, uint8_t* __restrict__ deleted,
                                   __nv_bfloat16* __restrict__ xbf16, uint32_t k_pad, __nv_bfloat16* __restrict__ xbf16_ctx,
                                   const uint32_t* __restrict__ inv_perm) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const uint32_t h = handles[i];
    deleted[h] = 1;
    xbf16[(uint64_t)h * k_pad] = __float2bfloat16(__int_as_float(0x7FC00000));
    if (xbf16_ctx) xbf16_ctx[(uint64_t)inv_perm[h] * k_pad] = __float2bfloat16(__int_as_float(0x7FC00000));  // context-sorted copy
  }
}
cudaError_t launch_apply_tombstones(const uint32_t* handles, uint32_t m, uint8_t* deleted, void* xbf16, uint32_t k_pad,
                                    void* xbf16_ctx, const uint32_t* inv_perm, cudaStream_t s) {
  if (m == 0) return cudaSuccess;
  k_apply_tombstones<<<(m + 255) / 256, 256, 0, s>>>(handles, m, deleted, (__nv_bfloat16*)xbf16, k_pad, (__nv_bfloat16*)xbf16_ctx,
                                                     inv_perm);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// distance_prepared on row pairs (src/metric.rs:219-224): one lane-octet per pair; lane j carries the
// reference's SIMD lane j (elements j, j+8, ...), then the reference's reduction order and scalar tail.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_distance(const float* __restrict__ a, const float* __restrict__ b,
                                                       uint64_t n, uint32_t dim, uint32_t ld, float* __restrict__ out) {
  uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  uint32_t j = threadIdx.x & 7;
  bool valid = pair < n;
  uint64_t p = valid ? pair : 0;
  const float* pa = a + p * ld;
  const float* pb = b + p * ld;
  uint32_t body = dim & ~7u;
  float acc = 0.0f;
  for (uint32_t e = j; e < body; e += 8) acc = __fmaf_rn(pa[e], pb[e], acc);
  float sum = octet_hsum(acc);
  for (uint32_t e = body; e < dim; ++e) sum = __fadd_rn(sum, __fmul_rn(pa[e], pb[e]));
  if (valid && j == 0) out[pair] = dim == 0 ? 2.0f : clamp02(__fsub_rn(1.0f, sum));
}

cudaError_t launch_pair_distance(const float* a, const float* b, uint64_t n, uint32_t dim, uint32_t ld, float* out,
                                 cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  uint64_t threads = n * 8;
  k_pair_distance<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(a, b, n, dim, ld, out);
  count_launch();
  return cudaGetLastError();
}

// dot_avx2(a[i], b[i]) (src/metric.rs:129-147) for n row pairs — with a == b this is the squared norm exactly as
// dot_and_norms_avx2 accumulates it (src/metric.rs:78-116: same FMA chain, same horizontal sum, same scalar tail).
__global__ void __launch_bounds__(256) k_pair_dot(const float* __restrict__ a, const float* __restrict__ b, uint64_t n,
                                                  uint32_t dim, uint32_t ld, float* __restrict__ out) {
  uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  uint32_t j = threadIdx.x & 7;
  bool valid = pair < n;
  uint64_t p = valid ? pair : 0;
  const float* pa = a + p * ld;
  const float* pb = b + p * ld;
  uint32_t body = dim & ~7u;
  float acc = 0.0f;
  for (uint32_t e = j; e < body; e += 8) acc = __fmaf_rn(pa[e], pb[e], acc);
  float sum = octet_hsum(acc);
  for (uint32_t e = body; e < dim; ++e) sum = __fadd_rn(sum, __fmul_rn(pa[e], pb[e]));
  if (valid && j == 0) out[pair] = sum;
}

cudaError_t launch_pair_dot(const float* a, const float* b, uint64_t n, uint32_t dim, uint32_t ld, float* out,
                            cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  uint64_t threads = n * 8;
  k_pair_dot<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(a, b, n, dim, ld, out);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fp32 exact scan (CUDA cores), bit-exact to the reference's dot_avx2.
//
// CTA = 8 warps, 16 queries x 256 rows. The 16 queries sit in shared memory transposed (qs[e][qi]) so one
// lane fetches the 16 query values of its element with four LDS.128. A warp covers 4 rows x 8 SIMD lanes
// per pass: lane (r, j) streams x[row][8t + j] (the octet reads one 32-byte sector per step) and keeps 16
// accumulators — the reference's lane-j partial sum for each of the 16 queries — so every x value is reused
// 16 times from a register.
// ---------------------------------------------------------------------------------------------------------
constexpr int EX_QT = 16;   // queries per CTA; 8 or 4 for very long rows (the 16 x dim query tile must fit shared memory)
constexpr int EX_ROWS = 256;

// One block of EX_ROWS rows against the QT queries staged in `qs` ([body][QT], transposed): D[qq][row] for the live
// (qq < nq) queries. `q` is the row-major query matrix the tile was staged from (scalar tail of rows with dim % 8 != 0).
template <int QT>
__device__ __forceinline__ void exact_dist_rowblock(const float* __restrict__ qs, const float* __restrict__ q, uint32_t nq,
                                                    uint32_t ldq, uint32_t q0, const float* __restrict__ x, uint64_t n,
                                                    uint32_t dim, uint32_t ldx, float* __restrict__ D, uint64_t ldD,
                                                    uint64_t rb, const ContextEpilogue& ce) {
  constexpr int EX_QT = QT;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t j = lane & 7, rsub = lane >> 3;
  const uint32_t body = dim & ~7u;
  const uint64_t row_base = rb * EX_ROWS;
  for (uint32_t pass = 0; pass < EX_ROWS / 32; ++pass) {
    uint64_t row = row_base + pass * 32 + warp * 4 + rsub;
    bool valid = row < n;
    const float* xr = x + (valid ? row : 0) * ldx;
    float acc[EX_QT];
#pragma unroll
    for (int i = 0; i < EX_QT; ++i) acc[i] = 0.0f;
#pragma unroll 4
    for (uint32_t e = j; e < body; e += 8) {
      float xv = __ldg(xr + e);
      const float4* qp = reinterpret_cast<const float4*>(qs + (size_t)e * EX_QT);
#pragma unroll
      for (int g = 0; g < EX_QT / 4; ++g) {
        const float4 a = qp[g];
        acc[4 * g + 0] = __fmaf_rn(xv, a.x, acc[4 * g + 0]);
        acc[4 * g + 1] = __fmaf_rn(xv, a.y, acc[4 * g + 1]);
        acc[4 * g + 2] = __fmaf_rn(xv, a.z, acc[4 * g + 2]);
        acc[4 * g + 3] = __fmaf_rn(xv, a.w, acc[4 * g + 3]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < EX_QT; ++qi) {
      float sum = octet_hsum(acc[qi]);
      uint32_t qq = q0 + qi;
      if (body < dim && qq < nq) {
        const float* qrow = q + (uint64_t)qq * ldq;
        for (uint32_t e = body; e < dim; ++e) sum = __fadd_rn(sum, __fmul_rn(xr[e], qrow[e]));
      }
      if ((qi & 7) == (int)j && valid && qq < nq) {
        float out;
        if (!ce.enabled) {
          out = clamp02(__fsub_rn(1.0f, sum));  // distance_prepared
        } else {
          // search_in_context (src/store.rs:362-376): members of the query's context only; x/q are RAW rows, so
          // `sum` is dot_and_norms' dot and the squared norms come precomputed in the same arithmetic.
          out = __int_as_float(0x7F800000);
          if (ce.ctx[row] == ce.want[qq] && !(ce.deleted && ce.deleted[row])) {
            const float denom = __fsqrt_rn(__fmul_rn(ce.na[row], ce.nb[qq]));
            float dist = 2.0f;                                   // zero vector: similarity undefined (metric.rs:181-184)
            if (!(denom <= 1.1920929e-7f)) {
              float c = __fdiv_rn(sum, denom);
              c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);      // f32::clamp: NaN passes through
              dist = __fsub_rn(1.0f, c);
            }
            out = combined_score_dev(dist, ce.ts_secs[row], ce.ts_nanos[row], ce.decay[row], ce.now_secs, ce.now_nanos,
                                     ce.w, ce.base);
          }
        }
        D[(uint64_t)qq * ldD + row] = out;
      }
    }
  }
}

// Stage queries [q0, q0 + QT) of `q` into qs[e][qi] (transposed; rows past nq are zero).
template <int QT>
__device__ __forceinline__ void exact_dist_stage_queries(float* qs, const float* __restrict__ q, uint32_t nq, uint32_t ldq,
                                                         uint32_t q0, uint32_t body) {
  for (uint32_t idx = threadIdx.x; idx < body * QT; idx += blockDim.x) {
    uint32_t e = idx / QT, qi = idx % QT;
    uint32_t qq = q0 + qi;
    qs[idx] = qq < nq ? q[(uint64_t)qq * ldq + e] : 0.0f;
  }
}

template <int QT>
__global__ void __launch_bounds__(256) k_exact_dist(const float* __restrict__ q, uint32_t nq, uint32_t ldq,
                                                    const float* __restrict__ x, uint64_t n, uint32_t dim,
                                                    uint32_t ldx, float* __restrict__ D, uint64_t ldD,
                                                    const uint32_t* __restrict__ active_count,
                                                    uint32_t active_base, ContextEpilogue ce) {
  extern __shared__ __align__(16) float qs[];  // [body][QT]
  if (active_count) {
    uint32_t ac = *active_count;
    ac = ac > active_base ? ac - active_base : 0;
    if (blockIdx.y * QT >= ac) return;
    nq = min(nq, ac);
  }
  const uint32_t q0 = blockIdx.y * QT;
  exact_dist_stage_queries<QT>(qs, q, nq, ldq, q0, dim & ~7u);
  __syncthreads();
  const uint64_t n_row_blocks = (n + EX_ROWS - 1) / EX_ROWS;
  for (uint64_t rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x)
    exact_dist_rowblock<QT>(qs, q, nq, ldq, q0, x, n, dim, ldx, D, ldD, rb, ce);
}

cudaError_t launch_exact_dist(const float* q, uint32_t nq, uint32_t ldq, const float* x, uint64_t n, uint32_t dim,
                              uint32_t ldx, float* D, uint64_t ldD, const uint32_t* active_count, uint32_t active_base,
                              cudaStream_t s, const ContextEpilogue* ce) {
  if (nq == 0 || n == 0) return cudaSuccess;
  const int qt = (size_t)(dim & ~7u) * 16 * 4 <= 200 * 1024 ? 16 : ((size_t)(dim & ~7u) * 8 * 4 <= 200 * 1024 ? 8 : 4);
  size_t smem = (size_t)(dim & ~7u) * qt * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;  // dim > 12,800
  // D rows are indexed by the query's position inside this launch.
  // The in-stream fallback of the tensor-core scan (active_count != nullptr) is almost always empty: keep its
  // grid small (row blocks are walked with a grid stride) so that an empty launch costs microseconds.
  uint64_t row_blocks = (n + EX_ROWS - 1) / EX_ROWS;
  if (active_count) row_blocks = std::min<uint64_t>(row_blocks, 296);
  dim3 grid((unsigned)row_blocks, (nq + qt - 1) / qt);
  ContextEpilogue none{};
  auto launch = [&](auto kernel) -> cudaError_t {
    // per device, so set on every launch (microseconds)
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 256, smem, s>>>(q, nq, ldq, x, n, dim, ldx, D, ldD, active_count, active_base, ce ? *ce : none);
    return cudaSuccess;
  };
  cudaError_t e = qt == 16 ? launch(k_exact_dist<16>) : (qt == 8 ? launch(k_exact_dist<8>) : launch(k_exact_dist<4>));
  if (e != cudaSuccess) return e;
  count_launch();
  return cudaGetLastError();
}

__global__ void k_gather_rows(const float* __restrict__ src, uint32_t ld, const uint32_t* __restrict__ list,
                              const uint32_t* __restrict__ count, uint32_t cap, float* __restrict__ dst) {
  uint32_t c = min(*count, cap);
  if (blockIdx.x >= c) return;
  const float* s = src + (uint64_t)list[blockIdx.x] * ld;
  float* d = dst + (uint64_t)blockIdx.x * ld;
  for (uint32_t e = threadIdx.x; e < ld; e += blockDim.x) d[e] = s[e];
}
cudaError_t launch_gather_rows(const float* src, uint32_t ld, const uint32_t* list, const uint32_t* count,
                               uint32_t cap, float* dst, cudaStream_t s) {
  k_gather_rows<<<cap, 128, 0, s>>>(src, ld, list, count, cap, dst);
  count_launch();
  return cudaGetLastError();
}

// Per-chunk counters of the tensor-core scan {rescored u64, flagged u32} folded into the call's totals
// {rescored u64, flagged u32, most flagged in one chunk u32}.
__global__ void k_fold_scan_counters(const unsigned char* chunk, unsigned char* total, uint32_t flag_cap, uint32_t* status) {
  const unsigned long long r = *reinterpret_cast<const unsigned long long*>(chunk);
  const uint32_t f = *reinterpret_cast<const uint32_t*>(chunk + 8);
  *reinterpret_cast<unsigned long long*>(total) += r;
  *reinterpret_cast<uint32_t*>(total + 8) += f;
  uint32_t* worst = reinterpret_cast<uint32_t*>(total + 12);
  if (f > *worst) *worst = f;
  if (status) {  // status points at word [1] of the index's status block: [1] unresolved, [2..3] rescored (u64), [4] flagged
    if (f > flag_cap) atomicAdd(status, f - flag_cap);
    atomicAdd(reinterpret_cast<unsigned long long*>(status + 1), r);
    atomicAdd(status + 3, f);
  }
}
cudaError_t launch_fold_scan_counters(const void* chunk, void* total, uint32_t flag_cap, uint32_t* status, cudaStream_t s) {
  k_fold_scan_counters<<<1, 1, 0, s>>>((const unsigned char*)chunk, (unsigned char*)total, flag_cap, status);
  count_launch();
  return cudaGetLastError();
}

__global__ void k_fill(float* p, uint64_t n, float v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
cudaError_t launch_fill(float* p, uint64_t n, float v, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_fill<<<148 * 4, 256, 0, s>>>(p, n, v);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// consolidate's pair test (src/store.rs:531-534): similarity = metric.similarity(a.data, b.data) on the RAW rows —
// cosine (src/metric.rs:163-186) or 0 when undefined. One lane octet per candidate pair of the self-join; the dot
// chain is dot_and_norms_avx2's, the squared norms come precomputed in the same arithmetic.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_similarity(const float* __restrict__ x, uint32_t ld, uint32_t dim,
                                                         const float* __restrict__ norm2,
                                                         const unsigned long long* __restrict__ pairs, uint32_t n_pairs,
                                                         float* __restrict__ sim) {
  uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  uint32_t j = threadIdx.x & 7;
  bool valid = p < n_pairs;
  unsigned long long pr = valid ? pairs[p] : 0ull;
  uint32_t ia = (uint32_t)(pr >> 32), ib = (uint32_t)pr;
  const float* pa = x + (uint64_t)ia * ld;
  const float* pb = x + (uint64_t)ib * ld;
  uint32_t body = dim & ~7u;
  float acc = 0.0f;
  if (valid)
    for (uint32_t e = j; e < body; e += 8) acc = __fmaf_rn(__ldg(pa + e), __ldg(pb + e), acc);
  float dot = octet_hsum(acc);
  if (valid) {
    for (uint32_t e = body; e < dim; ++e) dot = __fadd_rn(dot, __fmul_rn(pa[e], pb[e]));
    if (j == 0) {
      const float denom = __fsqrt_rn(__fmul_rn(norm2[ia], norm2[ib]));
      float c = 0.0f;  // unwrap_or(0.0)
      if (!(denom <= 1.1920929e-7f)) {
        c = __fdiv_rn(dot, denom);
        c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);
      }
      sim[p] = c;
    }
  }
}

cudaError_t launch_pair_similarity(const float* x, uint32_t ld, uint32_t dim, const float* norm2,
                                   const unsigned long long* pairs, uint32_t n_pairs, float* sim, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  uint64_t threads = (uint64_t)n_pairs * 8;
  k_pair_similarity<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(x, ld, dim, norm2, pairs, n_pairs, sim);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// apply_decay sweep (src/store.rs:428-458), one thread per memory: claim the interval (from, now], scale importance
// by exp(-rate * hours), clamp to [0, 1] (scale_importance, :114-128). The CAS gate of the reference protects
// against concurrent sweeps; one launch is one sweep.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_decay_sweep(float* __restrict__ importance, unsigned long long* __restrict__ decayed_through,
                              const unsigned long long* __restrict__ last_access, const float* __restrict__ decay_rate,
                              uint64_t n, float base_rate, unsigned long long now_nanos) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long prev = decayed_through[i];
    const unsigned long long la = last_access[i];
    const unsigned long long from = prev > la ? prev : la;
    if (now_nanos <= from) continue;
    decayed_through[i] = now_nanos;
    const float hours = __fdiv_rn(__fdiv_rn(__ull2float_rn(now_nanos - from), 1.0e9f), 3600.0f);
    const float r = decay_rate[i];
    const float rate = r > 0.0f ? r : base_rate;
    const float factor = (float)exp((double)__fmul_rn(-rate, hours));  // correctly rounded f32::exp
    float v = __fmul_rn(importance[i], factor);
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);                    // f32::clamp(0, 1): NaN passes through
    importance[i] = v;
  }
}

cudaError_t launch_decay_sweep(float* importance, unsigned long long* decayed_through, const unsigned long long* last_access,
                               const float* decay_rate, uint64_t n, float base_rate, unsigned long long now_nanos,
                               cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_decay_sweep<<<148 * 4, 256, 0, s>>>(importance, decayed_through, last_access, decay_rate, n, base_rate, now_nanos);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Top-k select: the k smallest (TotalF32 distance, handle) tuples of one row of D per CTA — the stable
// `sort_by(total_cmp)` + `take(k)` of tests/recall_test.rs:77-78 / tests/property_test.rs:206 without sorting
// the row. Threshold filter + shared-memory candidate buffer + bitonic flushes.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t SEL_THREADS = 256;
constexpr uint32_t SEL_BUF = 2048;
constexpr uint32_t SEL_STEP = SEL_THREADS * 4;

struct SelState {
  unsigned long long tau;
  uint32_t cnt;
};

// Sort keys[0 .. kp+cnt) (padded with MAX to a power of two), keep the first kp, refresh tau = k-th key.
__device__ __forceinline__ void select_flush(unsigned long long* keys, uint32_t kp, uint32_t k, SelState* st) {
  __syncthreads();
  uint32_t total = kp + st->cnt;
  uint32_t P = next_pow2(total);
  for (uint32_t i = total + threadIdx.x; i < P; i += blockDim.x) keys[i] = kKeyMax;
  block_bitonic_sort(keys, P);
  if (threadIdx.x == 0) {
    st->cnt = 0;
    st->tau = keys[k - 1];
  }
  __syncthreads();
}

// The k smallest (distance, handle) of one row of D -> ids/dist[0..k). All SEL_THREADS threads of the CTA.
__device__ __forceinline__ void select_row(const float* __restrict__ row, uint64_t n, const uint8_t* __restrict__ deleted,
                                           uint32_t k, uint32_t kp, unsigned long long* keys, SelState* st,
                                           uint32_t* __restrict__ ids, float* __restrict__ dist, int skip_inf) {
  for (uint32_t i = threadIdx.x; i < kp; i += blockDim.x) keys[i] = kKeyMax;
  if (threadIdx.x == 0) {
    st->cnt = 0;
    st->tau = kKeyMax;
  }
  __syncthreads();
  for (uint64_t base = 0; base < n; base += SEL_STEP) {
    unsigned long long tau = st->tau;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint64_t i = base + c * SEL_THREADS + threadIdx.x;
      if (i < n) {
        unsigned long long key = pack_key(row[i], (uint32_t)i);
        if (key < tau && !(deleted && deleted[i]) && !(skip_inf && (uint32_t)(key >> 32) == 0xFF800000u)) {  // +inf
          uint32_t slot = atomicAdd(&st->cnt, 1u);
          keys[kp + slot] = key;
        }
      }
    }
    __syncthreads();
    uint32_t pending = st->cnt;  // read between two barriers: no thread is appending
    __syncthreads();
    if (pending > SEL_BUF - SEL_STEP) select_flush(keys, kp, k, st);
  }
  select_flush(keys, kp, k, st);
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    unsigned long long key = keys[i];
    bool ok = key != kKeyMax;
    ids[i] = ok ? (uint32_t)key : CM_NO_ID_DEV;
    dist[i] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
  }
}

__global__ void __launch_bounds__(SEL_THREADS) k_select_from_dist(const float* __restrict__ D, uint64_t ldD, uint64_t n,
                                                                  const uint8_t* __restrict__ deleted, uint32_t k,
                                                                  uint32_t kp, uint32_t* __restrict__ ids,
                                                                  float* __restrict__ dist, uint32_t out_ld,
                                                                  const uint32_t* __restrict__ qmap,
                                                                  const uint32_t* __restrict__ active_count,
                                                                  uint32_t active_base, int skip_inf) {
  extern __shared__ __align__(16) unsigned long long keys[];  // next_pow2(kp + SEL_BUF)
  __shared__ SelState st;
  if (active_count && active_base + blockIdx.x >= *active_count) return;
  const uint32_t out_row = qmap ? qmap[blockIdx.x] : blockIdx.x;
  select_row(D + (uint64_t)blockIdx.x * ldD, n, deleted, k, kp, keys, &st, ids + (uint64_t)out_row * out_ld,
             dist + (uint64_t)out_row * out_ld, skip_inf);
}

cudaError_t launch_select_from_dist(const float* D, uint64_t ldD, uint32_t nq, uint64_t n, const uint8_t* deleted,
                                    uint32_t k, uint32_t* ids, float* dist, uint32_t out_ld, const uint32_t* qmap,
                                    const uint32_t* active_count, uint32_t active_base, cudaStream_t s, bool skip_inf) {
  if (nq == 0) return cudaSuccess;
  uint32_t kp = next_pow2(k < 32 ? 32 : k);
  size_t smem = (size_t)next_pow2(kp + SEL_BUF) * sizeof(unsigned long long);
  k_select_from_dist<<<nq, SEL_THREADS, smem, s>>>(D, ldD, n, deleted, k, kp, ids, dist, out_ld, qmap, active_count,
                                                   active_base, skip_inf ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// In-stream fp32 fallback of the tensor-core scan: ONE cooperative launch that re-runs every flagged query through
// the bit-exact fp32 scan + select, however many there are, without a host round trip — and costs one empty launch
// (a few microseconds) in the normal case of none. Per chunk of FB_QT flagged queries: gather their prepared rows,
// all CTAs fill D[FB_QT][n] (row blocks grid-strided), grid barrier, one CTA per query selects its k smallest,
// grid barrier.
// ---------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int FB_QT = 16;

struct FallbackArgs {
  const float* qn;            // all prepared queries of the call [nq][ldq]
  uint32_t ldq;
  const uint32_t* flag_list;  // [*flag_count] query indices
  const uint32_t* flag_count;
  float* qc;                  // compact copy of one chunk's queries [FB_QT][ldq]
  const float* x;
  uint64_t n;
  uint32_t dim, ldx;
  float* D;                   // [FB_QT][ldD]
  uint64_t ldD;
  const uint8_t* deleted;
  uint32_t k, kp;
  uint32_t* ids;              // [nq][k] — rows of the flagged queries are overwritten
  float* dist;
};

__global__ void __launch_bounds__(256) k_fallback_exact(FallbackArgs a) {
  extern __shared__ __align__(16) unsigned char fb_smem[];  // max(query tile, select keys)
  __shared__ SelState st;
  const uint32_t fc = *a.flag_count;
  if (fc == 0) return;  // uniform over the grid: nobody reaches a grid barrier
  cg::grid_group grid = cg::this_grid();
  const uint32_t body = a.dim & ~7u;
  const uint64_t n_row_blocks = (a.n + EX_ROWS - 1) / EX_ROWS;
  const ContextEpilogue none{};
  for (uint32_t c0 = 0; c0 < fc; c0 += FB_QT) {
    const uint32_t nqc = min((uint32_t)FB_QT, fc - c0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nqc * a.ldq; i += gridDim.x * blockDim.x)
      a.qc[i] = a.qn[(uint64_t)a.flag_list[c0 + i / a.ldq] * a.ldq + i % a.ldq];
    grid.sync();
    float* qs = reinterpret_cast<float*>(fb_smem);
    exact_dist_stage_queries<FB_QT>(qs, a.qc, nqc, a.ldq, 0, body);
    __syncthreads();
    for (uint64_t rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x)
      exact_dist_rowblock<FB_QT>(qs, a.qc, nqc, a.ldq, 0, a.x, a.n, a.dim, a.ldx, a.D, a.ldD, rb, none);
    grid.sync();
    if (blockIdx.x < nqc) {
      const uint32_t q = a.flag_list[c0 + blockIdx.x];
      select_row(a.D + (uint64_t)blockIdx.x * a.ldD, a.n, a.deleted, a.k, a.kp,
                 reinterpret_cast<unsigned long long*>(fb_smem), &st, a.ids + (uint64_t)q * a.k, a.dist + (uint64_t)q * a.k, 0);
    }
    grid.sync();
  }
}

size_t fallback_exact_dist_bytes(uint64_t n) { return (size_t)FB_QT * ((n + 3) & ~3ull) * sizeof(float); }
size_t fallback_exact_query_bytes(uint32_t ldq) { return (size_t)FB_QT * ldq * sizeof(float); }

cudaError_t launch_fallback_exact(const float* qn, uint32_t ldq, const uint32_t* flag_list, const uint32_t* flag_count,
                                  float* qc, const float* x, uint64_t n, uint32_t dim, uint32_t ldx, float* D,
                                  const uint8_t* deleted, uint32_t k, uint32_t* ids, float* dist, int sm_count,
                                  cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  FallbackArgs a;
  a.qn = qn;
  a.ldq = ldq;
  a.flag_list = flag_list;
  a.flag_count = flag_count;
  a.qc = qc;
  a.x = x;
  a.n = n;
  a.dim = dim;
  a.ldx = ldx;
  a.D = D;
  a.ldD = (n + 3) & ~3ull;
  a.deleted = deleted;
  a.k = k;
  a.kp = next_pow2(k < 32 ? 32 : k);
  a.ids = ids;
  a.dist = dist;
  const size_t smem = std::max((size_t)(dim & ~7u) * FB_QT * sizeof(float),
                               (size_t)next_pow2(a.kp + SEL_BUF) * sizeof(unsigned long long));
  if (smem > 200 * 1024) return cudaErrorInvalidValue;  // rows too long for a 16-query tile: the caller keeps the fp32 engine
  cudaError_t e = cudaFuncSetAttribute(k_fallback_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fallback_exact, 256, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::max(sm_count * std::min(per_sm, 2), FB_QT);  // every CTA co-resident (grid barrier)
  if ((int)grid > sm_count * per_sm) return cudaErrorInvalidValue;
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((void*)k_fallback_exact, dim3(grid), dim3(256), args, smem, s);
  count_launch();
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// K5 shard merge — ShardedRwLockHnsw::search's gather step (src/index/sharded_rwlock.rs:81-94): concat the
// per-shard lists, sort by (distance, global handle), truncate. global = local * n_shards + shard.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_merge_topk(const uint32_t* __restrict__ ids, const float* __restrict__ dist,
                                                    const unsigned long long* __restrict__ packed, uint32_t n_shards,
                                                    uint32_t nq, uint32_t len, uint32_t out_len, uint32_t P,
                                                    uint32_t* __restrict__ out_ids, float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned long long keys[];  // P >= n_shards * len
  const uint32_t qi = blockIdx.x;
  const uint32_t total = n_shards * len;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long key = kKeyMax;
    if (i < total) {
      uint32_t sh = i / len, jj = i % len;
      uint64_t o = ((uint64_t)sh * nq + qi) * len + jj;
      if (packed) {  // (TotalF32 distance << 32 | local handle) as cm_pack_topk_dev leaves it; all-ones = padding
        const unsigned long long pk = packed[o];
        if (pk != kKeyMax) key = (pk & 0xFFFFFFFF00000000ull) | ((uint32_t)pk * n_shards + sh);
      } else {
        uint32_t local = ids[o];
        if (local != CM_NO_ID_DEV) key = pack_key(dist[o], local * n_shards + sh);
      }
    }
    keys[i] = key;
  }
  block_bitonic_sort(keys, P);
  for (uint32_t i = threadIdx.x; i < out_len; i += blockDim.x) {
    unsigned long long key = i < P ? keys[i] : kKeyMax;
    bool ok = key != kKeyMax;
    out_ids[(uint64_t)qi * out_len + i] = ok ? (uint32_t)key : CM_NO_ID_DEV;
    out_dist[(uint64_t)qi * out_len + i] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
  }
}

cudaError_t launch_merge_topk(const uint32_t* ids, const float* dist, const unsigned long long* packed, uint32_t n_shards,
                              uint32_t nq, uint32_t len, uint32_t out_len, uint32_t* out_ids, float* out_dist,
                              cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  uint32_t P = next_pow2(n_shards * len < 32 ? 32 : n_shards * len);
  size_t smem = (size_t)P * sizeof(unsigned long long);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  {  // per device, so set on every launch (microseconds)
    cudaError_t e = cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  k_merge_topk<<<nq, P >= 512 ? 256 : 64, smem, s>>>(ids, dist, packed, n_shards, nq, len, out_len, P, out_ids, out_dist);
  count_launch();
  return cudaGetLastError();
}

// (distance, local handle) pairs -> one sortable u64 each, so that a sharded search exchanges ONE array per step.
__global__ void k_pack_topk(const uint32_t* __restrict__ ids, const float* __restrict__ dist, uint64_t n,
                            unsigned long long* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = ids[i] == CM_NO_ID_DEV ? kKeyMax : pack_key(dist[i], ids[i]);
}
cudaError_t launch_pack_topk(const uint32_t* ids, const float* dist, uint64_t n, unsigned long long* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_pack_topk<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(ids, dist, n, out);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// K4 temporal rerank — ChronoMind::search's tail (src/store.rs:327-344) with combined_score (:395-415).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rerank(const uint32_t* __restrict__ pool_ids, const float* __restrict__ pool_dist,
                                                uint32_t pool, uint32_t P, uint32_t k,
                                                const unsigned long long* __restrict__ ts_secs,
                                                const uint32_t* __restrict__ ts_nanos, const float* __restrict__ decay,
                                                float w, float base, unsigned long long now_secs, uint32_t now_nanos,
                                                uint32_t* __restrict__ out_ids, float* __restrict__ out_score,
                                                float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned long long keys[];  // P
  const uint64_t qo = (uint64_t)blockIdx.x * pool;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long key = kKeyMax;
    if (i < pool) {
      uint32_t id = pool_ids[qo + i];
      if (id != CM_NO_ID_DEV) {
        float sc = combined_score_dev(pool_dist[qo + i], ts_secs[id], ts_nanos[id], decay[id], now_secs, now_nanos, w, base);
        // Stable sort by score: the pool rank breaks ties (pool is ascending (distance, handle)).
        key = ((unsigned long long)total_key(sc) << 32) | i;
      }
    }
    keys[i] = key;
  }
  block_bitonic_sort(keys, P);
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    unsigned long long key = i < P ? keys[i] : kKeyMax;
    bool ok = key != kKeyMax;
    uint32_t rank = (uint32_t)key;
    uint64_t o = (uint64_t)blockIdx.x * k + i;
    out_ids[o] = ok ? pool_ids[qo + rank] : CM_NO_ID_DEV;
    out_score[o] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
    if (out_dist) out_dist[o] = ok ? pool_dist[qo + rank] : __int_as_float(0x7F800000);
  }
}

cudaError_t launch_rerank(const uint32_t* pool_ids, const float* pool_dist, uint32_t nq, uint32_t pool, uint32_t k,
                          const uint64_t* ts_secs, const uint32_t* ts_nanos, const float* decay, float w, float base,
                          uint64_t now_secs, uint32_t now_nanos, uint32_t* out_ids, float* out_score,
                          float* out_dist, cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  uint32_t P = next_pow2(pool < 32 ? 32 : pool);
  if ((size_t)P * 8 > 48 * 1024) {  // pools beyond 4096 need the opt-in shared-memory window (per device)
    cudaError_t e = cudaFuncSetAttribute(k_rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)P * 8));
    if (e != cudaSuccess) return e;
  }
  k_rerank<<<nq, 128, (size_t)P * 8, s>>>(pool_ids, pool_dist, pool, P, k, (const unsigned long long*)ts_secs, ts_nanos,
                                          decay, w, base, now_secs, now_nanos, out_ids, out_score, out_dist);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cm
