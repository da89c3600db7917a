// row_gather.cuh — gather corpus rows with the copy engine and evaluate distance_prepared out of shared memory.
//
// Used by the HNSW traversal (K3) and the exact scan's rescore (K2): both evaluate the reference's exact fp32
// distance for a handful of RANDOM rows per step. A register-staged gather keeps only a few loads per thread in
// flight (a dozen dependent DRAM round trips per 3 KB row); one `cp.async.bulk` (UBLKCP) per row lands the whole
// row in shared memory, all rows of a round in flight at once, at no register cost.
#pragma once
#include "device_common.cuh"

namespace cm {

constexpr uint32_t RG_ROW_PAD = 32;  // bytes; row stride = 32 (mod 128) => conflict-free LDS.64 across a warp's 8 rows

__device__ __forceinline__ uint32_t rg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rg_mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rg_smem_u32(bar)), "r"(1u) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Bounded wait on the copy barrier: a lost completion traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void rg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

// Stage rows x[ids[i]] (i < cnt <= blockDim.x) into rowbuf (stride row_stride bytes, 16-byte aligned) with one bulk
// copy per row and wait for all of them. Called by ALL threads of the CTA after a barrier that ordered the previous
// reads of rowbuf; `phase` is each thread's private copy of the barrier parity. ldx = row length in floats, a
// multiple of 8 (32-byte rows).
__device__ __forceinline__ void rg_stage_rows(const float* x, uint32_t ldx, unsigned long long* mbar,
                                              unsigned char* rowbuf, uint32_t row_stride, const uint32_t* ids,
                                              uint32_t cnt, uint32_t& phase) {
  const uint32_t tid = threadIdx.x;
  const uint32_t row_bytes = ldx * 4;
  const uint32_t bar = rg_smem_u32(mbar);
  if (tid == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * row_bytes) : "memory");
  if (tid < cnt) {
    // rowbuf was last read through the generic proxy: order those reads before the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const float* src = x + (uint64_t)ids[tid] * ldx;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(rg_smem_u32(rowbuf + (size_t)tid * row_stride)), "l"(src), "r"(row_bytes), "r"(bar)
        : "memory");
  }
  rg_mbar_wait(bar, phase);
  phase ^= 1;
}

// clamp(1 - dot_avx2(row, q), 0, 2) (src/metric.rs:129-147, 219-224) out of shared memory by one thread QUAD (4
// consecutive lanes); thread c of the quad carries the reference's SIMD lanes 2c, 2c+1, i.e. two FMA chains over
// elements 8t+2c, 8t+2c+1. All four threads return the same bits. Must be called by whole quads (inactive quads
// pass active = false and still take part in the shuffles).
__device__ __forceinline__ float rg_quad_dot(const float* qs, const unsigned char* row, uint32_t dim, uint32_t c, bool active);
__device__ __forceinline__ float rg_quad_distance(const float* qs, const unsigned char* row, uint32_t dim, uint32_t c,
                                                  bool active) {
  return clamp02(__fsub_rn(1.0f, rg_quad_dot(qs, row, dim, c, active)));
}

// dot_avx2(row, q) (src/metric.rs:129-147) by one thread quad — also the `dot` accumulator of dot_and_norms_avx2
// (:78-116: same FMA chain, same horizontal sum, same scalar tail). All four threads return the same bits.
__device__ __forceinline__ float rg_quad_dot(const float* qs, const unsigned char* row, uint32_t dim, uint32_t c,
                                             bool active) {
  const uint32_t body = dim & ~7u;
  const float* xr = reinterpret_cast<const float*>(row);
  float a0 = 0.0f, a1 = 0.0f;
  if (active) {
    const float2* xv = reinterpret_cast<const float2*>(xr) + c;
    const float2* qv = reinterpret_cast<const float2*>(qs) + c;
#pragma unroll 8
    for (uint32_t t = 0; t < body / 8; ++t) {
      const float2 x = xv[4 * t];
      const float2 q = qv[4 * t];
      a0 = __fmaf_rn(x.x, q.x, a0);
      a1 = __fmaf_rn(x.y, q.y, a1);
    }
  }
  // (a0+a4, a1+a5, a2+a6, a3+a7) -> (s0+s2, s1+s3) -> t0+t1 (src/metric.rs:137-141). Thread c holds lanes
  // (2c, 2c+1): lane j pairs with j+4 across threads c and c^2, then (s0,s1)|(s2,s3) across threads c and c^1.
  // IEEE addition commutes, so all four threads end with the same bits.
  const float s_lo = __fadd_rn(a0, __shfl_xor_sync(0xFFFFFFFFu, a0, 2));
  const float s_hi = __fadd_rn(a1, __shfl_xor_sync(0xFFFFFFFFu, a1, 2));
  const float t0 = __fadd_rn(s_lo, __shfl_xor_sync(0xFFFFFFFFu, s_lo, 1));  // s0 + s2
  const float t1 = __fadd_rn(s_hi, __shfl_xor_sync(0xFFFFFFFFu, s_hi, 1));  // s1 + s3
  float sum = __fadd_rn(t0, t1);
  if (active)
    for (uint32_t e = body; e < dim; ++e) sum = __fadd_rn(sum, __fmul_rn(xr[e], qs[e]));
  return sum;
}

}  // namespace cm
