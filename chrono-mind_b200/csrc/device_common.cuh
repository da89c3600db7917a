// device_common.cuh — device helpers shared by the kernels.
//
// Arithmetic contract: everything that feeds a distance or a score replays the reference's fp32 operation
// order with explicitly rounded intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn), so
// nvcc's default FMA contraction cannot change a single bit relative to src/metric.rs and src/store.rs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cm_internal.h"

namespace cm {

#define CM_NO_ID_DEV 0xFFFFFFFFu
constexpr unsigned long long kKeyMax = 0xFFFFFFFFFFFFFFFFull;

// total_key(): f32::total_cmp as an unsigned key — defined __host__ __device__ in cm_internal.h.
__device__ __forceinline__ float key_to_float(uint32_t t) {
  uint32_t b = (t & 0x80000000u) ? (t & 0x7FFFFFFFu) : ~t;
  return __uint_as_float(b);
}
// (TotalF32, u32) tuple packed so that u64 order == tuple order (src/index/lockfree_hnsw.rs:149-150).
__device__ __forceinline__ unsigned long long pack_key(float d, uint32_t id) {
  return ((unsigned long long)total_key(d) << 32) | id;
}

// f32::clamp(0, 2) — NaN passes through (src/metric.rs:223).
__device__ __forceinline__ float clamp02(float x) {
  if (x < 0.0f) return 0.0f;
  if (x > 2.0f) return 2.0f;
  return x;
}

// The reference's horizontal sum (src/metric.rs:137-141) over 8 lane-accumulators held by 8 consecutive
// lanes (lane & 7 == accumulator index): (a0+a4, a1+a5, a2+a6, a3+a7) -> (s0+s2, s1+s3) -> t0+t1.
// IEEE addition is commutative, so every lane of the octet ends with the same bits.
__device__ __forceinline__ float octet_hsum(float v) {
  v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, 4));
  v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, 2));
  v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, 1));
  return v;
}

// Temporal half of combined_score (src/store.rs:395-415): w * (1 - exp(-rate * age_hours)), every operation rounded
// like the reference's f32 code.
__device__ __forceinline__ float temporal_term_dev(unsigned long long ts_secs, uint32_t ts_nanos, float decay_rate,
                                                   unsigned long long now_secs, uint32_t now_nanos, float w,
                                                   float base_decay_rate) {
  // now.duration_since(ts).unwrap_or_default()
  unsigned long long dsecs = 0;
  uint32_t dnanos = 0;
  bool later = (now_secs > ts_secs) || (now_secs == ts_secs && now_nanos >= ts_nanos);
  if (later) {
    dsecs = now_secs - ts_secs;
    if (now_nanos >= ts_nanos) {
      dnanos = now_nanos - ts_nanos;
    } else {
      dsecs -= 1;
      dnanos = now_nanos + 1000000000u - ts_nanos;
    }
  }
  // Duration::as_secs_f32 = secs as f32 + nanos as f32 / 1e9
  float age_secs = __fadd_rn(__ull2float_rn(dsecs), __fdiv_rn(__uint2float_rn(dnanos), 1.0e9f));
  float age_hours = __fdiv_rn(age_secs, 3600.0f);
  float rate = decay_rate > 0.0f ? decay_rate : base_decay_rate;
  // f32::exp: evaluated in double and rounded once (correctly rounded fp32 result).
  float temporal_relevance = (float)exp((double)__fmul_rn(-rate, age_hours));
  return __fmul_rn(w, __fsub_rn(1.0f, temporal_relevance));
}

// combined_score (src/store.rs:395-415), shared by the rerank (K4) and the context scans.
__device__ __forceinline__ float combined_score_dev(float distance, unsigned long long ts_secs, uint32_t ts_nanos,
                                                    float decay_rate, unsigned long long now_secs,
                                                    uint32_t now_nanos, float w, float base_decay_rate) {
  float geo = __fmul_rn(__fsub_rn(1.0f, w), __fdiv_rn(distance, 2.0f));
  float tmp = temporal_term_dev(ts_secs, ts_nanos, decay_rate, now_secs, now_nanos, w, base_decay_rate);
  return __fadd_rn(geo, tmp);
}

// Ascending bitonic sort of n (power of two) u64 keys in shared memory by all threads of the block.
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* s, uint32_t n) {
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        uint32_t lo = 2 * i - (i & (stride - 1));
        uint32_t hi = lo + stride;
        bool up = (lo & size) == 0;
        unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

__host__ __device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace cm
