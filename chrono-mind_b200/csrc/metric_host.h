// metric_host.h — host-side CosineDistance arithmetic used by the graph builder and by index creation.
// Same operation order as src/metric.rs so that the host-built graph and the device-side distances agree
// bit for bit with the reference's: 8-lane FMA accumulate, lo+hi / movehl / lane0+lane1 reduction, scalar
// tail (src/metric.rs:129-147); sequential sum-of-squares, sqrt, true division (src/metric.rs:208-214).
#pragma once
#include <immintrin.h>

#include <cmath>
#include <cstdint>
#include <cstring>

namespace cm {

static inline float hsum_ref_order(__m256 v) {
  __m128 s4 = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
  __m128 s2 = _mm_add_ps(s4, _mm_movehl_ps(s4, s4));
  __m128 s1 = _mm_add_ss(s2, _mm_shuffle_ps(s2, s2, 1));
  return _mm_cvtss_f32(s1);
}

static inline float clamp_pass_nan(float x, float lo, float hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}

// distance_prepared for equal, non-zero lengths: clamp(1 - dot, 0, 2).
static inline float dist_prepared(const float* a, const float* b, uint32_t n) {
  __m256 acc = _mm256_setzero_ps();
  uint32_t body = n & ~7u;
  for (uint32_t i = 0; i < body; i += 8) acc = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), acc);
  float s = hsum_ref_order(acc);
  for (uint32_t i = body; i < n; ++i) s += a[i] * b[i];
  return clamp_pass_nan(1.0f - s, 0.0f, 2.0f);
}

// Four distances against one query with independent accumulator chains (hides the 4-cycle FMA latency);
// each result is bit-identical to dist_prepared.
static inline void dist_prepared_x4(const float* q, const float* const* rows, uint32_t n, float* out) {
  __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0;
  uint32_t body = n & ~7u;
  for (uint32_t i = 0; i < body; i += 8) {
    __m256 vq = _mm256_loadu_ps(q + i);
    a0 = _mm256_fmadd_ps(_mm256_loadu_ps(rows[0] + i), vq, a0);
    a1 = _mm256_fmadd_ps(_mm256_loadu_ps(rows[1] + i), vq, a1);
    a2 = _mm256_fmadd_ps(_mm256_loadu_ps(rows[2] + i), vq, a2);
    a3 = _mm256_fmadd_ps(_mm256_loadu_ps(rows[3] + i), vq, a3);
  }
  float s[4] = {hsum_ref_order(a0), hsum_ref_order(a1), hsum_ref_order(a2), hsum_ref_order(a3)};
  for (int r = 0; r < 4; ++r) {
    for (uint32_t i = body; i < n; ++i) s[r] += rows[r][i] * q[i];
    out[r] = clamp_pass_nan(1.0f - s[r], 0.0f, 2.0f);
  }
}

}  // namespace cm
