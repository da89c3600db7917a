// scan_tc.cu — K1/K2: exact brute-force kNN as a dense contraction on the 5th-gen tensor cores.
//
// Replaces the all-pairs loop of the reference's brute force (tests/recall_test.rs:70-79,
// tests/property_test.rs:201-206, src/store.rs:362-376) for large batches:
//
//   K1  scan_tc_kernel   S = Q · Xᵀ in bf16 on tcgen05 (fp32 accumulators in TMEM, operands staged by TMA into
//                        128B-swizzled shared memory). S is never written: the epilogue warps read each
//                        accumulator tile straight out of TMEM and keep, per query, only the rows whose bf16
//                        score can still be among the exact top-k (threshold select, see below).
//   K2  rescore_kernel   the few surviving rows per query are re-scored with the reference's exact fp32
//                        arithmetic (dot_avx2 order, src/metric.rs:129-147) and sorted by (distance, handle).
//
// Why the result is EXACT (not "bf16-approximate"): let e_r be the fp32 reference dot of row r and s_r the
// tensor-core bf16 score, |s_r - e_r| <= eps for unit vectors (bf16 rounding of both operands: 2·2^-9 relative
// per product, Cauchy-Schwarz => <= 2^-8 absolute; eps = kEps adds slack for accumulation order). The k-th
// largest s is <= (k-th largest e) + eps, so every row of the exact top-k has s_r >= s_(k) - 2·eps. The
// epilogue's threshold tau_q is always a LOWER bound of s_(k) - 2·eps (it is derived from k scores actually
// seen for that query), so no row of the exact top-k is ever filtered. K2 then orders the survivors with the
// exact arithmetic. If a query's candidate buffer overflows, or fewer than k candidates survive (NaN/Inf data),
// K2 flags the query and it is re-run through the fp32 scan — never silently degraded.
//
// Kernel organisation (one CTA per SM, persistent, 8 warps):
//   warp 0  TMA producer     (one lane)  — 4-stage ring of {Q tile 128x64, X tile 256x64} bf16, 48 KB per stage
//   warp 1  MMA issuer       (one lane)  — tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, 4 per stage
//   warp 2  TMEM allocator               — 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4-7 epilogue (128 threads)     — thread = one query row (TMEM lane); tcgen05.ld 32 columns at a time
// Work unit = (query tile m, strip of kStripTiles X tiles), ordered m-fastest so that the CTAs in flight stream
// the same few MB of X at the same time (X is read from HBM once per strip and hit in L2 by the other query tiles).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"
#include "scan_tc.h"

namespace cm {

constexpr uint32_t BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr uint32_t kStages = 4, kAccStages = 2;
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kStripTiles = 16;
constexpr uint32_t kThreads = 256;
constexpr float kEps = 0.0042f;  // >= 2^-8 + accumulation slack; margin is 2 * kEps

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a broken pipeline traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version=1 @46 | layout=2 @61).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 @4, A=BF16 @7, B=BF16 @10, K-major both,
// N>>3 @17, M>>4 @24.
__device__ __forceinline__ constexpr uint32_t make_idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct ScanArgs {
  uint32_t nq, m_tiles, n_tiles, k_blocks, k;
  uint32_t strip_begin, strip_end;  // this launch covers strips [strip_begin, strip_end)
  uint64_t n;
  const uint8_t* deleted;
  uint32_t* tau_key;             // [m_tiles*128] ordered-uint threshold per query (0 = none yet)
  uint32_t* cand_cnt;            // [m_tiles*128]
  unsigned long long* cand;      // [m_tiles*128][cap]  (ordered score key << 32 | row)
  uint32_t cap;
};

// Shared-memory carve-up (dynamic): [stages x (A|B)] 1024-aligned | local top-k lists | barriers | tmem ptr
struct ScanSmemTail {
  unsigned long long full[kStages], empty[kStages], tmem_full[kAccStages], tmem_empty[kAccStages];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, ScanArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries: align by hand (the launch reserves the slack).
  uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  float* lists = reinterpret_cast<float*>(tiles + kStages * kStageBytes);  // [k][128]
  ScanSmemTail* tail = reinterpret_cast<ScanSmemTail*>(tiles + kStages * kStageBytes + (size_t)a.k * 128 * 4);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t units = (a.strip_end - a.strip_begin) * a.m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 1);
    }
    for (uint32_t s = 0; s < kAccStages; ++s) {
      mbar_init(smem_u32(&tail->tmem_full[s]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[s]), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =========================================================================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
        const uint32_t m = u % a.m_tiles, strip = a.strip_begin + u / a.m_tiles;
        const uint32_t t0 = strip * kStripTiles, t1 = min(a.n_tiles, t0 + kStripTiles);
        for (uint32_t t = t0; t < t1; ++t) {
          for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
            mbar_wait(smem_u32(&tail->empty[stage]), phase ^ 1);
            const uint32_t bar = smem_u32(&tail->full[stage]);
            mbar_expect_tx(bar, kStageBytes);
            const uint32_t sa = smem_u32(tiles + stage * kStageBytes);
            tma_load_2d(sa, &map_q, bar, (int32_t)(kb * BK), (int32_t)(m * BM));
            tma_load_2d(sa + kStageBytesA, &map_x, bar, (int32_t)(kb * BK), (int32_t)(t * BN));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===========================================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
        const uint32_t strip = a.strip_begin + u / a.m_tiles;
        const uint32_t t0 = strip * kStripTiles, t1 = min(a.n_tiles, t0 + kStripTiles);
        for (uint32_t t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(&tail->tmem_empty[acc]), acc_phase ^ 1);  // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
            mbar_wait(smem_u32(&tail->full[stage]), phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + stage * kStageBytes);
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kStageBytesA);
#pragma unroll
            for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
              // advancing 16 bf16 (32 bytes) along K inside the 128-byte swizzle row: +2 in the >>4 start field
              umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
            }
            umma_commit(smem_u32(&tail->empty[stage]));  // frees the smem stage when these MMAs retire
            if (kb + 1 == a.k_blocks) umma_commit(smem_u32(&tail->tmem_full[acc]));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++acc == kAccStages) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: threshold select straight out of TMEM =======================================================
    const uint32_t ew = warp & 3;              // TMEM lane quarter this warp may access
    const uint32_t row_in_tile = ew * 32 + lane;
    const uint32_t et = (warp - 4) * 32 + lane;  // 0..127 index into the per-thread lists
    float* my_list = lists + et;               // my_list[i * 128]
    const uint32_t k = a.k;
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
      const uint32_t m = u % a.m_tiles, strip = a.strip_begin + u / a.m_tiles;
      const uint32_t t0 = strip * kStripTiles, t1 = min(a.n_tiles, t0 + kStripTiles);
      const uint32_t q = m * BM + row_in_tile;
      const bool q_valid = q < a.nq;
      // Per-strip local state: the k best scores this thread has seen that reached the emission threshold.
      uint32_t filled = 0;
      float local_kth = -INFINITY;   // k-th best of the local list once it holds k entries
      uint32_t min_pos = 0;
      float tau = INFINITY;          // emission threshold; +inf for padding queries (nothing passes)
      for (uint32_t t = t0; t < t1; ++t) {
        if (q_valid) {
          uint32_t tk = a.tau_key[q];
          float g = tk ? key_to_float(tk) : -INFINITY;
          tau = fmaxf(g, local_kth - 2.0f * kEps);
        }
        mbar_wait(smem_u32(&tail->tmem_full[acc]), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((ew * 32) << 16) + acc * BN;
        const uint64_t row0 = (uint64_t)t * BN;
        bool improved = false;
#pragma unroll 1
        for (uint32_t c = 0; c < BN; c += 32) {
          uint32_t r[32];
          __syncwarp();  // the slow path below diverges; tcgen05.ld is warp-collective
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          if (mx >= tau) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = __uint_as_float(r[i]);
              if (s >= tau) {
                const uint64_t row = row0 + c + i;
                if (row < a.n && !(a.deleted && a.deleted[row])) {
                  uint32_t slot = atomicAdd(&a.cand_cnt[q], 1u);
                  if (slot < a.cap) a.cand[(uint64_t)q * a.cap + slot] = ((unsigned long long)total_key(s) << 32) | (uint32_t)row;
                  if (filled < k) {
                    my_list[filled * 128] = s;
                    ++filled;
                    if (filled == k) {
                      local_kth = my_list[0];
                      min_pos = 0;
                      for (uint32_t j = 1; j < k; ++j) {
                        float v = my_list[j * 128];
                        if (v < local_kth) local_kth = v, min_pos = j;
                      }
                      improved = true;
                    }
                  } else if (s > local_kth) {
                    my_list[min_pos * 128] = s;
                    local_kth = my_list[0];
                    min_pos = 0;
                    for (uint32_t j = 1; j < k; ++j) {
                      float v = my_list[j * 128];
                      if (v < local_kth) local_kth = v, min_pos = j;
                    }
                    improved = true;
                  }
                  if (improved) tau = fmaxf(tau, local_kth - 2.0f * kEps);
                }
              }
            }
          }
        }
        // accumulator drained: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tail->tmem_empty[acc]));
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
        // publish the lower bound s_(k) - 2 eps for the other CTAs working on this query
        if (improved && q_valid) atomicMax(&a.tau_key[q], total_key(local_kth - 2.0f * kEps));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2 rescore: one CTA per query. Keep the candidates whose bf16 score reaches the FINAL threshold, evaluate
// distance_prepared exactly (octet per candidate, reference lane order), sort by (distance, handle).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t RS_THREADS = 256;
constexpr uint32_t RS_MAX = 1024;  // survivors handled per query

struct RescoreArgs {
  const float* qn;
  uint32_t ldq, nq, k, dim, ldx;
  const float* x;
  uint64_t n_live;
  const uint32_t* tau_key;
  const uint32_t* cand_cnt;
  const unsigned long long* cand;
  uint32_t cap;
  uint32_t* ids;
  float* dist;
  uint32_t* flag_count;   // queries that need the fp32 fallback
  uint32_t* flag_list;    // [flag_cap]
  uint32_t flag_cap;
  unsigned long long* rescored_total;
};

__global__ void __launch_bounds__(RS_THREADS) rescore_kernel(RescoreArgs a) {
  __shared__ __align__(16) unsigned long long keys[RS_MAX];
  __shared__ uint32_t surv[RS_MAX];
  __shared__ uint32_t n_surv;
  extern __shared__ __align__(16) float qs[];  // dim floats
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t cnt_raw = a.cand_cnt[q];
  const uint32_t cnt = min(cnt_raw, a.cap);
  const uint32_t tk = a.tau_key[q];
  if (tid == 0) n_surv = 0;
  for (uint32_t e = tid; e < a.dim; e += RS_THREADS) qs[e] = a.qn[(uint64_t)q * a.ldq + e];
  __syncthreads();
  for (uint32_t i = tid; i < cnt; i += RS_THREADS) {
    unsigned long long c = a.cand[(uint64_t)q * a.cap + i];
    if ((uint32_t)(c >> 32) >= tk) {
      uint32_t slot = atomicAdd(&n_surv, 1u);
      if (slot < RS_MAX) surv[slot] = (uint32_t)c;
    }
  }
  __syncthreads();
  const uint32_t ns = n_surv;
  const uint64_t need = a.n_live < a.k ? a.n_live : a.k;
  const bool bad = cnt_raw > a.cap || ns > RS_MAX || ns < need;
  if (bad) {
    if (tid == 0) {
      uint32_t slot = atomicAdd(a.flag_count, 1u);
      if (slot < a.flag_cap) a.flag_list[slot] = q;
    }
    // The fp32 fallback overwrites this row; if more queries are flagged than it absorbs the row stays
    // visibly invalid (CM_NO_ID / NaN) — never a plausible-looking wrong answer.
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
      a.ids[(uint64_t)q * a.k + i] = CM_NO_ID_DEV;
      a.dist[(uint64_t)q * a.k + i] = __int_as_float(0x7FC00000);
    }
    return;
  }
  const uint32_t P = next_pow2(ns < 32 ? 32 : ns);
  for (uint32_t i = tid; i < P; i += RS_THREADS) keys[i] = kKeyMax;
  __syncthreads();
  const uint32_t body = a.dim & ~7u;
  const uint32_t j = tid & 7, oct = tid >> 3;
  for (uint32_t base = 0; base < ns; base += RS_THREADS / 8) {
    uint32_t i = base + oct;
    bool active = i < ns;
    uint32_t row = active ? surv[i] : 0;
    const float* xr = a.x + (uint64_t)row * a.ldx;
    float acc = 0.0f;
    if (active) {
#pragma unroll 8
      for (uint32_t e = j; e < body; e += 8) acc = __fmaf_rn(__ldg(xr + e), qs[e], acc);
    }
    float sum = octet_hsum(acc);
    if (active) {
      for (uint32_t e = body; e < a.dim; ++e) sum = __fadd_rn(sum, __fmul_rn(__ldg(xr + e), qs[e]));
      if (j == 0) keys[i] = pack_key(clamp02(__fsub_rn(1.0f, sum)), row);
    }
  }
  block_bitonic_sort(keys, P);
  for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
    unsigned long long key = i < P ? keys[i] : kKeyMax;
    bool ok = key != kKeyMax;
    a.ids[(uint64_t)q * a.k + i] = ok ? (uint32_t)key : CM_NO_ID_DEV;
    a.dist[(uint64_t)q * a.k + i] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
  }
  if (tid == 0 && a.rescored_total) atomicAdd(a.rescored_total, (unsigned long long)ns);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2D bf16 row-major [rows][cols] tensor map, box = {64 cols, box_rows}, 128B swizzle.
static bool make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

uint64_t scan_tc_pad_rows(uint64_t n) { return (n + BN - 1) / BN * BN; }
uint32_t scan_tc_pad_queries(uint32_t nq) { return (nq + BM - 1) / BM * BM; }
uint32_t scan_tc_pad_k(uint32_t dim) { return (dim + BK - 1) / BK * BK; }

bool scan_tc_supported(const DeviceIndex& ix, uint32_t k) {
  return ix.xbf16 != nullptr && k >= 1 && k <= 64 && encode_fn() != nullptr;
}

cudaError_t launch_scan_tc(const DeviceIndex& ix, const ScanTcLaunch& l, cudaStream_t s) {
  CUtensorMap map_q, map_x;
  const uint32_t nq_pad = scan_tc_pad_queries(l.nq);
  if (!make_map(&map_q, l.q_bf16, nq_pad, ix.k_pad, BM)) return cudaErrorInvalidValue;
  if (!make_map(&map_x, ix.xbf16, ix.n_pad, ix.k_pad, BN)) return cudaErrorInvalidValue;
  ScanArgs a;
  a.nq = l.nq;
  a.m_tiles = nq_pad / BM;
  a.n_tiles = (uint32_t)(ix.n_pad / BN);
  a.k_blocks = ix.k_pad / BK;
  a.k = l.k;
  a.n = ix.n;
  a.deleted = ix.deleted;
  a.tau_key = l.tau_key;
  a.cand_cnt = l.cand_cnt;
  a.cand = l.cand;
  a.cap = l.cap;
  size_t smem = 1024 + (size_t)kStages * kStageBytes + (size_t)l.k * 128 * 4 + sizeof(ScanSmemTail) + 16;
  cudaError_t e = cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const uint32_t strips = (a.n_tiles + kStripTiles - 1) / kStripTiles;
  // Launch 1 warms the per-query thresholds on the first strip only (one unit per query tile), so that the
  // strips of launch 2 — many of which run concurrently for the same queries — start with a real threshold
  // instead of each emitting its own cold-start candidates.
  a.strip_begin = 0;
  a.strip_end = 1;
  unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ix.sm_count, (uint64_t)a.m_tiles);
  scan_tc_kernel<<<grid, kThreads, smem, s>>>(map_q, map_x, a);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess || strips == 1) return e;
  a.strip_begin = 1;
  a.strip_end = strips;
  grid = (unsigned)std::min<uint64_t>((uint64_t)ix.sm_count, (uint64_t)(strips - 1) * a.m_tiles);
  scan_tc_kernel<<<grid, kThreads, smem, s>>>(map_q, map_x, a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rescore(const DeviceIndex& ix, const RescoreLaunch& l, cudaStream_t s) {
  if (l.nq == 0) return cudaSuccess;
  RescoreArgs a;
  a.qn = l.qn;
  a.ldq = l.ldq;
  a.nq = l.nq;
  a.k = l.k;
  a.dim = ix.dim;
  a.ldx = ix.dim_pad;
  a.x = ix.x32;
  a.n_live = ix.live;
  a.tau_key = l.tau_key;
  a.cand_cnt = l.cand_cnt;
  a.cand = l.cand;
  a.cap = l.cap;
  a.ids = l.ids;
  a.dist = l.dist;
  a.flag_count = l.flag_count;
  a.flag_list = l.flag_list;
  a.flag_cap = l.flag_cap;
  a.rescored_total = l.rescored_total;
  rescore_kernel<<<l.nq, RS_THREADS, (size_t)ix.dim * 4, s>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cm
