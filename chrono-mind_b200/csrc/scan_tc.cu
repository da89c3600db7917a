// scan_tc.cu — K1/K2: exact brute-force kNN as a dense contraction on the 5th-gen tensor cores.
//
// Replaces the all-pairs loop of the reference's brute force (tests/recall_test.rs:70-79,
// tests/property_test.rs:201-206, src/store.rs:362-376) for large batches:
//
//   K1  scan_tc_kernel   S = Q · Xᵀ in bf16 on tcgen05 (cta_group::2: one 256x256 tile per CTA PAIR, fp32
//                        accumulators in TMEM, operands staged by TMA into 128B-swizzled shared memory). S is
//                        never written: the epilogue warps read each accumulator tile straight out of TMEM and
//                        keep, per query, only the rows whose bf16 score can still be among the exact top-k
//                        (threshold select, see below).
//   K2  rescore_kernel   per query: tighten the candidate set with the now-known k-th best bf16 score, re-score
//                        the few survivors with the reference's exact fp32 arithmetic (dot_avx2 order,
//                        src/metric.rs:129-147) and sort by (distance, handle).
//
// Why the result is EXACT (not "bf16-approximate"): let e_r be the fp32 reference dot of row r and s_r the
// tensor-core bf16 score, |s_r - e_r| <= eps for unit vectors (bf16 rounding of both operands: 2·2^-9 relative
// per product, Cauchy-Schwarz => <= 2^-8 absolute; eps_for_dim() adds a dim-dependent allowance for the fp32
// accumulation). With s_(k) / e_(k)
// the k-th largest score over the live rows: the k rows of largest s have e >= s_(k) - eps, so
// e_(k) >= s_(k) - eps, so every row of the exact top-k (e_r >= e_(k)) has s_r >= s_(k) - 2·eps.
//   * K1's threshold tau_q is always a LOWER bound of s_(k) - 2·eps: it is the k-th best of k scores of distinct
//     live rows actually seen for that query, minus 2·eps, and only ever grows (atomicMax). Every row with
//     s_r >= tau_q (final) is therefore in the candidate list, which contains the whole bf16 top-k.
//   * K2 reads s_(k) off the candidate list (its k-th largest key), keeps s_r >= s_(k) - 2·eps and orders those
//     with the exact arithmetic.
// If a query's candidate buffer overflows, or fewer than min(k, live) candidates survive (NaN/Inf queries),
// K2 flags the query and it is re-run through the fp32 scan — never silently degraded.
//
// The same kernel serves two more callers:
//   * GPU-assisted graph construction (cm_index_build_graph_gpu): corpus against itself in query batches, K2 in
//     `approx` mode (top-k by bf16 score, no exact rescore — the host re-measures every candidate);
//   * consolidate's all-pairs similarity test (cm_similar_pairs): `join` mode — Q == X, one fixed threshold, only
//     tiles on or above the diagonal, pairs (query < row) appended to one list.
//
// Kernel organisation (persistent, one CTA per SM, clusters of 2 = one CTA pair per TPC, 8 warps per CTA):
//   warp 0  TMA producer (one lane, both CTAs) — 16 KB operand blocks (128 rows x 64 bf16, 128B-swizzled): the first
//           q_res k-blocks of the CTA's 128 query rows are loaded ONCE per work unit and stay resident; the remaining
//           query blocks and every corpus block stream through a ring of `slots` blocks. One full / empty barrier
//           pair per k-block (not per block: a second commit per k-block cost 20 %); the pair's loads all complete
//           on the LEADER's full barrier
//   warp 1  MMA issuer   (one lane, leader CTA only) — tcgen05.mma.cta_group::2.kind::f16, M=256 N=256 K=16:
//           CTA r supplies Q rows [128r, 128r+128) and X rows [128r, 128r+128) of the tile and receives the
//           accumulator rows of ITS queries against all 256 X rows
//   warp 2  TMEM allocator — 512 columns = two 128x256 fp32 accumulators (double buffered) per CTA
//   warps 4-7 epilogue (128 threads) — thread = one query row (TMEM lane); tcgen05.ld 32 columns at a time
//   warps 8-11 (kEpi2 instantiations, rows <= 192 columns) a second epilogue quad: quad e drains accumulator stage e
// The per-thread running threshold is an exact top-k list (LocalTopK, k <= 32 or biased) or a score histogram over a
// fixed ladder of bounds (Ladder, kLadder instantiations, k > 32); either is a lower bound of s_(k) built from scores of
// k distinct live rows the thread has seen, and the threads of a query meet through tau_key (atomicMax).
// Operand traffic: all-streamed, each CTA fetches 32 KB per 2·128·256·64 FLOP (128 FLOP/B, vs 85 for a 1-CTA 128x256
// tile) = 64 B/clk per SM at full tensor rate, more than the L2 -> SM path delivers chip-wide (≈6.3-7.1 KB/clk =
// 43-48 B/clk per SM). With 7 of a 768-wide row's 12 query blocks resident the stream drops to 17/24 of that.
//
// Work unit = (query tile pair m, corpus split s): the pair scans tiles [s·T/S, (s+1)·T/S) for its 256 queries,
// so a thread's running top-k covers a long row range and its threshold converges after a few tiles. Units are
// ordered m-fastest: the clusters in flight stream the same corpus splits at the same time (X comes from HBM
// once per split and is hit in L2 by the other query tiles).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "row_gather.cuh"
#include "scan_tc.h"

namespace cm {

constexpr uint32_t BM = 128;        // query rows per CTA (TMEM lanes)
constexpr uint32_t PAIR_M = 256;    // query rows per CTA pair (UMMA M)
constexpr uint32_t BN = 256;        // corpus rows per tile (UMMA N)
constexpr uint32_t BN_HALF = 128;   // corpus rows each CTA stages
constexpr uint32_t BK = 64, UMMA_K = 16;
constexpr uint32_t kAccStages = 2;    // the epilogue derives stage and phase from a tile counter: stage = seq & 1
// Operand blocks: one k-block (BK columns) of a CTA's 128 query rows or of its 128 corpus rows = 16 KB, 128B-swizzled.
// A CTA keeps the first `q_res` k-blocks of its query rows RESIDENT for a whole work unit and streams everything else
// (the remaining query blocks and every corpus block) through a ring of `slots` such blocks.
constexpr uint32_t kBlockBytes = BM * BK * 2;
static_assert(BM == BN_HALF, "query and corpus blocks share the ring");
constexpr uint32_t kMaxSlots = 13;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kThreads = 256;      // 8 warps; the kEpi2 instantiations add a second epilogue warp-quad (384 threads)
constexpr uint32_t kThreadsEpi2 = 384;
constexpr uint32_t kCtxSortChunk = 16384;  // queries grouped by context per single-CTA sort (128 KB of keys)
// Bound on |tensor-core bf16 score - fp32 reference dot| for (at most) unit-length rows of `dim` elements:
//   operand rounding   both operands to bf16 (unit roundoff 2^-9): |sum(a^b^) - sum(ab)| <= (2^-8 + 2^-18) sum|ab| <= 0.0039101
//   accumulation       every UMMA K-step adds 16 exact products into an fp32 accumulator of magnitude <= 1; allowing a
//                      truncation of 2^-23 per addend that is <= 1.07e-7 * dim, plus the reference's own chain
//                      (dim/8 roundings of 2^-24); the 3.8e-7 * dim below is three times that.
// 0.0042 is what every dim <= 768 has used since round 1 (kept, so those candidate sets are unchanged); beyond that
// the allowance grows with dim instead of silently running out (dim 12,800: 0.0088). The margin is 2 * eps.
static inline float eps_for_dim(uint32_t dim) { return std::max(0.0042f, 0.0039139f + 3.8e-7f * (float)dim); }
constexpr uint32_t kMaxSmem = 232448;  // 227 KB opt-in limit per CTA

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier of any CTA of the cluster (shared::cluster address). CTA-scope release: what the
// waiter consumes is TMEM, ordered by tcgen05.wait::ld + tcgen05.fence, not generic memory — a cluster-scope
// release would drain every outstanding candidate store first (MEMBAR.ALL.GPU + ERRBAR).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
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

// TMA load issued by either CTA of a pair into ITS OWN shared memory; the bytes are accounted on `bar_cluster`
// (a shared::cluster address — the leader's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1)
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
// D[tmem] (+)= A[smem] · B[smem]ᵀ over the CTA pair; issued by one thread of the leader CTA.
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// When every MMA issued so far has retired: arrive (count 1) on the barrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
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

// r[i] for a run-time i without spilling r[] to local memory: a 5-level select tree (31 selects).
__device__ __forceinline__ float pick32(const uint32_t (&r)[32], uint32_t i) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (i & 1) ? r[2 * j + 1] : r[2 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = (i & 4) ? b[2 * j + 1] : b[2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j) d[j] = (i & 8) ? c[2 * j + 1] : c[2 * j];
  return __uint_as_float((i & 16) ? d[1] : d[0]);
}

struct ScanArgs {
  uint32_t nq, m_pairs, n_tiles, k_blocks, k, splits, slots, q_res;
  float eps2;                    // 2 * eps_for_dim(dim): the candidate margin
  uint32_t diag_no_hits;         // CM_SCAN_DIAG=1 (measurement only, results invalid): every threshold is +inf
  uint32_t slot_chunk;           // candidate slots reserved per atomic (8; 32 for k > 32: the reply is a ~1 us stall)
  uint32_t ladder_n;             // kLadder: buckets of the threshold ladder (shared memory [ladder_n][128] u16 replaces the lists)
  uint32_t* tau_key;             // [m_pairs*256] ordered-uint threshold per query (0 = none yet)
  uint32_t* cand_cnt;            // [m_pairs*256] slots reserved per query
  unsigned long long* cand;      // [m_pairs*256][cap]  (ordered score key << 32 | row); 0 = unused slot
  uint32_t cap;
  // threshold self-join (consolidate's all-pairs pass): fixed threshold, only row > query, pairs to one list
  uint32_t join;                 // 1: join mode (Q == X)
  float join_tau;                // bf16-score threshold (the caller already subtracted the error bound)
  unsigned long long* pairs;     // [pair_cap] (query << 32 | row)
  uint32_t* pair_count;          // pairs found (may exceed pair_cap: the caller retries with a larger buffer)
  uint32_t pair_cap;
  // predecessor mode (graph construction): query i of this launch is corpus row q_base + i (q_base a multiple of
  // 256) and only rows BEFORE it count — what `LockFreeHnsw::insert` can see when it links node q_base + i
  uint32_t pred;
  uint32_t q_base;
  // biased mode (kBiased instantiation, `search_in_context`): the ranking key of query q against row r is
  // score + bias[r], and only rows with row_ctx[r] == q_ctx[q] take part
  const float* bias;             // [n_pad]
  const uint32_t* row_ctx;       // [n_pad]
  const uint32_t* q_ctx;         // [nq]
  // per query-tile-pair tile range [pair_lo[m], pair_hi[m]) instead of [0, n_tiles): rows sorted by context, queries
  // grouped by context — a pair only scans the tiles that hold its queries' contexts
  const uint32_t* pair_lo;
  const uint32_t* pair_hi;
};

// Tile range [t0, t1) of work unit (m, sp).
__device__ __forceinline__ void unit_tiles(const ScanArgs& a, uint32_t m, uint32_t sp, uint32_t& t0, uint32_t& t1) {
  const uint32_t lo = a.pair_lo ? a.pair_lo[m] : 0u;
  const uint32_t len = a.pair_lo ? a.pair_hi[m] - lo : a.n_tiles;
  t0 = lo + (uint32_t)((uint64_t)sp * len / a.splits);
  t1 = lo + (uint32_t)((uint64_t)(sp + 1) * len / a.splits);
}

// Shared-memory carve-up (dynamic): [q_res resident query blocks | ring of `slots` blocks] 1024-aligned | local top-k
// lists [k][128] | barriers | tmem ptr.
// Both CTAs of a pair use the same offsets (the MMA descriptors and multicast commits rely on it).
struct ScanSmemTail {
  unsigned long long full[kMaxSlots], empty[kMaxSlots], tmem_full[kAccStages], tmem_empty[kAccStages];
  unsigned long long res_full, res_free;  // resident query blocks of the current unit: landed / no longer read
  uint32_t tmem_base;
};

// Running top-k (k largest scores) of one thread, unsorted, strided by 128 in shared memory.
struct LocalTopK {
  float* list;
  uint32_t k, filled, min_pos;
  float kth;  // smallest of the k entries once filled == k
  __device__ __forceinline__ void rescan() {  // (as a call it shrinks the kernel from 11.4k to 3.5k instructions and is 0.7-5 % slower)
    kth = list[0];
    min_pos = 0;
    for (uint32_t j = 1; j < k; ++j) {
      float v = list[j * 128];
      if (v < kth) kth = v, min_pos = j;
    }
  }
  // returns true when the k-th best changed
  __device__ __forceinline__ bool insert(float s) {
    if (filled < k) {
      list[filled * 128] = s;
      if (++filled == k) {
        rescan();
        return true;
      }
      return false;
    }
    if (s > kth) {
      list[min_pos * 128] = s;
      rescan();
      return true;
    }
    return false;
  }
};

// Running threshold of one thread for LARGE k (kLadder instantiations, k > 32): a histogram of the scores seen so far
// over a fixed ladder of bounds b(j) = j * 2 / n - 1, instead of the exact top-k list. "At least k distinct live rows
// scored >= b(j)" proves s_(k) >= b(j), so tau = b(j) - 2 eps is a valid lower bound of s_(k) - 2 eps exactly like the
// list's k-th best — only up to one ladder step (2 / n) looser, i.e. a few tens of per cent more candidates for K2 to
// drop. What it buys: recording a score is one 16-bit increment and raising the threshold one load per step, where the
// list pays an O(k) rescan per improving score with the whole warp waiting on the busiest lane (at k = 100 the
// epilogue of configs[4] spent ~5,500 of its ~6,000 cycles per tile there).
//   hist[i]  scores recorded in bucket i, recorded only while i > j (a score below b(j + 1) can never matter again)
//   above    recorded scores >= b(j + 1) = sum of hist[i] over i > j
// Only this thread's own bound is tracked: a higher threshold published by another unit of the same query simply
// filters the hits; the ladder catches up through empty buckets, one load each.
struct Ladder {
  unsigned short* hist;  // this thread's column of [n][128]
  uint32_t n, k, above;
  int j;                 // current bound b(j); -1: none yet
  float scale, step;     // n / 2, 2 / n
  float b_next;          // b(j + 1); -inf before the first bound, +inf at the top of the ladder
  __device__ __forceinline__ void reset() {
    for (uint32_t i = 0; i < n; ++i) hist[i * 128] = 0;
    above = 0;
    j = -1;
    b_next = -INFINITY;
  }
  __device__ __forceinline__ void record(float s) {  // s >= b_next, not NaN
    const uint32_t i = (uint32_t)fminf(fmaxf((s + 1.0f) * scale, 0.0f), (float)(n - 1));
    hist[i * 128] += 1;
    ++above;
  }
  // (s + 1) * scale may round a score a hair below b(i) up into bucket i: the 2^-20 covers that (<= 2^-23 absolute)
  __device__ __forceinline__ float tau(float eps2) const {
    return j >= 0 ? (float)j * step - 1.0f - eps2 - 9.5367431640625e-07f : -INFINITY;
  }
  __device__ __forceinline__ bool advance() {
    bool moved = false;
    while (above >= k && j < (int)n - 1) {
      ++j;                      // >= k recorded scores reach b(j)
      above -= hist[j * 128];   // what is left reaches b(j + 1)
      moved = true;
    }
    if (moved) b_next = j < (int)n - 1 ? (float)(j + 1) * step - 1.0f : INFINITY;
    return moved;
  }
};

template <bool kBiased, bool kPred, bool kLadder, bool kEpi2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEpi2 ? kThreadsEpi2 : kThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, ScanArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries: align by hand (the launch reserves the slack). The
  // dynamic window starts at the same offset in both CTAs of the pair, so the carve-up matches.
  uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t kRes = a.q_res, kSlots = a.slots;
  uint8_t* ring = tiles + kRes * kBlockBytes;
  float* lists = reinterpret_cast<float*>(ring + kSlots * kBlockBytes);  // [k][128]
  ScanSmemTail* tail = reinterpret_cast<ScanSmemTail*>(ring + kSlots * kBlockBytes +
                                                       (kEpi2 ? 2 : 1) * (kLadder ? (size_t)a.ladder_n * 128 * 2 : (size_t)a.k * 128 * 4));
  // biased mode: the current tile's 256 row biases and context ids, one slot per accumulator stage
  float* tile_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tail) + ((sizeof(ScanSmemTail) + 15) & ~(size_t)15));
  uint32_t* tile_ctx = reinterpret_cast<uint32_t*>(tile_bias + kAccStages * BN);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint32_t units = a.m_pairs * a.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < kMaxSlots; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);    // leader's producer: arrive.expect_tx for the pair's bytes
      mbar_init(smem_u32(&tail->empty[s]), 1);   // one multicast commit per use
    }
    mbar_init(smem_u32(&tail->res_full), 1);
    mbar_init(smem_u32(&tail->res_free), 1);
    for (uint32_t s = 0; s < kAccStages; ++s) {
      mbar_init(smem_u32(&tail->tmem_full[s]), 1);   // one multicast commit per tile
      mbar_init(smem_u32(&tail->tmem_empty[s]), 8);  // leader's copy: 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / TMA lands
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =============================================================================
    if (lane == 0) {
      // Ring bookkeeping (mirrored by the MMA issuer): k-block number i of this cluster's stream owns barrier pair
      // i % kMaxSlots and takes 1 granule (corpus block only: its query block is resident) or 2 (query block, then corpus
      // block) of the `kSlots`-granule ring, in order. A granule is free again when the k-block that used it has retired
      // (its MMAs' commit on `empty`), and k-blocks retire in order, so one cursor replays the frees.
      uint32_t issue = 0, retire = 0, gpos = 0, gfree = kSlots, res_units = 0;
      const uint32_t kbs = a.k_blocks;
      for (uint32_t u = cluster_id; u < units; u += n_clusters) {
        const uint32_t m = u % a.m_pairs, sp = u / a.m_pairs;
        uint32_t t0, t1_all;
        unit_tiles(a, m, sp, t0, t1_all);
        if (a.join) t0 = max(t0, u % a.m_pairs);  // self-join keeps row > query: tiles below the diagonal are skipped
        const uint32_t t1 = kPred ? min(t1_all, a.q_base / BN + m + 1) : t1_all;  // predecessors: up to the diagonal tile
        if (t0 >= t1) continue;
        const int32_t q_row = (int32_t)(m * PAIR_M + rank * BM);
        if (kRes) {
          // The unit's resident query blocks: once the previous unit's last MMA has retired (res_free), one bulk of
          // kRes blocks per CTA, accounted on the leader's res_full.
          if (res_units) mbar_wait(smem_u32(&tail->res_free), (res_units - 1) & 1);
          const uint32_t bar_local = smem_u32(&tail->res_full);
          if (rank == 0) mbar_expect_tx(bar_local, 2 * kRes * kBlockBytes);
          const uint32_t bar = map_to_cta(bar_local, 0);
          for (uint32_t kb = 0; kb < kRes; ++kb)
            tma_load_2d_pair(smem_u32(tiles + kb * kBlockBytes), &map_q, bar, (int32_t)(kb * BK), q_row);
          ++res_units;
        }
        for (uint32_t t = t0; t < t1; ++t) {
          const int32_t x_row = (int32_t)(t * BN + rank * BN_HALF);
          for (uint32_t kb = 0; kb < kbs; ++kb) {
            const uint32_t g = kb < kRes ? 1u : 2u;
            while (gfree < g || issue - retire >= kMaxSlots) {  // retire the oldest k-block in flight
              mbar_wait(smem_u32(&tail->empty[retire % kMaxSlots]), (retire / kMaxSlots) & 1);
              gfree += (retire % kbs) < kRes ? 1u : 2u;
              ++retire;
            }
            const uint32_t bar_local = smem_u32(&tail->full[issue % kMaxSlots]);
            if (rank == 0) mbar_expect_tx(bar_local, 2 * g * kBlockBytes);  // this k-block's bytes of both CTAs
            const uint32_t bar = map_to_cta(bar_local, 0);
            if (g == 2) {  // streamed query block first, then the corpus block
              tma_load_2d_pair(smem_u32(ring + gpos * kBlockBytes), &map_q, bar, (int32_t)(kb * BK), q_row);
              if (++gpos == kSlots) gpos = 0;
            }
            tma_load_2d_pair(smem_u32(ring + gpos * kBlockBytes), &map_x, bar, (int32_t)(kb * BK), x_row);
            if (++gpos == kSlots) gpos = 0;
            gfree -= g;
            ++issue;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) ==============================================================================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(PAIR_M, BN);
      uint32_t issue = 0, gpos = 0, acc = 0, acc_phase = 0, res_phase = 0;
      for (uint32_t u = cluster_id; u < units; u += n_clusters) {
        const uint32_t sp = u / a.m_pairs;
        uint32_t t0, t1;
        unit_tiles(a, u % a.m_pairs, sp, t0, t1);
        if (a.join) t0 = max(t0, u % a.m_pairs);  // self-join keeps row > query: tiles below the diagonal are skipped
        if (kPred) t1 = min(t1, a.q_base / BN + u % a.m_pairs + 1);  // predecessors: up to the diagonal tile
        if (t0 >= t1) continue;
        if (kRes) {
          mbar_wait(smem_u32(&tail->res_full), res_phase);  // this unit's resident query blocks have landed (both CTAs)
          res_phase ^= 1;
          tc_fence_after();
        }
        for (uint32_t t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(&tail->tmem_empty[acc]), acc_phase ^ 1);  // both epilogues drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
            const uint32_t bi = issue % kMaxSlots;
            mbar_wait(smem_u32(&tail->full[bi]), (issue / kMaxSlots) & 1);
            tc_fence_after();
            uint64_t adesc;
            if (kb < kRes) {
              adesc = make_smem_desc(smem_u32(tiles + kb * kBlockBytes));
            } else {
              adesc = make_smem_desc(smem_u32(ring + gpos * kBlockBytes));
              if (++gpos == kSlots) gpos = 0;
            }
            const uint64_t bdesc = make_smem_desc(smem_u32(ring + gpos * kBlockBytes));
            if (++gpos == kSlots) gpos = 0;
#pragma unroll
            for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
              // advancing 16 bf16 (32 bytes) along K inside the 128-byte swizzle row: +2 in the >>4 start field
              umma_f16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
            }
            umma_commit_pair(smem_u32(&tail->empty[bi]));  // retires the k-block in both CTAs when these MMAs are done
            if (kb + 1 == a.k_blocks) umma_commit_pair(smem_u32(&tail->tmem_full[acc]));
            ++issue;
          }
          if (++acc == kAccStages) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        if (kRes) umma_commit_pair(smem_u32(&tail->res_free));  // the resident blocks may be overwritten once these retire
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: threshold select straight out of TMEM =======================================================
    const uint32_t quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const uint32_t row_in_cta = quarter * 32 + lane;   // accumulator row == TMEM lane
    // kEpi2 (short rows: a tile's MMAs take a few hundred cycles and ONE warp per scheduler, with every ALU and memory
    // latency exposed, cannot keep up): two epilogue warp-quads, quad e draining accumulator stage e — consecutive tiles
    // alternate between them, so two tiles are filtered at a time and each scheduler has two warps to interleave. A
    // query's two threads keep separate running thresholds (each over the rows it saw: both valid lower bounds) and
    // meet through tau_key like the units of different clusters do.
    const uint32_t quad = kEpi2 ? (warp - 4) >> 2 : 0;
    constexpr bool kTuned = kEpi2 || kLadder;  // the shapes whose epilogue is the bound
    LocalTopK top;
    top.list = lists + (size_t)quad * a.k * 128 + row_in_cta;
    top.k = a.k;
    Ladder lad;
    lad.hist = reinterpret_cast<unsigned short*>(lists) + (size_t)quad * a.ladder_n * 128 + row_in_cta;
    lad.n = a.ladder_n;
    lad.k = a.k;
    lad.scale = 0.5f * (float)a.ladder_n;
    lad.step = 2.0f / (float)a.ladder_n;  // ladder_n is a power of two: every bound is exact in fp32
    const uint32_t empty_bar0 = map_to_cta(smem_u32(&tail->tmem_empty[0]), 0);
    const uint32_t empty_bar1 = map_to_cta(smem_u32(&tail->tmem_empty[1]), 0);
    uint32_t seq = 0;  // tiles this cluster has scanned so far: tile `seq` lands in accumulator stage seq & 1
    for (uint32_t u = cluster_id; u < units; u += n_clusters) {
      const uint32_t m = u % a.m_pairs, sp = u / a.m_pairs;
      uint32_t t0, t1;
      unit_tiles(a, m, sp, t0, t1);
      if (a.join) t0 = max(t0, m);  // self-join keeps row > query: tiles below the diagonal are skipped
      if (kPred) t1 = min(t1, a.q_base / BN + m + 1);  // predecessors: up to the diagonal tile
      const uint32_t q = m * PAIR_M + rank * BM + row_in_cta;
      const uint64_t q_row_id = (uint64_t)a.q_base + q;  // predecessor mode: this query IS corpus row q_base + q
      const bool q_valid = q < a.nq;
      uint32_t my_ctx = CM_NO_ID_DEV;  // biased mode: CM_NO_ID matches no row
      if (kBiased && q_valid) my_ctx = a.q_ctx[q];
      unsigned long long* my_cand = a.cand + (uint64_t)q * a.cap;
      // Per-unit state: the k best scores of this unit's rows, and a chunk of reserved candidate slots.
      top.filled = 0;
      top.min_pos = 0;
      top.kth = -INFINITY;
      if constexpr (kLadder) lad.reset();
      uint32_t slot_next = 0, slots_left = 0;
      // the next chunk of slots is reserved while half of the current one is still free, so the atomic's reply (a ~1 us
      // stall of the whole warp when waited for in place: 4 % of a small-corpus launch's samples) is back before it is needed
      uint32_t slot_ahead = 0;
      bool have_ahead = false;
      const uint32_t look_at = a.slot_chunk / 2;  // slots left in the current chunk when the next one is requested
      bool parked = false;
      float pend0 = -INFINITY, pend1 = -INFINITY, pend2 = -INFINITY, pend3 = -INFINITY;
      uint32_t tk_next = 0;     // the shared threshold, fetched one tile ahead (an L2 round trip per tile otherwise)
      bool tk_ready = false;
      for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t acc = seq & 1, acc_phase = (seq >> 1) & 1;
        ++seq;
        if (kEpi2 && acc != quad) continue;  // the other quad's tile
        float tau = INFINITY;  // padding queries: nothing passes
        if (q_valid && a.join) {
          tau = a.join_tau;
        } else if (q_valid) {
          const uint32_t tk = tk_ready ? tk_next : *reinterpret_cast<volatile uint32_t*>(a.tau_key + q);
          const float g = tk ? key_to_float(tk) : -INFINITY;
          tau = fmaxf(g, kLadder ? lad.tau(a.eps2) : (top.filled == top.k ? top.kth - a.eps2 : -INFINITY));
        }
        if (a.diag_no_hits) tau = INFINITY;
        const uint64_t row0 = (uint64_t)t * BN;
        const bool diag = kPred && t == a.q_base / BN + m;  // the only tile that holds rows at or after the query
        const float* tb = tile_bias + acc * BN;
        const uint32_t* tcx = tile_ctx + acc * BN;
        if (kBiased) {
          // Stage this tile's per-row terms. The 4 epilogue warps meet at a named barrier per tile, so a slot is
          // rewritten (two tiles later) only after every warp has finished reading it.
          float* wb = tile_bias + acc * BN;
          uint32_t* wc = tile_ctx + acc * BN;
          wb[row_in_cta] = a.bias[row0 + row_in_cta];
          wb[row_in_cta + 128] = a.bias[row0 + row_in_cta + 128];
          wc[row_in_cta] = a.row_ctx[row0 + row_in_cta];
          wc[row_in_cta + 128] = a.row_ctx[row0 + row_in_cta + 128];
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        mbar_wait(smem_u32(&tail->tmem_full[acc]), acc_phase);
        tc_fence_after();
        if (q_valid && !a.join) {  // next tile's view of the shared threshold: a tile stale at worst, still a lower bound
          tk_next = *reinterpret_cast<volatile uint32_t*>(a.tau_key + q);
          tk_ready = true;
        }
        const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + acc * BN;
        bool improved = false;
        // Cold start (no threshold yet): first build the running top-k from this tile WITHOUT emitting, so the
        // main pass below emits ~k rows instead of all 256. Tombstoned and padding rows score NaN (see
        // launch_to_bf16) and never enter a list or pass a threshold.
        // (biased mode: a sparse context may never fill its running top-k — after two tiles the few members are simply
        // emitted instead of paying the extra pass over the accumulator on every tile. On the context-SORTED layout a
        // tile is either foreign to the query or dense with members of its context: the pre-pass runs exactly on the
        // member tiles, wherever in the unit the context starts — without it the first tile of a context would be
        // emitted row by row through the divergent hit loop, ~50k cycles per tile with the whole warp waiting.)
        bool cold = q_valid && tau == -INFINITY;
        if (kBiased) cold = cold && (a.pair_lo ? (tcx[0] <= my_ctx && my_ctx <= tcx[BN - 1]) : t < t0 + 2);
        bool listed = a.join != 0;  // this tile's rows are already in the running top-k (join: there is none)
        if (__any_sync(0xFFFFFFFFu, cold)) {
#pragma unroll 1
          for (uint32_t c = 0; c < BN; c += 32) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(taddr + c, r);
            tmem_ld_wait();
            if (cold && !kBiased && !kLadder && a.k <= 16) {
              // Small k: only the maximum of every 8 columns is offered to the list — 32 scores of 32 distinct rows per
              // tile, plenty to fill it, and a k-th best of those is as valid a bound as one of all 256 (on average it is
              // the ~11th best of the tile instead of the 10th). The warp executes an insert whenever ANY lane has one,
              // which with all 256 scores was nearly every column: ~13k instructions per cold tile, now ~2k.
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float m = __int_as_float(0x7FC00000);
#pragma unroll
                for (int i = 8 * g; i < 8 * g + 8; ++i) {
                  const float s = __uint_as_float(r[i]);
                  if (!diag || row0 + c + i < q_row_id) m = fmaxf(m, s);  // fmaxf drops NaN (tombstoned / padding rows)
                }
                if (m == m) top.insert(m);
              }
            } else if (cold) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float s = __uint_as_float(r[i]);
                if (kBiased) s = tcx[c + i] == my_ctx ? s + tb[c + i] : __int_as_float(0x7FC00000);
                if (s == s && (!diag || row0 + c + i < q_row_id)) {
                  if constexpr (kLadder) lad.record(s);  // no bound yet: every live row counts
                  else top.insert(s);
                }
              }
            }
          }
          if (cold) {
            listed = true;
            if constexpr (kLadder) {
              if (lad.advance()) {
                tau = lad.tau(a.eps2);
                improved = true;
              }
            } else if (top.filled == top.k) {
              tau = top.kth - a.eps2;
              improved = true;
            }
          }
        }
        // Main pass, software pipelined: the next 32 columns are in flight while this chunk is filtered.
        auto filter_chunk = [&](uint32_t (&r)[32], uint32_t c) {
          if (kBiased) {
            // ranking key = score + bias[row] for the rows of this query's context, NaN for every other row (broadcast
            // reads: all lanes look at the same rows). Folding the membership test into the value keeps chunks
            // without a member on the fast path — with many small contexts that is nearly every chunk.
#pragma unroll
            for (int i = 0; i < 32; ++i)
              r[i] = tcx[c + i] == my_ctx ? __float_as_uint(__uint_as_float(r[i]) + tb[c + i]) : 0x7FC00000u;
          }
          // kTuned (epilogue-bound shapes): four independent chains each — a lone warp per scheduler issues a dependent
          // chain at ALU latency. The long-row kernel keeps the round-1 single chains (its epilogue has slack, and the
          // A/B on the 768-d headline put the tuned form 0.5 % behind).
          float mx;
          if constexpr (kTuned) {
            float mx0 = __uint_as_float(r[0]), mx1 = __uint_as_float(r[1]), mx2 = __uint_as_float(r[2]), mx3 = __uint_as_float(r[3]);
#pragma unroll
            for (int i = 4; i < 32; i += 4) {  // fmaxf drops NaN
              mx0 = fmaxf(mx0, __uint_as_float(r[i]));
              mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
            }
            mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          } else {
            mx = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
#pragma unroll
            for (int i = 2; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));  // fmaxf drops NaN
          }
          if (mx >= tau) {  // rare: ~1e-4 of the scores reach a converged threshold
            // (a two-level search — scan only the strided class whose partial maximum passes, 8 compares and a 7-select
            // pick — was measured 15-50 % SLOWER, also on the 768-d headline: four divergent regions instead of one)
            uint32_t mask = 0;
            if constexpr (kTuned) {
              uint32_t mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                mk0 |= (__uint_as_float(r[i]) >= tau) ? (1u << i) : 0u;
                mk1 |= (__uint_as_float(r[i + 1]) >= tau) ? (2u << i) : 0u;
                mk2 |= (__uint_as_float(r[i + 2]) >= tau) ? (4u << i) : 0u;
                mk3 |= (__uint_as_float(r[i + 3]) >= tau) ? (8u << i) : 0u;
              }
              mask = (mk0 | mk1) | (mk2 | mk3);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(r[i]) >= tau) ? (1u << i) : 0u;
            }
            while (mask) {
              const uint32_t i = __ffs(mask) - 1;
              mask &= mask - 1;
              const float s = pick32(r, i);
              if (diag && row0 + c + i >= q_row_id) continue;  // not a predecessor
              if (s >= tau && a.join) {
                const uint32_t row = (uint32_t)(row0 + c + i);
                if (row > q) {
                  const uint32_t at = atomicAdd(a.pair_count, 1u);
                  if (at < a.pair_cap) a.pairs[at] = ((unsigned long long)q << 32) | row;
                }
              } else if (s >= tau) {  // tau may have risen inside this loop
                if (slots_left == 0) {
                  slot_next = have_ahead ? slot_ahead : atomicAdd(&a.cand_cnt[q], a.slot_chunk);
                  have_ahead = false;
                  slots_left = a.slot_chunk;
                }
                if (slot_next < a.cap)
                  my_cand[slot_next] = ((unsigned long long)total_key(s) << 32) | (uint32_t)(row0 + c + i);
                ++slot_next;
                --slots_left;
                if (!have_ahead && slots_left == look_at) {
                  slot_ahead = atomicAdd(&a.cand_cnt[q], a.slot_chunk);
                  have_ahead = true;
                }
                // The running top-k is NOT updated per hit: an O(k) insert inside this divergent loop stalls the whole
                // warp for one lane's hit. The tile's FOUR BEST improving scores are parked in registers (pend0 >= pend1
                // >= pend2 >= pend3) and inserted after the tile by all lanes together. Beyond the first tiles a tile
                // improves a list at most a few times, and a score that is not offered only leaves the bound a little
                // lower for a tile or two (the list always holds scores of k distinct rows seen: a valid bound). An
                // earlier version flushed a full park in place — ~180 instructions with one lane active, up to dozens
                // of times per tile and warp while thresholds converge: 47 us of every launch (1M x 768: -0.5 %,
                // a 125k-row shard: -3 %).
                if constexpr (kLadder) {
                  if (!listed && s >= lad.b_next) lad.record(s);
                } else if (!listed) {
                  if (top.filled < top.k) {  // list not full yet (sparse tiles): cheap appends, no rescan until it fills
                    improved |= top.insert(s);
                    if (top.filled == top.k) tau = fmaxf(tau, top.kth - a.eps2);
                  } else if (s > top.kth && s > pend3) {
                    pend3 = s;
                    if (pend3 > pend2) { const float x = pend2; pend2 = pend3; pend3 = x; }
                    if (pend2 > pend1) { const float x = pend1; pend1 = pend2; pend2 = x; }
                    if (pend1 > pend0) { const float x = pend0; pend0 = pend1; pend1 = x; }
                    parked = true;
                  }
                }
              }
            }
            if constexpr (kLadder) {  // k recorded scores reach the next bound: raise the threshold right here
              if (lad.above >= lad.k && lad.advance()) {
                tau = fmaxf(tau, lad.tau(a.eps2));
                improved = true;
              }
            }
          }
        };
        {
          uint32_t ra[32], rb[32];
          __syncwarp();
          tmem_ld32(taddr, ra);
#pragma unroll 1
          for (uint32_t c = 0; c < BN; c += 64) {
            tmem_ld_wait();
            __syncwarp();  // filter_chunk diverges; tcgen05.ld is warp-collective
            tmem_ld32(taddr + c + 32, rb);
            filter_chunk(ra, c);
            tmem_ld_wait();
            __syncwarp();
            if (c + 64 < BN) tmem_ld32(taddr + c + 64, ra);
            filter_chunk(rb, c + 32);
          }
        }
        // accumulator drained: hand it back to the MMA warp of the leader CTA
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? empty_bar1 : empty_bar0);
        // parked scores -> running top-k (all lanes insert together: the cost is the busiest lane's, not the sum)
        if (parked) {  // empty places hold -inf: insert() turns them away
          improved |= top.insert(pend0);
          improved |= top.insert(pend1);
          improved |= top.insert(pend2);
          improved |= top.insert(pend3);
          parked = false;
          pend0 = pend1 = pend2 = pend3 = -INFINITY;
        }
        // publish the lower bound s_(k) - 2 eps for the other clusters working on this query
        if constexpr (kLadder) {
          if (improved && q_valid && lad.j >= 0) atomicMax(&a.tau_key[q], total_key(lad.tau(a.eps2)));
        } else if (improved && q_valid && top.filled == top.k) {
          atomicMax(&a.tau_key[q], total_key(top.kth - a.eps2));
        }
      }
      // reserved but unused candidate slots: mark empty
      if (have_ahead && slots_left == 0) slot_next = slot_ahead, slots_left = a.slot_chunk, have_ahead = false;
      if (have_ahead)  // a whole reserved chunk that was never started
        for (uint32_t i = 0; i < a.slot_chunk; ++i)
          if (slot_ahead + i < a.cap) my_cand[slot_ahead + i] = 0ull;
      while (slots_left) {
        if (slot_next < a.cap) my_cand[slot_next] = 0ull;
        ++slot_next;
        --slots_left;
      }
    }
  }

  // Neither CTA may exit (or free TMEM) while its peer can still multicast to / read from it.
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2 rescore: one CTA per query.
//   1. keep the candidates whose bf16 score reaches the FINAL scan threshold (everything the scan emitted
//      before the threshold converged drops out here);
//   2. sort them by score, read s_(k) — the exact k-th best bf16 score over the live rows — and keep the
//      prefix with s >= s_(k) - 2 eps;
//   3. evaluate distance_prepared exactly for that prefix (octet per candidate, reference lane order) and
//      sort by (distance, handle).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t RS_THREADS = 128;
constexpr uint32_t RS_ROWS = 16;     // most rows gathered per copy round (RS_THREADS / 4); fewer when rows are long
constexpr uint32_t RS2_MAX = 512;    // candidates re-scored per query

struct RescoreArgs {
  const float* qn;
  uint32_t ldq, nq, k, dim, ldx;
  const float* x;
  uint64_t n_live;
  const uint32_t* tau_key;
  const uint32_t* cand_cnt;
  const unsigned long long* cand;
  uint32_t cap;
  uint32_t rs_max;        // candidates above the final scan threshold handled per query (power of two)
  uint32_t approx;        // 1: emit the top-k by bf16 score (distance = 1 - score) without the exact rescore
  uint32_t rows;          // rows gathered per copy round, <= RS_ROWS
  float eps2;             // 2 * eps_for_dim(dim)
  uint32_t region;        // bytes of the shared keys / row-buffer region (128-byte multiple)
  uint32_t pred;          // predecessor mode: query q has only pred_base + q rows to choose from
  uint32_t pred_base;
  uint32_t debug;         // CM_RS_DEBUG: print why a query was flagged
  uint32_t* ids;
  float* dist;
  uint32_t* flag_count;   // queries that need the fp32 fallback
  uint32_t* flag_list;    // [flag_cap]
  uint32_t flag_cap;
  unsigned long long* rescored_total;
};

// Order statistics of a SMALL key set without a sort: thread i counts the keys below its own (all keys are distinct:
// a row is a candidate of a query at most once). Broadcast reads, no barriers — the block-wide bitonic sort spends
// ~15-28 barriers on the 13-30 keys a typical query has here, which was a third of this kernel's stall samples.
__device__ __forceinline__ uint32_t rank_among(const unsigned long long* keys, uint32_t n, unsigned long long mine) {
  uint32_t r = 0;
  for (uint32_t j = 0; j < n; ++j) r += keys[j] < mine ? 1u : 0u;
  return r;
}

__global__ void __launch_bounds__(RS_THREADS) rescore_kernel(RescoreArgs a) {
  extern __shared__ __align__(128) unsigned char rs_smem[];
  const uint32_t row_stride = a.ldx * 4 + RG_ROW_PAD;
  // The candidate keys are only needed until the rows to re-score are known, the row buffer only afterwards: they
  // share one region (a barrier separates the last read of `keys` from the first bulk copy), which is what lets six
  // queries instead of four be resident per SM.
  unsigned char* rowbuf = rs_smem;                                                       // [RS_ROWS][row_stride]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(rs_smem);             // [rs_max] ~(score key, row)
  unsigned long long* dkeys = reinterpret_cast<unsigned long long*>(rs_smem + a.region);  // [RS2_MAX] (distance, row)
  uint32_t* rows = reinterpret_cast<uint32_t*>(dkeys + RS2_MAX);                         // [RS2_MAX] rows to re-score
  float* qs = reinterpret_cast<float*>(rows + RS2_MAX);                                  // [ldq]
  __shared__ unsigned long long mbar, qbar;
  __shared__ unsigned long long kth_key;
  __shared__ uint32_t n_surv, n_keep;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  // The prepared query row arrives by ONE bulk copy that completes on its own barrier while the candidates are read
  // and ranked (a per-thread load loop here was 12 % of the kernel's stall samples, all of it exposed latency).
  const float* q_src = a.qn + (uint64_t)q * a.ldq;
  const bool bulk_q = !a.approx && (a.ldq & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.qn) & 15u) == 0;
  if (tid == 0) {
    n_surv = 0, n_keep = 0;
    kth_key = 0ull;
    rg_mbar_init(&mbar);
    if (bulk_q) {
      rg_mbar_init(&qbar);
      const uint32_t bytes = ((a.dim + 3u) & ~3u) * 4u;
      const uint32_t bar = rg_smem_u32(&qbar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(rg_smem_u32(qs)), "l"(q_src), "r"(bytes), "r"(bar) : "memory");
    }
  }
  const uint32_t cnt_raw = a.cand_cnt[q];
  const uint32_t cnt = min(cnt_raw, a.cap);
  const uint32_t tk = a.tau_key[q];
  if (!a.approx && !bulk_q)  // the bf16 ranking of `approx` never looks at the fp32 query
    for (uint32_t e = tid; e < a.dim; e += RS_THREADS) qs[e] = q_src[e];
  __syncthreads();
  for (uint32_t i = tid; i < cnt; i += RS_THREADS) {
    const unsigned long long c = a.cand[(uint64_t)q * a.cap + i];
    if (c != 0ull && (uint32_t)(c >> 32) >= tk) {
      uint32_t slot = atomicAdd(&n_surv, 1u);
      if (slot < a.rs_max) keys[slot] = ~c;  // ascending order of ~c == descending score
    }
  }
  __syncthreads();
  const uint32_t ns = n_surv;
  uint32_t need = (uint32_t)(a.n_live < a.k ? a.n_live : a.k);
  if (a.pred) need = min(a.k, a.pred_base + q);
  bool bad = cnt_raw > a.cap || ns > a.rs_max || ns < need;
  const bool small = ns <= RS_THREADS;   // one key per thread: ranks by counting instead of a sort
  uint32_t keep = 0, my_rank = 0;
  unsigned long long my_key = kKeyMax;
  if (!bad) {
    if (small) {
      if (tid < ns) {
        my_key = keys[tid];
        my_rank = rank_among(keys, ns, my_key);
        if (my_rank == a.k - 1) kth_key = my_key;  // s_(k): the k-th largest bf16 score (the whole bf16 top-k is in the list)
      }
    } else {
      const uint32_t P = next_pow2(ns);
      for (uint32_t i = ns + tid; i < P; i += RS_THREADS) keys[i] = kKeyMax;
      block_bitonic_sort(keys, P);
      if (tid == 0 && ns >= a.k) kth_key = keys[a.k - 1];
    }
    __syncthreads();
    uint32_t thr = 0;  // ns < k only when fewer than k rows are live: keep everything
    if (ns >= a.k) thr = total_key(key_to_float((uint32_t)((~kth_key) >> 32)) - a.eps2);
    if (small) {  // the kept set is a prefix of the score order: its size is the largest kept rank + 1
      if (tid < ns && (uint32_t)((~my_key) >> 32) >= thr) atomicMax(&n_keep, my_rank + 1);
    } else {
      for (uint32_t i = tid; i < ns; i += RS_THREADS)
        if ((uint32_t)((~keys[i]) >> 32) >= thr) atomicMax(&n_keep, i + 1);
    }
    __syncthreads();
    keep = n_keep;
    bad = !a.approx && keep > RS2_MAX;
  }
  if (bulk_q) rg_mbar_wait(rg_smem_u32(&qbar), 0);  // landed long ago; no CTA may exit with a copy into its memory pending
  if (bad) {
    if (tid == 0) {
      uint32_t slot = atomicAdd(a.flag_count, 1u);
      if (slot < a.flag_cap) a.flag_list[slot] = q;
      if (a.debug)
        printf("[rescore] query %u flagged: emitted %u (cap %u) above-threshold %u (max %u, need %u) keep %u tau_key %08x\n", q,
               cnt_raw, a.cap, ns, a.rs_max, need, keep, tk);
    }
    // The fp32 fallback overwrites this row; if more queries are flagged than it absorbs the row stays
    // visibly invalid (CM_NO_ID / NaN) — never a plausible-looking wrong answer.
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
      a.ids[(uint64_t)q * a.k + i] = CM_NO_ID_DEV;
      a.dist[(uint64_t)q * a.k + i] = __int_as_float(0x7FC00000);
    }
    return;
  }
  if (a.approx) {  // candidate generation for graph construction: bf16 ranking is enough, the host re-measures
    if (small) {
      if (tid < ns && my_rank < a.k) {
        const unsigned long long c = ~my_key;
        a.ids[(uint64_t)q * a.k + my_rank] = (uint32_t)c;
        a.dist[(uint64_t)q * a.k + my_rank] = clamp02(1.0f - key_to_float((uint32_t)(c >> 32)));
      }
      for (uint32_t i = ns + tid; i < a.k; i += RS_THREADS) {
        a.ids[(uint64_t)q * a.k + i] = CM_NO_ID_DEV;
        a.dist[(uint64_t)q * a.k + i] = __int_as_float(0x7F800000);
      }
      return;
    }
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
      const bool ok = i < ns;
      const unsigned long long c = ok ? ~keys[i] : 0ull;
      a.ids[(uint64_t)q * a.k + i] = ok ? (uint32_t)c : CM_NO_ID_DEV;
      a.dist[(uint64_t)q * a.k + i] = ok ? clamp02(1.0f - key_to_float((uint32_t)(c >> 32))) : __int_as_float(0x7F800000);
    }
    return;
  }
  // rows to re-score, in score order (keys share their region with the row buffer: read them all before the barrier)
  if (small) {
    if (tid < ns && my_rank < keep) rows[my_rank] = (uint32_t)(~my_key);
  } else {
    for (uint32_t i = tid; i < keep; i += RS_THREADS) rows[i] = (uint32_t)(~keys[i]);
  }
  __syncthreads();
  // exact distances: rounds of a.rows rows, bulk-copied into shared memory, one thread quad per row
  const uint32_t r_slot = tid >> 2, c4 = tid & 3;
  uint32_t phase = 0;
  for (uint32_t r0 = 0; r0 < keep; r0 += a.rows) {
    const uint32_t cntr = min(a.rows, keep - r0);
    rg_stage_rows(a.x, a.ldx, &mbar, rowbuf, row_stride, rows + r0, cntr, phase);
    const bool active = r_slot < cntr;
    const float d = rg_quad_distance(qs, rowbuf + (size_t)r_slot * row_stride, a.dim, c4, active);
    if (active && c4 == 0) dkeys[r0 + r_slot] = pack_key(d, rows[r0 + r_slot]);
    __syncthreads();  // rowbuf free for the next round; dkeys visible to the ordering below
  }
  // order by (TotalF32 distance, handle) and truncate to k
  if (keep <= RS_THREADS) {
    if (tid < keep) {
      const unsigned long long key = dkeys[tid];
      const uint32_t r = rank_among(dkeys, keep, key);
      if (r < a.k) {
        a.ids[(uint64_t)q * a.k + r] = (uint32_t)key;
        a.dist[(uint64_t)q * a.k + r] = key_to_float((uint32_t)(key >> 32));
      }
    }
    for (uint32_t i = keep + tid; i < a.k; i += RS_THREADS) {
      a.ids[(uint64_t)q * a.k + i] = CM_NO_ID_DEV;
      a.dist[(uint64_t)q * a.k + i] = __int_as_float(0x7F800000);
    }
  } else {
    const uint32_t P2 = next_pow2(keep);
    for (uint32_t i = keep + tid; i < P2; i += RS_THREADS) dkeys[i] = kKeyMax;
    block_bitonic_sort(dkeys, P2);
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
      unsigned long long key = i < P2 ? dkeys[i] : kKeyMax;
      bool ok = key != kKeyMax;
      a.ids[(uint64_t)q * a.k + i] = ok ? (uint32_t)key : CM_NO_ID_DEV;
      a.dist[(uint64_t)q * a.k + i] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
    }
  }
  if (tid == 0 && a.rescored_total) atomicAdd(a.rescored_total, (unsigned long long)keep);
}

// ---------------------------------------------------------------------------------------------------------
// `ChronoMind::search_in_context` (src/store.rs:353-377) on the tensor cores.
//
// The reference scores every live member r of the query's context with
//   score = (1 - w) * distance(r, q) / 2 + w * (1 - exp(-rate_r * age_r))          (combined_score, :395-415)
// and keeps the k smallest. With cos = 1 - distance:  score = (1 - w)/2 * (1 - (cos - c_r)),
//   c_r = 2 w (1 - exp(-rate_r * age_r)) / (1 - w) >= 0,
// so the k smallest scores are the k LARGEST cos - c_r: the exact scan's threshold select with a per-row additive
// bias -c_r (k_context_bias, one pass over the attributes per call) and a membership test on the hits
// (scan_tc_kernel<true>). The candidates are then decided by the reference's own arithmetic on the rows AS STORED
// (rescore_context_kernel): dot_and_norms' dot chain, sqrt(na * nb), true division, clamp, combined_score — bit for
// bit what the fp32 context scan computes, for a few dozen rows per query instead of all of them.
// Fewer than k candidates is NOT an error here: a threshold only ever rises once k members of the context have been
// seen, so a list shorter than k simply is the whole (live) context.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_context_bias(const unsigned long long* __restrict__ ts_secs, const uint32_t* __restrict__ ts_nanos,
                               const float* __restrict__ decay, uint64_t n, uint64_t n_pad, float w, float base,
                               unsigned long long now_secs, uint32_t now_nanos, float* __restrict__ bias,
                               const uint32_t* __restrict__ perm) {
  const float scale = __fdiv_rn(2.0f, __fsub_rn(1.0f, w));
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = (perm && i < n) ? perm[i] : i;  // bias of scan row i = attributes of handle perm[i]
    bias[i] = i < n ? -__fmul_rn(scale, temporal_term_dev(ts_secs[h], ts_nanos[h], decay[h], now_secs, now_nanos, w, base)) : 0.0f;
  }
}

// Context-sorted layout: group the queries of a call by context and give every query-tile pair the tile range that
// holds its contexts.
//   k_group_queries_by_context   one CTA per chunk of <= kCtxSortChunk queries: bitonic sort of (context, caller row)
//                                keys in shared memory -> order[i] = caller row of scan query i, qctx[i] = its context
//   k_context_pair_ranges        one thread per query-tile pair: binary search of the pair's first / last context in the
//                                context ids of the sorted rows -> [lo, hi) in corpus tiles (empty when no row has them)
constexpr uint32_t GQ_THREADS = 1024;
__global__ void __launch_bounds__(GQ_THREADS) k_group_queries_by_context(const uint32_t* __restrict__ want, uint32_t nq,
                                                                          uint32_t chunk, uint32_t* __restrict__ order,
                                                                          uint32_t* __restrict__ qctx) {
  extern __shared__ __align__(16) unsigned long long gq_keys[];
  const uint32_t q0 = blockIdx.x * chunk, nb = min(chunk, nq - q0);
  const uint32_t P = next_pow2(nb < 2 ? 2 : nb);
  for (uint32_t i = threadIdx.x; i < P; i += GQ_THREADS)
    gq_keys[i] = i < nb ? (((unsigned long long)want[q0 + i] << 32) | (q0 + i)) : kKeyMax;
  block_bitonic_sort(gq_keys, P);
  for (uint32_t i = threadIdx.x; i < nb; i += GQ_THREADS) {
    order[q0 + i] = (uint32_t)gq_keys[i];
    qctx[q0 + i] = (uint32_t)(gq_keys[i] >> 32);
  }
}

__global__ void k_context_pair_ranges(const uint32_t* __restrict__ qctx, uint32_t nq, uint32_t chunk,
                                      const uint32_t* __restrict__ row_ctx, uint32_t n, uint32_t n_pairs,
                                      uint32_t* __restrict__ pair_lo, uint32_t* __restrict__ pair_hi) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  // pairs are numbered per chunk: chunk c owns pairs [c * pairs_per_chunk, ...), its queries start at c * chunk
  const uint32_t ppc = (chunk + PAIR_M - 1) / PAIR_M;
  const uint32_t c = p / ppc, first = c * chunk + (p % ppc) * PAIR_M;
  const uint32_t chunk_end = min(nq, (c + 1) * chunk);
  uint32_t lo = 0, hi = 0;
  if (first < chunk_end) {
    const uint32_t c_first = qctx[first], c_last = qctx[min(first + PAIR_M, chunk_end) - 1];
    uint32_t a = 0, b = n;  // lower_bound(c_first)
    while (a < b) {
      const uint32_t mid = a + (b - a) / 2;
      if (row_ctx[mid] < c_first) a = mid + 1; else b = mid;
    }
    const uint32_t row_lo = a;
    b = n;                  // upper_bound(c_last), from row_lo on
    while (a < b) {
      const uint32_t mid = a + (b - a) / 2;
      if (row_ctx[mid] <= c_last) a = mid + 1; else b = mid;
    }
    if (a > row_lo) lo = row_lo / BN, hi = (a + BN - 1) / BN;
  }
  pair_lo[p] = lo;
  pair_hi[p] = hi;
}

struct ContextRescoreArgs {
  const float* q_raw;   // queries as given [nq][ldq]
  uint32_t ldq, nq, k, dim, ldx;
  const float* qnorm2;  // [nq] squared norms of the raw queries (dot chain)
  const float* xraw;    // rows as stored [n][ldx]
  const float* norm2_raw;
  const unsigned long long* ts_secs;
  const uint32_t* ts_nanos;
  const float* decay;
  float w, base;
  unsigned long long now_secs;
  uint32_t now_nanos;
  const uint32_t* tau_key;
  const uint32_t* cand_cnt;
  const unsigned long long* cand;
  uint32_t cap, rs_max, rows, region;
  float eps2;
  uint32_t* ids;
  float* score;
  uint32_t* flag_count;
  uint32_t* flag_list;
  uint32_t flag_cap;
  unsigned long long* rescored_total;
  const uint32_t* perm;     // candidate row (scan order) -> handle; nullptr: identity
  const uint32_t* q_order;  // scan query index -> the caller's query row; nullptr: identity
};

__global__ void __launch_bounds__(RS_THREADS) rescore_context_kernel(ContextRescoreArgs a) {
  extern __shared__ __align__(128) unsigned char rs_smem[];
  const uint32_t row_stride = a.ldx * 4 + RG_ROW_PAD;
  unsigned char* rowbuf = rs_smem;                                                        // shares its region with keys
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(rs_smem);              // [rs_max] ~(biased score key, row)
  unsigned long long* dkeys = reinterpret_cast<unsigned long long*>(rs_smem + a.region);  // [RS2_MAX] (score, row)
  uint32_t* rows = reinterpret_cast<uint32_t*>(dkeys + RS2_MAX);
  float* qs = reinterpret_cast<float*>(rows + RS2_MAX);
  __shared__ unsigned long long mbar, qbar;
  __shared__ unsigned long long kth_key;
  __shared__ uint32_t n_surv, n_keep;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t qo = a.q_order ? a.q_order[q] : q;  // the caller's row of this query (inputs and outputs live there)
  // the query row as given: one bulk copy completing on its own barrier while the candidates are read and ranked
  const float* q_src = a.q_raw + (uint64_t)qo * a.ldq;
  const bool bulk_q = (a.ldq & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.q_raw) & 15u) == 0;
  if (tid == 0) {
    n_surv = 0, n_keep = 0;
    kth_key = 0ull;
    rg_mbar_init(&mbar);
    if (bulk_q) {
      rg_mbar_init(&qbar);
      const uint32_t bytes = ((a.dim + 3u) & ~3u) * 4u;
      const uint32_t bar = rg_smem_u32(&qbar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(rg_smem_u32(qs)), "l"(q_src), "r"(bytes), "r"(bar) : "memory");
    }
  }
  const uint32_t cnt_raw = a.cand_cnt[q];
  const uint32_t cnt = min(cnt_raw, a.cap);
  const uint32_t tk = a.tau_key[q];
  if (!bulk_q)
    for (uint32_t e = tid; e < a.dim; e += RS_THREADS) qs[e] = q_src[e];
  __syncthreads();
  for (uint32_t i = tid; i < cnt; i += RS_THREADS) {
    const unsigned long long c = a.cand[(uint64_t)q * a.cap + i];
    if (c != 0ull && (uint32_t)(c >> 32) >= tk) {
      uint32_t slot = atomicAdd(&n_surv, 1u);
      if (slot < a.rs_max) keys[slot] = ~c;
    }
  }
  __syncthreads();
  const uint32_t ns = n_surv;
  bool bad = cnt_raw > a.cap || ns > a.rs_max;
  const bool small = ns <= RS_THREADS;  // ranks by counting instead of a sort (see rescore_kernel)
  uint32_t keep = 0, my_rank = 0;
  unsigned long long my_key = kKeyMax;
  if (!bad) {
    if (small) {
      if (tid < ns) {
        my_key = keys[tid];
        my_rank = rank_among(keys, ns, my_key);
        if (my_rank == a.k - 1) kth_key = my_key;
      }
    } else {
      const uint32_t P = next_pow2(ns);
      for (uint32_t i = ns + tid; i < P; i += RS_THREADS) keys[i] = kKeyMax;
      block_bitonic_sort(keys, P);
      if (tid == 0 && ns >= a.k) kth_key = keys[a.k - 1];
    }
    __syncthreads();
    uint32_t thr = 0;  // fewer than k candidates: the whole context, keep everything
    if (ns >= a.k) thr = total_key(key_to_float((uint32_t)((~kth_key) >> 32)) - a.eps2);
    if (small) {
      if (tid < ns && (uint32_t)((~my_key) >> 32) >= thr) atomicMax(&n_keep, my_rank + 1);
    } else {
      for (uint32_t i = tid; i < ns; i += RS_THREADS)
        if ((uint32_t)((~keys[i]) >> 32) >= thr) atomicMax(&n_keep, i + 1);
    }
    __syncthreads();
    keep = n_keep;
    bad = keep > RS2_MAX;
  }
  if (bulk_q) rg_mbar_wait(rg_smem_u32(&qbar), 0);  // no CTA may exit with a copy into its memory pending
  if (bad) {
    if (tid == 0) {
      uint32_t slot = atomicAdd(a.flag_count, 1u);
      if (slot < a.flag_cap) a.flag_list[slot] = qo;
    }
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {  // visibly invalid until the caller's fp32 redo replaces it
      a.ids[(uint64_t)qo * a.k + i] = CM_NO_ID_DEV;
      a.score[(uint64_t)qo * a.k + i] = __int_as_float(0x7FC00000);
    }
    return;
  }
  // rows to re-score in score order, as HANDLES from here on (keys share their region with the row buffer)
  if (small) {
    if (tid < ns && my_rank < keep) {
      const uint32_t r = (uint32_t)(~my_key);
      rows[my_rank] = a.perm ? a.perm[r] : r;
    }
  } else {
    for (uint32_t i = tid; i < keep; i += RS_THREADS) {
      const uint32_t r = (uint32_t)(~keys[i]);
      rows[i] = a.perm ? a.perm[r] : r;
    }
  }
  __syncthreads();
  const uint32_t r_slot = tid >> 2, c4 = tid & 3;
  const float nb = a.qnorm2[qo];
  uint32_t phase = 0;
  for (uint32_t r0 = 0; r0 < keep; r0 += a.rows) {
    const uint32_t cntr = min(a.rows, keep - r0);
    rg_stage_rows(a.xraw, a.ldx, &mbar, rowbuf, row_stride, rows + r0, cntr, phase);
    const bool active = r_slot < cntr;
    const float dot = rg_quad_dot(qs, rowbuf + (size_t)r_slot * row_stride, a.dim, c4, active);
    if (active && c4 == 0) {
      const uint32_t row = rows[r0 + r_slot];
      // CosineDistance::distance on the stored rows (src/metric.rs:163-195)
      const float denom = __fsqrt_rn(__fmul_rn(a.norm2_raw[row], nb));
      float dist = 2.0f;  // zero vector: similarity undefined
      if (!(denom <= 1.1920929e-7f)) {
        float c = __fdiv_rn(dot, denom);
        c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);  // f32::clamp: NaN passes through
        dist = __fsub_rn(1.0f, c);
      }
      const float sc = combined_score_dev(dist, a.ts_secs[row], a.ts_nanos[row], a.decay[row], a.now_secs, a.now_nanos, a.w, a.base);
      dkeys[r0 + r_slot] = pack_key(sc, row);
    }
    __syncthreads();
  }
  if (keep <= RS_THREADS) {
    if (tid < keep) {
      const unsigned long long key = dkeys[tid];
      const uint32_t r = rank_among(dkeys, keep, key);
      if (r < a.k) {
        a.ids[(uint64_t)qo * a.k + r] = (uint32_t)key;
        a.score[(uint64_t)qo * a.k + r] = key_to_float((uint32_t)(key >> 32));
      }
    }
    for (uint32_t i = keep + tid; i < a.k; i += RS_THREADS) {
      a.ids[(uint64_t)qo * a.k + i] = CM_NO_ID_DEV;
      a.score[(uint64_t)qo * a.k + i] = __int_as_float(0x7F800000);
    }
  } else {
    const uint32_t P2 = next_pow2(keep);
    for (uint32_t i = keep + tid; i < P2; i += RS_THREADS) dkeys[i] = kKeyMax;
    block_bitonic_sort(dkeys, P2);
    for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
      unsigned long long key = i < P2 ? dkeys[i] : kKeyMax;
      bool ok = key != kKeyMax;
      a.ids[(uint64_t)qo * a.k + i] = ok ? (uint32_t)key : CM_NO_ID_DEV;
      a.score[(uint64_t)qo * a.k + i] = ok ? key_to_float((uint32_t)(key >> 32)) : __int_as_float(0x7F800000);
    }
  }
  if (tid == 0 && a.rescored_total) atomicAdd(a.rescored_total, (unsigned long long)keep);
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
uint32_t scan_tc_pad_queries(uint32_t nq) { return (nq + PAIR_M - 1) / PAIR_M * PAIR_M; }
uint32_t scan_tc_pad_k(uint32_t dim) { return (dim + BK - 1) / BK * BK; }
float scan_tc_eps(uint32_t dim) { return eps_for_dim(dim); }
uint32_t scan_tc_cand_cap(uint32_t k) { return k <= 16 ? 4096u : (k <= 64 ? 8192u : 16384u); }

bool scan_tc_supported(const DeviceIndex& ix, uint32_t k) {
  return ix.xbf16 != nullptr && k >= 1 && k <= 128 && encode_fn() != nullptr;
}

// Corpus splits per query-tile pair: minimise waves x (tiles per unit + a few tiles of pipeline fill / cold start).
// Every split of a query starts cold (its own running top-k) and emits ~k·(1 + ln(rows/k)) candidates before its
// threshold converges, so the number of splits is also bounded by the candidate buffer: splits·k·16 <= cap.
static uint32_t pick_splits(uint32_t m_pairs, uint32_t n_tiles, uint32_t clusters, uint32_t k, uint32_t cap, uint32_t per_thread = 1) {
  uint32_t best = 1;
  uint64_t best_cost = ~0ull;
  // (per_thread = 2: two epilogue threads per query and unit, each with its own threshold over half of the unit's rows)
  const uint32_t max_splits = std::max(1u, cap / (16u * k * per_thread));
  for (uint32_t s = 1; s <= std::min<uint32_t>(std::min<uint32_t>(n_tiles, 128), max_splits); ++s) {
    uint64_t tiles_per = (n_tiles + s - 1) / s;
    uint64_t waves = ((uint64_t)m_pairs * s + clusters - 1) / clusters;
    uint64_t cost = waves * (tiles_per + 3);
    if (cost < best_cost) best_cost = cost, best = s;
  }
  return best;
}

cudaError_t launch_scan_tc(const DeviceIndex& ix, const ScanTcLaunch& l, cudaStream_t s) {
  CUtensorMap map_q, map_x;
  const uint32_t nq_pad = scan_tc_pad_queries(l.nq);
  if (!make_map(&map_q, l.q_bf16, nq_pad, ix.k_pad, BM)) return cudaErrorInvalidValue;
  if (!make_map(&map_x, l.x_bf16 ? l.x_bf16 : ix.xbf16, ix.n_pad, ix.k_pad, BN_HALF)) return cudaErrorInvalidValue;
  const uint32_t clusters = (uint32_t)std::max(1, ix.sm_count / 2);
  ScanArgs a;
  a.nq = l.nq;
  a.m_pairs = nq_pad / PAIR_M;
  a.n_tiles = (uint32_t)(ix.n_pad / BN);
  a.pred = l.pred ? 1u : 0u;
  a.q_base = l.q_base;
  if (l.pred) {
    if (l.q_base % BN != 0) return cudaErrorInvalidValue;
    a.n_tiles = std::min<uint32_t>(a.n_tiles, (l.q_base + nq_pad) / BN);  // no query of this launch sees rows beyond that
  }
  a.k_blocks = ix.k_pad / BK;
  a.k = l.k;
  a.splits = pick_splits(a.m_pairs, a.n_tiles, clusters, l.k, l.cap);
  a.eps2 = 2.0f * eps_for_dim(ix.dim);
  a.tau_key = l.tau_key;
  a.cand_cnt = l.cand_cnt;
  a.cand = l.cand;
  a.cap = l.cap;
  a.join = l.join ? 1u : 0u;
  a.join_tau = l.join_tau;
  a.pairs = l.pairs;
  a.pair_count = l.pair_count;
  a.pair_cap = l.pair_cap;
  if (l.join) a.splits = 1;  // units = query-tile pairs, each from its diagonal tile to the end
  const bool biased = l.bias != nullptr;
  a.bias = l.bias;
  a.row_ctx = l.row_ctx;
  a.q_ctx = l.q_ctx;
  a.pair_lo = l.pair_lo;
  a.pair_hi = l.pair_hi;
  // context-sorted ranges are short: size the splits for the expected range length, not for the whole corpus
  if (l.pair_lo) a.splits = pick_splits(a.m_pairs, std::max(1u, std::min(l.pair_tiles_avg, a.n_tiles)), clusters, l.k, l.cap);
  // Large k without a bias: the per-thread threshold comes from a score histogram (Ladder) instead of an exact top-k list.
  // 256 buckets (bound step 1/128) unless the operand ring needs the 32 KB more (rows of 8 or more k-blocks): 128 buckets.
  bool ladder = !biased && !l.join && l.k > 32;
  if (const char* env = getenv("CM_SCAN_LADDER")) ladder = ladder && atoi(env) != 0;
  a.ladder_n = ladder ? (a.k_blocks >= 8 ? 128u : 256u) : 0u;
  // Short rows (<= 3 k-blocks: a tile's MMAs take <= 1.5k cycles) are bound by the epilogue: second epilogue warp-quad.
  bool epi2 = !biased && !l.join && !l.pred && a.k_blocks <= 3;
  if (const char* env = getenv("CM_SCAN_EPI2")) epi2 = (epi2 && atoi(env) != 0) || (atoi(env) == 2 && !biased && !l.join && !l.pred);  // 2: every row length
  if (epi2) a.splits = pick_splits(a.m_pairs, a.n_tiles, clusters, l.k, l.cap, 2);
  if (const char* env = getenv("CM_SCAN_SPLITS"))  // experiments
    if (!l.join && !l.pair_lo && atoi(env) > 0) a.splits = std::min<uint32_t>((uint32_t)atoi(env), a.n_tiles);
  a.slot_chunk = l.k > 32 ? 32u : 8u;
  a.diag_no_hits = getenv("CM_SCAN_DIAG") != nullptr ? 1u : 0u;
  const size_t lists_bytes = (epi2 ? 2 : 1) * (ladder ? (size_t)a.ladder_n * 128 * 2 : (size_t)l.k * 128 * 4);
  const size_t fixed = 1024 + lists_bytes + sizeof(ScanSmemTail) + 32 + (biased ? (size_t)kAccStages * BN * 8 : 0);
  // Shared-memory budget in 16 KB operand blocks: `q_res` resident query blocks + a ring of `slots` streamed blocks.
  // Every resident block removes one query block per corpus tile from the L2 -> SM stream (all-streamed the kernel asks
  // for 64 B/clk per SM at full tensor rate, against ~43-48 B/clk per SM the L2 delivers chip-wide). The ring needs six
  // blocks to cover the L2 latency (sweep in profiles/r02_scan_qres_sweep.txt: 7 resident + 6 streamed is the best split
  // of the 13 blocks that fit beside k = 10 lists; 6 + 5 or 9 + 4 starve the MMA). CM_SCAN_QRES / CM_SCAN_SLOTS override.
  const size_t room = l.leave_room && l.k <= 16 ? 40 * 1024 : 0;  // K2 (~34 KB + 1 KB reserved) beside this kernel
  const uint32_t blocks_max = (uint32_t)std::min<size_t>(kMaxSlots + 16, (kMaxSmem - fixed - room) / kBlockBytes);
  if (blocks_max < 4) return cudaErrorInvalidValue;
  // (11 blocks beside a rescore CTA: 4 + 7 measured 11.10 ms against 11.19 for 5 + 6 and 11.17 all-streamed)
  const uint32_t ring_min = std::min(room ? 7u : 6u, blocks_max);
  uint32_t q_res = std::min<uint32_t>(std::min<uint32_t>(a.k_blocks, 7), blocks_max - ring_min), slots = 0;
  if (const char* env = getenv("CM_SCAN_QRES"))
    q_res = std::min<uint32_t>(std::min<uint32_t>(a.k_blocks, (uint32_t)std::max(0, atoi(env))), blocks_max - 3);
  slots = std::min<uint32_t>(kMaxSlots, blocks_max - q_res);
  if (const char* env = getenv("CM_SCAN_SLOTS")) {
    const uint32_t f = (uint32_t)atoi(env);
    if (f >= 3 && f <= slots) slots = f;
  }
  a.q_res = q_res;
  a.slots = slots;
  const size_t smem = fixed + (size_t)(q_res + slots) * kBlockBytes;
  auto kernel = biased ? scan_tc_kernel<true, false, false, false>
                : l.pred ? (ladder ? scan_tc_kernel<false, true, true, false> : scan_tc_kernel<false, true, false, false>)
                : epi2   ? (ladder ? scan_tc_kernel<false, false, true, true> : scan_tc_kernel<false, false, false, true>)
                         : (ladder ? scan_tc_kernel<false, false, true, false> : scan_tc_kernel<false, false, false, false>);
  if (biased && l.pred) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
  if (e != cudaSuccess) return e;
  const uint32_t units = a.m_pairs * a.splits;
  const unsigned grid = 2u * std::min(clusters, units);  // whole CTA pairs (__cluster_dims__(2,1,1))
  kernel<<<grid, epi2 ? kThreadsEpi2 : kThreads, smem, s>>>(map_q, map_x, a);
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
  a.rs_max = l.k <= 16 ? 2048u : (l.k <= 64 ? 4096u : 8192u);
  a.approx = l.approx ? 1u : 0u;
  a.pred = l.pred ? 1u : 0u;
  a.pred_base = l.pred_base;
  a.debug = getenv("CM_RS_DEBUG") != nullptr;
  a.eps2 = 2.0f * eps_for_dim(ix.dim);
  const size_t row_stride = (size_t)ix.dim_pad * 4 + RG_ROW_PAD;
  const size_t other = (size_t)RS2_MAX * 12 + (size_t)ix.dim_pad * 4;
  // ~24 KB of row buffer (8 rows of 768 floats; a typical query re-scores ~15-30 rows = 2-4 rounds): the kernel is
  // latency-bound per CTA, so queries in flight per SM matter more than rounds.
  a.rows = std::min<uint32_t>(RS_ROWS, std::max<uint32_t>(4u, (uint32_t)((24u * 1024u + row_stride - 1) / row_stride)));
  if (const char* env = getenv("CM_RS_ROWS")) a.rows = std::min<uint32_t>(RS_ROWS, std::max(1, atoi(env)));  // sweeps
  while (a.rows > 1 && a.rows * row_stride + other > 200 * 1024) --a.rows;   // very long rows: fewer per round
  a.region = (uint32_t)((std::max<size_t>(a.rows * row_stride, (size_t)a.rs_max * 8) + 127) & ~(size_t)127);
  const size_t smem = a.region + other;
  {  // per device, so set on every launch (microseconds)
    cudaError_t e = cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  rescore_kernel<<<l.nq, RS_THREADS, smem, s>>>(a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_context_bias(const DeviceIndex& ix, float w, float base, uint64_t now_secs, uint32_t now_nanos, float* bias,
                                cudaStream_t s, const uint32_t* perm) {
  if (ix.n_pad == 0) return cudaSuccess;
  k_context_bias<<<148 * 4, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(ix.ts_secs), ix.ts_nanos, ix.decay, ix.n,
                                          ix.n_pad, w, base, now_secs, now_nanos, bias, perm);
  count_launch();
  return cudaGetLastError();
}

uint32_t scan_tc_context_chunk() { return kCtxSortChunk; }

cudaError_t launch_group_queries_by_context(const uint32_t* want, uint32_t nq, uint32_t* order, uint32_t* qctx,
                                            cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  const uint32_t chunks = (nq + kCtxSortChunk - 1) / kCtxSortChunk;
  const size_t smem = (size_t)next_pow2(std::min(nq, kCtxSortChunk) < 2 ? 2 : std::min(nq, kCtxSortChunk)) * 8;
  cudaError_t e = cudaFuncSetAttribute(k_group_queries_by_context, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kCtxSortChunk * 8));
  if (e != cudaSuccess) return e;
  k_group_queries_by_context<<<chunks, GQ_THREADS, smem, s>>>(want, nq, kCtxSortChunk, order, qctx);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_context_pair_ranges(const DeviceIndex& ix, const uint32_t* qctx, uint32_t nq, uint32_t* pair_lo,
                                       uint32_t* pair_hi, cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  const uint32_t chunks = (nq + kCtxSortChunk - 1) / kCtxSortChunk;
  const uint32_t n_pairs = chunks * (kCtxSortChunk / PAIR_M);
  k_context_pair_ranges<<<(n_pairs + 127) / 128, 128, 0, s>>>(qctx, nq, kCtxSortChunk, ix.ctx_sorted, (uint32_t)ix.n, n_pairs,
                                                              pair_lo, pair_hi);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rescore_context(const DeviceIndex& ix, const ContextRescoreLaunch& l, cudaStream_t s) {
  if (l.nq == 0) return cudaSuccess;
  ContextRescoreArgs a;
  a.q_raw = l.q_raw;
  a.ldq = l.ldq;
  a.nq = l.nq;
  a.k = l.k;
  a.dim = ix.dim;
  a.ldx = ix.dim_pad;
  a.qnorm2 = l.qnorm2;
  a.xraw = ix.xraw;
  a.norm2_raw = ix.norm2_raw;
  a.ts_secs = reinterpret_cast<const unsigned long long*>(ix.ts_secs);
  a.ts_nanos = ix.ts_nanos;
  a.decay = ix.decay;
  a.w = l.w;
  a.base = l.base;
  a.now_secs = l.now_secs;
  a.now_nanos = l.now_nanos;
  a.tau_key = l.tau_key;
  a.cand_cnt = l.cand_cnt;
  a.cand = l.cand;
  a.cap = l.cap;
  a.ids = l.ids;
  a.score = l.score;
  a.flag_count = l.flag_count;
  a.flag_list = l.flag_list;
  a.flag_cap = l.flag_cap;
  a.rescored_total = l.rescored_total;
  a.perm = l.perm;
  a.q_order = l.q_order;
  a.rs_max = l.k <= 16 ? 2048u : (l.k <= 64 ? 4096u : 8192u);
  // the stored rows are not unit length and the reference divides by fp32 norms: a little slack on top of the scan's bound
  a.eps2 = 2.0f * (eps_for_dim(ix.dim) + 2e-5f);
  const size_t row_stride = (size_t)ix.dim_pad * 4 + RG_ROW_PAD;
  const size_t other = (size_t)RS2_MAX * 12 + (size_t)ix.dim_pad * 4;
  a.rows = std::min<uint32_t>(RS_ROWS, std::max<uint32_t>(4u, (uint32_t)((24u * 1024u + row_stride - 1) / row_stride)));
  while (a.rows > 1 && a.rows * row_stride + other > 200 * 1024) --a.rows;
  a.region = (uint32_t)((std::max<size_t>(a.rows * row_stride, (size_t)a.rs_max * 8) + 127) & ~(size_t)127);
  const size_t smem = a.region + other;
  cudaError_t e = cudaFuncSetAttribute(rescore_context_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  rescore_context_kernel<<<l.nq, RS_THREADS, smem, s>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cm
