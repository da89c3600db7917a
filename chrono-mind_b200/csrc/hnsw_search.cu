// hnsw_search.cu — K3: batched HNSW search, one CTA (4-8 warps) per query.
//
// Replaces `LockFreeHnsw::search` (src/index/lockfree_hnsw.rs:469-495) and `search_layer` (:140-194) for a
// batch of prepared queries over the frozen, flattened graph (padded CSR, see cm::DeviceIndex):
//   greedy descent through layers entry_top..1 with ef = 1, then the layer-0 beam with ef candidates.
//
// How the reference's two heaps map onto a GPU:
//   * `best` (max-heap, cap ef) is a SORTED list of u64 keys (TotalF32 distance << 32 | handle << 1 | expanded)
//     in shared memory. `frontier` needs no storage: every admitted-but-unexpanded node that could still be
//     popped is exactly an unexpanded entry of `best` (a node evicted from `best` is >= worst and would hit the
//     `dist > worst` break), so "pop the frontier" == "first unexpanded list entry" and the loop ends when none
//     is left. Per expansion the <=32 fresh neighbors are merged into the list by rank (each key computes its
//     position in the union). The rank merge equals the reference's sequential admit/evict whenever no group of
//     bit-identical distances straddles the cut at ef (proof in DESIGN.md §4): the reference admits on distance
//     only, strictly (`d < worst`, :182), and evicts by (distance, id), so only inside such a tie group does the
//     ORDER of arrival decide who stays. When the union's keys at ranks ef-1 and ef carry the same distance bits
//     the expansion is replayed sequentially in neighbor-list order by one thread (`tie replay`). An entry that
//     was admitted, never expanded, and evicted while its distance still equals the worst distance stays in the
//     reference's frontier and is expanded before the loop breaks (`dist > worst` is strict, :166-169): those
//     entries wait in a small `limbo` array and are popped once the list has no unexpanded entry left.
//   * `visited` (HashSet<u32>) is an open-addressing table in shared memory, tagged with an 8-bit epoch so it
//     never has to be cleared between layers or queries. Membership is exact. When a query would crowd the
//     table past 3/4 the CTA restarts that query on a larger per-CTA table in global memory.
//     At large ef that table is the biggest shared-memory consumer (30 * ef entries) and caps the queries in flight
//     per SM, so from ef = 48 up `visited` moves to HBM instead: one BYTE per node per resident CTA, holding the epoch
//     of the search that last visited the node (a plain load + store per neighbor — the lanes of warp 0 look at
//     distinct nodes; 32 scattered sectors per expansion, ~2 % of the row bytes, and no capacity limit at all).
//   * the fresh neighbors' rows are GATHERED BY THE COPY ENGINE: one `cp.async.bulk` (UBLKCP) per row, issued by
//     one thread each, lands the whole dim*4-byte row in shared memory and signals an mbarrier — all rows of a
//     round (up to HN_ROWS x 3 KB) are in flight at once and cost no registers. (A register-staged gather keeps
//     only ~8 loads per thread in flight: 12 dependent DRAM round trips per expansion, measured 18 us/expansion.)
//   * distance_prepared is then evaluated bit-exactly out of shared memory: a QUAD of threads owns one row;
//     thread c of the quad carries the reference's SIMD lanes 2c, 2c+1 (src/metric.rs:129-147), i.e. two FMA
//     chains over elements 8t+2c, 8t+2c+1 (one LDS.64 per step; rows are padded by 32 bytes so the 8 rows of a
//     warp hit distinct banks), then the reference's reduction order and scalar tail.
//
// Gather-bound: per expansion one 128-byte neighbor list and up to 32 rows of dim*4 bytes. The traversal is a
// dependent chain (pop -> list -> visited -> rows -> merge), so HBM is kept busy by bytes in flight per query
// (bulk copies) times queries in flight (shared memory — row buffer, visited table — decides: 4-5 CTAs per SM).
#include <algorithm>
#include <cstdlib>

#include "device_common.cuh"
#include "kernels.h"
#include "row_gather.cuh"

namespace cm {

constexpr int HN_THREADS = 256;      // most threads per CTA; the launcher uses 128 for small ef (see hnsw_threads_for)
constexpr uint32_t HN_ROWS = 16;       // most rows staged per round (HN_THREADS / 4); the launcher may pick fewer
constexpr uint32_t HN_ROW_PAD = RG_ROW_PAD;
constexpr uint32_t HN_LIMBO = 48;      // evicted, unexpanded entries tied with the worst distance (see header)

struct HnswKernelArgs {
  const float* x;
  uint32_t ldx, dim;
  const uint32_t* nbr0;
  const uint32_t* upper_base;
  const uint32_t* nbr_up;
  uint32_t cap0, capU;
  const uint8_t* deleted;
  uint32_t entry_id, entry_top;
  const float* q;
  uint32_t nq, ldq, ef;
  uint32_t degenerate;  // query length != dim: every distance is 2.0 (src/metric.rs:220-222)
  uint32_t* pool_ids;
  float* pool_dist;
  unsigned long long* counters;  // [0] dist_evals [1] expansions [2] list_entries [3] global-table restarts [4] failures
                                 // [5] tie replays [6] CTAs that lost a tied evictee (limbo overflow)
  uint32_t* work_counter;
  uint32_t table_entries;  // shared-memory visited table (any size >= 64)
  uint32_t* gtable;        // [gridDim.x] slabs of 1 << ghash_bits entries
  uint32_t ghash_bits;
  uint32_t rows;           // rows staged per round, <= HN_ROWS
  uint32_t* status;        // optional: bumped once per query that could not be answered like the reference
  unsigned char* vmap;     // visited byte maps in HBM, one slab of vmap_stride bytes per CTA (nullptr: shared-memory table)
  uint64_t vmap_stride;
  uint32_t* vmap_epoch;    // [gridDim.x] epoch each CTA's slab was last written with (persists across launches)
};

struct HnShared {
  unsigned long long mbar;   // completion barrier of the row copies
  unsigned long long cand[32];
  uint32_t fresh[32];
  uint32_t list_count;
  uint32_t fresh_count;
  int cur_pos;
  uint32_t cur_id;
  unsigned long long next_key;  // nearest unexpanded entry after the popped one (kKeyMax: none)
  uint32_t query;
  uint32_t visited_count;
  uint32_t overflow;
  uint32_t epoch;
  unsigned long long ctr_evals, ctr_exp, ctr_entries, ctr_replays;
  // tie handling (see the header comment): the union's keys at ranks ef-1 / ef of the current merge, and the
  // evicted-but-still-poppable entries (all carry the current worst distance)
  unsigned long long bk0, bk1;
  uint32_t limbo_count, limbo_lost;
  unsigned long long limbo[HN_LIMBO];
};

struct VisitedTable {
  uint32_t* slots;
  uint32_t size;  // entries (not necessarily a power of two: shared memory is what limits the CTAs per SM)
  unsigned char* bytes;  // byte-map mode (HBM): bytes[id] == epoch <=> visited in this search_layer call
};

__device__ __forceinline__ uint32_t key_handle(unsigned long long key) { return ((uint32_t)key) >> 1; }
__device__ __forceinline__ unsigned long long make_key(float d, uint32_t id) {
  return ((unsigned long long)total_key(d) << 32) | ((unsigned long long)id << 1);
}

// HashSet::insert — true when `id` was not present. Called by the lanes of warp 0 concurrently. The caller
// keeps the load below 3/4, so the probe always ends.
__device__ __forceinline__ bool visited_insert(const VisitedTable& t, uint32_t tag, uint32_t id) {
  if (t.bytes) {  // concurrent lanes hold distinct ids (a neighbor list has no duplicates): plain load + store
    const unsigned char e = (unsigned char)(tag >> 24);
    if (t.bytes[id] == e) return false;
    t.bytes[id] = e;
    return true;
  }
  uint32_t want = tag | id;
  uint32_t h = __umulhi(id * 2654435761u, t.size);  // multiply-shift range reduction: [0, size)
  for (;;) {
    uint32_t old = t.slots[h];
    if (old == want) return false;
    if ((old & 0xFF000000u) != tag) {  // empty, or left over from another epoch: claim it
      uint32_t prev = atomicCAS(&t.slots[h], old, want);
      if (prev == old) return true;
      if (prev == want) return false;
      continue;  // another lane took the slot for a different id of this epoch: look at it again
    }
    h = h + 1 == t.size ? 0 : h + 1;
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One search_layer call over shared state. `lists` holds two ping-pong buffers of `ef_cap` keys; returns which
// one holds the result (ascending). All threads of the CTA call this.
__device__ int search_layer_dev(const HnswKernelArgs& a, HnShared* sh, const float* qs, unsigned long long* lists,
                                uint32_t ef_cap, const VisitedTable& vt, uint32_t ef, uint32_t layer, uint32_t ep,
                                unsigned char* rowbuf, uint32_t row_stride, uint32_t& phase) {
  const uint64_t vt_bytes_len = a.vmap_stride;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t r_slot = tid >> 2, c4 = tid & 3;
  const uint32_t cap = layer == 0 ? a.cap0 : a.capU;
  const uint32_t table_size = vt.size;
  const uint32_t limit = vt.bytes ? 0xFFFFFFFFu - 64u : (table_size >> 1) + (table_size >> 2);  // the byte map never fills
  int buf = 0;

  // A fresh visited set per call (:148): bump the epoch; clear the table when the 8-bit tag wraps.
  if (tid == 0) {
    sh->epoch += 1;
    sh->visited_count = 0;
    sh->limbo_count = 0;
  }
  __syncthreads();
  if (sh->epoch == 256) {
    if (vt.bytes) {
      uint4* w = reinterpret_cast<uint4*>(vt.bytes);
      for (uint64_t i = tid; i < vt_bytes_len / 16; i += blockDim.x) w[i] = make_uint4(0, 0, 0, 0);
    } else {
      for (uint32_t i = tid; i < table_size; i += blockDim.x) vt.slots[i] = 0;
    }
    __syncthreads();
    if (tid == 0) sh->epoch = 1;
    __syncthreads();
  }
  const uint32_t tag = sh->epoch << 24;

  // Entry point: pushed to both heaps (:152-160).
  if (tid == 0) {
    visited_insert(vt, tag, ep);
    sh->visited_count = 1;
  }
  {
    if (tid == 0) sh->fresh[0] = ep;
    __syncthreads();
    if (!a.degenerate) rg_stage_rows(a.x, a.ldx, &sh->mbar, rowbuf, row_stride, sh->fresh, 1, phase);
    float d = a.degenerate ? 2.0f : rg_quad_distance(qs, rowbuf, a.dim, c4, r_slot == 0);
    if (tid == 0) {
      lists[0] = make_key(d, ep);
      sh->list_count = 1;
      sh->ctr_evals += 1;
    }
  }
  __syncthreads();

  for (;;) {
    unsigned long long* L = lists + (size_t)buf * ef_cap;
    // pop: first unexpanded entry of the ascending list == nearest frontier element (:165)
    if (warp == 0) {
      uint32_t cnt = sh->list_count;
      int pos = -1;
      for (uint32_t base = 0; base < cnt; base += 32) {
        uint32_t i = base + lane;
        bool un = i < cnt && !(L[i] & 1ull);
        uint32_t m = __ballot_sync(0xFFFFFFFFu, un);
        if (m) {
          pos = (int)(base + __ffs(m) - 1);
          break;
        }
      }
      // The next pop is either the following unexpanded entry or a neighbor admitted by this expansion: start
      // pulling the former's neighbor list into L2 now (the latter's follows once its distance is known).
      unsigned long long nxt = kKeyMax;
      if (pos >= 0 && layer == 0) {
        for (uint32_t base = (uint32_t)pos + 1; base < cnt; base += 32) {
          uint32_t i = base + lane;
          bool un = i < cnt && !(L[i] & 1ull);
          uint32_t m2 = __ballot_sync(0xFFFFFFFFu, un);
          if (m2) {
            nxt = L[base + __ffs(m2) - 1];
            break;
          }
        }
      }
      if (lane == 0) {
        sh->cur_pos = pos;
        sh->next_key = nxt;
        if (pos >= 0) {
          sh->cur_id = key_handle(L[pos]);
          L[pos] |= 1ull;
          sh->ctr_exp += 1;
        } else if (sh->limbo_count) {
          // The list has nothing left to expand, but the reference's frontier still holds evicted entries whose
          // distance equals the worst one: `dist > worst` (:166-169) is strict, so the nearest of them is expanded.
          uint32_t bi = 0;
          for (uint32_t i = 1; i < sh->limbo_count; ++i)
            if (sh->limbo[i] < sh->limbo[bi]) bi = i;
          sh->cur_id = key_handle(sh->limbo[bi]);
          sh->limbo[bi] = sh->limbo[--sh->limbo_count];
          sh->cur_pos = 0;
          sh->ctr_exp += 1;
        }
        if (nxt != kKeyMax) prefetch_l2(a.nbr0 + (uint64_t)key_handle(nxt) * a.cap0);
      }
    }
    __syncthreads();
    if (sh->cur_pos < 0 || sh->overflow) break;
    const uint32_t cur = sh->cur_id;
    const uint32_t* nlist = layer == 0 ? a.nbr0 + (uint64_t)cur * a.cap0
                                       : a.nbr_up + (uint64_t)(__ldg(a.upper_base + cur) + layer - 1) * a.capU;

    for (uint32_t c0 = 0; c0 < cap; c0 += 32) {
      // visited filter in list order (:174-177), warp 0
      if (warp == 0) {
        bool room = sh->visited_count + 32 <= limit;
        uint32_t i = c0 + lane;
        uint32_t nb = i < cap ? __ldg(nlist + i) : CM_NO_ID_DEV;
        bool present = nb != CM_NO_ID_DEV;
        bool is_new = room && present && visited_insert(vt, tag, nb);
        uint32_t pm = __ballot_sync(0xFFFFFFFFu, present);
        uint32_t nm = __ballot_sync(0xFFFFFFFFu, is_new);
        if (is_new) sh->fresh[__popc(nm & ((1u << lane) - 1))] = nb;
        __syncwarp();
        if (lane == 0) {
          uint32_t m = __popc(nm);
          sh->fresh_count = m;
          sh->bk0 = sh->bk1 = kKeyMax;
          sh->visited_count += m;
          sh->ctr_entries += __popc(pm);
          sh->ctr_evals += m;
          if (!room) sh->overflow = 1;
        }
      }
      __syncthreads();
      const uint32_t m = sh->fresh_count;
      if (sh->overflow) break;  // uniform: written before the barrier
      if (m == 0) {
        __syncthreads();
        continue;
      }
      // distances (:181): rounds of up to HN_ROWS rows — bulk-copied to shared memory, one thread quad per row
      for (uint32_t r0 = 0; r0 < m; r0 += a.rows) {
        const uint32_t cntr = min(a.rows, m - r0);
        if (!a.degenerate) rg_stage_rows(a.x, a.ldx, &sh->mbar, rowbuf, row_stride, sh->fresh + r0, cntr, phase);
        const bool active = r_slot < cntr;
        float d = a.degenerate ? 2.0f : rg_quad_distance(qs, rowbuf + (size_t)r_slot * row_stride, a.dim, c4, active);
        if (active && c4 == 0) sh->cand[r0 + r_slot] = make_key(d, sh->fresh[r0 + r_slot]);
        __syncthreads();  // cand visible to the merge; rowbuf free for the next round
      }
      if (warp == 0 && layer == 0) {  // nearest fresh neighbor pops before next_key: prefetch its list instead
        unsigned long long best = lane < m ? sh->cand[lane] : kKeyMax;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
          best = other < best ? other : best;
        }
        if (lane == 0 && best < sh->next_key) prefetch_l2(a.nbr0 + (uint64_t)key_handle(best) * a.cap0);
      }
      // admit + evict (:182-189) as a merge by rank into the other buffer
      {
        const uint32_t cnt = sh->list_count;
        unsigned long long* Lsrc = lists + (size_t)buf * ef_cap;
        unsigned long long* Ldst = lists + (size_t)(buf ^ 1) * ef_cap;
        for (uint32_t i = tid; i < cnt + m; i += blockDim.x) {
          unsigned long long key = i < cnt ? Lsrc[i] : sh->cand[i - cnt];
          uint32_t pos = 0;
          for (uint32_t c = 0; c < m; ++c) pos += sh->cand[c] < key;
          if (i < cnt) {
            pos += i;
          } else {  // list keys below `key`: binary search (keys are unique)
            uint32_t lo = 0, hi = cnt;
            while (lo < hi) {
              uint32_t mid = (lo + hi) >> 1;
              if (Lsrc[mid] < key)
                lo = mid + 1;
              else
                hi = mid;
            }
            pos += lo;
          }
          if (pos < ef) Ldst[pos] = key;
          if (pos + 1 == ef) sh->bk0 = key;  // the union's keys on either side of the cut
          if (pos == ef) sh->bk1 = key;
        }
        __syncthreads();
        // Equal distance bits on both sides of the cut: the reference's result depends on arrival order there
        // (strict `d < worst` admission, :182; eviction by (distance, id), :186-188). Replay this expansion
        // sequentially in neighbor-list order. Uniform branch (bk0 / bk1 were written before the barrier).
        const bool tie = sh->bk1 != kKeyMax && (uint32_t)(sh->bk0 >> 32) == (uint32_t)(sh->bk1 >> 32);
        if (tie) {
          for (uint32_t i = tid; i < cnt; i += blockDim.x) Ldst[i] = Lsrc[i];
          __syncthreads();
          if (tid == 0) {
            uint32_t n = cnt;
            for (uint32_t c = 0; c < m; ++c) {
              const unsigned long long key = sh->cand[c];
              uint32_t lim = n;  // list entries that stay
              if (n >= ef) {
                const unsigned long long worst = Ldst[n - 1];
                if (!((uint32_t)(key >> 32) < (uint32_t)(worst >> 32))) continue;  // not admitted (:182)
                if (!(worst & 1ull)) {  // evicted before it was expanded: still in the reference's frontier
                  if (sh->limbo_count < HN_LIMBO)
                    sh->limbo[sh->limbo_count++] = worst;
                  else
                    sh->limbo_lost = 1;
                }
                lim = n - 1;
              }
              uint32_t j = lim;
              while (j > 0 && Ldst[j - 1] > key) {
                Ldst[j] = Ldst[j - 1];
                --j;
              }
              Ldst[j] = key;
              n = lim + 1;
            }
            // evicted entries beyond the worst distance would hit the break: only the tied ones stay poppable
            const uint32_t wd = (uint32_t)(Ldst[n - 1] >> 32);
            uint32_t keep = 0;
            for (uint32_t i = 0; i < sh->limbo_count; ++i)
              if ((uint32_t)(sh->limbo[i] >> 32) == wd) sh->limbo[keep++] = sh->limbo[i];
            sh->limbo_count = keep;
            sh->list_count = n;
            sh->ctr_replays += 1;
          }
        } else if (tid == 0) {
          const uint32_t n = min(ef, cnt + m);
          sh->list_count = n;
          // the worst distance moved below the tied evictees: they can no longer be popped
          if (sh->limbo_count && (uint32_t)(Ldst[n - 1] >> 32) != (uint32_t)(sh->limbo[0] >> 32)) sh->limbo_count = 0;
        }
        buf ^= 1;
        __syncthreads();
      }
    }
    if (sh->overflow) break;
  }
  return buf;
}

__global__ void __launch_bounds__(HN_THREADS) k_hnsw_search(HnswKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t ef_cap = a.ef;
  float* qs = reinterpret_cast<float*>(smem_raw);
  size_t off = ((size_t)a.dim * 4 + 15) & ~(size_t)15;
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(smem_raw + off);
  off += (size_t)2 * ef_cap * 8;
  uint32_t* stable = reinterpret_cast<uint32_t*>(smem_raw + off);
  off += (size_t)4 * a.table_entries;
  off = (off + 15) & ~(size_t)15;
  HnShared* sh = reinterpret_cast<HnShared*>(smem_raw + off);
  off = (off + sizeof(HnShared) + 127) & ~(size_t)127;
  unsigned char* rowbuf = smem_raw + off;                 // HN_ROWS rows, 128-byte aligned
  const uint32_t row_stride = a.ldx * 4 + HN_ROW_PAD;

  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < a.table_entries; i += blockDim.x) stable[i] = 0;
  if (tid == 0) {
    sh->epoch = a.vmap ? a.vmap_epoch[blockIdx.x] : 0;  // the byte map outlives the launch: so does its epoch
    sh->ctr_evals = sh->ctr_exp = sh->ctr_entries = sh->ctr_replays = 0;
    sh->limbo_count = sh->limbo_lost = 0;
    rg_mbar_init(&sh->mbar);
  }
  uint32_t phase = 0;  // parity of the copy barrier (every thread waits on every round, so all copies agree)
  __syncthreads();

  for (;;) {
    if (tid == 0) sh->query = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t qi = sh->query;
    if (qi >= a.nq) break;
    for (uint32_t e = tid; e < a.dim; e += blockDim.x) qs[e] = a.degenerate ? 0.0f : a.q[(uint64_t)qi * a.ldq + e];

    bool use_global = false;
    int b = 0;
    for (;;) {
      VisitedTable vt;
      vt.bytes = nullptr;
      if (a.vmap) {
        vt.bytes = a.vmap + (size_t)blockIdx.x * a.vmap_stride;
        vt.slots = nullptr;
        vt.size = 0;
      } else if (use_global) {
        vt.slots = a.gtable + ((size_t)blockIdx.x << a.ghash_bits);
        vt.size = 1u << a.ghash_bits;
        for (uint32_t i = tid; i < vt.size; i += blockDim.x) vt.slots[i] = 0;
      } else {
        vt.slots = stable;
        vt.size = a.table_entries;
      }
      if (tid == 0) sh->overflow = 0;
      __syncthreads();

      uint32_t ep = a.entry_id;
      for (uint32_t layer = a.entry_top; layer >= 1 && !sh->overflow; --layer) {  // descend_layer (:196-201)
        int bb = search_layer_dev(a, sh, qs, lists, ef_cap, vt, 1, layer, ep, rowbuf, row_stride, phase);
        ep = key_handle(lists[(size_t)bb * ef_cap]);
        __syncthreads();
      }
      if (!sh->overflow) b = search_layer_dev(a, sh, qs, lists, ef_cap, vt, a.ef, 0, ep, rowbuf, row_stride, phase);
      __syncthreads();
      if (!sh->overflow) break;
      if (use_global) {  // cannot happen for sane ef; reported, never silent
        if (tid == 0) atomicAdd(&a.counters[4], 1ull);
        if (tid == 0 && a.status) atomicAdd(a.status, 1u);
        if (tid == 0) sh->list_count = 0;
        __syncthreads();
        break;
      }
      use_global = true;
      if (tid == 0) atomicAdd(&a.counters[3], 1ull);
    }
    if (use_global) {  // the shared table missed epochs while the global one was in use: start it over
      for (uint32_t i = tid; i < a.table_entries; i += blockDim.x) stable[i] = 0;
      if (tid == 0) sh->epoch = 0;
      __syncthreads();
    }
    const unsigned long long* L = lists + (size_t)b * ef_cap;

    // tombstone filter on the final list only (:488-492), order preserved; pad to ef.
    if (tid < 32) {
      uint32_t cnt = sh->list_count, outp = 0;
      for (uint32_t base = 0; base < cnt; base += 32) {
        uint32_t i = base + tid;
        unsigned long long key = i < cnt ? L[i] : 0;
        uint32_t id = key_handle(key);
        bool keep = i < cnt && !(a.deleted && a.deleted[id]);
        uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
          uint32_t o = outp + __popc(km & ((1u << tid) - 1));
          a.pool_ids[(uint64_t)qi * a.ef + o] = id;
          a.pool_dist[(uint64_t)qi * a.ef + o] = key_to_float((uint32_t)(key >> 32));
        }
        outp += __popc(km);
      }
      for (uint32_t o = outp + tid; o < a.ef; o += 32) {
        a.pool_ids[(uint64_t)qi * a.ef + o] = CM_NO_ID_DEV;
        a.pool_dist[(uint64_t)qi * a.ef + o] = __int_as_float(0x7F800000);
      }
    }
    __syncthreads();
  }
  if (tid == 0 && a.vmap) a.vmap_epoch[blockIdx.x] = sh->epoch;
  if (tid == 0 && a.counters) {
    atomicAdd(&a.counters[0], sh->ctr_evals);
    atomicAdd(&a.counters[1], sh->ctr_exp);
    atomicAdd(&a.counters[2], sh->ctr_entries);
    if (sh->ctr_replays) atomicAdd(&a.counters[5], sh->ctr_replays);
    if (sh->limbo_lost) {  // a tie group larger than HN_LIMBO was cut short somewhere in this CTA's queries: reported
      atomicAdd(&a.counters[6], 1ull);
      if (a.status) atomicAdd(a.status, 1u);
    }
  }
}

// Threads per CTA: the row distances need only 4 lanes per staged row; the rest of the CTA shortens the list merge, the
// table clears and the pop scan, which grow with ef — but every extra warp also lowers the queries resident per SM.
// Round-2 sweep on 1M x 768 (tools/diag_hnsw_rows.py, profiles/r02_diag_hnsw_sweep.txt; ms per 10k batch at 64/128/256
// threads, best rows, visited set where it is fastest): ef=32 3.92/3.84/4.12, ef=64 6.69/6.44/7.0, ef=128
// 12.06/11.62/13.4, ef=200 19.17/17.86/20.8 => 128 threads throughout the measured range.
static int hnsw_threads_for(uint32_t ef) {
  if (const char* env = getenv("CM_HNSW_THREADS")) {  // experiments only
    const int t = atoi(env);
    if (t == 64 || t == 128 || t == 256) return t;
  }
  return ef <= 256 ? 128 : HN_THREADS;
}

static size_t hnsw_smem_bytes(uint32_t dim, uint32_t ef, uint32_t table_entries, uint32_t rows) {
  const size_t dim_pad = (dim + 7) & ~7u;
  return (((size_t)dim * 4 + 15) & ~(size_t)15) + (size_t)2 * ef * 8 + (size_t)4 * table_entries + 16 + sizeof(HnShared) +
         128 + (size_t)rows * (dim_pad * 4 + HN_ROW_PAD);
}

// Rows per round vs CTAs per SM: fewer rows per round leave shared memory for more queries in flight but add a
// dependent copy round per expansion (an expansion evaluates ~17 of its 32 neighbors on average). Round-2 sweep
// (same tool; ms per 10k batch at rows = 12/10/8/6, 128 threads): ef=32 (table) -/4.22/3.84/3.92, ef=64 (byte map)
// -/-/6.60/6.44, ef=128 (byte map) -/13.39/12.09/11.62, ef=200 (byte map) 21.11/-/17.86/18.79 => 8 rows (~25 KB of
// row buffer at 768 floats) for small and very large ef, 6 in between; short rows take the full 16.
cudaError_t hnsw_grid_size(const DeviceIndex& ix, uint32_t ef, uint32_t table_entries, unsigned* grid, uint32_t* rows_out) {
  cudaError_t e = cudaFuncSetAttribute(k_hnsw_search, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return e;
  const uint32_t row_stride = ix.dim_pad * 4 + HN_ROW_PAD;
  const uint32_t budget = (ef < 48 || ef >= 160) ? 25u * 1024u : 19u * 1024u;  // 8 or 6 rows of 768 floats
  uint32_t rows = std::min(HN_ROWS, std::max(4u, budget / row_stride));
  if (const char* env = getenv("CM_HNSW_ROWS")) {  // experiments only
    uint32_t f = (uint32_t)atoi(env);
    if (f) rows = std::min(std::max(f, 1u), HN_ROWS);
  }
  while (rows > 1 && hnsw_smem_bytes(ix.dim, ef, table_entries, rows) > 220 * 1024) --rows;
  if (hnsw_smem_bytes(ix.dim, ef, table_entries, rows) > 220 * 1024) return cudaErrorInvalidValue;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hnsw_search, hnsw_threads_for(ef),
                                                    hnsw_smem_bytes(ix.dim, ef, table_entries, rows));
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  *grid = (unsigned)ix.sm_count * per_sm;
  *rows_out = rows;
  return cudaSuccess;
}

cudaError_t launch_hnsw_search(const DeviceIndex& ix, const HnswLaunch& l, cudaStream_t s) {
  if (l.nq == 0) return cudaSuccess;
  HnswKernelArgs a;
  a.x = ix.x32;
  a.ldx = ix.dim_pad;
  a.dim = ix.dim;
  a.nbr0 = ix.nbr0;
  a.upper_base = ix.upper_base;
  a.nbr_up = ix.nbr_up;
  a.cap0 = ix.cap0;
  a.capU = ix.capU;
  a.deleted = ix.deleted;
  a.entry_id = ix.entry_id;
  a.entry_top = ix.entry_top;
  a.q = l.q;
  a.nq = l.nq;
  a.ldq = l.ldq;
  a.ef = l.ef;
  a.degenerate = l.degenerate;
  a.pool_ids = l.pool_ids;
  a.pool_dist = l.pool_dist;
  a.counters = l.counters;
  a.work_counter = l.work_counter;
  a.table_entries = l.table_entries;
  a.gtable = l.gtable;
  a.ghash_bits = l.ghash_bits;
  a.rows = l.rows;
  a.status = l.status;
  a.vmap = l.vmap;
  a.vmap_stride = l.vmap_stride;
  a.vmap_epoch = l.vmap_epoch;
  cudaError_t e = cudaMemsetAsync(l.work_counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  unsigned grid = l.grid < l.nq ? l.grid : l.nq;
  k_hnsw_search<<<grid, hnsw_threads_for(l.ef), hnsw_smem_bytes(ix.dim, l.ef, l.table_entries, l.rows), s>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cm
