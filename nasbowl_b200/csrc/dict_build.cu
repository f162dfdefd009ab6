// Dictionary build: global LSD radix sort of 64-bit WL keys + unique + dense id assignment.
// Replaces `sorted(list(label_set))` and the running id counter of
// grakel_replace/weisfeiler_lehman.py:226-229,271-274 (ids here are local to the level and
// ordered by key; only the partition of nodes into label classes enters the kernel value).
//
// Sort: 8-bit digits, stable, three kernels per pass (per-tile digit counts -> exclusive scan
// of the digit-major count matrix -> ranked scatter).  A tile is 4096 keys: 8 warps x 16
// rounds x 32 lanes; ranks inside a round come from __match_any_sync, so there are no
// shared-memory atomics and the order of equal digits is preserved.
#include "nbw_common.cuh"

#include <stdlib.h>

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRounds = 16;
constexpr int kTile = kSortWarps * kRounds * 32;   // 4096

// ---------------------------------------------------------------- exclusive scan (u32)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) smem_warp[lane] = winc - w;
    if (lane == 31) smem_warp[32] = winc;
  }
  __syncthreads();
  total = smem_warp[32];
  uint32_t r = smem_warp[warp] + inc - v;
  __syncthreads();
  return r;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  uint32_t total;
  block_exclusive_scan(s, sw, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// pass 3 (or the only pass when one tile): scan inside the tile, add tile offset
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t* in, uint32_t* out, int64_t n,   // in may alias out
                  const uint32_t* tile_offsets) {
  __shared__ uint32_t sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, sw, total);
  ex += tile_offsets ? tile_offsets[blockIdx.x] : 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

size_t scan_scratch_elems(int64_t n) {
  size_t total = 0;
  while (n > kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    total += (size_t)n + 64;
  }
  return total + 64;
}

// Exclusive scan of n u32 (in-place allowed); tmp has scan_scratch_elems(n) elements.
int device_exclusive_scan(nbw_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n,
                          uint32_t* tmp, cudaStream_t st) {
  if (n <= 0) return NBW_OK;
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    scan_apply_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
    NBW_CHECK_LAUNCH(ctx);
    return NBW_OK;
  }
  uint32_t* sums = tmp;
  scan_reduce_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, sums);
  NBW_CHECK_LAUNCH(ctx);
  int rc = device_exclusive_scan(ctx, sums, sums, tiles, tmp + tiles + 64, st);
  if (rc) return rc;
  scan_apply_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

// ---------------------------------------------------------------- radix sort passes
// Per-warp digit counts of the warp's strip of the tile -> cnt[warp][digit].
__device__ __forceinline__ void strip_digit_counts(const uint64_t* __restrict__ keys, int64_t n,
                                                   int shift, uint32_t (*cnt)[256]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  const int64_t strip = (int64_t)blockIdx.x * kTile + (int64_t)warp * (kRounds * 32);
  for (int r = 0; r < kRounds; ++r) {
    const int64_t i = strip + r * 32 + lane;
    const uint32_t digit = (i < n) ? (uint32_t)((keys[i] >> shift) & 0xFFu) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    if (digit < 256u && (peers & ((1u << lane) - 1u)) == 0u) cnt[warp][digit] += __popc(peers);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_count_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int n_tiles,
                   uint32_t* __restrict__ counts /* [256][n_tiles] */) {
  __shared__ uint32_t cnt[kSortWarps][256];
  strip_digit_counts(keys, n, shift, cnt);
  __syncthreads();
  uint32_t s = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) s += cnt[w][threadIdx.x];
  counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s;
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                     int shift, int n_tiles, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t cnt[kSortWarps][256];
  strip_digit_counts(keys_in, n, shift, cnt);
  __syncthreads();
  {
    uint32_t base = offsets[(size_t)threadIdx.x * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = base;
      base += c;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t strip = (int64_t)blockIdx.x * kTile + (int64_t)warp * (kRounds * 32);
  for (int r = 0; r < kRounds; ++r) {
    const int64_t i = strip + r * 32 + lane;
    const bool valid = i < n;
    const uint64_t key = valid ? keys_in[i] : 0;
    const uint32_t digit = valid ? (uint32_t)((key >> shift) & 0xFFu) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid) {
      const uint32_t pos = cnt[warp][digit] + rank;
      keys_out[pos] = key;
      vals_out[pos] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    __syncwarp();
    if (valid && rank == 0) cnt[warp][digit] += __popc(peers);
    __syncwarp();
  }
}

// ---------------------------------------------------------------- unique / ids / verification
// ---------------------------------------------------------------- small inputs: one-CTA sort
// A training set is a few hundred cells: a few thousand keys.  The eight radix passes above are then
// ~30 launches of almost empty kernels per WL level; one CTA sorting (key, index) pairs in shared memory
// (bitonic network, lexicographic on (key, original index), so the result equals the stable radix
// order) does the same in one launch.
constexpr int kSmallSortMax = 8192;    // measured: at 16000 keys the radix passes win (1.2 vs 1.6 ms per fit)
__global__ void __launch_bounds__(1024)
small_sort_kernel(const uint64_t* __restrict__ keys, int n, int n_pad, uint64_t* __restrict__ out_k,
                  uint32_t* __restrict__ out_v) {
  extern __shared__ uint64_t sk[];
  uint32_t* sv = reinterpret_cast<uint32_t*>(sk + n_pad);
  for (int i = threadIdx.x; i < n_pad; i += 1024) {
    sk[i] = i < n ? keys[i] : ~0ull;      // padding sorts behind every real entry: same key at most, larger index
    sv[i] = (uint32_t)i;
  }
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pad >> 1); t += 1024) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const uint64_t ka = sk[i], kb = sk[l];
        const uint32_t va = sv[i], vb = sv[l];
        const bool gt = ka > kb || (ka == kb && va > vb);
        if (gt == ((i & k) == 0)) {
          sk[i] = kb; sk[l] = ka;
          sv[i] = vb; sv[l] = va;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += 1024) {
    out_k[i] = sk[i];
    out_v[i] = sv[i];
  }
}

__global__ void mark_heads_kernel(const uint64_t* __restrict__ sk, int64_t n, uint32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = sk[i];
    flags[i] = (k != NBW_INVALID_KEY && (i == 0 || sk[i - 1] != k)) ? 1u : 0u;
  }
}

// Exact comparison of two credentials (own id, multiset of successor ids).
template <int NMAX>
__device__ bool same_credential(const uint8_t* cells, const uint32_t* ids_prev, uint32_t a, uint32_t b) {
  if (ids_prev[a] != ids_prev[b]) return false;
  uint32_t la[NMAX], lb[NMAX];
  int na = 0, nb = 0;
  {
    const int64_t g = a / NMAX;
    const uint32_t m = CellTraits<NMAX>::succ(cells, g, a % NMAX);
    for (int j = 0; j < NMAX; ++j)
      if ((m >> j) & 1u) {
        uint32_t v = ids_prev[g * NMAX + j];
        int p = na++;
        while (p > 0 && la[p - 1] > v) { la[p] = la[p - 1]; --p; }
        la[p] = v;
      }
  }
  {
    const int64_t g = b / NMAX;
    const uint32_t m = CellTraits<NMAX>::succ(cells, g, b % NMAX);
    for (int j = 0; j < NMAX; ++j)
      if ((m >> j) & 1u) {
        uint32_t v = ids_prev[g * NMAX + j];
        int p = nb++;
        while (p > 0 && lb[p - 1] > v) { lb[p] = lb[p - 1]; --p; }
        lb[p] = v;
      }
  }
  if (na != nb) return false;
  for (int j = 0; j < na; ++j)
    if (la[j] != lb[j]) return false;
  return true;
}

typedef NbwDictCounters DictCounters;

__global__ void assign_ids_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ sv,
                                  const uint64_t* __restrict__ chks, const uint32_t* __restrict__ flags,
                                  const uint32_t* __restrict__ rank, int64_t n,
                                  const uint8_t* __restrict__ cells, int n_max,
                                  const uint32_t* __restrict__ ids_prev,
                                  uint32_t* __restrict__ ids, uint64_t* __restrict__ uniq_keys,
                                  uint64_t* __restrict__ uniq_chks, int64_t cap,
                                  DictCounters* __restrict__ counters, uint8_t* __restrict__ first_flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = sk[i];
    const uint32_t src = sv[i];
    if (i == n - 1) counters->n_unique = rank[i] + flags[i];
    if (k == NBW_INVALID_KEY) {
      ids[src] = NBW_INVALID_ID;
      if (first_flag) first_flag[src] = 0;
      continue;
    }
    const uint32_t head = flags[i];
    const uint32_t id = rank[i] + head - 1u;
    ids[src] = id;
    if (first_flag) first_flag[src] = (uint8_t)head;     // the sort is stable: the head of a run is the first occurrence
    if (head) {
      if ((int64_t)id < cap) {
        uniq_keys[id] = k;
        uniq_chks[id] = chks[src];
      }
    } else {
      const uint32_t prev = sv[i - 1];
      bool ok = chks[src] == chks[prev];
      if (ok && cells != nullptr && ids_prev != nullptr) {
        ok = (n_max == 8) ? same_credential<8>(cells, ids_prev, prev, src)
                          : same_credential<16>(cells, ids_prev, prev, src);
      }
      if (!ok) atomicOr(&counters->collision, 1u);
    }
  }
}

// Small inputs, all of it in ONE launch: the single-CTA sort above, then head flags, their block-wide exclusive scan,
// the ids, the dictionary and the credential check, without leaving shared memory (a training set's level used to be
// sort + heads + three scan launches + assign: six launches of near-empty kernels).
__global__ void __launch_bounds__(1024)
small_dict_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ chks, int n, int n_pad,
                  const uint8_t* __restrict__ cells, int n_max, const uint32_t* __restrict__ ids_prev,
                  uint32_t* __restrict__ ids, uint64_t* __restrict__ uniq_keys, uint64_t* __restrict__ uniq_chks, int64_t cap,
                  DictCounters* __restrict__ counters, uint8_t* __restrict__ first_flag) {
  extern __shared__ uint64_t sk[];
  uint32_t* sv = reinterpret_cast<uint32_t*>(sk + n_pad);
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t collided;
  const int tid = threadIdx.x;
  if (tid == 0) collided = 0u;
  for (int i = tid; i < n_pad; i += 1024) {
    sk[i] = i < n ? keys[i] : ~0ull;
    sv[i] = (uint32_t)i;
  }
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n_pad >> 1); t += 1024) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const uint64_t ka = sk[i], kb = sk[l];
        const uint32_t va = sv[i], vb = sv[l];
        const bool gt = ka > kb || (ka == kb && va > vb);
        if (gt == ((i & k) == 0)) {
          sk[i] = kb; sk[l] = ka;
          sv[i] = vb; sv[l] = va;
        }
      }
      __syncthreads();
    }
  }
  // heads of runs inside this thread's contiguous chunk, then an exclusive scan of the chunk counts over the block
  const int chunk = (n_pad + 1023) / 1024;
  const int lo = tid * chunk, hi = min(lo + chunk, n);
  uint32_t mine = 0;
  for (int i = lo; i < hi; ++i) {
    const uint64_t k = sk[i];
    mine += (k != NBW_INVALID_KEY && (i == 0 || sk[i - 1] != k)) ? 1u : 0u;
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = warp_sums[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += t;
    }
    warp_sums[tid] = w;       // inclusive over warps
  }
  __syncthreads();
  uint32_t rank = incl - mine + ((tid >> 5) ? warp_sums[(tid >> 5) - 1] : 0u);    // heads before this chunk
  if (tid == 1023) counters->n_unique = warp_sums[31];
  bool bad = false;
  for (int i = lo; i < hi; ++i) {
    const uint64_t k = sk[i];
    const uint32_t src = sv[i];
    if (k == NBW_INVALID_KEY) {
      ids[src] = NBW_INVALID_ID;
      if (first_flag) first_flag[src] = 0;
      continue;
    }
    const bool head = i == 0 || sk[i - 1] != k;
    rank += head ? 1u : 0u;
    const uint32_t id = rank - 1u;
    ids[src] = id;
    if (first_flag) first_flag[src] = head ? 1 : 0;     // (key, index) order: the head of a run is the first occurrence
    if (head) {
      if ((int64_t)id < cap) {
        uniq_keys[id] = k;
        uniq_chks[id] = chks[src];
      }
    } else {
      const uint32_t prev = sv[i - 1];
      bool ok = chks[src] == chks[prev];
      if (ok && cells != nullptr && ids_prev != nullptr)
        ok = (n_max == 8) ? same_credential<8>(cells, ids_prev, prev, src) : same_credential<16>(cells, ids_prev, prev, src);
      bad = bad || !ok;
    }
  }
  if (bad) atomicOr(&collided, 1u);
  __syncthreads();
  if (tid == 0) counters->collision = collided;
}

}  // namespace

// ---------------------------------------------------------------- hash pre-dedup for large key counts
// A WL level of a large batch has millions of keys but few distinct ones (the 200k-cell stress: 16M keys, 5 ..
// 23 493 labels per level).  Sorting all of them moves every key eight times; instead every key is looked up
// in an open-addressing table that lives in L2 (2 MB), only the DISTINCT keys are sorted (for the same
// deterministic, key-ordered ids), and one more sweep gives every node its id, verifies its credential against
// the first node of its class and flags first occurrences.  ~30 B of traffic per key instead of ~290.
namespace {
constexpr int64_t kHashDistinctCap = 65536;            // more distinct keys than this: radix path
constexpr int64_t kHashSlots = 4 * kHashDistinctCap;   // load factor <= 1/4
constexpr uint32_t kOverflowBit = 2u;

// A CTA keeps the slots of the keys it has met in a small insert-only shared-memory cache (with the smallest node
// index it saw for each), so that the few hot labels of a level are looked up in the global table once per CTA
// and not once per node; keys that lose the race for a cache entry take the global path.
constexpr int kCtaCache = 2048;
__global__ void __launch_bounds__(256)
hash_insert_kernel(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ tab,
                   uint32_t* __restrict__ first, uint32_t* __restrict__ node_slot, uint64_t* __restrict__ dkeys,
                   uint32_t* __restrict__ dslot, uint32_t* __restrict__ dcount, DictCounters* __restrict__ counters) {
  __shared__ unsigned long long c_key[kCtaCache];
  __shared__ uint32_t c_slot[kCtaCache];
  __shared__ uint32_t c_min[kCtaCache];
  constexpr uint32_t PENDING = 0xFFFFFFFFu;
  for (int e = threadIdx.x; e < kCtaCache; e += blockDim.x) {
    c_key[e] = (unsigned long long)NBW_INVALID_KEY;
    c_slot[e] = PENDING;
    c_min[e] = 0xFFFFFFFFu;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t mask = (uint32_t)(kHashSlots - 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t key = keys[i];
    if (key == NBW_INVALID_KEY) {
      node_slot[i] = 0xFFFFFFFFu;
      continue;
    }
    const uint64_t hk = nbw_mix_b(key);
    const int e = (int)(hk >> 44) & (kCtaCache - 1);
    unsigned long long ck = *(volatile unsigned long long*)(c_key + e);
    bool owner = false;
    if (ck == (unsigned long long)NBW_INVALID_KEY) {
      ck = atomicCAS(c_key + e, (unsigned long long)NBW_INVALID_KEY, (unsigned long long)key);
      owner = ck == (unsigned long long)NBW_INVALID_KEY;
      if (owner) ck = key;
    }
    if (!owner && ck == (unsigned long long)key) {
      const uint32_t s = *(volatile uint32_t*)(c_slot + e);
      if (s != PENDING) {
        node_slot[i] = s;
        if ((uint32_t)i < *(volatile uint32_t*)(c_min + e)) atomicMin(c_min + e, (uint32_t)i);
        continue;
      }
    }
    uint32_t slot = (uint32_t)(hk >> 20) & mask;
    bool placed = false;
    for (int probe = 0; probe < 512; ++probe) {
      unsigned long long cur = *(volatile unsigned long long*)(tab + slot);
      if (cur == (unsigned long long)NBW_INVALID_KEY) {
        cur = atomicCAS(tab + slot, (unsigned long long)NBW_INVALID_KEY, (unsigned long long)key);
        if (cur == (unsigned long long)NBW_INVALID_KEY) {
          const uint32_t pos = atomicAdd(dcount, 1u);
          if (pos < (uint32_t)kHashDistinctCap) {
            dkeys[pos] = key;
            dslot[pos] = slot;
          } else {
            atomicOr(&counters->collision, kOverflowBit);
          }
          placed = true;
          break;
        }
      }
      if (cur == (unsigned long long)key) { placed = true; break; }
      slot = (slot + 1u) & mask;
    }
    if (!placed) {
      atomicOr(&counters->collision, kOverflowBit);
      node_slot[i] = 0xFFFFFFFFu;
      continue;
    }
    node_slot[i] = slot;
    if (owner) {
      atomicMin(c_min + e, (uint32_t)i);
      __threadfence_block();
      *(volatile uint32_t*)(c_slot + e) = slot;
    } else if ((uint32_t)i < *(volatile uint32_t*)(first + slot)) {
      atomicMin(first + slot, (uint32_t)i);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kCtaCache; e += blockDim.x) {
    const uint32_t s = c_slot[e], m = c_min[e];
    if (s != PENDING && m != 0xFFFFFFFFu && m < *(volatile uint32_t*)(first + s)) atomicMin(first + s, m);
  }
}

// sorted distinct keys -> ids: slot_id[slot] = rank; the dictionary in key order
__global__ void rank_slots_kernel(const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                  const uint32_t* __restrict__ dslot, const uint32_t* __restrict__ first,
                                  const uint64_t* __restrict__ chks, uint32_t* __restrict__ slot_id,
                                  uint64_t* __restrict__ uniq_keys, uint64_t* __restrict__ uniq_chks, int64_t cap,
                                  int64_t n_distinct) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_distinct) return;
  const uint64_t key = skeys[r];
  if (key == NBW_INVALID_KEY) return;
  const uint32_t slot = dslot[svals[r]];
  slot_id[slot] = (uint32_t)r;
  if (r < cap) {
    uniq_keys[r] = key;
    uniq_chks[r] = chks[first[slot]];
  }
}

__global__ void assign_from_slots_kernel(const uint32_t* __restrict__ node_slot, const uint32_t* __restrict__ slot_id,
                                         const uint32_t* __restrict__ first, const uint64_t* __restrict__ chks, int64_t n,
                                         const uint8_t* __restrict__ cells, int n_max, const uint32_t* __restrict__ ids_prev,
                                         uint32_t* __restrict__ ids, const uint32_t* __restrict__ dcount,
                                         DictCounters* __restrict__ counters, uint8_t* __restrict__ first_flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (i == 0) counters->n_unique = *dcount;
    const uint32_t slot = node_slot[i];
    if (slot == 0xFFFFFFFFu) {
      ids[i] = NBW_INVALID_ID;
      if (first_flag) first_flag[i] = 0;
      continue;
    }
    ids[i] = slot_id[slot];
    const uint32_t rep = first[slot];
    if (first_flag) first_flag[i] = rep == (uint32_t)i ? 1 : 0;
    if (rep != (uint32_t)i) {
      bool ok = chks[i] == chks[rep];
      if (ok && cells != nullptr && ids_prev != nullptr)
        ok = (n_max == 8) ? same_credential<8>(cells, ids_prev, rep, (uint32_t)i) : same_credential<16>(cells, ids_prev, rep, (uint32_t)i);
      if (!ok) atomicOr(&counters->collision, 1u);
    }
  }
}
}  // namespace

// Enqueue one dictionary build (no sync, no read-back): the counters land in *counters (device).  `work` is a
// scratch region of nbw_dict_scratch_bytes(n_keys) bytes.
size_t nbw_dict_scratch_bytes(int64_t n_keys) {
  if (n_keys < kHashDistinctCap) n_keys = kHashDistinctCap;      // the distinct keys of the hash path are sorted here
  const int n_tiles = (int)((n_keys + kTile - 1) / kTile);
  const size_t n_counts = (size_t)256 * n_tiles;
  const size_t scan_tmp = scan_scratch_elems((int64_t)(n_counts > (size_t)n_keys ? n_counts : (size_t)n_keys));
  size_t need = 0;
  need += Carver::pad(sizeof(uint64_t) * n_keys) * 2;   // key ping-pong
  need += Carver::pad(sizeof(uint32_t) * n_keys) * 2;   // value ping-pong
  need += Carver::pad(sizeof(uint32_t) * n_counts);
  need += Carver::pad(sizeof(uint32_t) * n_keys) * 2;   // flags, rank
  need += Carver::pad(sizeof(uint32_t) * scan_tmp);
  // hash pre-dedup: table, first index and id per slot, slot per node, distinct keys / slots (+ their sort buffers
  // are the radix buffers above, sized for at least kHashDistinctCap keys)
  need += Carver::pad(sizeof(uint64_t) * kHashSlots) + 2 * Carver::pad(sizeof(uint32_t) * kHashSlots);
  need += Carver::pad(sizeof(uint32_t) * n_keys);
  need += Carver::pad(sizeof(uint64_t) * kHashDistinctCap) + Carver::pad(sizeof(uint32_t) * kHashDistinctCap) + 256;
  return need + 4096;
}

int nbw_dict_build_enqueue(nbw_ctx* ctx, const uint64_t* keys, const uint64_t* chks, int64_t n_keys, int key_bits,
                           const void* cells, int n_max, const uint32_t* ids_prev, uint32_t* ids, uint64_t* uniq_keys,
                           uint64_t* uniq_chks, int64_t uniq_cap, NbwDictCounters* counters, void* work, cudaStream_t st,
                           uint8_t* first_flag, bool allow_hash) {
  if (allow_hash && n_keys > kSmallSortMax && !getenv("NBW_DICT_RADIX") && !getenv("NBW_DICT_NOHASH")) {
    // ---- hash pre-dedup (see above)
    const int64_t size_n = n_keys > kHashDistinctCap ? n_keys : kHashDistinctCap;
    const size_t n_counts_full = (size_t)256 * (int)((size_n + kTile - 1) / kTile);
    const size_t scan_tmp_full = scan_scratch_elems((int64_t)(n_counts_full > (size_t)size_n ? n_counts_full : (size_t)size_n));
    Carver cv(work);
    uint64_t* kbuf[2] = {cv.take<uint64_t>(size_n), cv.take<uint64_t>(size_n)};
    uint32_t* vbuf[2] = {cv.take<uint32_t>(size_n), cv.take<uint32_t>(size_n)};
    uint32_t* counts = cv.take<uint32_t>(n_counts_full);
    cv.take<uint32_t>(size_n);   // flags (radix path)
    cv.take<uint32_t>(size_n);   // rank  (radix path)
    uint32_t* stmp = cv.take<uint32_t>(scan_tmp_full);
    unsigned long long* tab = cv.take<unsigned long long>(kHashSlots);
    uint32_t* first = cv.take<uint32_t>(kHashSlots);
    uint32_t* slot_id = cv.take<uint32_t>(kHashSlots);
    uint32_t* node_slot = cv.take<uint32_t>(n_keys);
    uint64_t* dkeys = cv.take<uint64_t>(kHashDistinctCap);
    uint32_t* dslot = cv.take<uint32_t>(kHashDistinctCap);
    uint32_t* dcount = cv.take<uint32_t>(1);
    NBW_CUDA(ctx, cudaMemsetAsync(counters, 0, sizeof(DictCounters), st));
    NBW_CUDA(ctx, cudaMemsetAsync(tab, 0xFF, sizeof(uint64_t) * kHashSlots, st));
    NBW_CUDA(ctx, cudaMemsetAsync(first, 0xFF, sizeof(uint32_t) * kHashSlots, st));
    NBW_CUDA(ctx, cudaMemsetAsync(dcount, 0, sizeof(uint32_t), st));
    const int grid = nbw_grid_for(ctx, n_keys, 256);
    hash_insert_kernel<<<grid, 256, 0, st>>>(keys, n_keys, tab, first, node_slot, dkeys, dslot, dcount, counters);
    NBW_CHECK_LAUNCH(ctx);
    // the number of distinct keys decides how they are sorted (one sync; this path only runs for large inputs)
    uint32_t hd = 0;
    DictCounters hc;
    NBW_CUDA(ctx, cudaMemcpyAsync(&hd, dcount, sizeof(hd), cudaMemcpyDeviceToHost, st));
    NBW_CUDA(ctx, cudaMemcpyAsync(&hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    NBW_CUDA(ctx, cudaStreamSynchronize(st));
    if (!(hc.collision & kOverflowBit) && hd <= (uint32_t)kHashDistinctCap) {
      const int64_t D = hd;
      const uint64_t* kin = dkeys;
      const uint32_t* vin = nullptr;
      if (D <= kSmallSortMax) {
        int n_pad = 2;
        while (n_pad < D) n_pad <<= 1;
        if (!(ctx->attr_set & NBW_ATTR_SMALL_SORT)) {
          NBW_CUDA(ctx, cudaFuncSetAttribute(small_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmallSortMax * (int)(sizeof(uint64_t) + sizeof(uint32_t))));
          ctx->attr_set |= NBW_ATTR_SMALL_SORT;
        }
        small_sort_kernel<<<1, 1024, (size_t)n_pad * (sizeof(uint64_t) + sizeof(uint32_t)), st>>>(dkeys, (int)D, n_pad, kbuf[0],
                                                                                                 vbuf[0]);
        NBW_CHECK_LAUNCH(ctx);
        kin = kbuf[0];
        vin = vbuf[0];
      } else {
        const int n_tiles_d = (int)((D + kTile - 1) / kTile);
        int cur = 0;
        for (int shift = 0; shift < key_bits; shift += 8) {
          radix_count_kernel<<<n_tiles_d, kSortThreads, 0, st>>>(kin, D, shift, n_tiles_d, counts);
          NBW_CHECK_LAUNCH(ctx);
          int rc2 = device_exclusive_scan(ctx, counts, counts, (int64_t)256 * n_tiles_d, stmp, st);
          if (rc2) return rc2;
          radix_scatter_kernel<<<n_tiles_d, kSortThreads, 0, st>>>(kin, vin, kbuf[cur], vbuf[cur], D, shift, n_tiles_d, counts);
          NBW_CHECK_LAUNCH(ctx);
          kin = kbuf[cur];
          vin = vbuf[cur];
          cur ^= 1;
        }
      }
      if (D > 0) {
        rank_slots_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(kin, vin, dslot, first, chks, slot_id, uniq_keys, uniq_chks,
                                                                      uniq_cap, D);
        NBW_CHECK_LAUNCH(ctx);
      }
      assign_from_slots_kernel<<<grid, 256, 0, st>>>(node_slot, slot_id, first, chks, n_keys, static_cast<const uint8_t*>(cells),
                                                    n_max, ids_prev, ids, dcount, counters, first_flag);
      NBW_CHECK_LAUNCH(ctx);
      return NBW_OK;
    }
    // more distinct keys than the table holds: sort them all (below)
  }
  const int n_tiles = (int)((n_keys + kTile - 1) / kTile);
  const size_t n_counts = (size_t)256 * n_tiles;
  const size_t scan_tmp = scan_scratch_elems((int64_t)(n_counts > (size_t)n_keys ? n_counts : (size_t)n_keys));
  Carver cv(work);
  uint64_t* kbuf[2] = {cv.take<uint64_t>(n_keys), cv.take<uint64_t>(n_keys)};
  uint32_t* vbuf[2] = {cv.take<uint32_t>(n_keys), cv.take<uint32_t>(n_keys)};
  uint32_t* counts = cv.take<uint32_t>(n_counts);
  uint32_t* flags = cv.take<uint32_t>(n_keys);
  uint32_t* rank = cv.take<uint32_t>(n_keys);
  uint32_t* stmp = cv.take<uint32_t>(scan_tmp);
  int rc;
  const uint64_t* kin = keys;
  const uint32_t* vin = nullptr;
  int cur = 0;
  const bool small = n_keys <= kSmallSortMax && !getenv("NBW_DICT_RADIX");
  if (small) {
    int n_pad = 2;
    while (n_pad < n_keys) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * (sizeof(uint64_t) + sizeof(uint32_t));
    if (!(ctx->attr_set & NBW_ATTR_SMALL_DICT)) {
      NBW_CUDA(ctx, cudaFuncSetAttribute(small_dict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmallSortMax * (int)(sizeof(uint64_t) + sizeof(uint32_t))));
      ctx->attr_set |= NBW_ATTR_SMALL_DICT;
    }
    small_dict_kernel<<<1, 1024, smem, st>>>(keys, chks, (int)n_keys, n_pad, static_cast<const uint8_t*>(cells), n_max, ids_prev,
                                             ids, uniq_keys, uniq_chks, uniq_cap, counters, first_flag);
    NBW_CHECK_LAUNCH(ctx);
    return NBW_OK;
  }
  for (int shift = 0; !small && shift < key_bits; shift += 8) {
    radix_count_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, n_keys, shift, n_tiles, counts);
    NBW_CHECK_LAUNCH(ctx);
    rc = device_exclusive_scan(ctx, counts, counts, (int64_t)n_counts, stmp, st);
    if (rc) return rc;
    radix_scatter_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, vin, kbuf[cur], vbuf[cur], n_keys, shift,
                                                          n_tiles, counts);
    NBW_CHECK_LAUNCH(ctx);
    kin = kbuf[cur];
    vin = vbuf[cur];
    cur ^= 1;
  }
  const int grid = nbw_grid_for(ctx, n_keys, 256);
  mark_heads_kernel<<<grid, 256, 0, st>>>(kin, n_keys, flags);
  NBW_CHECK_LAUNCH(ctx);
  rc = device_exclusive_scan(ctx, flags, rank, n_keys, stmp, st);
  if (rc) return rc;
  NBW_CUDA(ctx, cudaMemsetAsync(counters, 0, sizeof(DictCounters), st));
  assign_ids_kernel<<<grid, 256, 0, st>>>(kin, vin, chks, flags, rank, n_keys,
                                         static_cast<const uint8_t*>(cells), n_max, ids_prev, ids,
                                         uniq_keys, uniq_chks, uniq_cap, counters, first_flag);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_dict_build(nbw_ctx* ctx, const uint64_t* keys, const uint64_t* chks, int64_t n_keys,
                              int key_bits, const void* cells, int n_max, const uint32_t* ids_prev,
                              uint32_t* ids, uint64_t* uniq_keys, uint64_t* uniq_chks,
                              int64_t uniq_cap, int64_t* n_unique_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_dict_build");
  NBW_REQUIRE(ctx, n_keys >= 0 && n_keys < (int64_t)1 << 31, "n_keys out of range");
  NBW_REQUIRE(ctx, key_bits >= 8 && key_bits <= 64 && key_bits % 8 == 0, "key_bits must be a multiple of 8 in [8,64]");
  NBW_REQUIRE(ctx, n_unique_h != nullptr, "n_unique_h is NULL");
  NBW_REQUIRE(ctx, cells == nullptr || n_max == 8 || n_max == 16, "n_max must be 8 or 16");
  *n_unique_h = 0;
  if (n_keys == 0) return NBW_OK;
  NBW_REQUIRE(ctx, keys && chks && ids && uniq_keys && uniq_chks, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = nbw_scratch_ensure(ctx, nbw_dict_scratch_bytes(n_keys) + 1024, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  DictCounters* counters = cv.take<DictCounters>(1);
  void* work = static_cast<char*>(ctx->scratch) + 1024;
  rc = nbw_dict_build_enqueue(ctx, keys, chks, n_keys, key_bits, cells, n_max, ids_prev, ids, uniq_keys, uniq_chks, uniq_cap,
                              counters, work, st);
  if (rc) return rc;
  DictCounters host;
  NBW_CUDA(ctx, cudaMemcpyAsync(&host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  *n_unique_h = host.n_unique;
  if (host.collision)
    NBW_FAIL(ctx, NBW_ERR_HASH_COLLISION, "two distinct WL credentials share a key; retry with another seed");
  if ((int64_t)host.n_unique > uniq_cap)
    NBW_FAIL(ctx, NBW_ERR_CAPACITY, "dictionary has %u entries, capacity %lld", host.n_unique, (long long)uniq_cap);
  return NBW_OK;
}

// All WL levels of a batch in one call: relabel + dictionary build for levels 0..h back to back on the stream,
// ONE read-back of the dictionary sizes at the end (the per-level entry points sync once per level, which is
// most of a small training set's fit time).  ids: [h+1][n_keys] u32; uniq_keys / uniq_chks: [h+1][n_keys] u64
// (level l's dictionary at offset l * n_keys); dims_h: [h+1].
extern "C" int nbw_wl_fit(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, int h, uint64_t seed,
                          uint32_t* ids, uint64_t* uniq_keys, uint64_t* uniq_chks, int64_t* dims_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_wl_fit");
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "n_max must be 8 or 16 (got %d)", n_max);
  NBW_REQUIRE(ctx, n_cells > 0 && h >= 0 && h < NBW_MAX_LEVELS, "bad n_cells / h");
  NBW_REQUIRE(ctx, cells && ids && uniq_keys && uniq_chks && dims_h, "null pointer");
  const int64_t n_keys = n_cells * n_max;
  NBW_REQUIRE(ctx, n_keys < (int64_t)1 << 31, "n_keys out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t head = 1024 + 2 * Carver::pad(sizeof(uint64_t) * n_keys);
  int rc = nbw_scratch_ensure(ctx, head + nbw_dict_scratch_bytes(n_keys), st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  DictCounters* counters = cv.take<DictCounters>(NBW_MAX_LEVELS);
  cv.off = 1024;
  uint64_t* keys = cv.take<uint64_t>(n_keys);
  uint64_t* chks = cv.take<uint64_t>(n_keys);
  void* work = static_cast<char*>(ctx->scratch) + head;
  for (int l = 0; l <= h; ++l) {
    const uint32_t* prev = l > 0 ? ids + (size_t)(l - 1) * n_keys : nullptr;
    rc = nbw_wl_signatures_enqueue(ctx, cells, n_cells, n_max, prev, l > 0 ? uniq_keys + (size_t)(l - 1) * n_keys : nullptr,
                                   l, seed, keys, chks, st);
    if (rc) return rc;
    rc = nbw_dict_build_enqueue(ctx, keys, chks, n_keys, l == 0 ? 8 : 64, cells, n_max, prev, ids + (size_t)l * n_keys,
                                uniq_keys + (size_t)l * n_keys, uniq_chks + (size_t)l * n_keys, n_keys, counters + l, work, st);
    if (rc) return rc;
  }
  DictCounters host[NBW_MAX_LEVELS];
  NBW_CUDA(ctx, cudaMemcpyAsync(host, counters, sizeof(DictCounters) * (h + 1), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  for (int l = 0; l <= h; ++l) {
    dims_h[l] = host[l].n_unique;
    if (host[l].collision)
      NBW_FAIL(ctx, NBW_ERR_HASH_COLLISION, "two distinct WL credentials share a key at level %d; retry with another seed", l);
  }
  return NBW_OK;
}

// ids_out[i] = map[ids_in[i]] (NBW_INVALID_ID stays): renumbering of a level's labels, e.g. from the sorted order
// of a fresh dictionary build to the append-only numbering of a growing training dictionary.
namespace {
__global__ void remap_ids_kernel(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ map,
                                 uint32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = in[i];
    out[i] = v == NBW_INVALID_ID ? v : map[v];
  }
}
}  // namespace

extern "C" int nbw_remap_ids(nbw_ctx* ctx, const uint32_t* ids_in, int64_t n, const uint32_t* map, uint32_t* ids_out,
                             void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n >= 0, "n negative");
  if (n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, ids_in && map && ids_out, "null pointer");
  remap_ids_kernel<<<nbw_grid_for(ctx, n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids_in, n, map, ids_out);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}
