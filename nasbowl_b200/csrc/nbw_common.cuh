// Shared device/host helpers for libnbw (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "nbw.h"

struct nbw_ctx {
  int device;
  int sm_count;
  char err[512];
  void* scratch;
  size_t scratch_bytes;
  uint64_t launches;
  void* encode_tiled;   // cuTensorMapEncodeTiled, resolved lazily (gram_i8.cu)
  // host-buffer pipeline of nbw_score_pool_host (score_pool.cu), created on first use
  cudaStream_t s_in, s_out;
  cudaEvent_t ev_ready[2], ev_free[2], ev_done[2];
  cudaEvent_t ev_out;   // recorded on s_out after the last D2H copy of host scores
  int out_pending;      // host scores of the last nbw_score_pool_host / _mapped call not yet awaited
  void* stage[2];
  size_t stage_bytes;
  int pipe_ready;
  void* pinned_small;   // 16 KB of pinned host memory for small synchronous read-backs (top-k lists)
  // per-device launch state (cudaFuncSetAttribute applies to the CURRENT device, and one process may
  // hold a context per GPU: these flags must not be process-wide statics)
  uint32_t attr_set;    // bit i = opt-in shared-memory size already set for kernel i on this device
  int gram_max_pairs;   // co-resident CTA pairs of the Gram kernel on this device (0 = not queried yet)
  cudaEvent_t prof_ev[2];   // optional caller-owned events recorded right before / after the scoring kernel
  int packed_occ[64];   // resident CTAs per SM of each score_packed_kernel variant (0 = not queried yet)
  void* piece_cands;    // per-piece top-n lists of the host-buffer entry points (score_pool.cu)
  void* topk_blocks;    // per-CTA top-n lists of the fused scoring kernel (score_packed.cu), two alternating slots
  size_t topk_blocks_bytes;
  void* bin_order;      // cell order of the size-binned scoring pass (score_packed.cu) + its counters
  size_t bin_order_bytes;
  int topk_slot;
};
enum { NBW_ATTR_TRSV = 1u << 0, NBW_ATTR_SMALL_SORT = 1u << 1, NBW_ATTR_GRAM_SINGLE = 1u << 2, NBW_ATTR_FACTOR_NP = 1u << 3, NBW_ATTR_SMALL_DICT = 1u << 4,
       NBW_ATTR_GRAM_PAIR = 1u << 8 /* << output mode: bits 8..11 */ };

// NVTX range around a C-ABI entry point (visible in Nsight Systems / ncu --nvtx); header-only NVTX3,
// a no-op costing one predicted branch when no tool is attached.
struct NbwRange {
  explicit NbwRange(const char* name) { nvtxRangePushA(name); }
  ~NbwRange() { nvtxRangePop(); }
};

// Host scores written by the side stream are complete once this returns (called by every syncing
// top-k entry point and by nbw_wait_host_scores).
static inline cudaError_t nbw_await_host_scores(nbw_ctx* ctx) {
  if (!ctx->out_pending) return cudaSuccess;
  ctx->out_pending = 0;
  return cudaEventSynchronize(ctx->ev_out);
}

#define NBW_FAIL(ctx, code, ...)                                  \
  do {                                                            \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return (code);                                                \
  } while (0)

#define NBW_CUDA(ctx, expr)                                                          \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess)                                                          \
      NBW_FAIL(ctx, NBW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                  \
  } while (0)

#define NBW_CHECK_LAUNCH(ctx)          \
  do {                                 \
    (ctx)->launches++;                 \
    NBW_CUDA(ctx, cudaGetLastError()); \
  } while (0)

#define NBW_REQUIRE(ctx, cond, ...) \
  do {                              \
    if (!(cond)) NBW_FAIL(ctx, NBW_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// One entry of a top-n list (topk.cu; the 16-byte wire format of the per-rank top-n all-gather).
struct NbwCand {
  float v;
  int32_t pad;
  int64_t idx;   // -1 = empty
};
// Tournament merge of `count` device candidates into the k best, left on the device in *result (scratch
// memory, valid until the next scratch user).  topk.cu.
// `direct` (optional): when the input fits one chunk the list is written there and no scratch is touched.
int nbw_topk_merge_enqueue(nbw_ctx* ctx, const NbwCand* cands, int64_t count, int k, int distinct, NbwCand** result,
                           cudaStream_t st, NbwCand* direct = nullptr);

// Packed scoring path (score_packed.cu): applicability test and launch (k = 0: scores only).
bool nbw_packed_applicable(const nbw_model* M);
int nbw_score_packed(nbw_ctx* ctx, const nbw_model* M, const void* cells, int64_t n_cells, float* scores, double* mu64,
                     double* var64, double* score64, int k, int distinct, int64_t index_base, void* out_cands,
                     cudaStream_t st);

// WL relabelling of one level, enqueue only (wl_relabel.cu; used by nbw_wl_fit in dict_build.cu)
int nbw_wl_signatures_enqueue(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, const uint32_t* ids_prev,
                              const uint64_t* uniq_prev, int level, uint64_t seed, uint64_t* keys, uint64_t* chks,
                              cudaStream_t st);

// Sort / unique / dense ids of n_keys 64-bit keys, enqueue only (dict_build.cu): the counters land in *counters
// (device); `work` is scratch of nbw_dict_scratch_bytes(n_keys) bytes; first_flag[i] (optional) = 1 iff key i is
// the first occurrence of its value.
struct NbwDictCounters {
  uint32_t n_unique;
  uint32_t collision;
};
size_t nbw_dict_scratch_bytes(int64_t n_keys);
int nbw_dict_build_enqueue(nbw_ctx* ctx, const uint64_t* keys, const uint64_t* chks, int64_t n_keys, int key_bits,
                           const void* cells, int n_max, const uint32_t* ids_prev, uint32_t* ids, uint64_t* uniq_keys,
                           uint64_t* uniq_chks, int64_t uniq_cap, NbwDictCounters* counters, void* work, cudaStream_t st,
                           uint8_t* first_flag = nullptr, bool allow_hash = true);
// allow_hash: inputs above 8192 keys go through a hash pre-dedup (one stream sync inside to size the sort of the
// distinct keys; falls back to the full sort when there are more than 65536 of them)

// One-launch GP factorisation (gp_factor.cu)
bool nbw_factor_fused_ok(int64_t n);
int nbw_factor_fused(nbw_ctx* ctx, const double* K, int64_t ldk, int64_t n, double lik, const double* y, double* L,
                     int64_t ldl, double* Linv, int64_t ldi, double* W, double* alpha, double* z, double* stats,
                     int* info, cudaStream_t st);

// Scratch arena: grows on demand (syncs the stream before releasing the old block).
int nbw_scratch_ensure(nbw_ctx* ctx, size_t bytes, cudaStream_t stream);

struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

static inline int nbw_grid_for(const nbw_ctx* ctx, int64_t work_items, int per_block, int waves = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx->sm_count * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------------------
// Hashing of WL credentials.
// A credential is (own id, multiset of successor ids).  Every node hashes its id once into a
// 64-bit value ha and a 32-bit value hb (independent mixers, salted per level); the signature is
//   key   = fin64( own64(ha_v) + sum_{s in succ(v)} ha_s )      (sums mod 2^64 / 2^32:
//   chk32 = fin32( own32(hb_v) + sum_{s in succ(v)} hb_s )       order-independent, so no sort)
// where own64/own32 are cheap bijections (rotate + odd multiply) that keep "v with successor s"
// apart from "s with successor v".  Ids are 64-bit: a dense dictionary id (< 2^32) for labels
// known to the dictionary, UNSEEN_BIT | key for labels outside it.  96 signature bits decide
// pool-side membership; the training-side partition is verified exactly (dict_build.cu).
// What is hashed per node is the KEY of its previous label (level 0: the op code), never a batch-local id.
// ---------------------------------------------------------------------------------------
#define NBW_UNSEEN_BIT 0x8000000000000000ull
#define NBW_NOCLASS_BIT 0x4000000000000000ull   // | lane: a class id no other lane can have

__host__ __device__ __forceinline__ uint64_t nbw_mix_a(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ uint64_t nbw_mix_b(uint64_t x) {   // murmur3 fmix64
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint32_t nbw_mix32(uint32_t x) {   // murmur3 fmix32
  x ^= x >> 16; x *= 0x85ebca6bu;
  x ^= x >> 13; x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

struct LevelSalt {
  uint64_t a_nbr, a_own;
  uint32_t b_nbr, b_own;
};
__host__ __device__ __forceinline__ LevelSalt nbw_level_salt(uint64_t seed, int level) {
  uint64_t s = nbw_mix_a(seed + 0x9e3779b97f4a7c15ull * (uint64_t)(level + 1));
  LevelSalt t;
  t.a_nbr = nbw_mix_a(s ^ 0x01);
  t.a_own = nbw_mix_a(s ^ 0x02);
  const uint64_t b = nbw_mix_b(s ^ 0x03);
  t.b_nbr = (uint32_t)b;
  t.b_own = (uint32_t)(b >> 32);
  return t;
}
// per-node hashes of an id
__host__ __device__ __forceinline__ uint64_t nbw_hash_a(uint64_t id64, const LevelSalt& s) {
  return nbw_mix_a(id64 + s.a_nbr);
}
__host__ __device__ __forceinline__ uint32_t nbw_hash_b(uint64_t id64, const LevelSalt& s) {
  return nbw_mix32(((uint32_t)id64 * 0x9e3779b1u) ^ ((uint32_t)(id64 >> 32) * 0x85ebca77u) ^ s.b_nbr);
}
// signature from the node's own hashes and the sums over its successors
// (No finalising mix: the per-node hashes are already outputs of a full mixer, sums of them are uniform,
// and the pool-side kernel fetches them from a table per dictionary id — a final mix would be the only
// 64-bit mixing left on its critical path.)
__host__ __device__ __forceinline__ uint64_t nbw_key_from(uint64_t ha_own, uint64_t sum_a, const LevelSalt& s) {
  const uint64_t own = ((ha_own << 27) | (ha_own >> 37)) * 0x9e3779b97f4a7c15ull + s.a_own;
  const uint64_t k = own + sum_a;
  return k == NBW_INVALID_KEY ? k - 1 : k;
}
__host__ __device__ __forceinline__ uint32_t nbw_chk_from(uint32_t hb_own, uint32_t sum_b, const LevelSalt& s) {
  const uint32_t own = ((hb_own << 13) | (hb_own >> 19)) * 0xcc9e2d51u + s.b_own;
  return own + sum_b;
}

// ---------------------------------------------------------------------------------------
// Cell records: lane `node` of a sub-warp reads its own label and successor mask.
// The 32 lanes of a warp touch 32/NMAX consecutive records = one contiguous span.
// ---------------------------------------------------------------------------------------
template <int NMAX>
struct CellTraits;
template <>
struct CellTraits<8> {
  static constexpr int kStride = 16;
  static __device__ __forceinline__ uint32_t label(const uint8_t* cells, int64_t g, int node) {
    return __ldg(cells + g * 16 + node);
  }
  static __device__ __forceinline__ uint32_t succ(const uint8_t* cells, int64_t g, int node) {
    return __ldg(cells + g * 16 + 8 + node);
  }
};
template <>
struct CellTraits<16> {
  static constexpr int kStride = 48;
  static __device__ __forceinline__ uint32_t label(const uint8_t* cells, int64_t g, int node) {
    return __ldg(cells + g * 48 + node);
  }
  static __device__ __forceinline__ uint32_t succ(const uint8_t* cells, int64_t g, int node) {
    return __ldg(reinterpret_cast<const uint16_t*>(cells + g * 48 + 16) + node);
  }
};

// OR-reduce / sum-reduce over the NMAX lanes of a sub-warp (all 32 lanes must call).
template <int NMAX>
__device__ __forceinline__ uint32_t subwarp_or(uint32_t v) {
#pragma unroll
  for (int o = NMAX / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NMAX>
__device__ __forceinline__ int subwarp_sum(int v) {
#pragma unroll
  for (int o = NMAX / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NMAX>
__device__ __forceinline__ double subwarp_sum(double v) {
#pragma unroll
  for (int o = NMAX / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

// Sums of the successors' hashes.  Each round every lane fetches ITS next successor (per-lane
// source lane).  The first rounds are unrolled without any vote or branch (most nodes of
// the three search spaces have out-degree <= 2), so their shuffles are independent and overlap;
// higher degrees fall through to a voted loop.
template <int NMAX>
__device__ __forceinline__ void successor_sums(uint32_t succ, int sub_base, int node, uint64_t ha, uint32_t hb,
                                               uint64_t& sum_a, uint32_t& sum_b) {
  constexpr int kUnrolledSuccessorRounds = NMAX == 8 ? 2 : 4;   // measured: NB cells 2, DARTS cells 4
  sum_a = 0;
  sum_b = 0;
  uint32_t m = succ;
#pragma unroll
  for (int t = 0; t < kUnrolledSuccessorRounds; ++t) {
    const bool has = m != 0u;
    const int j = has ? (__ffs(m) - 1) : node;
    const uint64_t oa = shfl64(ha, sub_base + j);
    const uint32_t ob = __shfl_sync(0xffffffffu, hb, sub_base + j);
    sum_a += has ? oa : 0ull;
    sum_b += has ? ob : 0u;
    m &= m - 1u;      // 0 stays 0
  }
#pragma unroll 1
  for (int t = kUnrolledSuccessorRounds; t < NMAX; ++t) {
    if (!__any_sync(0xffffffffu, m != 0u)) break;
    const bool has = m != 0u;
    const int j = has ? (__ffs(m) - 1) : node;
    const uint64_t oa = shfl64(ha, sub_base + j);
    const uint32_t ob = __shfl_sync(0xffffffffu, hb, sub_base + j);
    sum_a += has ? oa : 0ull;
    sum_b += has ? ob : 0u;
    m &= m - 1u;
  }
}
