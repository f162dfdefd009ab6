// Top-n selection over the acquisition scores of a pool shard.
// Replaces propose_location's  np.unique(eis, return_index=True) + np.argpartition
// (bayesopt/acquisitions.py:101-105): the k largest DISTINCT values, each reported with the
// first index at which it occurs.  Selection is a tournament: every CTA extracts the top k of
// its 4096-score chunk by k rounds of (value desc, index asc) arg-max, the per-CTA lists are
// merged by the same kernel until one list is left.  "Distinct" merging is associative: a
// partial list always keeps the lowest index of each value it contains.
#include "nbw_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kTopThreads = 256;
constexpr int kTopItems = 16;
constexpr int kTopChunk = kTopThreads * kTopItems;

typedef NbwCand Cand;

__device__ __forceinline__ bool better(float v, int64_t i, float bv, int64_t bi) {
  // (v, i) beats (bv, bi): larger value, or equal value and lower index; empty loses
  if (i < 0) return false;
  if (bi < 0) return true;
  return v > bv || (v == bv && i < bi);
}

// (kWideThreads x kWideItems: one CTA that merges the per-CTA lists of a fused-top-n scoring launch in ONE launch)
constexpr int kWideThreads = 1024;
constexpr int kWideItems = 12;

template <bool FROM_CANDS, int THREADS = kTopThreads, int ITEMS = kTopItems>
__global__ void __launch_bounds__(THREADS)
topk_round_kernel(const float* __restrict__ scores, const Cand* __restrict__ cands, int64_t n, int k,
                  int distinct, int64_t index_base, Cand* __restrict__ out) {
  constexpr int CHUNK = THREADS * ITEMS;
  __shared__ float sv[THREADS / 32];
  __shared__ int64_t si[THREADS / 32];
  __shared__ float best_v_s;
  __shared__ int64_t best_i_s;
  float v[ITEMS];
  int64_t ix[ITEMS];
  const int64_t base = (int64_t)blockIdx.x * CHUNK;
#pragma unroll
  for (int t = 0; t < ITEMS; ++t) {
    const int64_t p = base + (int64_t)t * THREADS + threadIdx.x;
    v[t] = 0.f;
    ix[t] = -1;
    if (p < n) {
      if (FROM_CANDS) {
        v[t] = cands[p].v;
        ix[t] = cands[p].idx;
      } else {
        v[t] = scores[p];
        ix[t] = index_base + p;
      }
      if (v[t] != v[t]) ix[t] = -1;   // NaN scores never win
    }
  }
  float last_v = 0.f;
  int64_t last_i = -1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int round = 0; round < k; ++round) {
    float bv = 0.f;
    int64_t bi = -1;
#pragma unroll
    for (int t = 0; t < ITEMS; ++t) {
      bool eligible = ix[t] >= 0;
      if (eligible && last_i >= 0)
        eligible = distinct ? (v[t] < last_v) : (v[t] < last_v || (v[t] == last_v && ix[t] > last_i));
      if (eligible && better(v[t], ix[t], bv, bi)) { bv = v[t]; bi = ix[t]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < THREADS / 32 ? sv[lane] : 0.f;
      bi = lane < THREADS / 32 ? si[lane] : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        best_v_s = bv;
        best_i_s = bi;
        Cand c;
        c.v = bv; c.pad = 0; c.idx = bi;
        out[(int64_t)blockIdx.x * k + round] = c;
      }
    }
    __syncthreads();
    last_v = best_v_s;
    last_i = best_i_s;
    __syncthreads();
    if (last_i < 0) {   // chunk exhausted: pad the rest of the list with empties
      for (int r = round + 1 + threadIdx.x; r < k; r += THREADS) {
        Cand c;
        c.v = 0.f; c.pad = 0; c.idx = -1;
        out[(int64_t)blockIdx.x * k + r] = c;
      }
      break;
    }
  }
}

}  // namespace

// Tournament over `n` scores (or, when scores == NULL, over `n` candidates in `cands`); the
// final k-entry list ends up in *result (scratch memory, valid until the next scratch user).
static int topk_enqueue(nbw_ctx* ctx, const float* scores, const Cand* cands, int64_t n, int k, int distinct,
                        int64_t index_base, Cand** result, cudaStream_t st, Cand* direct = nullptr) {
  int64_t blocks = (n + kTopChunk - 1) / kTopChunk;
  if (cands != nullptr && direct != nullptr && n > kTopChunk && n <= (int64_t)kWideThreads * kWideItems) {
    // the block lists of one scoring launch (CTAs x k entries): one wide CTA instead of two passes
    topk_round_kernel<true, kWideThreads, kWideItems><<<1, kWideThreads, 0, st>>>(nullptr, cands, n, k, distinct, 0, direct);
    NBW_CHECK_LAUNCH(ctx);
    *result = direct;
    return NBW_OK;
  }
  if (blocks == 1 && direct != nullptr) {
    // one chunk: its list IS the result — written straight to the caller's buffer, no scratch involved
    // (which also makes this form safe on a side stream)
    if (scores)
      topk_round_kernel<false><<<1, kTopThreads, 0, st>>>(scores, nullptr, n, k, distinct, index_base, direct);
    else
      topk_round_kernel<true><<<1, kTopThreads, 0, st>>>(nullptr, cands, n, k, distinct, 0, direct);
    NBW_CHECK_LAUNCH(ctx);
    *result = direct;
    return NBW_OK;
  }
  const size_t list0 = (size_t)blocks * k;
  const size_t blocks1 = (list0 + kTopChunk - 1) / kTopChunk;
  const size_t list1 = blocks1 * (size_t)k;
  int rc = nbw_scratch_ensure(ctx, Carver::pad(sizeof(Cand) * list0) + Carver::pad(sizeof(Cand) * list1) + 1024, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  Cand* buf[2] = {cv.take<Cand>(list0), cv.take<Cand>(list1)};
  if (scores)
    topk_round_kernel<false><<<(unsigned)blocks, kTopThreads, 0, st>>>(scores, nullptr, n, k, distinct, index_base, buf[0]);
  else
    topk_round_kernel<true><<<(unsigned)blocks, kTopThreads, 0, st>>>(nullptr, cands, n, k, distinct, 0, buf[0]);
  NBW_CHECK_LAUNCH(ctx);
  int cur = 0;
  int64_t count = blocks * k;
  while (blocks > 1) {
    const int64_t nb = (count + kTopChunk - 1) / kTopChunk;
    topk_round_kernel<true><<<(unsigned)nb, kTopThreads, 0, st>>>(nullptr, buf[cur], count, k, distinct, 0, buf[cur ^ 1]);
    NBW_CHECK_LAUNCH(ctx);
    cur ^= 1;
    blocks = nb;
    count = nb * k;
  }
  *result = buf[cur];
  return NBW_OK;
}

int nbw_topk_merge_enqueue(nbw_ctx* ctx, const NbwCand* cands, int64_t count, int k, int distinct, NbwCand** result,
                           cudaStream_t st, NbwCand* direct) {
  return topk_enqueue(ctx, nullptr, cands, count, k, distinct, 0, result, st, direct);
}

static int topk_readback(nbw_ctx* ctx, const Cand* dev, int k, float* top_val_h, int64_t* top_idx_h, int* n_found_h,
                         cudaStream_t st) {
  // k <= 1024 entries of 16 bytes: read back through a pinned block owned by the context (a pageable
  // destination makes the driver stage the copy and costs several microseconds per pool)
  if (!ctx->pinned_small) NBW_CUDA(ctx, cudaHostAlloc(&ctx->pinned_small, sizeof(Cand) * 1024, cudaHostAllocDefault));
  Cand* host = static_cast<Cand*>(ctx->pinned_small);
  NBW_CUDA(ctx, cudaMemcpyAsync(host, dev, sizeof(Cand) * k, cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  NBW_CUDA(ctx, nbw_await_host_scores(ctx));   // D2H of host scores runs beside the selection
  int found = 0;
  for (int i = 0; i < k; ++i) {
    if (host[i].idx < 0) break;
    top_val_h[found] = host[i].v;
    top_idx_h[found] = host[i].idx;
    ++found;
  }
  *n_found_h = found;
  return NBW_OK;
}

extern "C" int nbw_topk(nbw_ctx* ctx, const float* scores, int64_t n, int k, int distinct, int64_t index_base,
                        float* top_val_h, int64_t* top_idx_h, int* n_found_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_topk");
  NBW_REQUIRE(ctx, k >= 1 && k <= 1024, "k must be in [1, 1024]");
  NBW_REQUIRE(ctx, n >= 0 && top_val_h && top_idx_h && n_found_h, "bad argument");
  *n_found_h = 0;
  if (n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, scores != nullptr, "scores is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Cand* res = nullptr;
  int rc = topk_enqueue(ctx, scores, nullptr, n, k, distinct, index_base, &res, st);
  if (rc) return rc;
  return topk_readback(ctx, res, k, top_val_h, top_idx_h, n_found_h, st);
}

extern "C" int nbw_topk_device(nbw_ctx* ctx, const float* scores, int64_t n, int k, int distinct,
                               int64_t index_base, void* out_cands, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_topk_device");
  NBW_REQUIRE(ctx, k >= 1 && k <= 1024, "k must be in [1, 1024]");
  NBW_REQUIRE(ctx, n >= 0 && out_cands, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {   // an empty shard contributes k empty entries (index -1)
    NBW_CUDA(ctx, cudaMemsetAsync(out_cands, 0xFF, sizeof(Cand) * k, st));
    return NBW_OK;
  }
  NBW_REQUIRE(ctx, scores != nullptr, "scores is NULL");
  Cand* res = nullptr;
  int rc = topk_enqueue(ctx, scores, nullptr, n, k, distinct, index_base, &res, st);
  if (rc) return rc;
  NBW_CUDA(ctx, cudaMemcpyAsync(out_cands, res, sizeof(Cand) * k, cudaMemcpyDeviceToDevice, st));
  return NBW_OK;
}

extern "C" int nbw_topk_merge(nbw_ctx* ctx, const void* cands, int64_t count, int k, int distinct,
                              float* top_val_h, int64_t* top_idx_h, int* n_found_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_topk_merge");
  NBW_REQUIRE(ctx, k >= 1 && k <= 1024, "k must be in [1, 1024]");
  NBW_REQUIRE(ctx, count >= 0 && top_val_h && top_idx_h && n_found_h, "bad argument");
  *n_found_h = 0;
  if (count == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cands != nullptr, "cands is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Cand* res = nullptr;
  int rc = topk_enqueue(ctx, nullptr, static_cast<const Cand*>(cands), count, k, distinct, 0, &res, st);
  if (rc) return rc;
  return topk_readback(ctx, res, k, top_val_h, top_idx_h, n_found_h, st);
}

extern "C" int nbw_cands_read(nbw_ctx* ctx, const void* cands, int k, float* top_val_h, int64_t* top_idx_h,
                              int* n_found_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, k >= 1 && k <= 1024 && cands && top_val_h && top_idx_h && n_found_h, "bad argument");
  return topk_readback(ctx, static_cast<const Cand*>(cands), k, top_val_h, top_idx_h, n_found_h,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int nbw_topk_merge_device(nbw_ctx* ctx, const void* cands, int64_t count, int k, int distinct,
                                     void* out_cands, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, k >= 1 && k <= 1024 && count >= 0 && out_cands, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (count == 0) {
    NBW_CUDA(ctx, cudaMemsetAsync(out_cands, 0xFF, sizeof(Cand) * k, st));
    return NBW_OK;
  }
  NBW_REQUIRE(ctx, cands != nullptr, "cands is NULL");
  Cand* res = nullptr;
  int rc = topk_enqueue(ctx, nullptr, static_cast<const Cand*>(cands), count, k, distinct, 0, &res, st,
                        static_cast<Cand*>(out_cands));
  if (rc) return rc;
  if (res != out_cands) NBW_CUDA(ctx, cudaMemcpyAsync(out_cands, res, sizeof(Cand) * k, cudaMemcpyDeviceToDevice, st));
  return NBW_OK;
}
