// Pool de-duplication by a Weisfeiler-Lehman isomorphism test on the device.
// Reproduces `allow_isomorphism=False` of the reference's mutation() (bayesopt/generate_test_graphs.py:497-520):
// a child isomorphic to its parent or to an earlier member of the evaluation pool is dropped.  The reference
// does a pairwise networkx isomorphism search on the host (its own comment there: a WL test "is essentially
// free" next to a WL kernel); here every cell gets a 128-bit signature from colour refinement over BOTH edge
// directions, run to depth n_max, and the pool is reduced with the dictionary sort/unique of dict_build.cu —
// the sort is stable, so the first cell of every class is the one that is kept.
#include "nbw_common.cuh"

namespace {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

template <int NMAX>
__global__ void __launch_bounds__(256)
cell_iso_signature_kernel(const uint8_t* __restrict__ cells, int64_t n_cells, int labeled, uint64_t seed,
                          uint64_t* __restrict__ keys, uint64_t* __restrict__ chks) {
  constexpr int GPW = 32 / NMAX;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int node = lane % NMAX;
  const int sub_base = (lane / NMAX) * NMAX;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t s1 = nbw_mix_a(seed ^ 0x11), s2 = nbw_mix_b(seed ^ 0x22), s3 = nbw_mix_a(seed ^ 0x33);
  for (int64_t g0 = warp_id * GPW; g0 < n_cells; g0 += n_warps * GPW) {
    const int64_t g = g0 + lane / NMAX;
    const bool in_range = g < n_cells;
    const uint32_t label = in_range ? CellTraits<NMAX>::label(cells, g, node) : 0xFFu;
    const bool present = label != 0xFFu;
    const uint32_t succ = (in_range && present) ? CellTraits<NMAX>::succ(cells, g, node) : 0u;
    uint32_t pred = 0u;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      const uint32_t sj = __shfl_sync(FULL, succ, sub_base + j);
      pred |= ((sj >> node) & 1u) ? (1u << j) : 0u;
    }
    // two independent 64-bit colourings (key / check word)
    uint64_t ca = present ? nbw_mix_a((labeled ? (uint64_t)label : 0ull) + s1) : 0ull;
    uint64_t cb = present ? nbw_mix_b((labeled ? (uint64_t)label : 0ull) + s2) : 0ull;
#pragma unroll 1
    for (int r = 0; r < NMAX; ++r) {
      const uint64_t ha = nbw_mix_a(ca + s3), hb = nbw_mix_b(cb + s3);
      uint64_t sa = 0, sb = 0, pa = 0, pb = 0;
#pragma unroll 1
      for (int j = 0; j < NMAX; ++j) {
        if (!__any_sync(FULL, ((succ | pred) >> j) != 0u)) break;
        const uint64_t oa = shfl64(ha, sub_base + j), ob = shfl64(hb, sub_base + j);
        if ((succ >> j) & 1u) { sa += oa; sb += ob; }
        if ((pred >> j) & 1u) { pa += oa; pb += ob; }
      }
      // successors and predecessors enter through different bijections: direction matters
      ca = nbw_mix_a(rotl64(ha, 27) * 0x9e3779b97f4a7c15ull + sa + rotl64(pa, 41) * 0xc2b2ae3d27d4eb4full);
      cb = nbw_mix_b(rotl64(hb, 19) * 0xff51afd7ed558ccdull + sb + rotl64(pb, 33) * 0x94d049bb133111ebull);
      if (!present) { ca = 0; cb = 0; }
    }
    // the cell: multiset of final colours + node / edge counts
    uint64_t fa = present ? nbw_mix_a(ca ^ s1) : 0ull, fb = present ? nbw_mix_b(cb ^ s2) : 0ull;
    uint32_t cnt = (present ? 1u : 0u) | ((uint32_t)__popc(succ) << 8);
#pragma unroll
    for (int o = NMAX / 2; o > 0; o >>= 1) {
      fa += shfl64(fa, lane ^ o);
      fb += shfl64(fb, lane ^ o);
      cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    if (in_range && node == 0) {
      uint64_t k = nbw_mix_a(fa + (uint64_t)cnt * 0x9e3779b97f4a7c15ull);
      if (k == NBW_INVALID_KEY) k -= 1;
      keys[g] = k;
      chks[g] = nbw_mix_b(fb + (uint64_t)cnt);
    }
  }
}

__global__ void count_kept_kernel(const uint8_t* __restrict__ keep, int64_t n, int64_t skip, unsigned long long* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned int acc = 0;
  for (int64_t i = skip + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += keep[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}

}  // namespace

extern "C" int nbw_pool_dedup(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, int labeled,
                              int64_t n_protected, uint64_t seed, uint8_t* keep, int64_t* n_classes_h,
                              int64_t* n_kept_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_pool_dedup");
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "n_max must be 8 or 16 (got %d)", n_max);
  NBW_REQUIRE(ctx, n_cells >= 0 && n_cells < ((int64_t)1 << 31) && n_protected >= 0 && n_protected <= n_cells, "bad counts");
  if (n_classes_h) *n_classes_h = 0;
  if (n_kept_h) *n_kept_h = 0;
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells && keep, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t head = 1024 + 4 * Carver::pad(sizeof(uint64_t) * n_cells) + Carver::pad(sizeof(uint32_t) * n_cells);
  int rc = nbw_scratch_ensure(ctx, head + nbw_dict_scratch_bytes(n_cells), st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  NbwDictCounters* counters = cv.take<NbwDictCounters>(1);
  unsigned long long* kept = cv.take<unsigned long long>(1);
  cv.off = 1024;
  uint64_t* keys = cv.take<uint64_t>(n_cells);
  uint64_t* chks = cv.take<uint64_t>(n_cells);
  uint64_t* uk = cv.take<uint64_t>(n_cells);
  uint64_t* uc = cv.take<uint64_t>(n_cells);
  uint32_t* ids = cv.take<uint32_t>(n_cells);
  void* work = static_cast<char*>(ctx->scratch) + head;
  const int grid = nbw_grid_for(ctx, n_cells, 256 / n_max);
  const uint8_t* c = static_cast<const uint8_t*>(cells);
  if (n_max == 8) cell_iso_signature_kernel<8><<<grid, 256, 0, st>>>(c, n_cells, labeled, seed, keys, chks);
  else cell_iso_signature_kernel<16><<<grid, 256, 0, st>>>(c, n_cells, labeled, seed, keys, chks);
  NBW_CHECK_LAUNCH(ctx);
  rc = nbw_dict_build_enqueue(ctx, keys, chks, n_cells, 64, nullptr, n_max, nullptr, ids, uk, uc, n_cells, counters, work,
                              st, keep, /*allow_hash=*/false);   // a pool has far more classes than the hash path holds
  if (rc) return rc;
  NBW_CUDA(ctx, cudaMemsetAsync(kept, 0, sizeof(unsigned long long), st));
  count_kept_kernel<<<nbw_grid_for(ctx, n_cells, 256), 256, 0, st>>>(keep, n_cells, n_protected, kept);
  NBW_CHECK_LAUNCH(ctx);
  NbwDictCounters host;
  unsigned long long hk = 0;
  NBW_CUDA(ctx, cudaMemcpyAsync(&host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaMemcpyAsync(&hk, kept, sizeof(hk), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  if (host.collision)
    NBW_FAIL(ctx, NBW_ERR_HASH_COLLISION, "two different cell signatures share a key; retry with another seed");
  if (n_classes_h) *n_classes_h = host.n_unique;
  if (n_kept_h) *n_kept_h = (int64_t)hk;
  return NBW_OK;
}
