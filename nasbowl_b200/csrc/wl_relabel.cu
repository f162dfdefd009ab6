// WL relabelling of packed cell batches: one sub-warp (n_max lanes) per cell, lane = node.
// Replaces the credential loop of grakel_replace/weisfeiler_lehman.py:258-296: instead of the
// string  str(own) + "," + str(sorted(successor labels))  each node gets two independent
// 64-bit multiset hashes of (own id, {successor ids}); nbw_dict_build turns them into dense
// ids and verifies that no two distinct credentials were merged.
#include "nbw_common.cuh"

template <int NMAX>
__global__ void __launch_bounds__(256)
wl_signature_kernel(const uint8_t* __restrict__ cells, int64_t n_cells,
                    const uint32_t* __restrict__ ids_prev, const uint64_t* __restrict__ uniq_prev, int level,
                    uint64_t seed, uint64_t* __restrict__ keys, uint64_t* __restrict__ chks) {
  constexpr int GPW = 32 / NMAX;   // cells per warp
  const int lane = threadIdx.x & 31;
  const int node = lane % NMAX;
  const int sub = lane / NMAX;
  const int sub_base = sub * NMAX;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const LevelSalt salt = nbw_level_salt(seed, level);

  for (int64_t g0 = warp_id * GPW; g0 < n_cells; g0 += n_warps * GPW) {
    const int64_t g = g0 + sub;
    const bool in_range = g < n_cells;
    uint32_t label = in_range ? CellTraits<NMAX>::label(cells, g, node) : 0xFFu;
    const bool present = label != 0xFFu;
    uint32_t succ = (in_range && present) ? CellTraits<NMAX>::succ(cells, g, node) : 0u;
    uint64_t key, chk;
    if (ids_prev == nullptr) {
      // level 0: the label is the op code itself (weisfeiler_lehman.py:223-229)
      key = present ? (uint64_t)label : NBW_INVALID_KEY;
      chk = 0;
    } else {
      // nodes that are an endpoint of no edge are not relabelled
      const uint32_t incoming = subwarp_or<NMAX>(succ);
      const bool endpoint = present && (succ != 0u || ((incoming >> node) & 1u));
      const uint32_t idp = in_range ? ids_prev[g * NMAX + node] : NBW_INVALID_ID;
      // A node is hashed through the KEY of its previous label (the dictionary entry of its previous id), not
      // through the id: signatures then depend on the rooted unfolding tree and the seed only, never on how a
      // particular batch numbered its labels — dictionaries of different fits can be matched key by key
      // (append-only growth of the training dictionary across BO iterations).
      const uint64_t id64 = idp != NBW_INVALID_ID ? __ldg(uniq_prev + idp) : 0ull;
      const uint64_t ha = nbw_hash_a(id64, salt);
      const uint32_t hb = nbw_hash_b(id64, salt);
      uint64_t sum_a;
      uint32_t sum_b;
      successor_sums<NMAX>(succ, sub_base, node, ha, hb, sum_a, sum_b);
      const bool counted = endpoint && idp != NBW_INVALID_ID;
      key = counted ? nbw_key_from(ha, sum_a, salt) : NBW_INVALID_KEY;
      chk = counted ? (uint64_t)nbw_chk_from(hb, sum_b, salt) : 0;
    }
    if (in_range) {
      keys[g * NMAX + node] = key;
      chks[g * NMAX + node] = chk;
    }
  }
}

extern "C" int nbw_wl_signatures(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max,
                                 const uint32_t* ids_prev, const uint64_t* uniq_keys_prev, int level, uint64_t seed,
                                 uint64_t* keys, uint64_t* chks, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_wl_signatures");
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "n_max must be 8 or 16 (got %d)", n_max);
  NBW_REQUIRE(ctx, n_cells >= 0 && level >= 0 && level < NBW_MAX_LEVELS, "bad n_cells/level");
  NBW_REQUIRE(ctx, (level == 0) == (ids_prev == nullptr) && (level == 0) == (uniq_keys_prev == nullptr),
              "ids_prev / uniq_keys_prev must be NULL exactly at level 0");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells && keys && chks, "null pointer");
  return nbw_wl_signatures_enqueue(ctx, cells, n_cells, n_max, ids_prev, uniq_keys_prev, level, seed, keys, chks,
                                   static_cast<cudaStream_t>(stream));
}

int nbw_wl_signatures_enqueue(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, const uint32_t* ids_prev,
                              const uint64_t* uniq_prev, int level, uint64_t seed, uint64_t* keys, uint64_t* chks,
                              cudaStream_t st) {
  const int cells_per_block = 256 / n_max;
  const int grid = nbw_grid_for(ctx, n_cells, cells_per_block);
  const uint8_t* c = static_cast<const uint8_t*>(cells);
  if (n_max == 8)
    wl_signature_kernel<8><<<grid, 256, 0, st>>>(c, n_cells, ids_prev, uniq_prev, level, seed, keys, chks);
  else
    wl_signature_kernel<16><<<grid, 256, 0, st>>>(c, n_cells, ids_prev, uniq_prev, level, seed, keys, chks);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}
