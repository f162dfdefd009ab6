// Vertex histograms: dense int8 feature rows Phi[g, :] from per-level dense label ids.
// Replaces VertexHistogram.parse_input (grakel_replace/vertex_histogram.py:98-209), which
// builds a float64 dense matrix through Counter() objects and fancy indexing.
//
// One CTA owns a tile of cells x a 1024-byte chunk of feature columns: the tile is
// assembled in shared memory (counts by warp-aggregation: each sub-warp compares its nodes'
// ids with shuffles, the first lane of each label class writes the class size) and is then
// streamed to HBM with 128-bit stores, so every byte of Phi is written exactly once, zeros
// included — no memset pass and no scattered global byte stores.
#include "nbw_common.cuh"

namespace {

constexpr int kHistThreads = 256;
constexpr int kChunk = 1024;   // feature columns (bytes) per CTA

struct HistLevels {
  int n_levels;
  int64_t col_off[NBW_MAX_LEVELS + 1];
  int64_t label_off[NBW_MAX_LEVELS + 1];
};

// Size of the node's label class inside its cell and its rank in it: dense 32-bit ids compared by
// one shuffle sweep over the sub-warp (a warp match gives the same answer but is much slower).
template <int NMAX>
__device__ __forceinline__ void class_count_rank32(uint32_t id, bool valid, int node, int sub_base, int& count, int& rank) {
  count = 0;
  rank = 0;
  const uint32_t mine = valid ? id : (0x80000000u | (uint32_t)node);   // uncounted nodes match nobody (ids < 2^31)
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    const uint32_t other = __shfl_sync(0xffffffffu, mine, sub_base + j);
    const bool same = valid && other == mine;
    count += same ? 1 : 0;
    rank += (same && j < node) ? 1 : 0;
  }
}

template <int NMAX>
__global__ void __launch_bounds__(kHistThreads)
hist_dense_kernel(const uint32_t* __restrict__ ids, int64_t n_cells, HistLevels lv,
                  const uint32_t* __restrict__ thermo_base, int oa, int8_t* __restrict__ phi,
                  int64_t ld_phi, int32_t* __restrict__ self_sim, int n_chunks) {
  constexpr int CPB = kHistThreads / NMAX;   // cells per block
  // consecutive blocks are the column chunks of ONE block of cells: its ids are read from HBM once and
  // from L2 by the other chunks (with the chunk index slow, every chunk re-streamed all ids: 13x the reads)
  const int chunk = (int)(blockIdx.x % (unsigned)n_chunks);
  const int64_t cell_block = blockIdx.x / (unsigned)n_chunks;
  __shared__ __align__(16) int8_t tile[CPB][kChunk];
  const int lane = threadIdx.x & 31;
  const int node = lane % NMAX;
  const int sub_base = (lane / NMAX) * NMAX;
  const int cell_local = threadIdx.x / NMAX;
  const int64_t g = cell_block * CPB + cell_local;
  const bool in_range = g < n_cells;
  const int64_t c_begin = lv.col_off[0] + (int64_t)chunk * kChunk;
  const int64_t c_end_all = lv.col_off[lv.n_levels];
  const int64_t c_end = c_begin + kChunk < c_end_all ? c_begin + kChunk : c_end_all;

  {
    uint4* t4 = reinterpret_cast<uint4*>(&tile[0][0]);
    for (int i = threadIdx.x; i < CPB * kChunk / 16; i += kHistThreads) t4[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  int self_acc = 0;
  for (int l = 0; l < lv.n_levels; ++l) {
    const uint32_t id = in_range ? ids[(int64_t)l * n_cells * NMAX + g * NMAX + node] : NBW_INVALID_ID;
    const bool valid = id != NBW_INVALID_ID;
    int count, rank;
    class_count_rank32<NMAX>(id, valid, node, sub_base, count, rank);
    if (valid) {
      if (!oa) {
        self_acc += count;   // sum over nodes of class size = sum over classes of size^2
        if (rank == 0) {
          const int64_t col = lv.col_off[l] + id;
          if (col >= c_begin && col < c_end) tile[cell_local][col - c_begin] = (int8_t)count;
        }
      } else {
        self_acc += 1;       // sum_v min(c, c) = number of counted nodes
        const int64_t col = lv.col_off[l] + thermo_base[lv.label_off[l] + id] + rank;
        if (col >= c_begin && col < c_end) tile[cell_local][col - c_begin] = 1;
      }
    }
  }
  if (chunk == 0 && self_sim != nullptr) {
    const int total = subwarp_sum<NMAX>(self_acc);
    if (in_range && node == 0) self_sim[g] = total;
  }
  __syncthreads();

  const int width16 = (int)((c_end - c_begin + 15) / 16);   // c_end - c_begin is a multiple of 16
  for (int i = threadIdx.x; i < CPB * width16; i += kHistThreads) {
    const int r = i / width16, c = i % width16;
    const int64_t gr = cell_block * CPB + r;
    if (gr < n_cells)
      *reinterpret_cast<uint4*>(phi + gr * ld_phi + c_begin + (int64_t)c * 16) =
          *reinterpret_cast<const uint4*>(&tile[r][c * 16]);
  }
}

// OA thermometer layout: per stacked label, the maximum count over the batch.
template <int NMAX>
__global__ void __launch_bounds__(kHistThreads)
max_count_kernel(const uint32_t* __restrict__ ids, int64_t n_cells, HistLevels lv,
                 uint32_t* __restrict__ max_count) {
  constexpr int GPW = 32 / NMAX;
  const int lane = threadIdx.x & 31;
  const int node = lane % NMAX;
  const int sub = lane / NMAX;
  const int sub_base = sub * NMAX;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g0 = warp_id * GPW; g0 < n_cells; g0 += n_warps * GPW) {
    const int64_t g = g0 + sub;
    const bool in_range = g < n_cells;
    for (int l = 0; l < lv.n_levels; ++l) {
      const uint32_t id = in_range ? ids[(int64_t)l * n_cells * NMAX + g * NMAX + node] : NBW_INVALID_ID;
      const bool valid = id != NBW_INVALID_ID;
      int count, rank;
      class_count_rank32<NMAX>(id, valid, node, sub_base, count, rank);
      if (valid && rank == 0) atomicMax(&max_count[lv.label_off[l] + id], (uint32_t)count);
    }
  }
}

// One block per level: exclusive scan of the level's max counts (levels are small).
__global__ void thermo_scan_kernel(const uint32_t* __restrict__ max_count32, HistLevels lv,
                                   uint32_t* __restrict__ thermo_base, uint8_t* __restrict__ max_count8,
                                   unsigned long long* __restrict__ level_width) {
  const int l = blockIdx.x;
  const int64_t lo = lv.label_off[l], hi = lv.label_off[l + 1];
  __shared__ uint32_t carry;
  __shared__ uint32_t wsum[33];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t base = lo; base < hi; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < hi ? max_count32[i] : 0;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < nwarps ? wsum[lane] : 0;
      uint32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      if (lane < nwarps) wsum[lane] = winc - w;
      if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    if (i < hi) {
      thermo_base[i] = carry + wsum[warp] + inc - v;
      max_count8[i] = (uint8_t)(v > 255u ? 255u : v);
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) level_width[l] = carry;
}

int fill_levels(nbw_ctx* ctx, HistLevels& lv, int n_levels, const int64_t* col_off_h,
                const int64_t* label_off_h) {
  NBW_REQUIRE(ctx, n_levels >= 1 && n_levels <= NBW_MAX_LEVELS, "n_levels out of range");
  lv.n_levels = n_levels;
  for (int l = 0; l <= n_levels; ++l) {
    lv.col_off[l] = col_off_h ? col_off_h[l] : 0;
    lv.label_off[l] = label_off_h ? label_off_h[l] : 0;
  }
  return NBW_OK;
}

}  // namespace

extern "C" int nbw_hist_dense_i8(nbw_ctx* ctx, const uint32_t* ids, int64_t n_cells, int n_max,
                                 int n_levels, const int64_t* col_off_h, const int64_t* label_off_h,
                                 const uint32_t* thermo_base, int oa, int8_t* phi, int64_t ld_phi,
                                 int32_t* self_sim, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_hist_dense_i8");
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "n_max must be 8 or 16");
  NBW_REQUIRE(ctx, col_off_h != nullptr, "col_off_h is NULL");
  NBW_REQUIRE(ctx, !oa || (thermo_base && label_off_h), "oa needs thermo_base and label_off_h");
  HistLevels lv;
  int rc = fill_levels(ctx, lv, n_levels, col_off_h, label_off_h);
  if (rc) return rc;
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, ids && phi, "null pointer");
  const int64_t width = lv.col_off[n_levels] - lv.col_off[0];
  NBW_REQUIRE(ctx, width > 0 && width % 16 == 0 && lv.col_off[0] % 16 == 0 && ld_phi % 16 == 0 &&
                       ld_phi >= lv.col_off[n_levels],
              "feature columns must be 16-byte aligned and fit ld_phi");
  NBW_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(phi) & 15) == 0, "phi must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpb = kHistThreads / n_max;
  const int64_t n_chunks = (width + kChunk - 1) / kChunk;
  const int64_t blocks = ((n_cells + cpb - 1) / cpb) * n_chunks;
  NBW_REQUIRE(ctx, blocks < ((int64_t)1 << 31), "histogram: too many tiles for one launch");
  const unsigned grid = (unsigned)blocks;
  if (n_max == 8)
    hist_dense_kernel<8><<<grid, kHistThreads, 0, st>>>(ids, n_cells, lv, thermo_base, oa, phi, ld_phi, self_sim,
                                                        (int)n_chunks);
  else
    hist_dense_kernel<16><<<grid, kHistThreads, 0, st>>>(ids, n_cells, lv, thermo_base, oa, phi, ld_phi, self_sim,
                                                         (int)n_chunks);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_oa_thermo_layout(nbw_ctx* ctx, const uint32_t* ids, int64_t n_cells, int n_max,
                                    int n_levels, const int64_t* label_off_h, uint32_t* thermo_base,
                                    uint8_t* max_count, int64_t* level_width_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "n_max must be 8 or 16");
  NBW_REQUIRE(ctx, label_off_h && level_width_h, "null host pointer");
  HistLevels lv;
  int rc = fill_levels(ctx, lv, n_levels, nullptr, label_off_h);
  if (rc) return rc;
  const int64_t total = lv.label_off[n_levels];
  for (int l = 0; l < n_levels; ++l) level_width_h[l] = 0;
  if (total == 0 || n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, ids && thermo_base && max_count, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = Carver::pad(sizeof(uint32_t) * total) + Carver::pad(sizeof(unsigned long long) * NBW_MAX_LEVELS) + 1024;
  rc = nbw_scratch_ensure(ctx, need, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  uint32_t* mc32 = cv.take<uint32_t>(total);
  unsigned long long* widths = cv.take<unsigned long long>(NBW_MAX_LEVELS);
  NBW_CUDA(ctx, cudaMemsetAsync(mc32, 0, sizeof(uint32_t) * total, st));
  const int grid = nbw_grid_for(ctx, n_cells, kHistThreads / n_max);
  if (n_max == 8)
    max_count_kernel<8><<<grid, kHistThreads, 0, st>>>(ids, n_cells, lv, mc32);
  else
    max_count_kernel<16><<<grid, kHistThreads, 0, st>>>(ids, n_cells, lv, mc32);
  NBW_CHECK_LAUNCH(ctx);
  thermo_scan_kernel<<<n_levels, 256, 0, st>>>(mc32, lv, thermo_base, max_count, widths);
  NBW_CHECK_LAUNCH(ctx);
  unsigned long long host[NBW_MAX_LEVELS];
  NBW_CUDA(ctx, cudaMemcpyAsync(host, widths, sizeof(unsigned long long) * n_levels, cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  for (int l = 0; l < n_levels; ++l) level_width_h[l] = (int64_t)host[l];
  return NBW_OK;
}
