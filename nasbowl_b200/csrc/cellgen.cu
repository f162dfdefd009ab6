// On-device synthetic pools shaped like the reference's samplers; one thread per cell,
// rejection sampling inside the thread, counter-based RNG keyed by (seed, cell index) so
// shards generated on different GPUs concatenate to the single-GPU pool.  Bit-identical to
// nasbowl_b200/synth.py (which documents the construction rules and cites the reference:
// bayesopt/generate_test_graphs.py:26-78,95-177,559-624; darts/darts_objectives.py:159-198;
// darts/utils.py:11-98).
#include "nbw_common.cuh"

namespace {

constexpr uint64_t kGold = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kGold2 = 0xD1B54A32D192ED03ull;

__device__ __forceinline__ uint64_t gen_draw(uint64_t state, uint32_t attempt, uint32_t t) {
  return nbw_mix_b(state + kGold2 * (uint64_t)(attempt * 64u + t + 1u));
}
__device__ __forceinline__ uint32_t below(uint64_t x, uint32_t n, int field) {
  return (uint32_t)((((x >> (8 * field)) & 0xFFull) * n) >> 8);
}

// Renumber kept nodes in ascending id order and write the packed record.
template <int NMAX>
__device__ void write_compact(uint8_t* rec, const uint8_t* ops, const uint32_t* succ, uint32_t keep, int n) {
  uint8_t labels[NMAX];
  uint32_t masks[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) { labels[i] = 0xFF; masks[i] = 0; }
  for (int v = 0; v < n; ++v) {
    if (!((keep >> v) & 1u)) continue;
    const int p = __popc(keep & ((1u << v) - 1u));
    labels[p] = ops[v];
    uint32_t m = 0;
    uint32_t s = succ[v] & keep;
    while (s) {
      const int j = __ffs(s) - 1;
      s &= s - 1;
      m |= 1u << __popc(keep & ((1u << j) - 1u));
    }
    masks[p] = m;
  }
  if (NMAX == 8) {
    for (int i = 0; i < 8; ++i) { rec[i] = labels[i]; rec[8 + i] = (uint8_t)masks[i]; }
  } else {
    for (int i = 0; i < 16; ++i) {
      rec[i] = labels[i];
      rec[16 + 2 * i] = (uint8_t)(masks[i] & 0xFF);
      rec[16 + 2 * i + 1] = (uint8_t)(masks[i] >> 8);
    }
  }
}

__device__ void gen_nb101(uint64_t state, uint8_t* rec) {
  for (uint32_t attempt = 0;; ++attempt) {
    const uint64_t x = gen_draw(state, attempt, 0), y = gen_draw(state, attempt, 1);
    uint32_t succ[7] = {0, 0, 0, 0, 0, 0, 0};
    int e = 0;
    for (int r = 0; r < 7; ++r)
      for (int c = r + 1; c < 7; ++c, ++e)
        if ((x >> e) & 1ull) succ[r] |= 1u << c;
    uint8_t ops[7];
    ops[0] = 0;
    ops[6] = 1;
    for (int i = 0; i < 5; ++i) ops[1 + i] = (uint8_t)(2 + below(y, 3, i));
    uint32_t fwd = 1u;
    for (int v = 1; v < 7; ++v) {
      bool hit = false;
      for (int u = 0; u < v; ++u) hit |= ((fwd >> u) & 1u) && ((succ[u] >> v) & 1u);
      if (hit) fwd |= 1u << v;
    }
    uint32_t bwd = 1u << 6;
    for (int v = 5; v >= 0; --v)
      if (succ[v] & bwd) bwd |= 1u << v;
    const uint32_t keep = fwd & bwd;
    if (keep == 0) continue;
    int n_edges = 0;
    for (int v = 0; v < 7; ++v)
      if ((keep >> v) & 1u) n_edges += __popc(succ[v] & keep);
    if (n_edges == 0 || n_edges > 9) continue;
    write_compact<8>(rec, ops, succ, keep, 7);
    return;
  }
}

// NB201 cell from its six edge choices (0..2 real ops, 3 skip_connect, 4 none); false = empty cell.
__device__ bool build_nb201(const uint32_t* choice, uint8_t* rec) {
  // template edges (0,1),(0,2),(0,4),(1,3),(1,5),(2,6),(3,6),(4,7),(5,7),(6,7)
  const uint32_t succ0[8] = {0x16u, 0x28u, 0x40u, 0x40u, 0x80u, 0x80u, 0x80u, 0x00u};
  const uint32_t pred0[8] = {0x00u, 0x01u, 0x01u, 0x02u, 0x01u, 0x02u, 0x0Cu, 0x70u};
  uint32_t succ[8];
  for (int v = 0; v < 8; ++v) succ[v] = succ0[v];
  uint32_t removed = 0;
  for (int i = 1; i < 7; ++i) {
    const uint32_t c = choice[i - 1];
    if (c >= 3) {
      removed |= 1u << i;
      if (c == 3)
        for (int u = 0; u < 8; ++u)
          if ((pred0[i] >> u) & 1u) succ[u] |= succ0[i];
    }
  }
  uint32_t keep = 0xFFu & ~removed;
  uint32_t has_in = 0;
  for (int v = 0; v < 8; ++v) {
    succ[v] = ((keep >> v) & 1u) ? (succ[v] & keep) : 0u;
    has_in |= succ[v];
  }
  uint32_t drop = 0;
  for (int v = 0; v < 8; ++v) {
    if (!((keep >> v) & 1u)) continue;
    if (v != 0 && !((has_in >> v) & 1u)) drop |= 1u << v;
    else if (v != 7 && succ[v] == 0) drop |= 1u << v;
  }
  keep &= ~drop;
  uint32_t any_edge = 0;
  for (int v = 0; v < 8; ++v) {
    succ[v] = ((keep >> v) & 1u) ? (succ[v] & keep) : 0u;
    any_edge |= succ[v];
  }
  if (keep == 0 || any_edge == 0) return false;
  uint8_t ops[8];
  ops[0] = 0;
  ops[7] = 1;
  for (int i = 0; i < 6; ++i) ops[1 + i] = (uint8_t)(2 + (choice[i] < 2 ? choice[i] : 2));
  write_compact<8>(rec, ops, succ, keep, 8);
  return true;
}

__device__ void gen_nb201(uint64_t state, uint8_t* rec, uint8_t* enc) {
  for (uint32_t attempt = 0;; ++attempt) {
    const uint64_t x = gen_draw(state, attempt, 0);
    uint32_t choice[6];
    for (int i = 0; i < 6; ++i) choice[i] = below(x, 5, i);
    if (!build_nb201(choice, rec)) continue;
    if (enc)
      for (int i = 0; i < 8; ++i) enc[i] = i < 6 ? (uint8_t)choice[i] : 0;
    return;
  }
}

// DARTS cell from its genotype: eight (op 0..6, input 0..5) pairs; false = an input of the cell is unused
// (is_valid_darts, darts/utils.py:148-153).  Inputs may point at a later block (the reference's _mutate
// draws them from all six sources, darts_objectives.py:217-227); the construction does not care.
__device__ bool build_darts(const uint32_t* op, const uint32_t* in, uint8_t* rec) {
  uint32_t used = 0;
  for (int k = 0; k < 8; ++k) used |= 1u << in[k];
  if ((used & 3u) != 3u) return false;
  const int n = 15;
  uint8_t ops[15];
  uint32_t succ[15];
  for (int v = 0; v < n; ++v) { ops[v] = 0xFF; succ[v] = 0; }
  ops[0] = 0; ops[1] = 1; ops[14] = 2;
  for (int i = 0; i < 4; ++i) {
    ops[3 * i + 2] = (uint8_t)(4 + op[2 * i]);
    ops[3 * i + 3] = (uint8_t)(4 + op[2 * i + 1]);
    ops[3 * i + 4] = 3;
    succ[3 * i + 2] |= 1u << (3 * i + 4);
    succ[3 * i + 3] |= 1u << (3 * i + 4);
  }
  for (int i = 0; i < 4; ++i)
    for (int off = 0; off < 2; ++off) {
      const uint32_t k = in[2 * i + off];
      const uint32_t src = k <= 1 ? k : (k - 2) * 3 + 4;
      succ[src] |= 1u << (3 * i + 2 + off);
    }
  for (int i = 0; i < 4; ++i) succ[3 * i + 4] |= 1u << 14;
  uint32_t keep = (1u << n) - 1u;
  for (int j = 0; j < n; ++j) {   // skip_connect contraction
    if (((keep >> j) & 1u) && ops[j] == 6) {
      const int out = __ffs(succ[j]) - 1;
      for (int u = 0; u < n; ++u)
        if (((keep >> u) & 1u) && ((succ[u] >> j) & 1u)) succ[u] = (succ[u] | (1u << out)) & ~(1u << j);
      succ[j] = 0;
      keep &= ~(1u << j);
    }
  }
  for (int j = 0; j < n; ++j) {   // removal sweep, sequential
    if (!((keep >> j) & 1u)) continue;
    if (ops[j] > 1) {
      bool has_in = false;
      for (int u = 0; u < n; ++u) has_in |= ((keep >> u) & 1u) && ((succ[u] >> j) & 1u);
      if (!has_in) { keep &= ~(1u << j); succ[j] = 0; }
    } else if ((succ[j] & keep) == 0) {
      keep &= ~(1u << j);
    }
  }
  write_compact<16>(rec, ops, succ, keep, n);
  return true;
}

__device__ void gen_darts(uint64_t state, uint8_t* rec, uint8_t* enc) {
  for (uint32_t attempt = 0;; ++attempt) {
    const uint64_t x = gen_draw(state, attempt, 0), y = gen_draw(state, attempt, 1);
    uint32_t op[8], in[8];
    for (int i = 0; i < 4; ++i) {
      uint32_t a = (uint32_t)((((y >> (16 * i)) & 0xFFull) * (uint32_t)(i + 2)) >> 8);
      uint32_t b = (uint32_t)((((y >> (16 * i + 8)) & 0xFFull) * (uint32_t)(i + 1)) >> 8);
      if (b >= a) ++b;
      op[2 * i] = below(x, 7, 2 * i);
      op[2 * i + 1] = below(x, 7, 2 * i + 1);
      in[2 * i] = a;
      in[2 * i + 1] = b;
    }
    if (!build_darts(op, in, rec)) continue;   // both cell inputs must be used
    if (enc)
      for (int k = 0; k < 8; ++k) { enc[2 * k] = (uint8_t)op[k]; enc[2 * k + 1] = (uint8_t)in[k]; }
    return;
  }
}

// BANANAS-style mutation of pruned NB101 cells (generate_test_graphs.py:437-456, 486-492):
// every possible edge flips with probability ~1/vertice, every interior op changes with
// probability ~1/(vertice-2); prune; reject <= 3 nodes or > 9 edges; after 64 failed attempts
// fall back to a random cell.  Bit-identical to synth.mutate_cells_host.
__global__ void __launch_bounds__(256)
mutate_nb101_kernel(const uint8_t* __restrict__ parents, int64_t n_parents, uint64_t seed, int64_t first_index,
                    int64_t n_children, uint8_t* __restrict__ cells) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_children; i += stride) {
    const int64_t idx = first_index + i;
    const uint8_t* par = parents + (idx % n_parents) * 16;
    uint8_t p_ops[8];
    uint32_t p_succ[8];
    int vertice = 0;
    for (int v = 0; v < 8; ++v) {
      p_ops[v] = par[v];
      p_succ[v] = par[8 + v];
      if (par[v] != 0xFF) ++vertice;
    }
    __align__(16) uint8_t rec[16];
    bool done = false;
    if (vertice > 2) {
      const uint64_t state = nbw_mix_a((seed ^ 0x6D75746174650000ull) + kGold * (uint64_t)(idx + 1));
      for (uint32_t attempt = 0; attempt < 64 && !done; ++attempt) {
        uint64_t d[5];
        for (int t = 0; t < 5; ++t) d[t] = gen_draw(state, attempt, t);
        uint32_t succ[8];
        uint8_t ops[8];
        for (int v = 0; v < 8; ++v) { succ[v] = p_succ[v]; ops[v] = p_ops[v]; }
        int e = 0;
        for (int src = 0; src < vertice - 1; ++src)
          for (int dst = src + 1; dst < vertice; ++dst, ++e)
            if (below(d[e >> 3], (uint32_t)vertice, e & 7) == 0) succ[src] ^= 1u << dst;
        for (int ind = 1; ind < vertice - 1; ++ind) {
          if (below(d[3], (uint32_t)(vertice - 2), ind - 1) == 0) {
            const uint32_t pick = below(d[4], 2, ind - 1);
            uint8_t avail[2];
            int na = 0;
            for (uint8_t o = 2; o <= 4; ++o)
              if (o != p_ops[ind]) avail[na++] = o;
            ops[ind] = avail[pick];
          }
        }
        const int out = vertice - 1;
        uint32_t fwd = 1u;
        for (int v = 1; v < vertice; ++v) {
          bool hit = false;
          for (int u = 0; u < v; ++u) hit |= ((fwd >> u) & 1u) && ((succ[u] >> v) & 1u);
          if (hit) fwd |= 1u << v;
        }
        uint32_t bwd = 1u << out;
        for (int v = out - 1; v >= 0; --v)
          if (succ[v] & bwd) bwd |= 1u << v;
        const uint32_t keep = fwd & bwd;
        if (keep == 0) continue;
        int n_edges = 0;
        for (int v = 0; v < vertice; ++v)
          if ((keep >> v) & 1u) n_edges += __popc(succ[v] & keep);
        if (__popc(keep) <= 3 || n_edges > 9) continue;
        write_compact<8>(rec, ops, succ, keep, vertice);
        done = true;
      }
    }
    if (!done) gen_nb101(nbw_mix_a((seed ^ 0x66616C6C6261636Bull) + kGold * (uint64_t)(idx + 1)), rec);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(rec);
    *reinterpret_cast<uint4*>(cells + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(256)
generate_cells_kernel(int family, uint64_t seed, int64_t first_index, int64_t n_cells, uint8_t* __restrict__ cells,
                      uint8_t* __restrict__ enc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int rec_bytes = family == 2 ? 48 : 16;
  const int enc_bytes = family == 2 ? 16 : 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
    const uint64_t state = nbw_mix_a(seed + kGold * (uint64_t)(first_index + i + 1));
    __align__(16) uint8_t rec[48];
    uint8_t e[16];
    if (family == 0) gen_nb101(state, rec);
    else if (family == 1) gen_nb201(state, rec, e);
    else gen_darts(state, rec, e);
    uint4* dst = reinterpret_cast<uint4*>(cells + i * rec_bytes);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(rec);
    for (int q = 0; q < rec_bytes / 16; ++q)
      dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    if (enc && family != 0)
      for (int q = 0; q < enc_bytes; ++q) enc[i * enc_bytes + q] = e[q];
  }
}

// BANANAS-style mutation on the ENCODINGS of NB201 / DARTS cells (the pruned cell does not determine its
// parent architecture).  Child i edits parent i % n_parents; up to 64 attempts, then a fresh random cell.
//   NB201 (generate_test_graphs.py:412-431): each of the six ops changes with probability ~1/5 to one of
//   the other four, uniformly; an empty child is retried.
//   DARTS (darts/darts_objectives.py:201-245, mutate_main_only): `edits` sequential edits; a coin picks op
//   or wiring, a position 0..7; op: uniform over the seven ops; wiring: uniform over the six sources,
//   redrawn while it equals the sibling's or its own current source; children that leave a cell input
//   unused are retried (:258-260).
// Bit-identical to synth.mutate_encoded_host.
__global__ void __launch_bounds__(256)
mutate_encoded_kernel(int family, const uint8_t* __restrict__ parents, int64_t n_parents, uint64_t seed,
                      int64_t first_index, int64_t n_children, int edits, uint8_t* __restrict__ cells,
                      uint8_t* __restrict__ enc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int rec_bytes = family == 2 ? 48 : 16;
  const int enc_bytes = family == 2 ? 16 : 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_children; i += stride) {
    const int64_t idx = first_index + i;
    const uint8_t* par = parents + (idx % n_parents) * enc_bytes;
    const uint64_t state = nbw_mix_a((seed ^ 0x6D75746174650000ull) + kGold * (uint64_t)(idx + 1));
    __align__(16) uint8_t rec[48];
    uint8_t e[16];
    bool done = false;
    for (uint32_t attempt = 0; attempt < 64 && !done; ++attempt) {
      if (family == 1) {
        const uint64_t d0 = gen_draw(state, attempt, 0), d1 = gen_draw(state, attempt, 1);
        uint32_t choice[6];
        for (int k = 0; k < 6; ++k) {
          choice[k] = par[k];
          if (below(d0, 5, k) == 0) {
            const uint32_t pick = below(d1, 4, k);          // index among the four other ops
            choice[k] = pick < par[k] ? pick : pick + 1;
          }
        }
        done = build_nb201(choice, rec);
        if (done)
          for (int k = 0; k < 8; ++k) e[k] = k < 6 ? (uint8_t)choice[k] : 0;
      } else {
        uint32_t op[8], in[8];
        for (int k = 0; k < 8; ++k) { op[k] = par[2 * k]; in[k] = par[2 * k + 1]; }
        bool ok = true;
        for (int ed = 0; ed < edits && ok; ++ed) {
          const uint64_t d0 = gen_draw(state, attempt, 2 * ed), d1 = gen_draw(state, attempt, 2 * ed + 1);
          const uint32_t coin = below(d0, 2, 0), pos = below(d0, 8, 1);
          if (coin == 1) {
            op[pos] = below(d0, 7, 2);
          } else {
            const uint32_t sibling = in[pos ^ 1u];
            bool found = false;
            for (int t = 0; t < 13 && !found; ++t) {       // fields 3..7 of d0, then 0..7 of d1
              const uint32_t c = t < 5 ? below(d0, 6, 3 + t) : below(d1, 6, t - 5);
              if (c != sibling && c != in[pos]) { in[pos] = c; found = true; }
            }
            ok = found;
          }
        }
        done = ok && build_darts(op, in, rec);
        if (done)
          for (int k = 0; k < 8; ++k) { e[2 * k] = (uint8_t)op[k]; e[2 * k + 1] = (uint8_t)in[k]; }
      }
    }
    if (!done) {
      const uint64_t fb = nbw_mix_a((seed ^ 0x66616C6C6261636Bull) + kGold * (uint64_t)(idx + 1));
      if (family == 1) gen_nb201(fb, rec, e);
      else gen_darts(fb, rec, e);
    }
    uint4* dst = reinterpret_cast<uint4*>(cells + i * rec_bytes);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(rec);
    for (int q = 0; q < rec_bytes / 16; ++q)
      dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    if (enc)
      for (int q = 0; q < enc_bytes; ++q) enc[i * enc_bytes + q] = e[q];
  }
}

}  // namespace

extern "C" int nbw_generate_cells(nbw_ctx* ctx, int family, uint64_t seed, int64_t first_index, int64_t n_cells,
                                  void* cells, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_generate_cells");
  NBW_REQUIRE(ctx, family >= 0 && family <= 2, "family must be 0 (NB101), 1 (NB201) or 2 (DARTS)");
  NBW_REQUIRE(ctx, n_cells >= 0 && first_index >= 0, "bad range");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells != nullptr && (reinterpret_cast<uintptr_t>(cells) & 15) == 0, "cells must be 16-byte aligned");
  generate_cells_kernel<<<nbw_grid_for(ctx, n_cells, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      family, seed, first_index, n_cells, static_cast<uint8_t*>(cells), nullptr);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_generate_encoded(nbw_ctx* ctx, int family, uint64_t seed, int64_t first_index, int64_t n_cells,
                                    void* cells, uint8_t* enc, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, family == 1 || family == 2, "encodings exist for family 1 (NB201: 8 bytes) and 2 (DARTS: 16 bytes)");
  NBW_REQUIRE(ctx, n_cells >= 0 && first_index >= 0, "bad range");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells && enc && (reinterpret_cast<uintptr_t>(cells) & 15) == 0, "null / misaligned pointer");
  generate_cells_kernel<<<nbw_grid_for(ctx, n_cells, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      family, seed, first_index, n_cells, static_cast<uint8_t*>(cells), enc);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_mutate_encoded(nbw_ctx* ctx, int family, const uint8_t* parent_enc, int64_t n_parents,
                                  uint64_t seed, int64_t first_index, int64_t n_children, int edits, void* cells,
                                  uint8_t* child_enc, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, family == 1 || family == 2, "encoded mutation is for family 1 (NB201) and 2 (DARTS); NB101 uses nbw_mutate_cells");
  NBW_REQUIRE(ctx, n_parents > 0 && n_children >= 0 && first_index >= 0, "bad range");
  NBW_REQUIRE(ctx, family == 1 || (edits >= 1 && edits <= 16), "DARTS mutation takes 1..16 edits");
  if (n_children == 0) return NBW_OK;
  NBW_REQUIRE(ctx, parent_enc && cells && (reinterpret_cast<uintptr_t>(cells) & 15) == 0, "null / misaligned pointer");
  mutate_encoded_kernel<<<nbw_grid_for(ctx, n_children, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      family, parent_enc, n_parents, seed, first_index, n_children, edits, static_cast<uint8_t*>(cells), child_enc);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_mutate_cells(nbw_ctx* ctx, int family, const void* parents, int64_t n_parents, uint64_t seed,
                                int64_t first_index, int64_t n_children, void* cells, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_mutate_cells");
  NBW_REQUIRE(ctx, family == 0, "nbw_mutate_cells edits pruned NAS-Bench-101 cells; NB201 / DARTS cells are edited on their encodings (nbw_mutate_encoded)");
  NBW_REQUIRE(ctx, n_parents > 0 && n_children >= 0 && first_index >= 0, "bad range");
  if (n_children == 0) return NBW_OK;
  NBW_REQUIRE(ctx, parents && cells && (reinterpret_cast<uintptr_t>(cells) & 15) == 0, "null / misaligned pointer");
  mutate_nb101_kernel<<<nbw_grid_for(ctx, n_children, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(parents), n_parents, seed, first_index, n_children, static_cast<uint8_t*>(cells));
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}
