// Fused pool scoring, second generation ("packed" path): WL relabelling -> dictionary lookup -> GP
// posterior -> acquisition -> local top-n, one launch for the whole candidate pool.
// Replaces, per candidate, GraphGP.predict (bayesopt/gp.py:240-283: a WL refit on train ∪ {candidate}
// and an (n+1)^2 Gram per kernel of the sum, kernels/combine_kernels.py:36-56),
// GraphExpectedImprovement.eval / GraphUpperConfidentBound.eval (bayesopt/acquisitions.py:59-80,130-138)
// and the np.unique + argpartition selection of propose_location (:101-105).
//
// Same mathematics as score_pool.cu (prefix form: a node is described by its deepest dictionary label,
// the posterior is a sum over node pairs of H = B'^T K_lam^-1 B'), different execution:
//   * cells are packed DENSELY into the warp: a warp takes as many consecutive cells as fit its 32
//     lanes (NAS-Bench cells average 4.4-5.1 nodes: 6-7 cells per pass instead of 4);
//   * nodes that are in the dictionary FETCH their hashes (nbw_model.id_hash) instead of mixing
//     64-bit integers; nodes outside it need no hash at all, because
//   * label classes inside a cell are refined exactly from (previous class, multiset of the successors'
//     previous classes) — 4 bits of count per class — at every level, not only on the stable suffix;
//   * the dictionary probe is branch-free: one 16-byte load of four 32-bit tags, one of the entry;
//   * a SUM of WL kernels is evaluated in one pass: part p reads its own record of the candidate and
//     contributes one stacked label index per node; pairs run over (node, part) x (node, part);
//   * every lane keeps the k best scores it produced; CTAs merge them at kernel end.
#include "nbw_common.cuh"

#include <math.h>
#include <stdlib.h>

#include <type_traits>

namespace {

constexpr int kPackThreads = 256;
constexpr int kPackWarps = kPackThreads / 32;
constexpr int kTopMax = 8;   // fused top-n handles k <= 8 (propose_location's default is 5)

struct PartParams {
  const uint16_t* level0_table;
  const uint4* id_hash[NBW_MAX_LEVELS];
  const uint4* tags[NBW_MAX_LEVELS];
  const uint4* entries[NBW_MAX_LEVELS];
  uint32_t bucket_mask[NBW_MAX_LEVELS];
  uint64_t a_own[NBW_MAX_LEVELS];
  uint32_t b_own[NBW_MAX_LEVELS];
  int32_t loff[NBW_MAX_LEVELS];     // stacked offset of level l inside this part
  double lw[NBW_MAX_LEVELS];        // level weights (kernel weight x layer weight)
  double wsuf[NBW_MAX_LEVELS];      // sum of lw over the levels deeper than l
  const int32_t* child;
  const int32_t* final_label;
  const uint32_t* oa_exc;           // histogram intersection: (first exception item << 5 | max count) per label
  int32_t h, first_stable, undirected, rec_offset, label_base, compact;
  int32_t oa_d2_base;               // histogram intersection: first rank-2 chain item (0 = none)
  uint8_t palette[16];
};

struct PackedParams {
  PartParams part[NBW_MAX_PARTS];
  const double* H;
  const double* w_mu;
  int64_t ld_h;
  int32_t zero_slot, rec_stride, acq, pad;
  double lik, y_mean, y_std, incumbent, xi, beta;
};

template <int NMAX> struct SigOf;
template <> struct SigOf<8> { typedef uint32_t type; };
template <> struct SigOf<16> { typedef uint64_t type; };
__device__ __forceinline__ uint32_t shfl_sig(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ uint64_t shfl_sig(uint64_t v, int src) { return shfl64(v, src); }

__device__ __forceinline__ double shfl_down_f64(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }

// bit i = byte i of w is not 0xFF
__device__ __forceinline__ uint32_t present_bits4(uint32_t w) {
  return ((__vcmpne4(w, 0xFFFFFFFFu) & 0x08040201u) * 0x01010101u) >> 24;
}

// Posterior scaling + acquisition for one candidate (fp64): gp.py:272-280, acquisitions.py:68-77,136.
__device__ __forceinline__ double finish_score(const PackedParams& P, double mu_acc, double q_acc, double kcc,
                                               double& mu, double& var) {
  const double inv_kcc = kcc > 0.0 ? 1.0 / kcc : 0.0;   // a cell without nodes falls back to the prior
  const double mu_n = mu_acc * sqrt(inv_kcc);
  const double q = q_acc * inv_kcc;
  const double var_n = fmax(1.0 + P.lik - q, P.lik);    // gp.py:272-276 (diagonal)
  const double sd = sqrt(var_n) * P.y_std;              // gp.py:278-279
  var = sd * sd;                                        // gp.py:280
  mu = mu_n * P.y_std + P.y_mean;                       // gp.py:277
  if (P.acq == NBW_ACQ_UCB) return mu + P.beta * sd;    // acquisitions.py:136
  if (P.acq == NBW_ACQ_MEAN) return mu;
  const double d = mu - P.incumbent - P.xi;             // acquisitions.py:71-74
  const double u = d / sd;
  const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
  const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
  double score = sd * pdf + d * cdf;
  if (P.acq == NBW_ACQ_AEI) score *= 1.0 - sqrt(P.lik) / sqrt(P.lik + var);   // :75-77
  return score;
}

__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) {
  if (i < 0) return false;
  if (bi < 0) return true;
  return v > bv || (v == bv && i < bi);
}

// Exception item of a node with 0-based rank r0 > 0 inside its label class (nbw_model.oa_exc)
__device__ __forceinline__ int oa_item(uint32_t info, int r0) {
  const int mc = (int)(info & 31u);
  return (int)(info >> 5) + (r0 < mc ? r0 : mc) - 1;
}

template <int NMAX, int NP, bool WEIGHTED, bool TOPK, bool OA = false, bool WIRE = false>
__global__ void __launch_bounds__(kPackThreads, NMAX == 8 ? 5 : 4)
score_packed_kernel(const __grid_constant__ PackedParams P, const uint8_t* __restrict__ cells, int64_t n_cells,
                    int64_t per_warp, float* __restrict__ scores, double* __restrict__ mu64,
                    double* __restrict__ var64, double* __restrict__ score64, NbwCand* __restrict__ block_cands,
                    int k_top, int distinct, int64_t index_base) {
  constexpr int PEEK = NMAX == 8 ? 8 : 4;     // cells inspected per pass
  constexpr unsigned FULL = 0xffffffffu;
  using SigT = typename SigOf<NMAX>::type;
  using KccT = typename std::conditional<WEIGHTED, double, int>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t c0 = warp_id * per_warp;
  const int64_t c1 = (c0 + per_warp < n_cells) ? c0 + per_warp : n_cells;
  const int64_t stride = P.rec_stride;
  const int zero_slot = P.zero_slot;
  const double* __restrict__ H = P.H;
  const int64_t ldh = P.ld_h;

  // Results are staged until (almost) every lane holds a candidate: the fp64 acquisition epilogue then runs
  // with the warp full instead of once per pass with a few lanes.  Staging slots and the lanes' top-n lists
  // live in shared memory (one column per thread, conflict-free): they are touched once per pass / per flush,
  // and keeping them out of the register file buys a fourth and fifth resident CTA.
  __shared__ double sh_mu[kPackThreads], sh_q[kPackThreads], sh_k[kPackThreads];
  __shared__ int sh_g[kPackThreads];
  __shared__ float sh_tv[TOPK ? kTopMax : 1][kPackThreads];
  __shared__ int sh_ti[TOPK ? kTopMax : 1][kPackThreads];
  // histogram intersection: the exception item of this lane's node at every level (a prefix of the levels: ranks
  // do not grow along a node's chain), one column per thread
  __shared__ int exc_s[OA ? NBW_MAX_LEVELS : 1][kPackThreads];
  static_assert(!OA || NP == 1, "the packed histogram-intersection path takes one kernel");
  const int tid = threadIdx.x;
  // succ_lut[mask]: positions of the four lowest set bits of an 8-bit successor mask (4 bits each; positions
  // beyond the population repeat a harmless 0) | the remaining bits << 16
  __shared__ uint32_t succ_lut[NMAX == 8 ? 256 : 1];
  if (NMAX == 8) {
    uint32_t m = (uint32_t)tid, pk = 0u;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (m) pk |= (uint32_t)(__ffs(m) - 1) << (4 * t);
      m &= m - 1u;
    }
    succ_lut[tid] = pk | (m << 16);
    __syncthreads();
  }
  int staged = 0;
  sh_g[tid] = -1;
  if constexpr (TOPK) {
#pragma unroll
    for (int j = 0; j < kTopMax; ++j) { sh_tv[j][tid] = 0.f; sh_ti[j][tid] = -1; }
  }
  __syncwarp();

  auto flush = [&]() {
    const int gl = sh_g[tid];                 // index of the staged candidate inside this launch's pool
    if (gl >= 0) {
      double mu, var;
      const double sc = finish_score(P, sh_mu[tid], sh_q[tid], sh_k[tid], mu, var);
      const float scf = (float)sc;
      if (scores) scores[gl] = scf;
      if (mu64) mu64[gl] = mu;
      if (var64) var64[gl] = var;
      if (score64) score64[gl] = sc;
      if constexpr (TOPK) if (scf == scf) {
        // sorted insert into this lane's list: value desc, index asc; in distinct mode a value already in the
        // list keeps its lowest index
        bool skip = false;
        if (distinct) {
#pragma unroll
          for (int j = 0; j < kTopMax; ++j) {
            const int tj = sh_ti[j][tid];
            if (tj >= 0 && sh_tv[j][tid] == scf) { sh_ti[j][tid] = gl < tj ? gl : tj; skip = true; }
          }
        }
        if (!skip && cand_better(scf, gl, sh_tv[k_top - 1][tid], sh_ti[k_top - 1][tid])) {
          float cv = scf;
          int ci = gl;
#pragma unroll
          for (int j = 0; j < kTopMax; ++j) {
            if (j < k_top && ci >= 0) {
              const float ov = sh_tv[j][tid];
              const int oi = sh_ti[j][tid];
              if (cand_better(cv, ci, ov, oi)) {
                sh_tv[j][tid] = cv; sh_ti[j][tid] = ci;
                cv = ov; ci = oi;
              }
            }
          }
        }
      }
    }
    sh_g[tid] = -1;
    staged = 0;
  };

  for (int64_t g = c0; g < c1;) {
    // ---------------------------------------------------------------- pack cells into the warp
    int allot = 0;
    const bool peek = lane < PEEK && g + lane < c1;
    if (peek) {
      const uint8_t* rec = cells + (g + lane) * stride;
      uint32_t pm = 0;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (NMAX == 16 && WIRE) {
          // 24-byte wire record: sixteen 4-bit palette indices, 15 = no node
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(rec + P.part[p].rec_offset));
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            pm |= (((w.x >> (4 * v)) & 15u) != 15u) ? (1u << v) : 0u;
            pm |= (((w.y >> (4 * v)) & 15u) != 15u) ? (1u << (8 + v)) : 0u;
          }
        } else if (NMAX == 8) {
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(rec + P.part[p].rec_offset));
          pm |= present_bits4(w.x) | (present_bits4(w.y) << 4);
        } else {
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(rec + P.part[p].rec_offset));
          pm |= present_bits4(w.x) | (present_bits4(w.y) << 4) | (present_bits4(w.z) << 8) | (present_bits4(w.w) << 12);
        }
      }
      allot = pm ? 32 - __clz(pm) : 1;     // lanes up to the highest node; an empty cell still owns one lane
    }
    int end = allot;
#pragma unroll
    for (int o = 1; o < PEEK; o <<= 1) {
      const int t = __shfl_up_sync(FULL, end, o);
      if (lane >= o) end += t;
    }
    const bool fits = peek && end <= 32;
    const int m = __popc(__ballot_sync(FULL, fits));          // `fits` is a prefix of the lanes; m >= 1
    const int begin = end - allot;
    const unsigned starts = __reduce_or_sync(FULL, fits ? (1u << begin) : 0u);
    const int total = __shfl_sync(FULL, end, m - 1);
    const unsigned below = starts & (FULL >> (31 - lane));
    const bool active = lane < total;
    const int k = active ? __popc(below) - 1 : 0;
    const int sub_base = active ? 31 - __clz(below) : 0;
    const int node = lane - sub_base;
    const int my_n = __shfl_sync(FULL, allot, k);
    const int64_t my_g = g + k;

    int deep[NP];
    int n_exc = 0;              // OA: exception levels of this lane's node
    KccT kacc = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const PartParams& Q = P.part[p];
      const uint8_t* rec = cells + my_g * stride + Q.rec_offset;
      constexpr bool wire16 = NMAX == 16 && WIRE;       // 24-byte wire records are a compile-time variant
      uint32_t label = (active && !wire16) ? (uint32_t)__ldg(rec + node) : 0xFFu;
      bool present = label != 0xFFu;
      uint32_t succ = 0u;
      if (present)
        succ = (NMAX == 8 ? (uint32_t)__ldg(rec + 8 + node) : (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(rec + 16) + node)) &
               ((1u << my_n) - 1u);     // edges to absent nodes would reach into a neighbouring cell's lanes
      if (wire16 && active) {
        // 24-byte wire record (nbw_model.compact, n_max = 16): my 4-bit palette index, my row of the upper triangle
        const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(rec));
        const uint32_t idx = ((node < 8 ? w0.x : w0.y) >> (4 * (node & 7))) & 15u;
        if (idx != 15u) {
          label = Q.palette[idx];
          present = true;
          if (node < 15) {
            const uint2 t0 = __ldg(reinterpret_cast<const uint2*>(rec + 8)), t1 = __ldg(reinterpret_cast<const uint2*>(rec + 16));
            const uint64_t lo = ((uint64_t)t0.y << 32) | t0.x, hi = ((uint64_t)t1.y << 32) | t1.x;
            const int off = node * (31 - node) / 2, len = 15 - node;
            uint64_t x = off < 64 ? (lo >> off) : (hi >> (off - 64));
            if (off > 0 && off < 64) x |= hi << (64 - off);
            succ = (((uint32_t)x & ((1u << len) - 1u)) << (node + 1)) & ((1u << my_n) - 1u);
          }
        }
      }
      if (Q.undirected) {
        // kernels/weisfilerlehman.py:127-128: neighbourhood = successors and predecessors
        uint32_t pred = 0u;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
          const uint32_t sj = __shfl_sync(FULL, succ, (sub_base + j) & 31);
          pred |= (j < my_n && ((sj >> node) & 1u)) ? (1u << j) : 0u;
        }
        succ = present ? (succ | pred) : 0u;
      }
      // nodes that are an endpoint of no edge are not relabelled
      const unsigned has_pred = __reduce_or_sync(FULL, succ << sub_base);
      const bool endpoint = present && (succ != 0u || ((has_pred >> lane) & 1u));
      const unsigned epm = __ballot_sync(FULL, endpoint) >> sub_base;

      // ---- level 0: op code -> dense id (weisfeiler_lehman.py:430-441 in transform form)
      const uint32_t d0 = present ? (uint32_t)__ldg(Q.level0_table + label) : 0xFFFFu;
      bool in_dict = present && d0 != 0xFFFFu;
      uint32_t dense_prev = in_dict ? d0 : 0u;
      // nodes of my cell (bit j) that share my label at the current level.  Level 0: equal op codes, found
      // with eight ballots over the bits of the code (exact for any u8 code) instead of a shuffle sweep.
      uint32_t cls;
      {
        const unsigned cell_lanes = ((my_n >= 32 ? 0u : (1u << my_n)) - 1u) << sub_base;
        unsigned same = __ballot_sync(FULL, present) & cell_lanes;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const unsigned vote = __ballot_sync(FULL, (label >> b) & 1u);
          same &= ((label >> b) & 1u) ? vote : ~vote;
        }
        cls = present ? (same >> sub_base) : 0u;
      }
      int cnt = __popc(cls);
      KccT kss = 0;
      // self-similarity: sum over labels of count^2 (linear) or of min(count, count) = count (intersection)
      if (present) kss = WEIGHTED ? (KccT)(Q.lw[0] * (OA ? 1 : cnt)) : (KccT)(OA ? 1 : cnt);
      int dp = in_dict ? Q.label_base + Q.loff[0] + (int)d0 : zero_slot;
      int ec = 0;                 // levels 0 .. ec-1 carry an exception item
      // a node of rank 2 at level 0 keeps rank 2 wherever it has an exception: one chain item (nbw_model.oa_d2_base)
      bool twin = false;
      int twin_item = zero_slot;
      if constexpr (OA) {
        const int r0 = __popc(cls & ((1u << node) - 1u));
        int item = zero_slot;
        if (in_dict && r0 > 0) {
          twin = r0 == 1 && Q.oa_d2_base != 0;
          if (twin) {
            twin_item = Q.oa_d2_base + Q.loff[0] + (int)d0;
          } else {
            item = oa_item(__ldg(Q.oa_exc + Q.loff[0] + (int)d0), r0);
            ec = 1;
          }
        }
        exc_s[0][tid] = item;
      }

      // The successors of a node do not change from level to level: their lanes are resolved once.  8-node
      // cells look the (up to four) lowest successor positions up in a 256-entry table built at kernel start
      // (find-first-set runs on the quarter-rate XU pipe).
      constexpr int kCached = 4;
      int sl[kCached];
      uint32_t rest;
      if (NMAX == 8) {
        const uint32_t pk = succ_lut[succ & 0xFFu];
#pragma unroll
        for (int t = 0; t < kCached; ++t) sl[t] = sub_base + (int)((pk >> (4 * t)) & 7u);
        rest = pk >> 16;                      // successors beyond the fourth (bitmask)
      } else {
        rest = succ;
#pragma unroll
        for (int t = 0; t < kCached; ++t) {
          sl[t] = rest ? (sub_base + __ffs(rest) - 1) : lane;
          rest &= rest - 1u;
        }
      }
      const int ns = __popc(succ);
      const bool more2 = __any_sync(FULL, ns > 2);
      const bool more4 = __any_sync(FULL, rest != 0u);
      const unsigned succ_lanes = succ << sub_base;        // my successors as lanes of the warp

      // ---- levels 1..h
      const int h = Q.h, fs = Q.first_stable;
#pragma unroll 1
      for (int l = 1; l <= h; ++l) {
        const bool stable = fs != 0 && l >= fs;   // nbw_model.child: membership is inherited, no signature
        // every successor's previous label is in the dictionary: one ballot instead of per-successor bookkeeping
        const bool all_in = (succ_lanes & __ballot_sync(FULL, present && !in_dict)) == 0u;
        // what a node contributes to its predecessors: one count in the 4-bit field of its previous class
        // (the class representative is the lowest member: 2^rep -> 2^(4 rep) by two squarings on the FMA pipe)
        SigT sigbit;
        {
          const uint32_t low = cls & (0u - cls);
          const uint32_t sq = low * low;
          sigbit = (SigT)sq * (SigT)sq;
        }
        // ... and, on hashed levels, the hashes of its previous id, FETCHED (nbw_model.id_hash)
        uint32_t ha_lo = 0u, ha_hi = 0u, hb = 0u;
        if (!stable) {
          const uint4 e = __ldg(Q.id_hash[l] + dense_prev);
          ha_lo = e.x; ha_hi = e.y; hb = e.z;
        }
        SigT sig = 0;           // multiset of the successors' previous classes, 4 bits of count per class
        uint64_t sum_a = 0;
        uint32_t sum_b = 0;
#define NBW_SUCC_ROUND(HAS, SRC)                                                               \
  {                                                                                            \
    const bool has = (HAS);                                                                    \
    const SigT osb = shfl_sig(sigbit, (SRC));                                                  \
    if (has) sig += osb;                                                                       \
    if (!stable) {                                                                             \
      const uint32_t olo = __shfl_sync(FULL, ha_lo, (SRC));                                    \
      const uint32_t ohi = __shfl_sync(FULL, ha_hi, (SRC));                                    \
      const uint32_t ob = __shfl_sync(FULL, hb, (SRC));                                        \
      if (has) { sum_a += ((uint64_t)ohi << 32) | olo; sum_b += ob; }                          \
    }                                                                                          \
  }
        NBW_SUCC_ROUND(ns > 0, sl[0])
        NBW_SUCC_ROUND(ns > 1, sl[1])
        if (more2) {
          NBW_SUCC_ROUND(ns > 2, sl[2])
          NBW_SUCC_ROUND(ns > 3, sl[3])
          if (more4) {
            uint32_t ms = rest;
#pragma unroll 1
            while (__any_sync(FULL, ms != 0u)) {
              const int j = ms ? (sub_base + __ffs(ms) - 1) : lane;
              NBW_SUCC_ROUND(ms != 0u, j)
              ms &= ms - 1u;
            }
          }
        }
#undef NBW_SUCC_ROUND
        // a credential can only be in the training dictionary if all its ids are
        const bool cand = endpoint && in_dict && all_in;
        bool seen = false;
        uint32_t dense = 0u;
        if (!stable) {
          // branch-free probe: every lane computes a key (lanes that are no candidates from id 0) and
          // fetches one bucket of four tags; only a tag hit fetches the entry
          LevelSalt salt;
          salt.a_own = Q.a_own[l];
          salt.b_own = Q.b_own[l];
          const uint64_t key = nbw_key_from(((uint64_t)ha_hi << 32) | ha_lo, sum_a, salt);
          const uint32_t chk = nbw_chk_from(hb, sum_b, salt);
          const uint32_t tag = (uint32_t)key, khi = (uint32_t)(key >> 32);
          const uint32_t b = khi & Q.bucket_mask[l];
          const uint4 t4 = __ldg(Q.tags[l] + b);
          const int idx = t4.x == tag ? 0 : t4.y == tag ? 1 : t4.z == tag ? 2 : t4.w == tag ? 3 : -1;
          if (cand && idx >= 0) {
            const uint4 en = __ldg(Q.entries[l] + 4 * (size_t)b + idx);
            if (en.x == khi && en.y == chk) { seen = true; dense = en.z; }
          }
        } else if (cand) {
          seen = true;
          dense = (uint32_t)__ldg(Q.child + Q.loff[l - 1] + (int)dense_prev);
        }
        // classes refine exactly: same previous class, both relabelled, same successor-class multiset
        const uint32_t cls_in = cls;
        uint32_t rem = endpoint ? (cls & epm & ~(1u << node)) : 0u;
        cls = endpoint ? (1u << node) : 0u;
#pragma unroll 1
        while (__any_sync(FULL, rem != 0u)) {
          const bool has = rem != 0u;
          const int j = has ? (__ffs(rem) - 1) : node;
          const SigT other = shfl_sig(sig, (sub_base + j) & 31);
          cls |= (has && other == sig) ? (1u << j) : 0u;
          rem &= rem - 1u;
        }
        const bool changed = cls != cls_in || seen != in_dict;
        in_dict = seen;
        dense_prev = dense;
        cnt = __popc(cls);
        if (endpoint) kss += WEIGHTED ? (KccT)(Q.lw[l] * (OA ? 1 : cnt)) : (KccT)(OA ? 1 : cnt);
        if (seen) dp = Q.label_base + Q.loff[l] + (int)dense;
        if constexpr (OA) {
          const int r0 = __popc(cls & ((1u << node) - 1u));
          int item = zero_slot;
          if (seen && r0 > 0) {
            if (twin) {
              twin_item = Q.oa_d2_base + Q.loff[l] + (int)dense;
            } else {
              item = oa_item(__ldg(Q.oa_exc + Q.loff[l] + (int)dense), r0);
              ec = l + 1;
            }
          }
          exc_s[l][tid] = item;
        }
        if (stable && l < h && !__any_sync(FULL, changed)) {
          // A stable level reproduced (classes, membership) for every cell of this warp: all deeper levels
          // apply the same map, so they reproduce it too; only the labels move down the child chain.
          if (endpoint) kss += WEIGHTED ? (KccT)(Q.wsuf[l] * (OA ? 1 : cnt)) : (KccT)((h - l) * (OA ? 1 : cnt));
          if constexpr (OA) {
            // ... and a node that is not the first of its class keeps its rank: one exception item per deeper level
            const int r0 = __popc(cls & ((1u << node) - 1u));
            if (seen && r0 > 0) {
              if (twin) {
                twin_item = Q.oa_d2_base + __ldg(Q.final_label + Q.loff[l] + (int)dense) - Q.label_base;
              } else {
                uint32_t dn = dense;
                for (int l2 = l + 1; l2 <= h; ++l2) {
                  dn = (uint32_t)__ldg(Q.child + Q.loff[l2 - 1] + (int)dn);
                  exc_s[l2][tid] = oa_item(__ldg(Q.oa_exc + Q.loff[l2] + (int)dn), r0);
                }
                ec = h + 1;
              }
            }
          }
          if (seen) dp = __ldg(Q.final_label + Q.loff[l] + (int)dense);
          break;
        }
      }
      deep[p] = dp;
      if (OA && twin) {
        exc_s[0][tid] = twin_item;
        ec = 1;
      }
      n_exc = ec;
      kacc += kss;
    }

    // ---------------------------------------------------------------- posterior (prefix form)
    double mu_acc = 0.0, q_acc = 0.0;
    if constexpr (OA) {
      // items of a node: its deepest label + its exception items; the sum runs over all item pairs of the cell
      __syncwarp();
      const int own = deep[0];
      if (own != zero_slot) {
        mu_acc += __ldg(P.w_mu + own);
        q_acc += __ldg(H + (int64_t)own * ldh + own);
      }
      for (int a = 0; a < n_exc; ++a) {
        const int xa = exc_s[a][tid];
        mu_acc += __ldg(P.w_mu + xa);
        const double* __restrict__ hrow = H + (int64_t)xa * ldh;
        double acc = own != zero_slot ? __ldg(hrow + own) : 0.0;
        for (int b = a + 1; b < n_exc; ++b) acc += __ldg(hrow + exc_s[b][tid]);
        q_acc += 2.0 * acc + __ldg(hrow + xa);
      }
#pragma unroll 1
      for (int r = 1; r <= NMAX / 2; ++r) {
        int t = node + r;
        t -= (t >= my_n) ? my_n : 0;
        const bool use = active && 2 * r <= my_n;
        const double f = (2 * r == my_n) ? 1.0 : 2.0;
        const int pl = (sub_base + t) & 31;
        const int pd = __shfl_sync(FULL, own, pl);
        const int pec = __shfl_sync(FULL, n_exc, pl);
        if (use) {
          const int ptid = tid - lane + pl;
          double acc = 0.0;
          if (own != zero_slot) {
            const double* __restrict__ hrow = H + (int64_t)own * ldh;
            if (pd != zero_slot) acc += __ldg(hrow + pd);
            for (int b = 0; b < pec; ++b) acc += __ldg(hrow + exc_s[b][ptid]);
          }
          for (int a = 0; a < n_exc; ++a) {
            const double* __restrict__ hrow = H + (int64_t)exc_s[a][tid] * ldh;
            if (pd != zero_slot) acc += __ldg(hrow + pd);
            for (int b = 0; b < pec; ++b) acc += __ldg(hrow + exc_s[b][ptid]);
          }
          q_acc += f * acc;
        }
      }
      __syncwarp();
    } else {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (deep[p] != zero_slot) {
        mu_acc += __ldg(P.w_mu + deep[p]);
        const double* __restrict__ hrow = H + (int64_t)deep[p] * ldh;
#pragma unroll
        for (int p2 = p; p2 < NP; ++p2)
          if (deep[p2] != zero_slot) {
            const double gv = __ldg(hrow + deep[p2]);
            q_acc += (p2 == p) ? gv : 2.0 * gv;
          }
      }
    }
#pragma unroll
    for (int r = 1; r <= NMAX / 2; ++r) {
      // unordered node pairs at cyclic distance r: the rotation meets each once (twice when 2r == my_n)
      int t = node + r;
      t -= (t >= my_n) ? my_n : 0;
      const bool use = active && 2 * r <= my_n;
      const double f = (2 * r == my_n) ? 1.0 : 2.0;
      int pd[NP];
#pragma unroll
      for (int p2 = 0; p2 < NP; ++p2) pd[p2] = __shfl_sync(FULL, deep[p2], (sub_base + t) & 31);
      if (use) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          if (deep[p] != zero_slot) {
            const double* __restrict__ hrow = H + (int64_t)deep[p] * ldh;
#pragma unroll
            for (int p2 = 0; p2 < NP; ++p2)
              if (pd[p2] != zero_slot) q_acc += f * __ldg(hrow + pd[p2]);
          }
        }
      }
    }
    }
    // segmented sums over the lanes of a cell: totals land on its first lane, in an order that
    // depends on the node index only (scores do not depend on how cells were packed)
    double kcc = (double)kacc;
#pragma unroll
    for (int o = 1; o < NMAX; o <<= 1) {
      const double tm = shfl_down_f64(mu_acc, o), tq = shfl_down_f64(q_acc, o), tk = shfl_down_f64(kcc, o);
      if (node + o < my_n) { mu_acc += tm; q_acc += tq; kcc += tk; }
    }

    // ---------------------------------------------------------------- stage the m finished cells
    {
      const int c = lane - staged;
      const bool take = c >= 0 && c < m;
      const int src = __shfl_sync(FULL, begin, take ? c : 0);
      const double s_mu = __shfl_sync(FULL, mu_acc, src);
      const double s_q = __shfl_sync(FULL, q_acc, src);
      const double s_k = __shfl_sync(FULL, kcc, src);
      if (take) { sh_mu[tid] = s_mu; sh_q[tid] = s_q; sh_k[tid] = s_k; sh_g[tid] = (int)(g + c); }
    }
    staged += m;
    g += m;
    if (staged > 32 - PEEK) flush();
  }
  if (staged > 0) flush();

  if constexpr (TOPK) {
    // ---- CTA tournament over the lanes' lists: k rounds of (value desc, index asc) arg-max with the
    // distinct-by-value rule of acquisitions.py:101-105; the list of the block goes to block_cands
    __shared__ float sv[kPackWarps];
    __shared__ int si[kPackWarps];
    __shared__ float best_v_s;
    __shared__ int best_i_s;
    const int warp = threadIdx.x >> 5;
    float last_v = 0.f;
    int last_i = -1;
    for (int round = 0; round < k_top; ++round) {
      float bv = 0.f;
      int bi = -1;
#pragma unroll
      for (int j = 0; j < kTopMax; ++j) {
        const float tvj = sh_tv[j][tid];
        const int tij = sh_ti[j][tid];
        bool eligible = j < k_top && tij >= 0;
        if (eligible && last_i >= 0)
          eligible = distinct ? (tvj < last_v) : (tvj < last_v || (tvj == last_v && tij > last_i));
        if (eligible && cand_better(tvj, tij, bv, bi)) { bv = tvj; bi = tij; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        bv = lane < kPackWarps ? sv[lane] : 0.f;
        bi = lane < kPackWarps ? si[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(FULL, bv, o);
          const int oi = __shfl_xor_sync(FULL, bi, o);
          if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
          best_v_s = bv;
          best_i_s = bi;
          NbwCand c;
          c.v = bv;
          c.pad = 0;
          c.idx = bi >= 0 ? index_base + bi : -1;
          block_cands[(int64_t)blockIdx.x * k_top + round] = c;
        }
      }
      __syncthreads();
      last_v = best_v_s;
      last_i = best_i_s;
      __syncthreads();
      if (last_i < 0) {   // block exhausted: the rest of its list is empty
        for (int r = round + 1 + threadIdx.x; r < k_top; r += kPackThreads) {
          NbwCand c;
          c.v = 0.f;
          c.pad = 0;
          c.idx = -1;
          block_cands[(int64_t)blockIdx.x * k_top + r] = c;
        }
        break;
      }
    }
  }
}

// ------------------------------------------------------------------ one THREAD per cell (8-node records)
// The lane-per-node kernel above spends most of its instructions on moving data between lanes (shuffles, votes,
// segmented reductions, rounds that run to the widest node of the warp).  For 8-node cells a single thread can
// hold a whole cell: labels, adjacency rows and label classes are bytes of 64-bit registers, per-node arrays
// are statically indexed registers, and the only dynamically indexed data — what a node contributes to its
// predecessors (fetched hashes + class count) — sits in a per-thread column of shared memory (bank =
// thread, whatever the node: conflict-free).  A warp scores 32 cells per pass with no inter-lane traffic at
// all; divergence (cells of different sizes) costs less than the idle lanes it replaces.
constexpr int kCellThreads = 128;

// bit j of byte i of the result = bit i of byte j of x: transpose of an 8 x 8 bit matrix (successor rows ->
// predecessor rows)
__device__ __forceinline__ uint64_t transpose8x8(uint64_t x) {
  uint64_t t;
  t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull; x ^= t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x ^= t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x ^= t ^ (t << 28);
  return x;
}
// bit i = byte i of w is not zero
__device__ __forceinline__ uint32_t nonzero_bits4(uint32_t w) {
  return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t eq_bits8(uint64_t w, uint32_t byte) {
  const uint32_t rep = byte * 0x01010101u;
  const uint32_t lo = ((__vcmpeq4((uint32_t)w, rep) & 0x08040201u) * 0x01010101u) >> 24;
  const uint32_t hi = ((__vcmpeq4((uint32_t)(w >> 32), rep) & 0x08040201u) * 0x01010101u) >> 24;
  return lo | (hi << 4);
}

// ---------------------------------------------------------------- size-binned order of a pool
// The thread-per-cell kernel steps over the eight node slots of every cell; a warp skips a slot only if none of its
// 32 cells has that node.  Cells arrive in arbitrary order (NB cells average ~5 of 8 nodes), so almost every warp pays
// for eight.  A counting sort of the cell INDICES by node count (largest first) gives every warp cells of one size:
// two small kernels over the records, then the scoring kernel walks the pool through that order.  Scores, indices and
// the top-n are those of the original positions.
constexpr int kBinMax = 8 * NBW_MAX_PARTS + 1;
constexpr int64_t kBinMinCells = 1 << 16;      // below this the two extra launches cost more than they save

__device__ __forceinline__ int cell_nodes(const uint8_t* __restrict__ rec, bool compact) {
  if (compact) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(rec));
    const uint32_t f = r.x & 0xFFFFFFu;                       // eight 3-bit palette indices, 7 = no node
    const uint32_t all = f & (f >> 1) & (f >> 2) & 0x249249u;  // bit 3v set iff field v == 7
    return 8 - __popc(all);
  }
  const uint2 w = __ldg(reinterpret_cast<const uint2*>(rec));
  return __popc(present_bits4(w.x) | (present_bits4(w.y) << 4));
}

template <bool SCATTER>
__global__ void __launch_bounds__(256)
cell_bin_kernel(const uint8_t* __restrict__ cells, int64_t n_cells, int64_t stride, int np, int4 offs, int compact,
                uint32_t* __restrict__ counters /* [kBinMax] counts, then [kBinMax] cursors */, uint32_t* __restrict__ order) {
  __shared__ uint32_t hist[kBinMax];
  __shared__ uint32_t base[kBinMax];
  if (threadIdx.x < kBinMax) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int key = -1;
  uint32_t slot = 0;
  if (g < n_cells) {
    const uint8_t* rec = cells + g * stride;
    key = cell_nodes(rec + offs.x, compact);
    if (np > 1) key += cell_nodes(rec + offs.y, compact);
    if (np > 2) key += cell_nodes(rec + offs.z, compact);
    key = 8 * np - key;                                        // largest cells first
    slot = atomicAdd(&hist[key], 1u);
  }
  __syncthreads();
  if (!SCATTER) {
    if (threadIdx.x < kBinMax && hist[threadIdx.x]) atomicAdd(&counters[threadIdx.x], hist[threadIdx.x]);
    return;
  }
  if (threadIdx.x < kBinMax) {
    // start of my bin = counts of the bins before it (a handful of values: every CTA adds them up itself)
    uint32_t start = 0;
    for (int b = 0; b < (int)threadIdx.x; ++b) start += counters[b];
    base[threadIdx.x] = hist[threadIdx.x] ? start + atomicAdd(&counters[kBinMax + threadIdx.x], hist[threadIdx.x]) : 0u;
  }
  __syncthreads();
  if (key >= 0) order[base[key] + slot] = (uint32_t)g;
}

template <int NP, bool WEIGHTED, bool TOPK, bool OA = false>
__global__ void __launch_bounds__(kCellThreads, OA ? 6 : 0)
score_cell_kernel(const __grid_constant__ PackedParams P, const uint8_t* __restrict__ cells, int64_t n_cells,
                  float* __restrict__ scores, double* __restrict__ mu64, double* __restrict__ var64,
                  double* __restrict__ score64, NbwCand* __restrict__ block_cands, int k_top, int distinct,
                  int64_t index_base, const uint32_t* __restrict__ order) {
  constexpr unsigned FULL = 0xffffffffu;
  using KccT = typename std::conditional<WEIGHTED, double, int>::type;
  static_assert(!OA || NP == 1, "the packed histogram-intersection path takes one kernel");
  // histogram intersection: the exception items of this thread's cell (nbw_model.oa_exc), at most one per node and
  // level.  Dynamically indexed, so the list lives in local memory (L1-resident, interleaved across the warp):
  // shared memory would cost the resident CTAs these latency-bound gathers need.
  int exs[OA ? 7 * NBW_MAX_LEVELS : 1];
  __shared__ uint4 ent[8][kCellThreads];     // per node: {ha_lo, ha_hi, hb, class count field}
  __shared__ float sh_tv[TOPK ? kTopMax : 1][kCellThreads];
  __shared__ int sh_ti[TOPK ? kTopMax : 1][kCellThreads];
  const int tid = threadIdx.x;
  const int zero_slot = P.zero_slot;
  const double* __restrict__ H = P.H;
  const int64_t ldh = P.ld_h;
  const int64_t stride = P.rec_stride;
  if constexpr (TOPK) {
#pragma unroll
    for (int j = 0; j < kTopMax; ++j) { sh_tv[j][tid] = 0.f; sh_ti[j][tid] = -1; }
  }

  for (int64_t gi = (int64_t)blockIdx.x * kCellThreads + tid; gi < n_cells; gi += (int64_t)gridDim.x * kCellThreads) {
    const int64_t g = order ? (int64_t)__ldg(order + gi) : gi;      // size-binned walk (cell_bin_kernel) or pool order
    int deep[NP][8];
    int ne = 0;                  // OA: exception items collected so far
    KccT kacc = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const PartParams& Q = P.part[p];
      uint64_t labw, succw;
      if (Q.compact) {
        // 8-byte wire record (nbw_model.compact): 3-bit palette indices + the strict upper triangle of the adjacency
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(cells + g * stride + Q.rec_offset));
        const uint64_t w = ((uint64_t)r.y << 32) | r.x;
        uint64_t pal = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) pal |= (uint64_t)(c == 7 ? 0xFFu : Q.palette[c]) << (8 * c);
        labw = 0;
        succw = 0;
        uint32_t tri = (uint32_t)(w >> 24);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const uint32_t c = (uint32_t)(w >> (3 * v)) & 7u;
          labw |= ((pal >> (8 * c)) & 0xFFull) << (8 * v);
          if (v < 7) {
            const uint32_t row = (tri & ((1u << (7 - v)) - 1u)) << (v + 1);
            succw |= (uint64_t)row << (8 * v);
            tri >>= (7 - v);
          }
        }
      } else {
        const uint4 rec = __ldg(reinterpret_cast<const uint4*>(cells + g * stride + Q.rec_offset));
        labw = ((uint64_t)rec.y << 32) | rec.x;
        succw = ((uint64_t)rec.w << 32) | rec.z;
      }
      const uint32_t pm = present_bits4((uint32_t)labw) | (present_bits4((uint32_t)(labw >> 32)) << 4);   // present nodes
      // adjacency rows, restricted to present nodes on both ends
      {
        uint64_t rowmask = 0;                       // byte v = 0xFF for present v
#pragma unroll
        for (int v = 0; v < 8; ++v) rowmask |= ((pm >> v) & 1u) ? (0xFFull << (8 * v)) : 0ull;
        succw &= rowmask & (0x0101010101010101ull * pm);
      }
      if (Q.undirected) succw |= transpose8x8(succw);   // kernels/weisfilerlehman.py:127-128
      // nodes that are an endpoint of no edge are not relabelled
      uint32_t tgt;
      {
        uint64_t x = succw | (succw >> 32);
        x |= x >> 16;
        x |= x >> 8;
        tgt = (uint32_t)x & 0xFFu;
      }
      const uint32_t em = pm & (tgt | nonzero_bits4((uint32_t)succw) | (nonzero_bits4((uint32_t)(succw >> 32)) << 4));

      // ---- level 0: op code -> dense id; classes = equal op codes
      uint32_t dense[8];
      uint32_t im = 0;           // nodes whose current label is in the dictionary
      uint64_t clsw = 0;         // byte v = nodes that share v's label at the current level
      int* dp = deep[p];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const uint32_t lab = (uint32_t)(labw >> (8 * v)) & 0xFFu;
        const bool present = (pm >> v) & 1u;
        const uint32_t d0 = present ? (uint32_t)__ldg(Q.level0_table + lab) : 0xFFFFu;
        const bool seen = present && d0 != 0xFFFFu;
        dense[v] = seen ? d0 : 0u;
        im |= seen ? (1u << v) : 0u;
        dp[v] = seen ? Q.label_base + Q.loff[0] + (int)d0 : zero_slot;
        if (present) clsw |= (uint64_t)(eq_bits8(labw, lab) & pm) << (8 * v);
      }
      // self-similarity: sum over labels of count^2 (linear) or of min(count, count) = count (intersection)
      const int cnt0 = OA ? __popc(pm) : __popcll(clsw);
      KccT kss = WEIGHTED ? (KccT)(Q.lw[0] * cnt0) : (KccT)cnt0;
      // exception items of the current level: nodes in the dictionary that are not the first of their class
      // a node of rank 2 at level 0 keeps rank 2 wherever it has an exception: it carries ONE chain item, named after
      // its label at the last such level (nbw_model.oa_d2_base)
      uint32_t twins = 0;
      int twin_item[8];
      auto emit_exceptions = [&](int l) {
#pragma unroll
        for (int v = 1; v < 8; ++v) {
          const int r0 = __popc((uint32_t)(clsw >> (8 * v)) & ((1u << v) - 1u));
          if (((im >> v) & 1u) && r0 > 0) {
            if (l == 0 && r0 == 1 && Q.oa_d2_base != 0) twins |= 1u << v;
            if ((twins >> v) & 1u) {
              twin_item[v] = Q.oa_d2_base + Q.loff[l] + (int)dense[v];
            } else {
              exs[ne] = oa_item(__ldg(Q.oa_exc + Q.loff[l] + (int)dense[v]), r0);
              ++ne;
            }
          }
        }
      };
      if constexpr (OA) emit_exceptions(0);

      // ---- levels 1..h
      const int h = Q.h, fs = Q.first_stable;
#pragma unroll 1
      for (int l = 1; l <= h; ++l) {
        const bool stable = fs != 0 && l >= fs;
        // what every node contributes to its predecessors
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          if ((pm >> v) & 1u) {
            const uint32_t c = (uint32_t)(clsw >> (8 * v)) & 0xFFu;
            const uint32_t low = c & (0u - c);
            const uint32_t sq = low * low;
            uint4 e = make_uint4(0u, 0u, 0u, sq * sq);        // one count in the 4-bit field of the class representative
            if (!stable && ((im >> v) & 1u)) {
              const uint4 hsh = __ldg(Q.id_hash[l] + dense[v]);
              e.x = hsh.x; e.y = hsh.y; e.z = hsh.z;
            }
            ent[v][tid] = e;
          }
        }
        uint32_t sig[8];
        uint32_t im_new = 0;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          sig[v] = 0u;
          if ((em >> v) & 1u) {
            const uint32_t sv = (uint32_t)(succw >> (8 * v)) & 0xFFu;
            uint64_t sum_a = 0;
            uint32_t sum_b = 0, sg = 0, m = sv;
            while (m) {
              const int j = __ffs(m) - 1;
              m &= m - 1u;
              const uint4 e = ent[j][tid];
              sum_a += ((uint64_t)e.y << 32) | e.x;
              sum_b += e.z;
              sg += e.w;
            }
            sig[v] = sg;
            // a credential can only be in the training dictionary if all its ids are
            const bool cand = ((im >> v) & 1u) && (sv & ~im) == 0u;
            if (cand) {
              bool seen = false;
              uint32_t dn = 0u;
              if (!stable) {
                const uint4 own = ent[v][tid];
                LevelSalt salt;
                salt.a_own = Q.a_own[l];
                salt.b_own = Q.b_own[l];
                const uint64_t key = nbw_key_from(((uint64_t)own.y << 32) | own.x, sum_a, salt);
                const uint32_t chk = nbw_chk_from(own.z, sum_b, salt);
                const uint32_t tag = (uint32_t)key, khi = (uint32_t)(key >> 32);
                const uint32_t b = khi & Q.bucket_mask[l];
                const uint4 t4 = __ldg(Q.tags[l] + b);
                const int idx = t4.x == tag ? 0 : t4.y == tag ? 1 : t4.z == tag ? 2 : t4.w == tag ? 3 : -1;
                if (idx >= 0) {
                  const uint4 en = __ldg(Q.entries[l] + 4 * (size_t)b + idx);
                  if (en.x == khi && en.y == chk) { seen = true; dn = en.z; }
                }
              } else {
                seen = true;
                dn = (uint32_t)__ldg(Q.child + Q.loff[l - 1] + (int)dense[v]);
              }
              if (seen) {
                im_new |= 1u << v;
                dense[v] = dn;
                dp[v] = Q.label_base + Q.loff[l] + (int)dn;
              }
            }
          }
        }
        // classes refine exactly: same previous class, both relabelled, same successor-class multiset
        uint64_t cls_new = 0;
        {
          // only cells that still have a class of two or more nodes need the pairwise comparison
          const uint32_t lo = (uint32_t)clsw, hi = (uint32_t)(clsw >> 32);
          const bool multi = ((lo & __vsubus4(lo, 0x01010101u)) | (hi & __vsubus4(hi, 0x01010101u))) != 0u;
          if (multi) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              if ((em >> v) & 1u) {
                const uint32_t c = (uint32_t)(clsw >> (8 * v)) & em;
                uint32_t nc = 1u << v;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  if (u != v) nc |= (((c >> u) & 1u) && sig[u] == sig[v]) ? (1u << u) : 0u;
                cls_new |= (uint64_t)nc << (8 * v);
              }
            }
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) cls_new |= ((em >> v) & 1u) ? (1ull << (9 * v)) : 0ull;
          }
        }
        const bool changed = cls_new != clsw || im_new != im;
        clsw = cls_new;
        im = im_new;
        const int cnt = OA ? __popc(em) : __popcll(clsw);
        kss += WEIGHTED ? (KccT)(Q.lw[l] * cnt) : (KccT)cnt;
        if constexpr (OA) emit_exceptions(l);
        if (stable && l < h && !changed) {
          // this cell reproduced (classes, membership) at a stable level: every deeper level will too; only the
          // labels move down the child chain (nbw_model.final_label)
          kss += WEIGHTED ? (KccT)(Q.wsuf[l] * cnt) : (KccT)((h - l) * cnt);
          if constexpr (OA) {
            // ... and a node that is not the first of its class keeps its rank: one exception item per deeper level
#pragma unroll
            for (int v = 1; v < 8; ++v) {
              const int r0 = __popc((uint32_t)(clsw >> (8 * v)) & ((1u << v) - 1u));
              if (((im >> v) & 1u) && r0 > 0) {
                if ((twins >> v) & 1u) {
                  twin_item[v] = Q.oa_d2_base + __ldg(Q.final_label + Q.loff[l] + (int)dense[v]) - Q.label_base;
                } else {
                  uint32_t dn = dense[v];
                  for (int l2 = l + 1; l2 <= h; ++l2) {
                    dn = (uint32_t)__ldg(Q.child + Q.loff[l2 - 1] + (int)dn);
                    exs[ne] = oa_item(__ldg(Q.oa_exc + Q.loff[l2] + (int)dn), r0);
                    ++ne;
                  }
                }
              }
            }
          }
#pragma unroll
          for (int v = 0; v < 8; ++v)
            if ((im >> v) & 1u) dp[v] = __ldg(Q.final_label + Q.loff[l] + (int)dense[v]);
          break;
        }
      }
      if constexpr (OA) {
#pragma unroll
        for (int v = 1; v < 8; ++v)
          if ((twins >> v) & 1u) {
            exs[ne] = twin_item[v];
            ++ne;
          }
      }
      kacc += kss;
    }

    // ---------------------------------------------------------------- posterior (prefix form), all in registers
    double mu_acc = 0.0, q_acc = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int du = deep[p][u];
        if (du != zero_slot) {
          mu_acc += __ldg(P.w_mu + du);
          const double* __restrict__ hrow = H + (int64_t)du * ldh;
#pragma unroll
          for (int p2 = p; p2 < NP; ++p2)
#pragma unroll
            for (int v = (p2 == p ? u : 0); v < 8; ++v) {
              const int dv = deep[p2][v];
              if (dv != zero_slot) {
                const double gv = __ldg(hrow + dv);
                q_acc += (p2 == p && v == u) ? gv : 2.0 * gv;
              }
            }
        }
      }
    if constexpr (OA) {
      for (int a = 0; a < ne; ++a) {
        const int xa = exs[a];
        mu_acc += __ldg(P.w_mu + xa);
        const double* __restrict__ hrow = H + (int64_t)xa * ldh;
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int du = deep[0][u];
          if (du != zero_slot) acc += __ldg(hrow + du);
        }
        for (int b = a + 1; b < ne; ++b) acc += __ldg(hrow + exs[b]);
        q_acc += 2.0 * acc + __ldg(hrow + xa);
      }
    }
    double mu, var;
    const double sc = finish_score(P, mu_acc, q_acc, (double)kacc, mu, var);
    const float scf = (float)sc;
    if (scores) scores[g] = scf;
    if (mu64) mu64[g] = mu;
    if (var64) var64[g] = var;
    if (score64) score64[g] = sc;
    if constexpr (TOPK) if (scf == scf) {
      const int gl = (int)g;
      bool skip = false;
      if (distinct) {
#pragma unroll
        for (int j = 0; j < kTopMax; ++j) {
          const int tj = sh_ti[j][tid];
          if (tj >= 0 && sh_tv[j][tid] == scf) { sh_ti[j][tid] = gl < tj ? gl : tj; skip = true; }
        }
      }
      if (!skip && cand_better(scf, gl, sh_tv[k_top - 1][tid], sh_ti[k_top - 1][tid])) {
        float cv = scf;
        int ci = gl;
#pragma unroll
        for (int j = 0; j < kTopMax; ++j) {
          if (j < k_top && ci >= 0) {
            const float ov = sh_tv[j][tid];
            const int oi = sh_ti[j][tid];
            if (cand_better(cv, ci, ov, oi)) {
              sh_tv[j][tid] = cv; sh_ti[j][tid] = ci;
              cv = ov; ci = oi;
            }
          }
        }
      }
    }
  }

  if constexpr (TOPK) {
    // CTA tournament over the threads' lists (see score_packed_kernel)
    __shared__ float sv[kCellThreads / 32];
    __shared__ int si[kCellThreads / 32];
    __shared__ float best_v_s;
    __shared__ int best_i_s;
    const int lane = tid & 31, warp = tid >> 5;
    float last_v = 0.f;
    int last_i = -1;
    for (int round = 0; round < k_top; ++round) {
      float bv = 0.f;
      int bi = -1;
#pragma unroll
      for (int j = 0; j < kTopMax; ++j) {
        const float tvj = sh_tv[j][tid];
        const int tij = sh_ti[j][tid];
        bool eligible = j < k_top && tij >= 0;
        if (eligible && last_i >= 0)
          eligible = distinct ? (tvj < last_v) : (tvj < last_v || (tvj == last_v && tij > last_i));
        if (eligible && cand_better(tvj, tij, bv, bi)) { bv = tvj; bi = tij; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        bv = lane < kCellThreads / 32 ? sv[lane] : 0.f;
        bi = lane < kCellThreads / 32 ? si[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(FULL, bv, o);
          const int oi = __shfl_xor_sync(FULL, bi, o);
          if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
          best_v_s = bv;
          best_i_s = bi;
          NbwCand c;
          c.v = bv;
          c.pad = 0;
          c.idx = bi >= 0 ? index_base + bi : -1;
          block_cands[(int64_t)blockIdx.x * k_top + round] = c;
        }
      }
      __syncthreads();
      last_v = best_v_s;
      last_i = best_i_s;
      __syncthreads();
      if (last_i < 0) {
        for (int r = round + 1 + tid; r < k_top; r += kCellThreads) {
          NbwCand c;
          c.v = 0.f;
          c.pad = 0;
          c.idx = -1;
          block_cands[(int64_t)blockIdx.x * k_top + r] = c;
        }
        break;
      }
    }
  }
}

// ------------------------------------------------------------------ dictionary forms of the packed path
__global__ void id_hash_kernel(const uint64_t* __restrict__ uniq_prev, int64_t n_ids, LevelSalt salt,
                               uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ids) return;
  const uint64_t k = uniq_prev[i];
  const uint64_t ha = nbw_hash_a(k, salt);
  const uint32_t hb = nbw_hash_b(k, salt);
  out[i] = make_uint4((uint32_t)ha, (uint32_t)(ha >> 32), hb, 0u);
}

__global__ void bucket_insert_kernel(const uint64_t* __restrict__ uniq_keys, const uint64_t* __restrict__ uniq_chks,
                                     int64_t n_unique, uint32_t* __restrict__ tags, uint4* __restrict__ entries,
                                     uint32_t mask, uint32_t* __restrict__ fill, uint32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_unique) return;
  const uint64_t key = uniq_keys[i];
  const uint32_t tag = (uint32_t)key, khi = (uint32_t)(key >> 32);
  const uint32_t b = khi & mask;
  const uint32_t pos = atomicAdd(fill + b, 1u);
  if (pos >= 4u) {
    atomicOr(flags, 1u);
    return;
  }
  tags[4 * (size_t)b + pos] = tag;
  entries[4 * (size_t)b + pos] = make_uint4(khi, (uint32_t)uniq_chks[i], (uint32_t)i, 0u);
}

// two keys of one bucket with the same tag (or a real tag 0 behind an empty slot) would make the first
// tag match ambiguous: report it, the caller rebuilds with more buckets
__global__ void bucket_check_kernel(const uint32_t* __restrict__ tags, const uint32_t* __restrict__ fill,
                                    int64_t n_buckets, uint32_t* __restrict__ flags) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buckets) return;
  const uint32_t f = fill[b] < 4u ? fill[b] : 4u;
  bool dup = false;
  for (uint32_t i = 0; i < f; ++i)
    for (uint32_t j = i + 1; j < f; ++j) dup = dup || tags[4 * b + i] == tags[4 * b + j];
  if (dup) atomicOr(flags, 2u);
}

// ------------------------------------------------------------------ host side
int fill_part(nbw_ctx* ctx, const nbw_model* M, PartParams& Q) {
  NBW_REQUIRE(ctx, M->level0_table != nullptr, "model.level0_table is NULL");
  NBW_REQUIRE(ctx, M->h >= 0 && M->h < NBW_MAX_LEVELS, "model.h out of range");
  NBW_REQUIRE(ctx, M->first_stable == 0 || (M->first_stable >= 3 && M->first_stable <= M->h && M->child && M->final_label),
              "model.first_stable needs child and final_label tables");
  memset(&Q, 0, sizeof(Q));
  Q.level0_table = M->level0_table;
  Q.h = M->h;
  Q.first_stable = M->first_stable;
  Q.undirected = M->undirected;
  Q.rec_offset = M->rec_offset;
  Q.label_base = M->label_base;
  Q.compact = M->compact;
  memcpy(Q.palette, M->palette, 16);
  Q.child = M->child;
  Q.final_label = M->final_label;
  Q.oa_exc = M->oa_exc;
  Q.oa_d2_base = M->oa_d2_base;
  double suffix = 0.0;
  for (int l = M->h; l >= 0; --l) {
    Q.lw[l] = M->weighted ? M->level_weight[l] : 1.0;
    Q.wsuf[l] = suffix;
    suffix += Q.lw[l];
  }
  for (int l = 0; l <= M->h; ++l) {
    Q.loff[l] = M->label_off[l];
    if (l == 0) continue;
    const LevelSalt s = nbw_level_salt(M->seed, l);
    Q.a_own[l] = s.a_own;
    Q.b_own[l] = s.b_own;
    const bool hashed = M->first_stable == 0 || l < M->first_stable;
    if (hashed) {
      NBW_REQUIRE(ctx, M->id_hash[l] && M->bucket_tags[l] && M->bucket_entries[l], "model level %d lacks the packed-path tables", l);
      NBW_REQUIRE(ctx, ((M->bucket_mask[l] + 1) & M->bucket_mask[l]) == 0, "bucket count must be a power of two");
    }
    Q.id_hash[l] = static_cast<const uint4*>(M->id_hash[l]);
    Q.tags[l] = static_cast<const uint4*>(M->bucket_tags[l]);
    Q.entries[l] = static_cast<const uint4*>(M->bucket_entries[l]);
    Q.bucket_mask[l] = M->bucket_mask[l];
  }
  return NBW_OK;
}

template <int NMAX, int NP, bool WEIGHTED, bool TOPK, bool OA = false, bool WIRE = false>
int launch_variant(nbw_ctx* ctx, int variant, const PackedParams& P, const uint8_t* cells, int64_t n, float* scores,
                   double* mu64, double* var64, double* score64, int k, int distinct, int64_t index_base,
                   NbwCand* out_cands, cudaStream_t st) {
  auto kern = score_packed_kernel<NMAX, NP, WEIGHTED, TOPK, OA, WIRE>;
  if (!ctx->packed_occ[variant]) {
    int occ = 0;
    NBW_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPackThreads, 0));
    NBW_REQUIRE(ctx, occ >= 1, "score_packed_kernel does not fit this device");
    ctx->packed_occ[variant] = occ;
  }
  // one resident wave of long-lived warps, each walking a contiguous share of the pool
  constexpr int kMinCellsPerWarp = NMAX == 8 ? 32 : 8;
  int64_t grid = (int64_t)ctx->sm_count * ctx->packed_occ[variant];
  const int64_t need = (n + (int64_t)kPackWarps * kMinCellsPerWarp - 1) / ((int64_t)kPackWarps * kMinCellsPerWarp);
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  const int64_t warps = grid * kPackWarps;
  const int64_t per_warp = (n + warps - 1) / warps;
  NbwCand* blocks = nullptr;
  if (TOPK) {
    // two slots, alternating: the merge of launch i may still read its lists while launch i + 1 writes
    const size_t slot_bytes = sizeof(NbwCand) * (size_t)ctx->sm_count * 32 * kTopMax;
    const size_t bytes = 2 * slot_bytes;
    if (ctx->topk_blocks_bytes < bytes) {
      if (ctx->topk_blocks) {
        NBW_CUDA(ctx, cudaStreamSynchronize(st));
        NBW_CUDA(ctx, cudaFree(ctx->topk_blocks));
        ctx->topk_blocks = nullptr;
        ctx->topk_blocks_bytes = 0;
      }
      NBW_CUDA(ctx, cudaMalloc(&ctx->topk_blocks, bytes));
      ctx->topk_blocks_bytes = bytes;
    }
    ctx->topk_slot ^= 1;
    blocks = reinterpret_cast<NbwCand*>(static_cast<char*>(ctx->topk_blocks) + (size_t)ctx->topk_slot * slot_bytes);
  }
  if (ctx->prof_ev[0]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[0], st));
  kern<<<(unsigned)grid, kPackThreads, 0, st>>>(P, cells, n, per_warp, scores, mu64, var64, score64, blocks, k, distinct,
                                               index_base);
  NBW_CHECK_LAUNCH(ctx);
  if (ctx->prof_ev[1]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[1], st));
  if (TOPK) {
    NbwCand* res = nullptr;
    int rc = nbw_topk_merge_enqueue(ctx, blocks, grid * k, k, distinct, &res, st, out_cands);
    if (rc) return rc;
    if (res != out_cands) NBW_CUDA(ctx, cudaMemcpyAsync(out_cands, res, sizeof(NbwCand) * k, cudaMemcpyDeviceToDevice, st));
  }
  return NBW_OK;
}

template <int NP, bool WEIGHTED, bool TOPK, bool OA = false>
int launch_cell_variant(nbw_ctx* ctx, int variant, const PackedParams& P, const uint8_t* cells, int64_t n, float* scores,
                        double* mu64, double* var64, double* score64, int k, int distinct, int64_t index_base,
                        NbwCand* out_cands, cudaStream_t st) {
  auto kern = score_cell_kernel<NP, WEIGHTED, TOPK, OA>;
  if (!ctx->packed_occ[variant]) {
    int occ = 0;
    NBW_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCellThreads, 0));
    NBW_REQUIRE(ctx, occ >= 1, "score_cell_kernel does not fit this device");
    ctx->packed_occ[variant] = occ;
  }
  int64_t grid = (int64_t)ctx->sm_count * ctx->packed_occ[variant];
  const int64_t need = (n + kCellThreads - 1) / kCellThreads;
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  NbwCand* blocks = nullptr;
  if (TOPK) {
    const size_t slot_bytes = sizeof(NbwCand) * (size_t)ctx->sm_count * 32 * kTopMax;
    const size_t bytes = 2 * slot_bytes;
    if (ctx->topk_blocks_bytes < bytes) {
      if (ctx->topk_blocks) {
        NBW_CUDA(ctx, cudaStreamSynchronize(st));
        NBW_CUDA(ctx, cudaFree(ctx->topk_blocks));
        ctx->topk_blocks = nullptr;
        ctx->topk_blocks_bytes = 0;
      }
      NBW_CUDA(ctx, cudaMalloc(&ctx->topk_blocks, bytes));
      ctx->topk_blocks_bytes = bytes;
    }
    ctx->topk_slot ^= 1;
    blocks = reinterpret_cast<NbwCand*>(static_cast<char*>(ctx->topk_blocks) + (size_t)ctx->topk_slot * slot_bytes);
    NBW_REQUIRE(ctx, (size_t)grid * kTopMax * sizeof(NbwCand) <= slot_bytes, "too many CTAs for the top-n block lists");
  }
  if (ctx->prof_ev[0]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[0], st));
  // large device-resident pools are walked in size-binned order (cell_bin_kernel); pools read over PCIe are not
  // (the pre-pass would fetch every record a second time)
  const uint32_t* order = nullptr;
  if (n >= kBinMinCells && !getenv("NBW_NO_BINNING")) {
    cudaPointerAttributes pa;
    const bool on_device = cudaPointerGetAttributes(&pa, cells) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
    if (!on_device) (void)cudaGetLastError();
    if (on_device) {
      const size_t need = 512 + sizeof(uint32_t) * (size_t)n;
      if (ctx->bin_order_bytes < need) {
        if (ctx->bin_order) {
          NBW_CUDA(ctx, cudaStreamSynchronize(st));
          NBW_CUDA(ctx, cudaFree(ctx->bin_order));
          ctx->bin_order = nullptr;
          ctx->bin_order_bytes = 0;
        }
        NBW_CUDA(ctx, cudaMalloc(&ctx->bin_order, need + need / 4));
        ctx->bin_order_bytes = need + need / 4;
      }
      uint32_t* counters = static_cast<uint32_t*>(ctx->bin_order);
      uint32_t* ord = counters + 128;
      NBW_CUDA(ctx, cudaMemsetAsync(counters, 0, 512, st));
      const int4 offs = make_int4(P.part[0].rec_offset, NP > 1 ? P.part[1].rec_offset : 0, NP > 2 ? P.part[2].rec_offset : 0, 0);
      const unsigned bgrid = (unsigned)((n + 255) / 256);
      cell_bin_kernel<false><<<bgrid, 256, 0, st>>>(cells, n, P.rec_stride, NP, offs, P.part[0].compact, counters, ord);
      NBW_CHECK_LAUNCH(ctx);
      cell_bin_kernel<true><<<bgrid, 256, 0, st>>>(cells, n, P.rec_stride, NP, offs, P.part[0].compact, counters, ord);
      NBW_CHECK_LAUNCH(ctx);
      order = ord;
    }
  }
  kern<<<(unsigned)grid, kCellThreads, 0, st>>>(P, cells, n, scores, mu64, var64, score64, blocks, k, distinct, index_base, order);
  NBW_CHECK_LAUNCH(ctx);
  if (ctx->prof_ev[1]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[1], st));
  if (TOPK) {
    NbwCand* res = nullptr;
    int rc = nbw_topk_merge_enqueue(ctx, blocks, grid * k, k, distinct, &res, st, out_cands);
    if (rc) return rc;
    if (res != out_cands) NBW_CUDA(ctx, cudaMemcpyAsync(out_cands, res, sizeof(NbwCand) * k, cudaMemcpyDeviceToDevice, st));
  }
  return NBW_OK;
}

template <bool TOPK>
int launch_cell_parts(nbw_ctx* ctx, int np, bool weighted, const PackedParams& P, const uint8_t* cells, int64_t n,
                      float* scores, double* mu64, double* var64, double* score64, int k, int distinct, int64_t index_base,
                      NbwCand* out_cands, cudaStream_t st) {
  const int base = 20 + (TOPK ? 4 : 0);
#define NBW_ARGS P, cells, n, scores, mu64, var64, score64, k, distinct, index_base, out_cands, st
  if (np == 1 && !weighted) return launch_cell_variant<1, false, TOPK>(ctx, base + 0, NBW_ARGS);
  if (np == 1) return launch_cell_variant<1, true, TOPK>(ctx, base + 1, NBW_ARGS);
  if (np == 2) return launch_cell_variant<2, true, TOPK>(ctx, base + 2, NBW_ARGS);
  if (np == 3) return launch_cell_variant<3, true, TOPK>(ctx, base + 3, NBW_ARGS);
#undef NBW_ARGS
  return -1;     // four parts: the lane-per-node kernel
}

template <int NMAX, bool TOPK, bool WIRE = false>
int launch_parts(nbw_ctx* ctx, int np, bool weighted, const PackedParams& P, const uint8_t* cells, int64_t n,
                 float* scores, double* mu64, double* var64, double* score64, int k, int distinct, int64_t index_base,
                 NbwCand* out_cands, cudaStream_t st) {
  // occupancy cache slots: 0..15 8-node, 16..31 16-node, 44.. 16-node wire records
  const int base = (WIRE ? 44 : (NMAX == 16 ? 16 : 0)) + (TOPK ? 8 : 0);
#define NBW_ARGS P, cells, n, scores, mu64, var64, score64, k, distinct, index_base, out_cands, st
  if (np == 1 && !weighted) return launch_variant<NMAX, 1, false, TOPK, false, WIRE>(ctx, base + 0, NBW_ARGS);
  if (np == 1) return launch_variant<NMAX, 1, true, TOPK, false, WIRE>(ctx, base + 1, NBW_ARGS);
  if (np == 2) return launch_variant<NMAX, 2, true, TOPK, false, WIRE>(ctx, base + 2, NBW_ARGS);
  if (np == 3) return launch_variant<NMAX, 3, true, TOPK, false, WIRE>(ctx, base + 3, NBW_ARGS);
  if (np == 4) return launch_variant<NMAX, 4, true, TOPK, false, WIRE>(ctx, base + 4, NBW_ARGS);
#undef NBW_ARGS
  NBW_FAIL(ctx, NBW_ERR_UNSUPPORTED, "the fused pool path takes sums of up to %d kernels", NBW_MAX_PARTS);
}

}  // namespace

// Is this model scored by the packed kernel?  (Linear kernels in prefix form with the packed-path tables.)
bool nbw_packed_applicable(const nbw_model* M) {
  if (!M || !M->H || !M->w_mu_prefix) return false;
  if (M->oa && (!M->oa_exc || M->next)) return false;   // histogram intersection: item form, single kernel
  if (getenv("NBW_SCORE_LEGACY")) return false;     // diagnostic switch: first-generation kernel (single part only)
  const bool hashed1 = M->h >= 1 && (M->first_stable == 0 || 1 < M->first_stable);
  return !hashed1 || M->bucket_tags[1] != nullptr;
}

int nbw_score_packed(nbw_ctx* ctx, const nbw_model* M, const void* cells, int64_t n_cells, float* scores, double* mu64,
                     double* var64, double* score64, int k, int distinct, int64_t index_base, void* out_cands,
                     cudaStream_t st) {
  PackedParams P;
  memset(&P, 0, sizeof(P));
  int np = 0;
  bool weighted = false;
  const int n_max = M->n_max;
  NBW_REQUIRE(ctx, n_max == 8 || n_max == 16, "model.n_max must be 8 or 16");
  NBW_REQUIRE(ctx, n_cells < ((int64_t)1 << 31), "at most 2^31 - 1 cells per launch");
  const int rec = n_max == 8 ? 16 : 48;
  for (const nbw_model* part = M; part; part = static_cast<const nbw_model*>(part->next)) {
    NBW_REQUIRE(ctx, np < NBW_MAX_PARTS, "the fused pool path takes sums of up to %d kernels", NBW_MAX_PARTS);
    NBW_REQUIRE(ctx, part->n_max == n_max && (!part->oa || (part == M && !part->next && part->oa_exc)),
                "parts of a sum must share n_max and be linear WL kernels");
    int rc = fill_part(ctx, part, P.part[np]);
    if (rc) return rc;
    weighted = weighted || part->weighted != 0;
    ++np;
  }
  weighted = weighted || np > 1;
  P.H = M->H;
  P.w_mu = M->w_mu_prefix;
  P.ld_h = M->ld_h;
  P.zero_slot = M->n_labels;
  P.rec_stride = M->rec_stride > 0 ? M->rec_stride : rec;
  NBW_REQUIRE(ctx, P.ld_h >= (int64_t)P.zero_slot + 1, "model.ld_h must be >= n_labels + 1");
  NBW_REQUIRE(ctx, P.rec_stride % (M->compact ? 8 : 16) == 0, "rec_stride must be a multiple of the record size");
  const bool compact = M->compact != 0;
  const int crec = n_max == 8 ? 8 : 24;      // wire records: 8 bytes (8-node cells) / 24 bytes (16-node cells)
  for (int p = 0; p < np; ++p) {
    NBW_REQUIRE(ctx, (P.part[p].compact != 0) == compact, "all parts of a sum must use the same record form");
    if (compact)
      NBW_REQUIRE(ctx, P.part[p].rec_offset % 8 == 0 && P.part[p].rec_offset + crec <= P.rec_stride,
                  "part %d: compact records are %d bytes", p, crec);
    else
      NBW_REQUIRE(ctx, P.part[p].rec_offset % 16 == 0 && P.part[p].rec_offset + rec <= P.rec_stride,
                  "part %d: rec_offset outside the candidate record", p);
  }
  NBW_REQUIRE(ctx, !compact || n_max == 16 || np <= 3, "8-byte compact records are scored by the thread-per-cell kernel (sums of up to 3 kernels)");
  NBW_REQUIRE(ctx, M->acq >= NBW_ACQ_EI && M->acq <= NBW_ACQ_MEAN, "unknown acquisition");
  P.acq = M->acq;
  P.lik = M->lik; P.y_mean = M->y_mean; P.y_std = M->y_std;
  P.incumbent = M->incumbent; P.xi = M->xi; P.beta = M->beta;
  const uint8_t* c = static_cast<const uint8_t*>(cells);
  NbwCand* oc = static_cast<NbwCand*>(out_cands);
  if (k > 0) NBW_REQUIRE(ctx, k <= kTopMax, "fused top-n handles k <= %d", kTopMax);
  if (M->oa) {
    // histogram intersection in item form (nbw_model.oa_exc): single kernel, unit weights; variants 32.. of the
    // occupancy cache
    NBW_REQUIRE(ctx, !weighted, "the packed histogram-intersection path takes unit level weights");
    NBW_REQUIRE(ctx, !(compact && n_max == 16), "24-byte wire records are not wired to the histogram-intersection kernels");
#define NBW_ARGS P, c, n_cells, scores, mu64, var64, score64, k, distinct, index_base, oc, st
#define NBW_ARGS0 P, c, n_cells, scores, mu64, var64, score64, 0, 0, 0, nullptr, st
    if (n_max == 8 && (compact || !getenv("NBW_SCORE_LANES"))) {
      if (k > 0) return launch_cell_variant<1, false, true, true>(ctx, 32, NBW_ARGS);
      return launch_cell_variant<1, false, false, true>(ctx, 33, NBW_ARGS0);
    }
    if (n_max == 8) {
      if (k > 0) return launch_variant<8, 1, false, true, true>(ctx, 34, NBW_ARGS);
      return launch_variant<8, 1, false, false, true>(ctx, 35, NBW_ARGS0);
    }
    if (k > 0) return launch_variant<16, 1, false, true, true>(ctx, 36, NBW_ARGS);
    return launch_variant<16, 1, false, false, true>(ctx, 37, NBW_ARGS0);
#undef NBW_ARGS
#undef NBW_ARGS0
  }
  if (n_max == 8 && np <= 3 && (compact || !getenv("NBW_SCORE_LANES"))) {
    // 8-node records: one thread per cell (NBW_SCORE_LANES=1 keeps the lane-per-node kernel for comparison)
    const int rc = k > 0 ? launch_cell_parts<true>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, k, distinct,
                                                   index_base, oc, st)
                         : launch_cell_parts<false>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, 0, 0, 0,
                                                    nullptr, st);
    if (rc >= 0) return rc;
  }
  if (k > 0) {
    if (n_max == 8) return launch_parts<8, true>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, k, distinct, index_base, oc, st);
    if (compact) return launch_parts<16, true, true>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, k, distinct, index_base, oc, st);
    return launch_parts<16, true>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, k, distinct, index_base, oc, st);
  }
  if (n_max == 8) return launch_parts<8, false>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, 0, 0, 0, nullptr, st);
  if (compact) return launch_parts<16, false, true>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, 0, 0, 0, nullptr, st);
  return launch_parts<16, false>(ctx, np, weighted, P, c, n_cells, scores, mu64, var64, score64, 0, 0, 0, nullptr, st);
}

extern "C" int nbw_model_build_id_hash(nbw_ctx* ctx, const uint64_t* uniq_keys_prev, int64_t n_ids, uint64_t seed,
                                       int level, void* out, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n_ids >= 0 && level >= 1 && level < NBW_MAX_LEVELS, "bad n_ids / level");
  if (n_ids == 0) return NBW_OK;
  NBW_REQUIRE(ctx, out != nullptr && uniq_keys_prev != nullptr, "null pointer");
  id_hash_kernel<<<(unsigned)((n_ids + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      uniq_keys_prev, n_ids, nbw_level_salt(seed, level), static_cast<uint4*>(out));
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_model_build_buckets(nbw_ctx* ctx, const uint64_t* uniq_keys, const uint64_t* uniq_chks,
                                       int64_t n_unique, void* tags, void* entries, int64_t n_buckets, uint32_t* flags,
                                       void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n_buckets > 0 && (n_buckets & (n_buckets - 1)) == 0 && n_buckets <= ((int64_t)1 << 28),
              "n_buckets must be a power of two");
  NBW_REQUIRE(ctx, n_unique >= 0 && tags && entries && flags, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = nbw_scratch_ensure(ctx, Carver::pad(sizeof(uint32_t) * (size_t)n_buckets) + 1024, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  uint32_t* fill = cv.take<uint32_t>(n_buckets);
  NBW_CUDA(ctx, cudaMemsetAsync(fill, 0, sizeof(uint32_t) * (size_t)n_buckets, st));
  NBW_CUDA(ctx, cudaMemsetAsync(tags, 0, 16 * (size_t)n_buckets, st));
  NBW_CUDA(ctx, cudaMemsetAsync(entries, 0xFF, 64 * (size_t)n_buckets, st));
  if (n_unique == 0) return NBW_OK;
  NBW_REQUIRE(ctx, uniq_keys && uniq_chks, "null pointer");
  bucket_insert_kernel<<<(unsigned)((n_unique + 255) / 256), 256, 0, st>>>(
      uniq_keys, uniq_chks, n_unique, static_cast<uint32_t*>(tags), static_cast<uint4*>(entries),
      (uint32_t)(n_buckets - 1), fill, flags);
  NBW_CHECK_LAUNCH(ctx);
  bucket_check_kernel<<<(unsigned)((n_buckets + 255) / 256), 256, 0, st>>>(static_cast<const uint32_t*>(tags), fill,
                                                                          n_buckets, flags);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_score_pool_topk(nbw_ctx* ctx, const nbw_model* model_h, const void* cells, int64_t n_cells,
                                   float* scores, int k, int distinct, int64_t index_base, void* out_cands,
                                   void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_score_pool_topk");
  NBW_REQUIRE(ctx, model_h != nullptr && out_cands != nullptr, "null pointer");
  NBW_REQUIRE(ctx, k >= 1 && n_cells >= 0, "bad k / n_cells");
  if (k > kTopMax || !nbw_packed_applicable(model_h))
    NBW_FAIL(ctx, NBW_ERR_UNSUPPORTED, "fused top-n needs the packed scoring path and k <= %d", kTopMax);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_cells == 0) {   // an empty shard contributes k empty entries (index -1)
    NBW_CUDA(ctx, cudaMemsetAsync(out_cands, 0xFF, sizeof(NbwCand) * k, st));
    return NBW_OK;
  }
  NBW_REQUIRE(ctx, cells != nullptr, "cells is NULL");
  return nbw_score_packed(ctx, model_h, cells, n_cells, scores, nullptr, nullptr, nullptr, k, distinct, index_base,
                          out_cands, st);
}
