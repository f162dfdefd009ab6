// Fused pool scoring: WL relabelling -> dictionary lookup -> GP posterior -> acquisition,
// one launch for the whole candidate pool.  Replaces, per candidate, GraphGP.predict
// (bayesopt/gp.py:240-283: a full WL refit on train ∪ {candidate} and an (n+1)^2 Gram) and
// GraphExpectedImprovement.eval / GraphUpperConfidentBound.eval
// (bayesopt/acquisitions.py:59-80,130-138).
//
// Structure used: the posterior of a candidate c depends only on its
// histogram restricted to the training vocabulary, phi_c, and on its own self-similarity
// k_cc over ALL its labels.  With B = diag(K_raw,ii^-1/2) Phi_train the cross-covariance is
// k_s = B phi_c / sqrt(k_cc), hence
//     mu_c  = phi_c . (B^T alpha) / sqrt(k_cc)
//     q_c   = k_s^T K_lam^-1 k_s = phi_c^T (B^T K_lam^-1 B) phi_c / k_cc = phi_c^T G phi_c / k_cc
// phi_c has at most n_nodes * (h+1) non-zeros, so q_c is a sum of <= (n_nodes (h+1))^2
// entries of the L2-resident D x D matrix G — no O(n D) or O(n^2) work per candidate.
//
// One sub-warp (n_max lanes) per cell, lane = node; successor labels travel by shuffles;
// the pool is read once (16 / 48 B per cell) and 4 B per cell are written.
#include "nbw_common.cuh"

#include <math.h>

namespace {

constexpr int kScoreThreads = 256;

__host__ __device__ __forceinline__ uint32_t slot_of(uint64_t key) { return (uint32_t)(key >> 17); }

// ---------------------------------------------------------------- lookup table build
__global__ void table_insert_kernel(const uint64_t* __restrict__ uniq_keys, const uint64_t* __restrict__ uniq_chks,
                                    int64_t n_unique, uint4* table, uint32_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_unique) return;
  const uint64_t key = uniq_keys[i];
  uint32_t slot = slot_of(key) & mask;
  for (uint32_t it = 0; it <= mask; ++it) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(table + slot);
    const unsigned long long old = atomicCAS(kp, (unsigned long long)NBW_INVALID_KEY, (unsigned long long)key);
    if (old == (unsigned long long)NBW_INVALID_KEY) {
      uint32_t* w = reinterpret_cast<uint32_t*>(table + slot);
      w[2] = (uint32_t)uniq_chks[i];
      w[3] = (uint32_t)i;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ bool table_probe(const uint4* __restrict__ table, uint32_t mask, uint64_t key,
                                            uint32_t chk, uint32_t& id) {
  uint32_t slot = slot_of(key) & mask;
  // first probe without a loop: at load factor <= 1/2 it almost always decides
  uint4 s = __ldg(table + slot);
  uint64_t k = ((uint64_t)s.y << 32) | s.x;
  if (k != key && k != NBW_INVALID_KEY) {
    for (uint32_t it = 0; it < mask; ++it) {
      slot = (slot + 1) & mask;
      s = __ldg(table + slot);
      k = ((uint64_t)s.y << 32) | s.x;
      if (k == key || k == NBW_INVALID_KEY) break;
    }
  }
  // dictionary keys are unique, so a key hit with a different check word is a label outside it
  id = s.w;
  return k == key && s.z == chk;
}

// Label classes inside a cell, tracked level by level.  WL labels refine: two nodes can share a
// level-l label only if they shared the level-(l-1) label, so after level 0 (one shuffle sweep over
// the op codes) each node compares itself only with the members of its previous class — usually
// none.  (__match_any_sync on the 64-bit ids gives the same answer but costs ~40 % of this kernel.)
// `cls` is a bitmask over the sub-warp's lanes (bit j = node j shares my label), 0 for uncounted nodes.
template <int NMAX>
__device__ __forceinline__ uint32_t classes_level0(uint32_t label, bool valid, int node, int sub_base) {
  uint32_t cls = 0;
  const uint32_t mine = valid ? label : (0x100u + (uint32_t)node);   // uncounted nodes match nobody
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    const uint32_t other = __shfl_sync(0xffffffffu, mine, sub_base + j);
    cls |= (valid && other == mine) ? (1u << j) : 0u;
  }
  return cls;
}
template <int NMAX>
__device__ __forceinline__ uint32_t classes_refine(uint32_t cls_prev, uint64_t id64, bool valid, int node, int lane,
                                                   int sub_base) {
  const uint64_t cid = valid ? id64 : (NBW_NOCLASS_BIT | (uint64_t)lane);
  uint32_t rem = valid ? (cls_prev & ~(1u << node)) : 0u;
  uint32_t cls = valid ? (1u << node) : 0u;
#pragma unroll 1
  for (int t = 0; t < NMAX; ++t) {
    if (!__any_sync(0xffffffffu, rem != 0u)) break;
    const bool has = rem != 0u;
    const int j = has ? (__ffs(rem) - 1) : node;
    const uint64_t other = shfl64(cid, sub_base + j);
    cls |= (has && other == cid) ? (1u << j) : 0u;
    rem &= rem - 1u;
  }
  return cls;
}

__device__ __forceinline__ int effective_column(const nbw_model& M, int l, bool seen, uint32_t dense, int rank) {
  if (!seen) return -1;
  if (!M.oa) return M.col_off[l] + (int)dense;
  const int sl = M.label_off[l] + (int)dense;
  const int mc = (int)__ldg(M.max_count + sl);
  return rank < mc ? M.col_off[l] + (int)__ldg(M.thermo_base + sl) + rank : -1;
}

// Posterior scaling + acquisition for one candidate (fp64): gp.py:272-280, acquisitions.py:68-77,136.
__device__ __forceinline__ void finish_candidate(const nbw_model& M, double mu_acc, double q_acc, int kcc_i, int64_t g,
                                                 float* __restrict__ scores, double* __restrict__ mu64,
                                                 double* __restrict__ var64, double* __restrict__ score64) {
  if (g < 0) return;
  const double inv_kcc = kcc_i > 0 ? 1.0 / (double)kcc_i : 0.0;   // a cell without nodes falls back to the prior
  const double mu_n = mu_acc * sqrt(inv_kcc);
  const double q = q_acc * inv_kcc;
  const double var_n = fmax(1.0 + M.lik - q, M.lik);    // gp.py:272-276 (diagonal)
  const double sd = sqrt(var_n) * M.y_std;              // gp.py:278-279
  const double var = sd * sd;                           // gp.py:280
  const double mu = mu_n * M.y_std + M.y_mean;          // gp.py:277
  double score;
  if (M.acq == NBW_ACQ_UCB) {
    score = mu + M.beta * sd;                           // acquisitions.py:136
  } else if (M.acq == NBW_ACQ_MEAN) {
    score = mu;
  } else {
    const double d = mu - M.incumbent - M.xi;           // acquisitions.py:71-74
    const double u = d / sd;
    const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
    const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
    score = sd * pdf + d * cdf;
    if (M.acq == NBW_ACQ_AEI) score *= 1.0 - sqrt(M.lik) / sqrt(M.lik + var);   // :75-77
  }
  scores[g] = (float)score;
  if (mu64) mu64[g] = mu;
  if (var64) var64[g] = var;
  if (score64) score64[g] = score;
}

// class-count signature of the stable levels: 4 bits per previous class (a node has fewer than NMAX successors)
template <int NMAX> struct StableSig;
template <> struct StableSig<8> { typedef uint32_t type; };
template <> struct StableSig<16> { typedef uint64_t type; };

struct ScoreParams {
  nbw_model M;
  LevelSalt salt[NBW_MAX_LEVELS];   // per-level hash salts, derived from M.seed on the host
};

// MODE 0: general path (G over (node, level) pairs; required for oa)
// MODE 1: features only (nbw_pool_features)
// MODE 2: prefix path (H over node pairs; linear kernel only)
template <int NMAX, int NL, int MODE>
__global__ void __launch_bounds__(kScoreThreads, MODE == 2 ? 5 : 3)
score_pool_kernel(const ScoreParams P, const uint8_t* __restrict__ cells, int64_t n_cells,
                  float* __restrict__ scores, double* __restrict__ mu64, double* __restrict__ var64,
                  double* __restrict__ score64, int8_t* __restrict__ phi, int64_t ld_phi,
                  int32_t* __restrict__ self_sim) {
  constexpr int GPW = 32 / NMAX;
  const nbw_model& M = P.M;
  const int lane = threadIdx.x & 31;
  const int node = lane % NMAX;
  const int sub = lane / NMAX;
  const int sub_base = sub * NMAX;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // staging area of the batched epilogue: 32 / GPW = NMAX rounds fill all 32 lanes
  double st_mu = 0.0, st_q = 0.0;
  int st_k = 1, staged = 0;
  int64_t st_g = -1;

  for (int64_t g0 = warp_id * GPW; g0 < n_cells; g0 += n_warps * GPW) {
    const int64_t g = g0 + sub;
    const bool in_range = g < n_cells;
    const uint32_t label = in_range ? CellTraits<NMAX>::label(cells, g, node) : 0xFFu;
    const bool present = label != 0xFFu;
    const uint32_t succ = (in_range && present) ? CellTraits<NMAX>::succ(cells, g, node) : 0u;
    const uint32_t incoming = subwarp_or<NMAX>(succ);
    const bool endpoint = present && (succ != 0u || ((incoming >> node) & 1u));

    constexpr bool FEATURES_ONLY = MODE == 1;
    constexpr bool PREFIX = MODE == 2;
    int col[NL];
    int cnt[NL];   // class size of the node's label at level l (FEATURES_ONLY / linear self-sim)
    int rk[NL];
    int kss = 0;
    int deep = M.n_labels;   // PREFIX: stacked index of the deepest dictionary label (zero slot if none)

    // ---- level 0: op code -> dense id (weisfeiler_lehman.py:430-441 in transform form)
    uint64_t id64;
    uint32_t cls;   // lanes of the sub-warp that share this node's label at the current level
    bool in_dict;   // label of the previous level is in the dictionary: only then can the next one be
    {
      const uint32_t d0 = present ? (uint32_t)__ldg(M.level0_table + label) : 0xFFFFu;
      const bool seen = present && d0 != 0xFFFFu;
      in_dict = seen;
      id64 = (uint64_t)label;      // the level-0 key of a label is its op code, in or out of the dictionary
      cls = classes_level0<NMAX>(label, present, node, sub_base);
      cnt[0] = __popc(cls);
      rk[0] = __popc(cls & ((1u << node) - 1u));
      if (PREFIX) { if (seen) deep = M.label_off[0] + (int)d0; }
      else col[0] = effective_column(M, 0, seen, d0, rk[0]);
      if (present) kss += M.oa ? 1 : cnt[0];
    }
    // ---- levels 1..h
    uint32_t dense_prev = 0;   // dictionary id at the previous level (meaningful while in_dict)
    bool settled = false;      // warp-uniform: a stable level changed neither classes nor membership
#pragma unroll
    for (int l = 1; l < NL; ++l) {
      if (l >= 3 && M.first_stable != 0 && l >= M.first_stable) {
        if (settled) {
          // Every stable level applies the same map to (classes, membership); it reproduced its input
          // for all cells of this warp, so it will again: only the labels move down the child chain.
          const bool seen = in_dict;
          const uint32_t dense = seen ? (uint32_t)__ldg(M.child + M.label_off[l - 1] + (int)dense_prev) : 0u;
          dense_prev = dense;
          cnt[l] = cnt[l - 1];
          rk[l] = rk[l - 1];
          if (PREFIX) { if (seen) deep = M.label_off[l] + (int)dense; }
          else col[l] = endpoint ? effective_column(M, l, seen, dense, rk[l]) : -1;
          if (endpoint) kss += M.oa ? 1 : cnt[l];
          continue;
        }
        const uint32_t cls_in = cls;
        const bool dict_in = in_dict;
        // Stable suffix (nbw_model.child): the training dictionary no longer refines, so membership
        // is inherited — this node and all its successors were in the dictionary one level up — and
        // the label is the only child of the previous one.  The classes inside the cell still refine
        // (labels outside the dictionary can): exactly, from the multiset of the successors'
        // previous classes, four bits of count per class.
        using Sig = typename StableSig<NMAX>::type;
        const uint32_t word = (in_dict ? 0x80000000u : 0u) | (cls ? (uint32_t)(__ffs(cls) - 1) : 0u);
        Sig sig = 0;
        bool all_in = true;
        uint32_t m = succ;
        constexpr int kUnrolled = NMAX == 8 ? 2 : 4;      // as in successor_sums: most nodes have few successors
#pragma unroll
        for (int t = 0; t < kUnrolled; ++t) {
          const bool has = m != 0u;
          const int j = has ? (__ffs(m) - 1) : node;
          const uint32_t w = __shfl_sync(0xffffffffu, word, sub_base + j);
          sig += has ? ((Sig)1 << (4 * (w & 31u))) : (Sig)0;
          all_in = all_in && (!has || (w >> 31));
          m &= m - 1u;
        }
#pragma unroll 1
        for (int t = kUnrolled; t < NMAX; ++t) {
          if (!__any_sync(0xffffffffu, m != 0u)) break;
          const bool has = m != 0u;
          const int j = has ? (__ffs(m) - 1) : node;
          const uint32_t w = __shfl_sync(0xffffffffu, word, sub_base + j);
          sig += has ? ((Sig)1 << (4 * (w & 31u))) : (Sig)0;
          all_in = all_in && (!has || (w >> 31));
          m &= m - 1u;
        }
        const bool seen = endpoint && in_dict && all_in;
        const uint32_t dense = seen ? (uint32_t)__ldg(M.child + M.label_off[l - 1] + (int)dense_prev) : 0u;
        in_dict = seen;
        dense_prev = dense;
        cls = classes_refine<NMAX>(cls, (uint64_t)sig, endpoint, node, lane, sub_base);
        settled = !__any_sync(0xffffffffu, cls != cls_in || in_dict != dict_in);
        cnt[l] = __popc(cls);
        rk[l] = __popc(cls & ((1u << node) - 1u));
        if (PREFIX) { if (seen) deep = M.label_off[l] + (int)dense; }
        else col[l] = endpoint ? effective_column(M, l, seen, dense, rk[l]) : -1;
        if (endpoint) kss += M.oa ? 1 : cnt[l];
        continue;
      }
      const LevelSalt salt = P.salt[l];
      const uint64_t ha = nbw_hash_a(id64, salt);
      const uint32_t hb = nbw_hash_b(id64, salt);
      uint64_t sum_a;
      uint32_t sum_b;
      successor_sums<NMAX>(succ, sub_base, node, ha, hb, sum_a, sum_b);
      const uint64_t key = nbw_key_from(ha, sum_a, salt);
      const uint32_t chk = nbw_chk_from(hb, sum_b, salt);
      uint32_t dense = 0;
      bool seen = false;
      if (endpoint && in_dict)
        seen = table_probe(static_cast<const uint4*>(M.table[l]), M.table_mask[l], key, chk, dense);
      in_dict = seen;
      dense_prev = dense;
      id64 = key;                  // labels are hashed through their keys (see wl_relabel.cu)
      cls = classes_refine<NMAX>(cls, id64, endpoint, node, lane, sub_base);
      cnt[l] = __popc(cls);
      rk[l] = __popc(cls & ((1u << node) - 1u));
      if (PREFIX) { if (endpoint && seen) deep = M.label_off[l] + (int)dense; }
      else col[l] = endpoint ? effective_column(M, l, seen, dense, rk[l]) : -1;
      if (endpoint) kss += M.oa ? 1 : cnt[l];
    }
    const int kcc_i = subwarp_sum<NMAX>(kss);

    if (FEATURES_ONLY) {
      if (in_range) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          if (col[l] >= 0) {
            if (M.oa) phi[g * ld_phi + col[l]] = 1;
            else if (rk[l] == 0) phi[g * ld_phi + col[l]] = (int8_t)cnt[l];
          }
        }
        if (node == 0 && self_sim) self_sim[g] = kcc_i;
      }
      continue;
    }

    double mu_acc = 0.0, q_acc = 0.0;
    if (PREFIX) {
      // ---- posterior, prefix form: one gather per node pair against H (see nbw_model.H)
      const int zero_slot = M.n_labels;
      const double* __restrict__ hrow = M.H + (int64_t)deep * M.ld_h;
      const bool own = deep != zero_slot;
      if (own) {
        mu_acc = __ldg(M.w_mu_prefix + deep);
        q_acc = __ldg(hrow + deep);
      }
#pragma unroll
      for (int r = 1; r <= NMAX / 2; ++r) {
        const int pd = __shfl_sync(0xffffffffu, deep, sub_base + ((node + r) % NMAX));
        if (own && pd != zero_slot) {
          const double g = __ldg(hrow + pd);
          q_acc += (r == NMAX / 2) ? g : 2.0 * g;
        }
      }
    } else {
    // ---- posterior: mean and quadratic form in feature space.
    // Labels outside the dictionary (and absent nodes) are mapped to the all-zero row/column
    // n_features of G and w_mu, so every gather below is unconditional.
    const int zero_col = M.n_features;
    const double* __restrict__ Gp = M.G;
    const int ldg = M.ld_g;
    int pcol[NL];
    const double* grow[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      pcol[l] = col[l] >= 0 ? col[l] : zero_col;
      grow[l] = Gp + (int64_t)pcol[l] * ldg;
      mu_acc += __ldg(M.w_mu + pcol[l]);
    }
    {
      // r = 0: the node with itself; G is symmetric, so off-diagonal level pairs count twice
      double part[NL];
#pragma unroll
      for (int la = 0; la < NL; ++la) part[la] = 0.0;
#pragma unroll
      for (int lb = 0; lb < NL; ++lb)
#pragma unroll
        for (int la = 0; la <= lb; ++la) {
          const double g = __ldg(grow[la] + pcol[lb]);
          part[la] += (la == lb) ? g : 2.0 * g;
        }
#pragma unroll
      for (int la = 0; la < NL; ++la) q_acc += part[la];
    }
#pragma unroll
    for (int r = 1; r <= NMAX / 2; ++r) {
      // ordered pairs (node, node + r): r < NMAX/2 also stands for (node + r, node)
      const int partner = sub_base + ((node + r) % NMAX);
      double part[NL];
#pragma unroll
      for (int la = 0; la < NL; ++la) part[la] = 0.0;
#pragma unroll
      for (int lb = 0; lb < NL; ++lb) {
        const int pc = __shfl_sync(0xffffffffu, pcol[lb], partner);
#pragma unroll
        for (int la = 0; la < NL; ++la) part[la] += __ldg(grow[la] + pc);
      }
      double sum = 0.0;
#pragma unroll
      for (int la = 0; la < NL; ++la) sum += part[la];
      q_acc += (r == NMAX / 2) ? sum : 2.0 * sum;
    }
    }
    mu_acc = subwarp_sum<NMAX>(mu_acc);
    q_acc = subwarp_sum<NMAX>(q_acc);

    // ---- stage (mu, q, k_cc) of this round's GPW cells on lanes [slot*GPW, slot*GPW + GPW):
    // the fp64 acquisition epilogue (sqrt, exp, erfc, divisions) then runs once per 32 cells
    // with every lane busy instead of once per round with GPW lanes busy.
    {
      const int src = (lane % GPW) * NMAX;   // any lane of sub-warp (lane % GPW) holds its totals
      const double s_mu = __shfl_sync(0xffffffffu, mu_acc, src);
      const double s_q = __shfl_sync(0xffffffffu, q_acc, src);
      const int s_k = __shfl_sync(0xffffffffu, kcc_i, src);
      if (lane / GPW == staged) {
        const int64_t gs = g0 + (lane % GPW);
        st_mu = s_mu;
        st_q = s_q;
        st_k = s_k;
        st_g = gs < n_cells ? gs : -1;
      }
    }
    if (++staged == NMAX) {
      finish_candidate(M, st_mu, st_q, st_k, st_g, scores, mu64, var64, score64);
      st_g = -1;
      staged = 0;
    }
  }
  if (MODE != 1 && staged > 0) finish_candidate(M, st_mu, st_q, st_k, st_g, scores, mu64, var64, score64);
}

template <int NMAX, int MODE>
int launch_levels(nbw_ctx* ctx, const nbw_model& M, const uint8_t* cells, int64_t n, float* scores,
                  double* mu64, double* var64, double* score64, int8_t* phi, int64_t ld_phi,
                  int32_t* self_sim, cudaStream_t st) {
  const int grid = nbw_grid_for(ctx, n, kScoreThreads / NMAX, 16);
  ScoreParams P;
  P.M = M;
  for (int l = 0; l < NBW_MAX_LEVELS; ++l) P.salt[l] = nbw_level_salt(M.seed, l);
#define NBW_LAUNCH_NL(NLV)                                                                         \
  case NLV:                                                                                         \
    score_pool_kernel<NMAX, NLV, MODE><<<grid, kScoreThreads, 0, st>>>(P, cells, n, scores, mu64, var64, \
                                                                    score64, phi, ld_phi, self_sim); \
    break;
  if (ctx->prof_ev[0]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[0], st));
  switch (M.h + 1) {
    NBW_LAUNCH_NL(1)
    NBW_LAUNCH_NL(2)
    NBW_LAUNCH_NL(3)
    NBW_LAUNCH_NL(4)
    NBW_LAUNCH_NL(5)
    NBW_LAUNCH_NL(6)
    NBW_LAUNCH_NL(7)
    NBW_LAUNCH_NL(8)
    default:
      NBW_FAIL(ctx, NBW_ERR_INVALID, "h out of range");
  }
#undef NBW_LAUNCH_NL
  NBW_CHECK_LAUNCH(ctx);
  if (ctx->prof_ev[1]) NBW_CUDA(ctx, cudaEventRecord(ctx->prof_ev[1], st));
  return NBW_OK;
}

int check_model(nbw_ctx* ctx, const nbw_model* M, bool need_posterior) {
  NBW_REQUIRE(ctx, M != nullptr, "model is NULL");
  NBW_REQUIRE(ctx, M->n_max == 8 || M->n_max == 16, "model.n_max must be 8 or 16");
  NBW_REQUIRE(ctx, M->h >= 0 && M->h < NBW_MAX_LEVELS, "model.h out of range");
  NBW_REQUIRE(ctx, M->level0_table != nullptr, "model.level0_table is NULL");
  for (int l = 1; l <= M->h; ++l) {
    NBW_REQUIRE(ctx, M->table[l] != nullptr, "model.table[%d] is NULL", l);
    NBW_REQUIRE(ctx, ((M->table_mask[l] + 1) & M->table_mask[l]) == 0, "table capacity must be a power of two");
  }
  NBW_REQUIRE(ctx, !M->oa || (M->thermo_base && M->max_count), "oa model needs thermo_base/max_count");
  NBW_REQUIRE(ctx, M->n_features > 0, "model.n_features must be positive");
  NBW_REQUIRE(ctx, M->first_stable == 0 || (M->first_stable >= 3 && M->first_stable <= M->h && M->child != nullptr),
              "model.first_stable must be 0 or in [3, h] with a child table");
  if (need_posterior && !nbw_packed_applicable(M) && !(M->H != nullptr && !M->oa)) {
    NBW_REQUIRE(ctx, M->G && M->w_mu && M->ld_g >= M->n_features + 1, "model.G / w_mu missing or ld_g < n_features + 1");
  }
  if (need_posterior) NBW_REQUIRE(ctx, M->acq >= NBW_ACQ_EI && M->acq <= NBW_ACQ_MEAN, "unknown acquisition");
  return NBW_OK;
}

// the first-generation kernel scores one kernel with unit level weights on plain records
bool legacy_ok(const nbw_model* M) {
  return M->next == nullptr && !M->weighted && !M->undirected && M->rec_offset == 0 && !M->compact &&
         (M->rec_stride == 0 || M->rec_stride == (M->n_max == 8 ? 16 : 48));
}

}  // namespace

extern "C" int nbw_model_build_table(nbw_ctx* ctx, const uint64_t* uniq_keys, const uint64_t* uniq_chks,
                                     int64_t n_unique, void* table, int64_t capacity, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, capacity > 0 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
  NBW_REQUIRE(ctx, n_unique >= 0 && capacity >= 2 * n_unique && capacity <= ((int64_t)1 << 31),
              "capacity must be >= 2 * n_unique");
  NBW_REQUIRE(ctx, table != nullptr, "table is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NBW_CUDA(ctx, cudaMemsetAsync(table, 0xFF, (size_t)capacity * 16, st));
  if (n_unique == 0) return NBW_OK;
  NBW_REQUIRE(ctx, uniq_keys && uniq_chks, "null pointer");
  table_insert_kernel<<<(unsigned)((n_unique + 255) / 256), 256, 0, st>>>(
      uniq_keys, uniq_chks, n_unique, static_cast<uint4*>(table), (uint32_t)(capacity - 1));
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_score_pool(nbw_ctx* ctx, const nbw_model* model_h, const void* cells, int64_t n_cells,
                              float* scores, double* mu64, double* var64, double* score64, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_score_pool");
  int rc = check_model(ctx, model_h, true);
  if (rc) return rc;
  NBW_REQUIRE(ctx, n_cells >= 0, "n_cells negative");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells && scores, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* c = static_cast<const uint8_t*>(cells);
  if (nbw_packed_applicable(model_h))
    return nbw_score_packed(ctx, model_h, cells, n_cells, scores, mu64, var64, score64, 0, 0, 0, nullptr, st);
  NBW_REQUIRE(ctx, legacy_ok(model_h), "sums of kernels / weighted levels / per-part records need the packed scoring path");
  const bool prefix = model_h->H != nullptr && !model_h->oa;
  if (prefix)
    NBW_REQUIRE(ctx, model_h->w_mu_prefix && model_h->n_labels > 0 && model_h->ld_h >= model_h->n_labels + 1,
                "model.H needs w_mu_prefix, n_labels and ld_h >= n_labels + 1");
  if (model_h->n_max == 8) {
    if (prefix) return launch_levels<8, 2>(ctx, *model_h, c, n_cells, scores, mu64, var64, score64, nullptr, 0, nullptr, st);
    return launch_levels<8, 0>(ctx, *model_h, c, n_cells, scores, mu64, var64, score64, nullptr, 0, nullptr, st);
  }
  if (prefix) return launch_levels<16, 2>(ctx, *model_h, c, n_cells, scores, mu64, var64, score64, nullptr, 0, nullptr, st);
  return launch_levels<16, 0>(ctx, *model_h, c, n_cells, scores, mu64, var64, score64, nullptr, 0, nullptr, st);
}

extern "C" int nbw_pool_features(nbw_ctx* ctx, const nbw_model* model_h, const void* cells, int64_t n_cells,
                                 int8_t* phi, int64_t ld_phi, int32_t* self_sim, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_pool_features");
  int rc = check_model(ctx, model_h, false);
  if (rc) return rc;
  NBW_REQUIRE(ctx, n_cells >= 0 && ld_phi >= model_h->n_features, "bad shape");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells && phi, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NBW_CUDA(ctx, cudaMemsetAsync(phi, 0, (size_t)n_cells * ld_phi, st));
  const uint8_t* c = static_cast<const uint8_t*>(cells);
  if (model_h->n_max == 8)
    return launch_levels<8, 1>(ctx, *model_h, c, n_cells, nullptr, nullptr, nullptr, nullptr, phi, ld_phi, self_sim, st);
  return launch_levels<16, 1>(ctx, *model_h, c, n_cells, nullptr, nullptr, nullptr, nullptr, phi, ld_phi, self_sim, st);
}

namespace {
int score_dispatch(nbw_ctx* ctx, const nbw_model& M, const uint8_t* dev, int64_t n, float* scores, cudaStream_t st,
                   int k = 0, int distinct = 0, int64_t index_base = 0, void* out_cands = nullptr) {
  if (nbw_packed_applicable(&M))
    return nbw_score_packed(ctx, &M, dev, n, scores, nullptr, nullptr, nullptr, k, distinct, index_base, out_cands, st);
  NBW_REQUIRE(ctx, k == 0, "fused top-n needs the packed scoring path");
  NBW_REQUIRE(ctx, legacy_ok(&M), "sums of kernels / weighted levels / per-part records need the packed scoring path");
  const bool prefix = M.H != nullptr && !M.oa;
  if (M.n_max == 8)
    return prefix ? launch_levels<8, 2>(ctx, M, dev, n, scores, nullptr, nullptr, nullptr, nullptr, 0, nullptr, st)
                  : launch_levels<8, 0>(ctx, M, dev, n, scores, nullptr, nullptr, nullptr, nullptr, 0, nullptr, st);
  return prefix ? launch_levels<16, 2>(ctx, M, dev, n, scores, nullptr, nullptr, nullptr, nullptr, 0, nullptr, st)
                : launch_levels<16, 0>(ctx, M, dev, n, scores, nullptr, nullptr, nullptr, nullptr, 0, nullptr, st);
}

int pipe_init(nbw_ctx* ctx) {
  if (ctx->pipe_ready) return NBW_OK;
  NBW_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  NBW_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    NBW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
    NBW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
    NBW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    ctx->stage[i] = nullptr;
  }
  NBW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming));
  ctx->out_pending = 0;
  ctx->stage_bytes = 0;
  ctx->pipe_ready = 1;
  return NBW_OK;
}

// The D2H copies of the host scores run on s_out BESIDE whatever the caller queues next on `stream`
// (normally the top-k selection); nbw_topk / nbw_topk_merge / nbw_wait_host_scores await them.  A new
// pipeline call first orders `stream` behind a still pending copy, because it rewrites scores_dev.
int pipe_begin(nbw_ctx* ctx, cudaStream_t st) {
  int rc = pipe_init(ctx);
  if (rc) return rc;
  if (ctx->out_pending) NBW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_out, 0));
  return NBW_OK;
}
int pipe_end(nbw_ctx* ctx, bool copied) {
  if (copied) {
    NBW_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->s_out));
    ctx->out_pending = 1;
  }
  return NBW_OK;
}
// device buffer for the per-piece top-n lists of the host-buffer entry points (64 pieces x 8 entries)
int piece_cands(nbw_ctx* ctx, NbwCand** out) {
  if (!ctx->piece_cands) NBW_CUDA(ctx, cudaMalloc(&ctx->piece_cands, sizeof(NbwCand) * 64 * 8));
  *out = static_cast<NbwCand*>(ctx->piece_cands);
  return NBW_OK;
}
int merge_pieces(nbw_ctx* ctx, NbwCand* lists, int pieces, int k, int distinct, void* out_cands, cudaStream_t st) {
  NbwCand* res = nullptr;
  int rc = nbw_topk_merge_enqueue(ctx, lists, (int64_t)pieces * k, k, distinct, &res, st, static_cast<NbwCand*>(out_cands));
  if (rc) return rc;
  if (res != out_cands) NBW_CUDA(ctx, cudaMemcpyAsync(out_cands, res, sizeof(NbwCand) * k, cudaMemcpyDeviceToDevice, st));
  return NBW_OK;
}
}  // namespace

static int score_pool_mapped_impl(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                  float* scores_h, float* scores_dev, int pieces, int k, int distinct,
                                  int64_t index_base, void* out_cands, void* stream, bool write_through = false);
static int score_pool_host_impl(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                float* scores_h, float* scores_dev, int chunks, int k, int distinct,
                                int64_t index_base, void* out_cands, void* stream);

extern "C" int nbw_score_pool_mapped(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                     float* scores_h, float* scores_dev, int pieces, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_score_pool_mapped");
  return score_pool_mapped_impl(ctx, model_h, cells_h, n_cells, scores_h, scores_dev, pieces, 0, 0, 0, nullptr, stream);
}
extern "C" int nbw_score_pool_host(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                   float* scores_h, float* scores_dev, int chunks, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_score_pool_host");
  return score_pool_host_impl(ctx, model_h, cells_h, n_cells, scores_h, scores_dev, chunks, 0, 0, 0, nullptr, stream);
}
extern "C" int nbw_propose_host(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                float* scores_h, float* scores_dev, int mapped, int pieces, int k, int distinct,
                                int64_t index_base, void* out_cands, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_propose_host");
  NBW_REQUIRE(ctx, k >= 1 && k <= 8 && out_cands != nullptr, "k must be in [1, 8] and out_cands non-NULL");
  if (n_cells == 0) {
    NBW_CUDA(ctx, cudaMemsetAsync(out_cands, 0xFF, sizeof(NbwCand) * k, static_cast<cudaStream_t>(stream)));
    return NBW_OK;
  }
  if (mapped) return score_pool_mapped_impl(ctx, model_h, cells_h, n_cells, scores_h, scores_dev, pieces, k, distinct, index_base, out_cands, stream, mapped == 2);
  return score_pool_host_impl(ctx, model_h, cells_h, n_cells, scores_h, scores_dev, pieces, k, distinct, index_base, out_cands, stream);
}

// Mapped form: the kernel reads pinned host records through the device alias of the host pointer.
// (Letting the kernel also store the scores into mapped host memory was measured and is slower than
// one D2H copy of the [M] f32 vector: 0.74 ms vs 0.60 ms per 1M candidates.)
static int score_pool_mapped_impl(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                  float* scores_h, float* scores_dev, int pieces, int k, int distinct,
                                  int64_t index_base, void* out_cands, void* stream, bool write_through) {
  int rc = check_model(ctx, model_h, true);
  if (rc) return rc;
  NBW_REQUIRE(ctx, n_cells >= 0 && pieces >= 1 && pieces <= 64, "bad n_cells / pieces");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells_h && scores_dev, "null pointer");
  cudaPointerAttributes attr;
  NBW_CUDA(ctx, cudaPointerGetAttributes(&attr, cells_h));
  NBW_REQUIRE(ctx, attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr,
              "nbw_score_pool_mapped needs pinned host memory (use nbw_score_pool_host for pageable buffers)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = pipe_begin(ctx, st);
  if (rc) return rc;
  const int64_t rec = model_h->rec_stride > 0 ? model_h->rec_stride : (model_h->n_max == 8 ? 16 : 48);
  const uint8_t* src = static_cast<const uint8_t*>(attr.devicePointer);
  const int64_t per = (n_cells + pieces - 1) / pieces;
  NbwCand* lists = nullptr;
  int used = 0;
  if (k > 0) {
    rc = piece_cands(ctx, &lists);
    if (rc) return rc;
  }
  // write_through: the kernel also stores its scores through the mapped alias of scores_h — no D2H copy behind the
  // kernel (with the thread-per-cell kernel the stores are one coalesced 128-byte line per warp, which PCIe takes
  // at full rate; the lane-per-node kernel of the first build scattered 4-byte stores and lost to the copy)
  float* sink = scores_dev;
  if (write_through) {
    NBW_REQUIRE(ctx, scores_h != nullptr, "write-through needs scores_h");
    cudaPointerAttributes oa;
    NBW_CUDA(ctx, cudaPointerGetAttributes(&oa, scores_h));
    NBW_REQUIRE(ctx, oa.type == cudaMemoryTypeHost && oa.devicePointer != nullptr, "write-through needs pinned scores_h");
    sink = static_cast<float*>(oa.devicePointer);
  }
  if (scores_h && !write_through) {
    NBW_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
    NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[0], 0));
  }
  for (int c = 0; c < pieces; ++c) {
    const int64_t lo = (int64_t)c * per, hi = (lo + per < n_cells) ? lo + per : n_cells;
    if (lo >= hi) break;
    // (a single piece writes its list straight into out_cands: no merge launch)
    rc = score_dispatch(ctx, *model_h, src + lo * rec, hi - lo, sink + lo, st, k, distinct, index_base + lo,
                        lists ? (pieces == 1 ? static_cast<NbwCand*>(out_cands) : lists + (size_t)c * k) : nullptr);
    if (rc) return rc;
    ++used;
    if (scores_h && !write_through) {
      NBW_CUDA(ctx, cudaEventRecord(ctx->ev_done[c & 1], st));
      NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[c & 1], 0));
      NBW_CUDA(ctx, cudaMemcpyAsync(scores_h + lo, scores_dev + lo, (size_t)(hi - lo) * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  if (k > 0 && pieces > 1) {
    rc = merge_pieces(ctx, lists, used, k, distinct, out_cands, st);
    if (rc) return rc;
  }
  return pipe_end(ctx, scores_h != nullptr && !write_through);
}

// Host-buffer form: the pool lives in host memory (pinned for full overlap).  Chunk i+1's H2D copy,
// chunk i's kernel and chunk i-1's D2H copy of the scores run on three streams.
static int score_pool_host_impl(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                                float* scores_h, float* scores_dev, int chunks, int k, int distinct,
                                int64_t index_base, void* out_cands, void* stream) {
  int rc = check_model(ctx, model_h, true);
  if (rc) return rc;
  NBW_REQUIRE(ctx, n_cells >= 0 && chunks >= 1 && chunks <= 64, "bad n_cells / chunks");
  if (n_cells == 0) return NBW_OK;
  NBW_REQUIRE(ctx, cells_h && scores_dev, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rec = model_h->rec_stride > 0 ? model_h->rec_stride : (model_h->n_max == 8 ? 16 : 48);
  rc = pipe_begin(ctx, st);
  if (rc) return rc;
  const int64_t per = (n_cells + chunks - 1) / chunks;
  const size_t need = (size_t)per * rec;
  if (need > ctx->stage_bytes) {
    NBW_CUDA(ctx, cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) {
      if (ctx->stage[i]) NBW_CUDA(ctx, cudaFree(ctx->stage[i]));
      NBW_CUDA(ctx, cudaMalloc(&ctx->stage[i], need));
    }
    ctx->stage_bytes = need;
  }
  // the side streams start after whatever the caller queued on `stream`
  NBW_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
  NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_done[0], 0));
  NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[0], 0));
  const uint8_t* src = static_cast<const uint8_t*>(cells_h);
  NbwCand* lists = nullptr;
  int used = 0;
  if (k > 0) {
    rc = piece_cands(ctx, &lists);
    if (rc) return rc;
  }
  for (int c = 0; c < chunks; ++c) {
    const int64_t lo = (int64_t)c * per, hi = (lo + per < n_cells) ? lo + per : n_cells;
    if (lo >= hi) break;
    const int b = c & 1;
    if (c >= 2) NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_free[b], 0));   // kernel of chunk c-2 has read it
    NBW_CUDA(ctx, cudaMemcpyAsync(ctx->stage[b], src + lo * rec, (size_t)(hi - lo) * rec, cudaMemcpyHostToDevice, ctx->s_in));
    NBW_CUDA(ctx, cudaEventRecord(ctx->ev_ready[b], ctx->s_in));
    NBW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_ready[b], 0));
    const uint8_t* dev = static_cast<const uint8_t*>(ctx->stage[b]);
    rc = score_dispatch(ctx, *model_h, dev, hi - lo, scores_dev + lo, st, k, distinct, index_base + lo,
                        lists ? lists + (size_t)c * k : nullptr);
    if (rc) return rc;
    ++used;
    NBW_CUDA(ctx, cudaEventRecord(ctx->ev_free[b], st));
    if (scores_h) {
      NBW_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
      NBW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[b], 0));
      NBW_CUDA(ctx, cudaMemcpyAsync(scores_h + lo, scores_dev + lo, (size_t)(hi - lo) * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  if (k > 0) {
    rc = merge_pieces(ctx, lists, used, k, distinct, out_cands, st);
    if (rc) return rc;
  }
  return pipe_end(ctx, scores_h != nullptr);
}
