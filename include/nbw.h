/*
 * nbw.h — C-ABI of the B200-native NAS-BOWL surrogate-scoring path (libnbw.so).
 *
 * The reference (xingchenwan/nasbowl) is pure Python and has no FFI; the entry points below
 * are what a maintainer would bind (ctypes) to replace, one for one, the
 * Python functions of the hot path.  Each entry cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns an nbw_status (0 = OK) and never throws; nbw_last_error() gives
 *     the message of the last failure on that context;
 *   - pointers are DEVICE pointers unless the name ends in _h (host);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and, unless stated
 *     ("syncs"), the call returns without waiting;
 *   - one nbw_ctx per GPU / rank, used from the thread that owns the CUDA context; the
 *     context owns a scratch arena that grows on demand (nbw_reserve pre-sizes it);
 *   - cells are packed records (nasbowl_b200/cells.py): n_max = 8 -> 16-byte record
 *     labels[8] u8 | succ[8] u8, n_max = 16 -> 48-byte record labels[16] u8 | succ[16] u16;
 *     label 0xFF = no node; succ[v] bit j = edge v->j (successor neighbourhood,
 *     grakel_replace/weisfeiler_lehman.py:266-267).
 */
#ifndef NBW_H_
#define NBW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbw_ctx nbw_ctx;

typedef enum {
  NBW_OK = 0,
  NBW_ERR_INVALID = 1,        /* bad argument (AssertionError / TypeError in the reference) */
  NBW_ERR_CUDA = 2,           /* CUDA runtime / driver failure */
  NBW_ERR_NOT_PD = 3,         /* Cholesky pivot <= 0 (RuntimeError in bayesopt/utils.py:127-137) */
  NBW_ERR_HASH_COLLISION = 4, /* two distinct WL credentials share a 64-bit key: re-run with another seed */
  NBW_ERR_CAPACITY = 5,       /* caller-provided output buffer too small */
  NBW_ERR_UNSUPPORTED = 6
} nbw_status;

#define NBW_MAX_LEVELS 8       /* WL depth h <= 7 */
#define NBW_INVALID_ID 0xFFFFFFFFu
#define NBW_INVALID_KEY 0xFFFFFFFFFFFFFFFFull

/* Lifecycle and diagnostics: no counterpart in the reference (its state lives in Python objects,
 * bayesopt/gp.py:28-73); one context per GPU replaces the module-level numpy / torch state. */
int nbw_version(void);
int nbw_create(nbw_ctx** out, int device);
int nbw_destroy(nbw_ctx* ctx);
const char* nbw_last_error(const nbw_ctx* ctx);
int nbw_reserve(nbw_ctx* ctx, uint64_t scratch_bytes);
/* NVTX range markers for the host layer (no reference counterpart; SURVEY.md §5 tracing row): the ranges
 * "fit", "score_state", "broadcast", "propose_location" bracket the kernels in an Nsight timeline. */
int nbw_range_push(const char* name);
int nbw_range_pop(void);
/* Measurement hook: two caller-owned cudaEvent_t (NULL, NULL to clear) that the next scoring entry point records
 * on its stream immediately before and after the SCORING KERNEL launch (not the small top-n merge behind it):
 * bench.py's per-launch kernel duration "measured live with CUDA events on the launching stream". */
int nbw_profile_events(nbw_ctx* ctx, void* ev_start, void* ev_stop);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t nbw_launch_count(const nbw_ctx* ctx);

/* ---------------------------------------------------------------------------------------
 * (1) WL relabelling — replaces the credential loop of WeisfeilerLehman.parse_input,
 *     grakel_replace/weisfeiler_lehman.py:258-296.
 * Level 0 (ids_prev == NULL): key = op code of each node, chk = 0.
 * Level l >= 1: key/chk = two independent multiset hashes of
 *     (label_{l-1}(v), {label_{l-1}(s) : s in successors(v)}), salted by (seed, level), where a label enters
 *     through its KEY uniq_keys_prev[ids_prev[.]] (the level-(l-1) dictionary of nbw_dict_build; level 0: the op
 *     code): signatures depend on the cell and the seed only, not on how a batch numbered its labels;
 *     nodes that are not an endpoint of any edge get NBW_INVALID_KEY.
 * keys/chks: [N * n_max] u64.  ids_prev: [N * n_max] u32 (NBW_INVALID_ID = not counted).
 */
int nbw_wl_signatures(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max,
                      const uint32_t* ids_prev, const uint64_t* uniq_keys_prev, int level, uint64_t seed,
                      uint64_t* keys, uint64_t* chks, void* stream);

/* ---------------------------------------------------------------------------------------
 * (2) Dictionary build: sort + unique that compacts keys into dense label ids —
 *     replaces `sorted(list(label_set))` + id assignment, weisfeiler_lehman.py:226-229,271-274.
 *     Three regimes with the same result: up to 8192 keys one CTA does everything in shared memory (one launch);
 *     above, a hash pre-dedup finds the distinct keys and only those are sorted (one extra stream sync inside);
 *     more than 65536 distinct keys take eight LSD radix passes over all keys.
 * ids[i] = rank of keys[i] among the distinct valid keys (NBW_INVALID_ID for invalid keys);
 * uniq_keys/uniq_chks[0..D) = the dictionary in ascending key order (capacity `uniq_cap`).
 * key_bits: number of significant low key bits (8 for level 0, 64 otherwise).
 * Verifies that equal keys carry equal chks, and (when cells != NULL) that nodes sharing a
 * key have identical (own id, successor-id multiset) tuples — the partition is then exact,
 * not probabilistic.  Syncs the stream; writes the dictionary size to *n_unique_h.
 */
int nbw_dict_build(nbw_ctx* ctx, const uint64_t* keys, const uint64_t* chks, int64_t n_keys,
                   int key_bits, const void* cells, int n_max, const uint32_t* ids_prev,
                   uint32_t* ids, uint64_t* uniq_keys, uint64_t* uniq_chks, int64_t uniq_cap,
                   int64_t* n_unique_h, void* stream);

/* (1) + (2) for every WL level of a batch in one call — the whole relabelling loop of
 * WeisfeilerLehman.parse_input (weisfeiler_lehman.py:223-296): levels 0..h are relabelled and their dictionaries
 * built back to back on the stream, with ONE read-back of the dictionary sizes at the end (batches above 8192
 * node slots add one sync per level, see (2)).
 * ids: [h+1][N * n_max] u32; uniq_keys / uniq_chks: [h+1][N * n_max] u64 (level l's sorted dictionary at offset
 * l * N * n_max); dims_h: [h+1] dictionary sizes.  Syncs once.  NBW_ERR_HASH_COLLISION as for nbw_dict_build. */
int nbw_wl_fit(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, int h, uint64_t seed,
               uint32_t* ids, uint64_t* uniq_keys, uint64_t* uniq_chks, int64_t* dims_h, void* stream);

/* ids_out[i] = map[ids_in[i]] (NBW_INVALID_ID is kept): renumber the labels of one level — from the sorted order
 * of a fresh nbw_dict_build to the APPEND-ONLY numbering of a training dictionary that grows across BO iterations
 * (the reference re-derives all ids from scratch on every fit, weisfeiler_lehman.py:223-296; keys are canonical
 * here, so the labels of the previous fit keep their ids and columns and new labels are appended). */
int nbw_remap_ids(nbw_ctx* ctx, const uint32_t* ids_in, int64_t n, const uint32_t* map, uint32_t* ids_out,
                  void* stream);

/* ---------------------------------------------------------------------------------------
 * (3) Vertex histograms — replaces VertexHistogram.parse_input,
 *     grakel_replace/vertex_histogram.py:98-209 (dense Phi of small integer counts).
 * ids: [n_levels][N * n_max] u32 stacked level-major.  Level l occupies columns
 * [col_off_h[l], col_off_h[l+1]) of the row-major int8 matrix phi [N x ld_phi]
 * (col_off_h and label_off_h have n_levels + 1 entries; label_off_h is only read when oa != 0);
 * every byte of phi[:, col_off_h[0] .. col_off_h[n_levels]) is written (pad columns = 0).
 *   oa == 0: phi[g, col_off[l] + id] = count of id in cell g at level l.
 *   oa != 0: thermometer code of the same counts (min(a,b) = sum_t [a>=t][b>=t]), so that
 *            Phi Phi^T equals the histogram-intersection kernel of vertex_histogram.py:233-249;
 *            thermo_base[l-stacked id] gives the first column of each label.
 * self_sim[g] (i32) = K_raw(g, g) = sum of squared counts (oa: number of counted nodes).
 */
int nbw_hist_dense_i8(nbw_ctx* ctx, const uint32_t* ids, int64_t n_cells, int n_max,
                      int n_levels, const int64_t* col_off_h, const int64_t* label_off_h,
                      const uint32_t* thermo_base, int oa, int8_t* phi, int64_t ld_phi,
                      int32_t* self_sim, void* stream);

/* OA support (histogram intersection sum_f min(x_f, y_f), vertex_histogram.py:233-249, evaluated
 * exactly as a dot product of thermometer codes): per-label maximum count over the batch ->
 * thermometer column bases.
 * level_off_h[l] = first stacked label index of level l (level_off_h[n_levels] = total labels).
 * thermo_base[i] (u32, i < total) = exclusive prefix sum of max counts *within its level*;
 * level_width_h[l] = sum of max counts of level l.  Syncs. */
int nbw_oa_thermo_layout(nbw_ctx* ctx, const uint32_t* ids, int64_t n_cells, int n_max,
                         int n_levels, const int64_t* level_off_h, uint32_t* thermo_base,
                         uint8_t* max_count, int64_t* level_width_h, void* stream);

/* ---------------------------------------------------------------------------------------
 * (4) Integer Gram — replaces VertexHistogram._calculate_kernel_matrix
 *     (vertex_histogram.py:211-259) and the level sum of weisfeiler_lehman.py:316-327.
 * C[i, j] = sum_{k in [k0, k1)} A[i, k] * B[j, k], int8 x int8 -> int32, exact.
 * A: [m x lda] int8 row-major, B: [n x ldb] int8 row-major, C: [m x ldc] int32 row-major.
 * tcgen05.mma kind::i8 with TMA-fed operands and TMEM accumulators; k0, k1, lda, ldb must be
 * multiples of 128 and A, B 128-byte aligned.  impl: 0 = tcgen05 (product), 1 = SIMT dp4a
 * cross-check kernel (test use).
 */
int nbw_gram_i8(nbw_ctx* ctx, const int8_t* A, int64_t m, int64_t lda, const int8_t* B,
                int64_t n, int64_t ldb, int64_t k0, int64_t k1, int32_t* C, int64_t ldc,
                int impl, void* stream);

/* The same product with the epilogue chosen per launch and a symmetric tile schedule (the Gram of a batch with
 * itself, vertex_histogram.py:243: K = Phi Phi^T is symmetric, computing both triangles doubles the work).
 *   out_mode    what leaves TMEM: s32 (C int32), u8 / s16 with saturation (*overflow |= 1 when a value did not
 *               fit: entries of NAS-cell Grams are <= (h+1) n_max^2), or fp32 K_ij * scale_a[i] * scale_b[j] — the
 *               cosine normalisation of kernels/combine_kernels.py:54-56 with scale = 1/sqrt(K_ii) fused in.
 *   upper_only  A = rows [row_offset, row_offset + m) of the matrix whose rows are B: 256 x 256 tiles that lie
 *               strictly below the diagonal of the global matrix are neither computed nor written;
 *               nbw_sym_mirror fills them from their mirror images afterwards (or the consumer uses symmetry).
 * ldc counts elements of the output type.  Always the CTA-pair tcgen05 kernel. */
typedef enum { NBW_GRAM_S32 = 0, NBW_GRAM_U8 = 1, NBW_GRAM_S16 = 2, NBW_GRAM_F32_NORM = 3 } nbw_gram_out;
int nbw_gram_i8_ex(nbw_ctx* ctx, const int8_t* A, int64_t m, int64_t lda, const int8_t* B, int64_t n,
                   int64_t ldb, int64_t k0, int64_t k1, void* C, int64_t ldc, int out_mode,
                   int upper_only, int64_t row_offset, const float* scale_a, const float* scale_b,
                   uint32_t* overflow, void* stream);
/* C[i][j] = C[j][i] for all i > j of a square matrix [n x ldc] of 1-, 2- or 4-byte elements (the lower
 * triangle of a symmetric Gram computed with upper_only; vertex_histogram.py:243). */
int nbw_sym_mirror(nbw_ctx* ctx, void* C, int64_t n, int64_t ldc, int elem_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (5) GP linear algebra in fp64 — replaces kernels/combine_kernels.py:36-56 (weighted sum +
 *     cosine normalisation), bayesopt/utils.py:120-140 (compute_pd_inverse) and the matrix
 *     products of GraphGP.fit / predict (bayesopt/gp.py:173-219, 271-276).
 * All matrices are row-major fp64 with explicit leading dimensions.
 */
/* The Gram of every single WL level of a SMALL training set in one launch (VertexHistogram._calculate_kernel_matrix,
 * vertex_histogram.py:243, per level — what the depth grid search of GraphGP.fit, bayesopt/gp.py:493-538, consumes):
 * out[l][i][j] = sum over the feature columns [col_off[l], col_off[l+1]) of phi[i][k] * phi[j][k], exact int32.
 * phi: [n x ld_phi] int8 (ld_phi and the column offsets multiples of 128, as nbw_hist_dense_i8 lays them out),
 * col_off: DEVICE int32 [n_levels + 1], out: n_levels matrices of ld_out columns, level_stride elements apart.
 * SIMT (dp4a); meant for n up to a few hundred, where one tensor-core launch per level costs more than the arithmetic.
 * Enqueues only. */
int nbw_level_grams_i32(nbw_ctx* ctx, const int8_t* phi, int64_t n, int64_t ld_phi, const int32_t* col_off,
                        int n_levels, int32_t* out, int64_t ld_out, int64_t level_stride, void* stream);

/* out[i,j] = beta*out[i,j] + w * K[i,j] (int32 -> fp64 accumulate): the level sum with
 * layer_weights of weisfeiler_lehman.py:316-327 when the levels are weighted separately */
int nbw_gram_accumulate(nbw_ctx* ctx, const int32_t* K, int64_t ldk, int64_t m, int64_t n,
                        double w, double beta, double* out, int64_t ldo, void* stream);
/* Y = a*X + b*Y on m x n fp64 blocks (weighted kernel sum, combine_kernels.py:40) */
int nbw_axpby_f64(nbw_ctx* ctx, const double* X, int64_t ldx, int64_t m, int64_t n, double a,
                  double b, double* Y, int64_t ldy, void* stream);
/* s[i] = 1/sqrt(K[i,i]);  K[i,j] *= s[i]*s[j]   (combine_kernels.py:54-56) */
int nbw_gram_normalize(nbw_ctx* ctx, double* K, int64_t ld, int64_t n, double* inv_sqrt_diag,
                       void* stream);
/* torch.cholesky of K + jitter*I (bayesopt/utils.py:129-131):
 * factor A + shift*I = L L^T (lower, in `L`, upper part zeroed).  *info_h = 0 on success,
 * k+1 if pivot k is not positive.  Writes log|A + shift I| to *logdet_h.  Syncs. */
int nbw_chol_f64(nbw_ctx* ctx, const double* A, int64_t lda, int64_t n, double shift,
                 double* L, int64_t ldl, double* logdet_h, int* info_h, void* stream);
/* Everything GraphGP.fit needs from one factorisation of K + lik*I (compute_pd_inverse,
 * bayesopt/utils.py:120-140, plus the terms of the marginal likelihood :102-117 and of its
 * gradient w.r.t. lik): L (lower), Linv = L^-1, alpha = (K + lik I)^-1 y and
 *   stats_h[0] = log|K + lik I|,  stats_h[1] = y^T (K + lik I)^-1 y,  stats_h[2] = tr (K + lik I)^-1,
 *   stats_h[3] = |alpha|^2   (stats_h has 4 entries).
 * Linv = alpha = NULL asks for the likelihood terms only (stats_h[0], stats_h[1]; the others are 0):
 * the inverse is skipped and z = L^-1 y comes from one forward substitution — what the grid search over
 * h needs per candidate (bayesopt/gp.py:521-527).
 * Syncs.  Returns NBW_ERR_NOT_PD (and *info_h = failing pivot + 1) when the factorisation breaks
 * down; the caller retries with 10x the jitter like utils.py:127-133. */
int nbw_gp_factor(nbw_ctx* ctx, const double* K, int64_t ldk, int64_t n, double lik,
                  const double* y, double* L, int64_t ldl, double* Linv, int64_t ldi,
                  double* alpha, double* stats_h, int* info_h, void* stream);
/* The whole grid search over the WL depth in one launch (_grid_search_wl_kernel, bayesopt/gp.py:493-538): for
 * every candidate c the UN-normalised Gram K_raw(h_c) = sum of the first n_levels_h[c] int32 level Grams
 * (level l at level_grams + l * level_stride, [n x ldg] each) + lik I is factored by its own thread-block
 * cluster; stats_h[2c] = log|K_raw(h_c) + lik I|, stats_h[2c + 1] = y^T (.)^-1 y, info_h[c] = 0 or failing
 * pivot + 1 (the caller retries that candidate with 10x the jitter, utils.py:127-133).  Up to 8 candidates.
 * Syncs once. */
int nbw_gp_grid(nbw_ctx* ctx, const int32_t* level_grams, int64_t ldg, int64_t level_stride, int64_t n,
                const int* n_levels_h, int n_cands, double lik, const double* y, double* stats_h,
                int* info_h, void* stream);

/* Gradient of the reference's marginal likelihood with respect to the weights of a SUM of kernels
 * (GraphGP.fit trains them by autograd: bayesopt/gp.py:141-210 through combine_kernels.py:36-56 and
 * utils.py:102-140).  Kinv = (N + lik I)^-1, alpha = Kinv y, N = normalised sum, inv_sqrt_diag from
 * nbw_gram_normalize, R_h = HOST array of n_kernels (<= 64) device pointers to raw fp64 Grams [n x n]
 * (the form is linear in R, so per-level Grams give the layer-weight gradients the same way).
 * Writes d nlml'/d w_i to grad_h[0..n_kernels).  Syncs. */
int nbw_kernel_weight_grad(nbw_ctx* ctx, const double* Kinv, int64_t ldk, const double* alpha,
                           const double* N, int64_t ldn, const double* inv_sqrt_diag, int64_t n,
                           const double* const* R_h, int64_t ldr, int n_kernels, double* grad_h,
                           void* stream);

/* Linv = L^-1 (lower triangular); Linv^T Linv is torch.cholesky_inverse (bayesopt/utils.py:139) */
int nbw_trtri_lower_f64(nbw_ctx* ctx, const double* L, int64_t ldl, int64_t n, double* Linv,
                        int64_t ldi, void* stream);
/* C = alpha * op(A) op(B) + beta * C;  op(X) = X or X^T; C is m x n, inner dimension k
 * (the products K_s^T K_i y and K_s^T K_i K_s of GraphGP.predict, bayesopt/gp.py:271-276) */
int nbw_gemm_f64(nbw_ctx* ctx, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
                 double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc, void* stream);
/* B[i, j] = s[i] * sqrt_w * phi[i, k0 + j]  (int8 -> fp64), j < k1-k0: the training features with the
 * cosine normalisation of combine_kernels.py:54-56 folded in (rows of K_s = B phi_c / sqrt(k_cc)) */
int nbw_scale_features(nbw_ctx* ctx, const int8_t* phi, int64_t ld_phi, int64_t n, int64_t k0,
                       int64_t k1, const double* s, double sqrt_w, double* B, int64_t ldb,
                       void* stream);
/* Ancestor map of one WL level: parent[ids_cur[i]] = ids_prev[i] for every counted node i
 * (all nodes sharing a label share the ancestor, weisfeiler_lehman.py:266-267). */
int nbw_label_parents(nbw_ctx* ctx, const uint32_t* ids_prev, const uint32_t* ids_cur,
                      int64_t n_keys, uint32_t* parent, void* stream);
/* Prefix-summed features over ancestor chains (see nbw_model.H) — the feature-space form of
 * K_s^T K_i K_s (bayesopt/gp.py:272-276) when labels refine level by level:
 *   Bp[i, label_off[0] + z]  = B[i, col_off[0] + z]
 *   Bp[i, label_off[l] + z]  = Bp[i, label_off[l-1] + parent[label_off[l] + z]] + B[i, col_off[l] + z]
 * B: [n x ldb] fp64 (feature-column layout), Bp: [n x ldp] fp64 (stacked-label layout);
 * parent: stacked u32 array (entries of level 0 unused); offsets have n_levels + 1 entries. */
int nbw_prefix_features(nbw_ctx* ctx, const double* B, int64_t ldb, int64_t n, int n_levels,
                        const int64_t* col_off_h, const int64_t* label_off_h,
                        const uint32_t* parent, double* Bp, int64_t ldp, void* stream);
/* cov[i,j] = (sqrt(max(Kss[i,j] + (i==j)*lik - Q[i,j], lik)) * y_std)^2   (gp.py:272-280) */
int nbw_posterior_cov_finish(nbw_ctx* ctx, const double* Kss, int64_t ldk, const double* Q,
                             int64_t ldq, int64_t m, double lik, double y_std, double* cov,
                             int64_t ldc, void* stream);

/* ---------------------------------------------------------------------------------------
 * (6) Fused pool scoring — replaces GraphGP.predict (bayesopt/gp.py:240-283, diagonal) +
 *     GraphExpectedImprovement.eval / GraphUpperConfidentBound.eval
 *     (bayesopt/acquisitions.py:59-80, 130-138) for a whole candidate pool in one launch.
 *
 * The fitted surrogate is handed over as a flat model descriptor (all device pointers):
 * per WL level a dictionary (open-addressing table of 16-byte slots built by
 * nbw_model_build_table), the feature-space posterior operators
 *     w_mu = B^T alpha          [D]
 *     G    = B^T K_lam^-1 B     [D x D]      with B = diag(1/sqrt(K_raw,ii)) Phi_train
 * and the scalars of the posterior / acquisition.  For a candidate c with (restricted)
 * feature vector phi_c and self-similarity k_cc over ALL its labels:
 *     mu  = (phi_c . w_mu / sqrt(k_cc)) * y_std + y_mean
 *     var = (sqrt(max(1 + lik - phi_c^T G phi_c / k_cc, lik)) * y_std)^2
 */
typedef enum { NBW_ACQ_EI = 0, NBW_ACQ_AEI = 1, NBW_ACQ_UCB = 2, NBW_ACQ_MEAN = 3 } nbw_acq;

typedef struct {
  int32_t n_max;                         /* 8 or 16 */
  int32_t h;                             /* WL depth: levels 0..h */
  int32_t oa;                            /* 0 linear, 1 histogram intersection */
  int32_t acq;                           /* nbw_acq */
  uint64_t seed;                         /* the seed the dictionary was built with */
  const uint16_t* level0_table;          /* [256] op code -> dense level-0 id, 0xFFFF = unseen */
  const void* table[NBW_MAX_LEVELS];     /* level l>=1: slots {u64 key; u32 chk_lo; u32 id} */
  uint32_t table_mask[NBW_MAX_LEVELS];   /* capacity-1 (capacity is a power of two) */
  int32_t col_off[NBW_MAX_LEVELS];       /* first feature column of level l (oa==0: + id) */
  int32_t label_off[NBW_MAX_LEVELS];     /* first stacked label index of level l */
  const uint32_t* thermo_base;           /* oa: stacked label -> first column within level */
  const uint8_t* max_count;              /* oa: stacked label -> max training count */
  int32_t n_features;                    /* D */
  int32_t ld_g;                          /* >= D + 1 */
  const double* G;                       /* [(D+1) x ld_g]; row D and column D are all zero: the */
  const double* w_mu;                    /* [D+1], w_mu[D] = 0   landing slot of unseen labels */
  /* Prefix form (oa == 0 only, optional: H == NULL selects the general G path).  A WL label of
   * level l determines its ancestors at levels < l, and a label can only be in the dictionary if
   * its ancestor is, so a node is fully described by its DEEPEST dictionary label s (stacked
   * label index).  H and w_mu_prefix are G and w_mu summed over the ancestor chains:
   *   H[s, s'] = sum_{a in chain(s), b in chain(s')} G[col(a), col(b)],
   * which turns phi^T G phi into a sum over node PAIRS instead of (node, level) pairs. */
  int32_t n_labels;                      /* T = total stacked labels of levels 0..h */
  int32_t ld_h;                          /* >= T + 1 */
  const double* H;                       /* [(T+1) x ld_h], row/column T all zero */
  const double* w_mu_prefix;             /* [T+1], entry T = 0 */
  /* Stable suffix of the WL refinement (optional; first_stable = 0 disables it).  Once the training
   * dictionary stops refining (D_{l-1} == D_{l-2}, l - 2 >= 1) every deeper parent map is a bijection:
   * a pool node is in the dictionary at level l >= first_stable iff it and all its successors were at
   * level l-1, and its label is then child[label_off[l-1] + id_{l-1}] — no signature, no probe.  Label
   * classes inside the cell (needed for k_cc) are refined exactly from the previous classes. */
  const int32_t* child;                  /* stacked [T]: level-(l-1) label -> its only level-l label */
  int32_t first_stable;                  /* levels l >= first_stable (>= 3) take the fast path; 0 = none */
  int32_t oa_d2_base;                    /* oa item form: first stacked index of the rank-2 chain items (0 = none), see oa_exc */
  double lik, y_mean, y_std, incumbent, xi, beta;
  /* ---- packed scoring path (csrc/score_packed.cu; optional: bucket_tags[1] == NULL keeps the first-generation
   * kernel).  Same dictionaries in a branch-free form:
   *   id_hash[l]        [D_{l-1}] 16-byte entries {u64 ha; u32 hb; u32 0}: the per-node hashes of every level-(l-1)
   *                     dictionary id salted for level l (nbw_model_build_id_hash) — pool nodes FETCH them;
   *   bucket_tags[l]    [n_buckets] 16-byte buckets of four 32-bit tags (low word of the key, 0 = empty);
   *   bucket_entries[l] [4 * n_buckets] 16-byte entries {u32 key_hi; u32 chk; u32 id; u32 0} (nbw_model_build_buckets);
   *   final_label       stacked [T]: a label of a stable level -> GLOBAL stacked index of its descendant at level h.
   * Sum of kernels (kernels/combine_kernels.py:36-56): the parts of the sum form a chain through `next`
   * (HOST pointer, NULL = last).  A candidate is `rec_stride` bytes (0 = one record); part p reads its cell
   * record at byte `rec_offset` of it ("normal | reduce": two records per candidate).  H / w_mu_prefix /
   * n_labels / ld_h and the scalars are taken from the FIRST part and span the stacked labels of all
   * parts; label_base is the first stacked index of this part.  level_weight[l] = kernel weight x WL layer
   * weight of level l (weisfeiler_lehman.py:316-327); weighted == 0 promises they are all 1 (k_cc stays an
   * integer).  undirected != 0 symmetrises the adjacency in the kernel (kernels/weisfilerlehman.py:127-128). */
  const void* id_hash[NBW_MAX_LEVELS];
  const void* bucket_tags[NBW_MAX_LEVELS];
  const void* bucket_entries[NBW_MAX_LEVELS];
  uint32_t bucket_mask[NBW_MAX_LEVELS];  /* n_buckets - 1 (power of two) */
  const int32_t* final_label;
  double level_weight[NBW_MAX_LEVELS];
  int32_t weighted, undirected, rec_offset, rec_stride, label_base;
  /* compact != 0: the part's record is the WIRE form of a cell (n_max = 8: 8 bytes) — half the PCIe bytes of a
   * pool that arrives in host memory.  Little-endian u64: bits [3v, 3v+3) = palette index of node v's op code
   * (7 = no node), bits [24, 52) = the strict upper triangle of the adjacency, row by row (edge u -> v, u < v,
   * at bit 24 + u (15 - u) / 2 + (v - u - 1)): cells whose nodes are numbered topologically, at most 7 distinct
   * ops.  palette[c] = the op code (level-0 key) of index c.  The kernel expands it in its prologue.
   * n_max = 16: a 24-byte record of three little-endian u64 words — word 0 = sixteen 4-bit palette indices (15 = no
   * node, at most 15 distinct ops), words 1-2 = the 120 bits of the strict upper triangle, row by row (edge u -> v,
   * u < v, at bit u (31 - u) / 2 + (v - u - 1) of the 128-bit field): half the bytes of the 48-byte record. */
  int32_t compact;
  uint8_t palette[16];                   /* (n_max = 8 uses the first 8 entries) */
  /* Histogram intersection on the packed path (oa != 0, optional: NULL keeps the general G path).  In thermometer
   * form (vertex_histogram.py:233-249: sum_label min(count, count') = sum_(label, t) [count >= t][count' >= t]) every
   * (node, level) of a candidate activates ONE column: (its label, its rank among the cell's nodes with that label).
   * Ranks are non-increasing along a node's ancestor chain, so a node is its rank-1 chain — described by its deepest
   * dictionary label, like the linear kernel — plus one EXCEPTION item per level at which its rank exceeds 1:
   *   e(label, r) = column (label, r) [if r <= max training count of the label] - column (label, 1).
   * H / w_mu_prefix then span n_labels = T + X stacked items: the T chain items (label_off) followed by the
   * exception items, and the posterior is the same sum over item pairs.
   * oa_exc[label_off[l] + id] = (stacked index of e(label, 2)) << 5 | max training count (<= 16); the item of rank r
   * is that index + min(r, max + 1) - 2.  Single part only.
   * Twins — two nodes with the same label at every level — are the common case, and a node whose rank is 2 at level 0
   * has rank 2 at every level it has an exception at (ranks do not grow): when oa_d2_base != 0 such a node carries ONE
   * item instead, oa_d2_base + (stacked label at its last exception level) = the sum of e(., 2) along that label's
   * ancestor chain. */
  const uint32_t* oa_exc;
  const void* next;                      /* const nbw_model* of the next part (host memory) */
} nbw_model;

#define NBW_MAX_PARTS 4                  /* kernels in a sum the fused pool path accepts */

/* Build one level's lookup table from the sorted dictionary of nbw_dict_build — the device form of
 * self._inv_labels[i], which WeisfeilerLehman.transform consults per node (weisfeiler_lehman.py:434-466).
 * table: capacity slots of 16 bytes, capacity = power of two >= 2 * n_unique. */
int nbw_model_build_table(nbw_ctx* ctx, const uint64_t* uniq_keys, const uint64_t* uniq_chks,
                          int64_t n_unique, void* table, int64_t capacity, void* stream);

/* Packed-path dictionary forms (see nbw_model): replace self._inv_labels[i] lookups of
 * WeisfeilerLehman.transform (weisfeiler_lehman.py:434-466).
 * nbw_model_build_id_hash: out[d] = hashes of the key of level-(level-1) dictionary id d (d < n_ids), salted for `level`.
 * nbw_model_build_buckets: four-entry buckets over the sorted dictionary of nbw_dict_build; n_buckets a
 * power of two.  *flags (device u32, caller zeroes) gets bit 0 when a bucket overflows and bit 1 when two
 * keys of one bucket share a tag: the caller rebuilds with more buckets.  Both enqueue only. */
int nbw_model_build_id_hash(nbw_ctx* ctx, const uint64_t* uniq_keys_prev, int64_t n_ids, uint64_t seed, int level,
                            void* out, void* stream);
int nbw_model_build_buckets(nbw_ctx* ctx, const uint64_t* uniq_keys, const uint64_t* uniq_chks,
                            int64_t n_unique, void* tags, void* entries, int64_t n_buckets,
                            uint32_t* flags, void* stream);

/* out[i, j] = sum_k sign(k) * B[i, col(k)] over the K entries idx[j * K + k]: v >= 0 adds column v, v < 0 subtracts
 * column ~v, 0x7FFFFFFF is skipped.  Builds the item features of the packed histogram-intersection path (rank-1
 * chains and exception items, see nbw_model.oa_exc) from the scaled thermometer features of the training set —
 * the feature-space form of VertexHistogram._calculate_kernel_matrix's min-sum (vertex_histogram.py:233-249).
 * B: [n x ldb] f64, idx: [J x K] i32 (device), out: [n x ldo] f64.  Enqueues only. */
int nbw_combine_columns(nbw_ctx* ctx, const double* B, int64_t n, int64_t ldb, const int32_t* idx, int64_t J, int K,
                        double* out, int64_t ldo, void* stream);

/* GraphGP.predict (bayesopt/gp.py:240-283, diagonal) + acquisition eval (acquisitions.py:59-80,130-138)
 * over the pool.  scores: [M] f32 (required).  mu64/var64/score64: optional [M] f64 outputs (NULL to skip). */
int nbw_score_pool(nbw_ctx* ctx, const nbw_model* model_h, const void* cells, int64_t n_cells,
                   float* scores, double* mu64, double* var64, double* score64, void* stream);

/* nbw_score_pool fused with the local top-n selection of propose_location (acquisitions.py:94-113):
 * every lane keeps the k best (distinct) scores it produced, CTAs merge them at kernel end, one small
 * tournament merges the per-CTA lists.  out_cands: k 16-byte entries {f32 value; i32 pad; i64 index}
 * (index_base added; -1 = empty) left on the device — what a rank contributes to the top-n all-gather.
 * k <= 8 and the packed path only (NBW_ERR_UNSUPPORTED otherwise: use nbw_score_pool + nbw_topk_device).
 * scores may be NULL when only the selection is wanted.  Enqueues only. */
int nbw_score_pool_topk(nbw_ctx* ctx, const nbw_model* model_h, const void* cells, int64_t n_cells,
                        float* scores, int k, int distinct, int64_t index_base, void* out_cands,
                        void* stream);

/* Host-buffer form of nbw_score_pool — the call a host-side plugin makes from propose_location
 * (bayesopt/acquisitions.py:94-113), which receives its candidates in host memory: `cells_h` are packed
 * records in HOST memory (pinned memory gives full overlap), `scores_h` (optional) receives the
 * acquisition values in host memory, `scores_dev` [M] keeps them on the device for nbw_topk.  The
 * pool is cut into `chunks` pieces; the H2D copy of piece i+1, the kernel on piece i and the D2H copy
 * of piece i-1 overlap on internal streams.  Enqueues only.  The D2H copies run BESIDE whatever is
 * queued next on `stream` (normally the top-n selection on scores_dev): scores_h is complete after the
 * next nbw_topk / nbw_topk_merge on this context or after nbw_wait_host_scores. */
int nbw_score_pool_host(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h,
                        int64_t n_cells, float* scores_h, float* scores_dev, int chunks,
                        void* stream);

/* Mapped form of the same call (acquisitions.py:94-113): `cells_h` must be PINNED host memory (cudaHostAlloc / cudaHostRegister); the kernel
 * reads the records straight over PCIe through the mapped pointer — no staging copy, no per-chunk
 * launches.  The pool is cut into `pieces` (1..64) kernels only so that the D2H copy of piece i's
 * scores overlaps the kernel on piece i+1 (pieces = 1 measured best at 1M).  Pays only when the kernel is
 * slower than the PCIe transfer of its records (16-byte records, h >= 5: 0.57 vs 0.61 ms per 1M); a
 * transfer-bound pool is faster through the DMA pipeline of nbw_score_pool_host.  Pageable `cells_h` is refused (NBW_ERR_INVALID): use
 * nbw_score_pool_host for it.  Same stream semantics as nbw_score_pool_host. */
int nbw_score_pool_mapped(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h,
                          int64_t n_cells, float* scores_h, float* scores_dev, int pieces,
                          void* stream);

/* Block until the host scores of the last nbw_score_pool_host / nbw_score_pool_mapped call have
 * landed (no-op when nothing is pending).  The eis array propose_location returns
 * (bayesopt/acquisitions.py:111-113) must be complete before the caller reads it. */
int nbw_wait_host_scores(nbw_ctx* ctx);

/* Pool-side features only — WeisfeilerLehman.transform(..., return_embedding_only=True)
 * (weisfeiler_lehman.py:364-516), used by feature_value / forward_t: restricted dense feature rows
 * phi [M x ld_phi] int8 (columns < n_features) and self-similarity k_cc [M] i32. */
int nbw_pool_features(nbw_ctx* ctx, const nbw_model* model_h, const void* cells,
                      int64_t n_cells, int8_t* phi, int64_t ld_phi, int32_t* self_sim,
                      void* stream);

/* ---------------------------------------------------------------------------------------
 * Top-n selection — replaces propose_location's np.unique + np.argpartition
 * (bayesopt/acquisitions.py:101-105): the k largest DISTINCT score values, each with the
 * first (lowest) index at which it occurs; NaN scores are ignored.  Results are sorted by
 * descending value; *n_found_h <= k.  index_base is added to the indices (pool shard offset).
 * distinct == 0 gives plain top-k (ties by lowest index).  Syncs.
 */
int nbw_topk(nbw_ctx* ctx, const float* scores, int64_t n, int k, int distinct,
             int64_t index_base, float* top_val_h, int64_t* top_idx_h, int* n_found_h,
             void* stream);

/* The same selection (acquisitions.py:101-105) with the result left on the device (no sync): out_cands[k] 16-byte entries
 * {f32 value; i32 pad; i64 index}, index -1 = empty.  This is what each rank contributes to the
 * top-n all-gather. */
int nbw_topk_device(nbw_ctx* ctx, const float* scores, int64_t n, int k, int distinct,
                    int64_t index_base, void* out_cands, void* stream);
/* Merge `count` such entries (e.g. the all-gathered lists of every rank) into the global top k
 * (acquisitions.py:101-105 applied to the union):
 * same distinct-by-value rule, ties to the lowest index.  Syncs. */
int nbw_topk_merge(nbw_ctx* ctx, const void* cands, int64_t count, int k, int distinct,
                   float* top_val_h, int64_t* top_idx_h, int* n_found_h, void* stream);

/* nbw_topk_merge without the read-back (acquisitions.py:101-105 applied to the union of the ranks' lists): the
 * merged list stays on the device in out_cands[k]; enqueues only.  Up to 4096 entries are merged by one CTA
 * without touching the context's scratch arena, so the call may run on a side stream beside a scoring launch. */
int nbw_topk_merge_device(nbw_ctx* ctx, const void* cands, int64_t count, int k, int distinct,
                          void* out_cands, void* stream);

/* Read a k-entry device list (nbw_topk_device / nbw_score_pool_topk / nbw_propose_host) back to the host:
 * the indices propose_location returns (acquisitions.py:105-113).  Syncs, and awaits pending host scores. */
int nbw_cands_read(nbw_ctx* ctx, const void* cands, int k, float* top_val_h, int64_t* top_idx_h,
                   int* n_found_h, void* stream);

/* propose_location (bayesopt/acquisitions.py:94-113) on a pool in HOST memory in one call: the staged
 * (mapped == 0, nbw_score_pool_host) or mapped (mapped == 1: nbw_score_pool_mapped; mapped == 2: the kernel also
 * writes the scores through the mapped alias of a pinned scores_h, no D2H copy, scores_dev untouched) pipeline with the local top-n of
 * every piece fused into its scoring kernel and merged on the device: out_cands as for
 * nbw_score_pool_topk.  k <= 8, packed path only.  Enqueues only; same host-score semantics. */
int nbw_propose_host(nbw_ctx* ctx, const nbw_model* model_h, const void* cells_h, int64_t n_cells,
                     float* scores_h, float* scores_dev, int mapped, int pieces, int k, int distinct,
                     int64_t index_base, void* out_cands, void* stream);

/* ---------------------------------------------------------------------------------------
 * Synthetic pools shaped like the reference's samplers, generated on device
 * (bayesopt/generate_test_graphs.py:543-641; darts/darts_objectives.py:159-198 +
 * darts/utils.py:11-98).  family: 0 = NAS-Bench-101 (n_max 8), 1 = NAS-Bench-201 (n_max 8),
 * 2 = DARTS (n_max 16).  Cell i depends only on (seed, first_index + i): shards generated on
 * different GPUs concatenate to the single-GPU pool.  Op codes are the fixed family
 * vocabularies listed in nasbowl_b200/synth.py.
 */
int nbw_generate_cells(nbw_ctx* ctx, int family, uint64_t seed, int64_t first_index,
                       int64_t n_cells, void* cells, void* stream);

/* BANANAS-style mutation pool (bayesopt/generate_test_graphs.py:393-539, NB101 branch): child i is
 * an edit of parents[(first_index + i) % n_parents] — every possible edge flips with probability
 * ~1/vertice, every interior op changes with probability ~1/(vertice-2), then `prune`; children with
 * <= 3 nodes or > 9 edges are re-drawn (64 attempts, then a random cell).  The reference's
 * isomorphism de-duplication (a host-side graph search) is not reproduced.  family must be 0. */
int nbw_mutate_cells(nbw_ctx* ctx, int family, const void* parents, int64_t n_parents, uint64_t seed,
                     int64_t first_index, int64_t n_children, void* cells, void* stream);

/* Pool de-duplication — `allow_isomorphism=False` of mutation() (bayesopt/generate_test_graphs.py:497-520: a child
 * isomorphic to its parent or to an earlier pool member is dropped; the reference searches pairwise with networkx).
 * Every cell gets a 128-bit Weisfeiler-Lehman signature (colour refinement over both edge directions to depth
 * n_max; labeled == 0 ignores the op codes like the reference's is_isomorphic call without node_match, labeled != 0
 * also requires equal ops); keep[i] = 1 iff no cell j < i has the same signature.  The first n_protected cells
 * are references (parents, already evaluated cells): they count as "earlier" but are not counted in *n_kept_h.
 * *n_classes_h = number of distinct signatures.  Syncs.  NBW_ERR_HASH_COLLISION: retry with another seed. */
int nbw_pool_dedup(nbw_ctx* ctx, const void* cells, int64_t n_cells, int n_max, int labeled,
                   int64_t n_protected, uint64_t seed, uint8_t* keep, int64_t* n_classes_h,
                   int64_t* n_kept_h, void* stream);

/* NB201 / DARTS cells with their ENCODINGS (the pruned cell does not determine the architecture it came
 * from, and mutation edits the architecture): NB201 = six edge choices (0..2 ops, 3 skip_connect, 4 none)
 * in 8 bytes; DARTS = eight (op 0..6, input 0..5) pairs in 16 bytes.  Same cells as nbw_generate_cells
 * for the same (seed, index).  Replaces random_sampling (generate_test_graphs.py:543-641) /
 * random_sample_darts (darts/darts_objectives.py:159-198) when the parents of a later mutation are needed. */
int nbw_generate_encoded(nbw_ctx* ctx, int family, uint64_t seed, int64_t first_index, int64_t n_cells,
                         void* cells, uint8_t* enc, void* stream);

/* BANANAS-style mutation on encodings: child i edits parent i % n_parents.  NB201: _banana_mutate
 * (generate_test_graphs.py:412-431); DARTS: _mutate with `edits` edits, mutate_main_only, invalid children
 * redrawn (darts/darts_objectives.py:201-245,258-260).  child_enc (optional) receives the children's
 * encodings.  `edits` is ignored for NB201. */
int nbw_mutate_encoded(nbw_ctx* ctx, int family, const uint8_t* parent_enc, int64_t n_parents,
                       uint64_t seed, int64_t first_index, int64_t n_children, int edits, void* cells,
                       uint8_t* child_enc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBW_H_ */
