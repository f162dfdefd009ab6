// GP factorisation in ONE launch: Cholesky of K + lik I with pivot check and log-determinant, z = L^-1 y,
// and (optionally) the triangular inverse, alpha = K_lam^-1 y, tr K_lam^-1 and |alpha|^2 — everything
// compute_pd_inverse + compute_log_marginal_likelihood (bayesopt/utils.py:102-140) and the lambda gradient of
// GraphGP.fit (bayesopt/gp.py:173-210) need — and, batched, the whole grid search over the WL depth
// (bayesopt/gp.py:493-538): one thread-block CLUSTER per candidate h, all candidates in the same launch, each
// building its K_raw(h) = sum_{l <= h} Phi_l Phi_l^T from the int32 level Grams.
//
// The blocked factorisation this replaces (linalg_f64.cu) is a chain of two launches per 32-column panel
// whose cost is the serial dependency of the chain, not flops: 0.19 ms at n = 150, 1.1 ms at n = 1000, and the
// h grid ran those chains one after another with a host sync each.  Here a cluster of C CTAs (C = 1..16 by n)
// walks the panels with hardware cluster barriers between the phases; the matrix lives in L2 (all traffic
// .cg); y rides along as an extra matrix row, so z = L^-1 y falls out of the panel solves.
#include "nbw_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

constexpr int FB = 32;     // panel width
constexpr int FT = 256;    // threads per CTA
constexpr int kMaxProblems = 8;

struct FactorProblem {
  const double* K;          // fp64 source [n x ldk] ...
  int64_t ldk;
  const int32_t* G;         // ... or the sum of the first n_levels int32 level Grams [n_levels][n x ldg]
  int64_t ldg, g_stride;
  int32_t n_levels, pad;
  double shift;
  double* L;                // [n x ldl] factor (lower, upper part zeroed); also the workspace
  int64_t ldl;
  double* Linv;             // optional [n x ldi]
  int64_t ldi;
  double* W;                // workspace of the inverse merges [n x n] (when Linv)
  double* alpha;            // optional [n]
  double* alpha_acc;        // [n] accumulator of alpha (zeroed here), needed when Linv is requested
  double* z;                // [n]: y on entry of the panel loop, L^-1 y at the end
  double* stats;            // [4] log|K + shift I|, y^T (K + shift I)^-1 y, tr (.)^-1, |alpha|^2
  int* info;                // 0, or failing pivot + 1
};
struct FactorBatch {
  FactorProblem p[kMaxProblems];
  const double* y;
  int64_t n;
};

__device__ __forceinline__ uint32_t fac_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t fac_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// every thread of every CTA of the cluster; global writes before it are visible to the whole cluster after it
__device__ __forceinline__ void fac_cluster_sync() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < FT / 32 ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;    // valid on warp 0
}

// C[64 x 64 tile] (+)= alpha * A[m x k] * B^T or B, generic small tile product out of L2, used by the trailing
// update (syrk form) and by the inverse merges.  256 threads, 4 x 4 outputs each.
struct TileBuf {
  double a[16][64 + 2];
  double b[16][64 + 2];
};

__global__ void __launch_bounds__(FT)
gp_factor_kernel(const __grid_constant__ FactorBatch B) {
  __shared__ TileBuf tb;                     // 16.9 KB
  __shared__ double l11[2][FB][FB + 1];      // 16.9 KB: panel factor (slot 0) / two diagonal blocks being inverted
  __shared__ double inv_diag[FB];
  __shared__ double colbuf[FB];
  __shared__ double red[FT / 32];
  const uint32_t C = fac_cluster_size(), rank = fac_cluster_rank();
  const FactorProblem& Q = B.p[blockIdx.x / C];
  const int64_t n = B.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t gt = (int64_t)rank * FT + tid, gstride = (int64_t)C * FT;
  double* __restrict__ L = Q.L;
  const int64_t ldl = Q.ldl;

  // ---------------------------------------------------------------- build A + shift I (lower), z = y
  for (int64_t t = gt; t < n * n; t += gstride) {
    const int64_t i = t / n, j = t % n;
    double v = 0.0;
    if (j <= i) {
      if (Q.K) {
        v = Q.K[i * Q.ldk + j];
      } else {
        int64_t acc = 0;
        for (int l = 0; l < Q.n_levels; ++l) acc += Q.G[(int64_t)l * Q.g_stride + i * Q.ldg + j];
        v = (double)acc;
      }
      if (i == j) v += Q.shift;
    }
    __stcg(L + i * ldl + j, v);
    if (Q.Linv) __stcg(Q.Linv + i * Q.ldi + j, 0.0);
  }
  for (int64_t i = gt; i < n; i += gstride) {
    __stcg(Q.z + i, B.y[i]);
    if (Q.Linv) __stcg(Q.alpha_acc + i, 0.0);
  }
  if (gt == 0) {
    Q.stats[0] = Q.stats[1] = Q.stats[2] = Q.stats[3] = 0.0;
    *Q.info = 0;
  }
  fac_cluster_sync();

  // ---------------------------------------------------------------- panels
  double logdet = 0.0;                       // meaningful on warp 0 (every CTA computes the same value)
  for (int64_t k0 = 0; k0 < n; k0 += FB) {
    const int nb = (int)((n - k0) < FB ? (n - k0) : FB);
    // (a) the diagonal block, factored warp-synchronously by warp 0 of EVERY CTA (no barrier needed before the
    // row solves): lane = row, the column travels by shuffle, rsqrt + multiplies instead of sqrt / divide chains
    if (warp == 0) {
      // Right-looking in registers (lane = row), the scaled pivot column broadcast through shared memory: one
      // store + one warp sync per step, then every lane reads the entries it needs (same address for all
      // lanes of a read: a broadcast) and updates its row with INDEPENDENT fmas.  (The shuffle form of this
      // block — 31 - j 64-bit shuffles per step on a single warp — was the 35 us panel step of the first build.)
      double a[FB];
      double piv = 1.0;                       // this lane's pivot L_jj^2 (one log per lane per panel, off the chain)
#pragma unroll
      for (int c = 0; c < FB; ++c) a[c] = (lane < nb && c <= lane && c < nb) ? __ldcg(L + (k0 + lane) * ldl + k0 + c) : 0.0;
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        if (j < nb) {
          if (lane == j) {
            const double ajj = a[j];
            if (!(ajj > 0.0) && rank == 0) atomicCAS(Q.info, 0, (int)(k0 + j + 1));
            const double rinv = rsqrt(ajj);
            piv = ajj;
            a[j] = ajj * rinv;
            inv_diag[j] = rinv;
          }
          __syncwarp();
          const double rinv = inv_diag[j];
          if (lane > j) a[j] *= rinv;
          colbuf[lane] = (lane >= j) ? a[j] : 0.0;
          __syncwarp();
#pragma unroll
          for (int c = j + 1; c < FB; ++c)
            if (lane >= c) a[c] = fma(-a[j], colbuf[c], a[c]);
          __syncwarp();
        }
      }
      if (lane >= nb) inv_diag[lane] = 0.0;
      {
        double lg = log(piv);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
        logdet += lg;
      }
#pragma unroll
      for (int c = 0; c < FB; ++c) {
        l11[0][lane][c] = (c <= lane) ? a[c] : 0.0;
        if (rank == 0 && lane < nb && c <= lane && c < nb) __stcg(L + (k0 + lane) * ldl + k0 + c, a[c]);
      }
    }
    __syncthreads();
    // (b) the rows below it (and the y row, index n): X = A21 L11^-T, right-looking per row
    for (int64_t i = k0 + nb + gt; i <= n; i += gstride) {
      double* row = (i < n) ? L + i * ldl + k0 : Q.z + k0;
      double x[FB];
#pragma unroll
      for (int j = 0; j < FB; ++j) x[j] = (j < nb) ? __ldcg(row + j) : 0.0;
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        if (j < nb) {
          const double xj = x[j] * inv_diag[j];
          x[j] = xj;
#pragma unroll
          for (int t = j + 1; t < FB; ++t) x[t] = fma(-xj, l11[0][t][j], x[t]);
        }
      }
#pragma unroll
      for (int j = 0; j < FB; ++j)
        if (j < nb) __stcg(row + j, x[j]);
    }
    const int64_t base = k0 + nb, rest = n - base;
    if (rest <= 0) break;
    fac_cluster_sync();
    // (c) trailing update A22 -= P P^T on the lower triangle (64 x 64 tiles dealt round-robin to the CTAs) ...
    const int64_t T = (rest + 63) / 64, n_tiles = T * (T + 1) / 2;
    for (int64_t tile = rank; tile < n_tiles; tile += C) {
      // tile -> (ti >= tj)
      int64_t ti = (int64_t)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
      while (ti * (ti + 1) / 2 > tile) --ti;
      while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
      const int64_t tj = tile - ti * (ti + 1) / 2;
      const int64_t i0 = base + ti * 64, j0 = base + tj * 64;
      const int tx = tid & 15, ty = tid >> 4;
      double acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
      for (int kk0 = 0; kk0 < nb; kk0 += 16) {
        __syncthreads();
        for (int t = tid; t < 64 * 16; t += FT) {
          const int kk = t % 16, r = t / 16;
          const int64_t gi = i0 + r, gj = j0 + r;
          tb.a[kk][r] = (gi < n && kk0 + kk < nb) ? __ldcg(L + gi * ldl + k0 + kk0 + kk) : 0.0;
          tb.b[kk][r] = (gj < n && kk0 + kk < nb) ? __ldcg(L + gj * ldl + k0 + kk0 + kk) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          double av[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) av[r] = tb.a[kk][ty * 4 + r];
#pragma unroll
          for (int c = 0; c < 4; ++c) bv[c] = tb.b[kk][tx * 4 + c];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t gi = i0 + ty * 4 + r;
        if (gi >= n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int64_t gj = j0 + tx * 4 + c;
          if (gj <= gi) __stcg(L + gi * ldl + gj, __ldcg(L + gi * ldl + gj) - acc[r][c]);
        }
      }
    }
    // ... and of the y row: z_j -= sum_t P_jt z_seg,t
    for (int64_t j = base + gt; j < n; j += gstride) {
      double pr[FB], zs[FB];
#pragma unroll
      for (int t = 0; t < FB; ++t) {          // 64 independent loads in flight, then the dot product
        pr[t] = (t < nb) ? __ldcg(L + j * ldl + k0 + t) : 0.0;
        zs[t] = (t < nb) ? __ldcg(Q.z + k0 + t) : 0.0;
      }
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int t = 0; t < FB; t += 2) {
        acc0 = fma(pr[t], zs[t], acc0);
        acc1 = fma(pr[t + 1], zs[t + 1], acc1);
      }
      __stcg(Q.z + j, __ldcg(Q.z + j) - (acc0 + acc1));
    }
    fac_cluster_sync();
  }
  fac_cluster_sync();
  // ---------------------------------------------------------------- likelihood terms
  {
    double zz = 0.0;
    if (rank == 0)
      for (int64_t i = tid; i < n; i += FT) { const double v = __ldcg(Q.z + i); zz += v * v; }
    const double t = block_sum(zz, red);
    if (rank == 0 && tid == 0) {
      Q.stats[0] = logdet;
      Q.stats[1] = t;
    }
  }
  if (!Q.Linv) return;

  // ---------------------------------------------------------------- triangular inverse by recursive doubling
  // 32 x 32 diagonal blocks: lane j of a warp solves L x = e_j; two warps per CTA work at a time
  double* __restrict__ X = Q.Linv;
  const int64_t ldx = Q.ldi;
  const int64_t n_blk = (n + FB - 1) / FB;
  for (int64_t b0 = (int64_t)rank * 2; b0 < n_blk; b0 += (int64_t)C * 2) {
    __syncthreads();
    if (warp < 2 && b0 + warp < n_blk) {
      const int64_t kb = (b0 + warp) * FB;
      const int nb = (int)((n - kb) < FB ? (n - kb) : FB);
      for (int r = 0; r < FB; ++r)
        l11[warp][r][lane] = (r < nb && lane <= r) ? __ldcg(L + (kb + r) * ldl + kb + lane) : (r == lane ? 1.0 : 0.0);
      __syncwarp();
      l11[warp][lane][FB] = 1.0 / l11[warp][lane][lane];      // reciprocal pivots in the pad column: no divide chain
      __syncwarp();
      double x[FB];
#pragma unroll
      for (int p = 0; p < FB; ++p) x[p] = (p == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int p = 0; p < FB; ++p) {
        const double xp = x[p] * l11[warp][p][FB];
        x[p] = xp;
#pragma unroll
        for (int t = p + 1; t < FB; ++t) x[t] = fma(-xp, l11[warp][t][p], x[t]);
      }
      if (lane < nb) {
#pragma unroll
        for (int p = 0; p < FB; ++p)
          if (p < nb) __stcg(X + (kb + p) * ldx + kb + lane, (p >= lane) ? x[p] : 0.0);
      }
    }
  }
  fac_cluster_sync();
  // merge neighbouring s-blocks: [L11 0; L21 L22]^-1 = [X11 0; -X22 L21 X11  X22]; phase 0: W = L21 X11,
  // phase 1: X21 = -X22 W.  64 x 64 output tiles of every pair dealt round-robin to the CTAs.
  double* __restrict__ W = Q.W;
  for (int64_t s = FB; s < n; s *= 2) {
    const int64_t pairs = (n + 2 * s - 1) / (2 * s);
    const int64_t ts = (s + 63) / 64;
    for (int phase = 0; phase < 2; ++phase) {
      for (int64_t w = rank; w < pairs * ts * ts; w += C) {
        const int64_t pr = w / (ts * ts), tr = (w / ts) % ts, tc = w % ts;
        const int64_t c0 = pr * 2 * s, r0 = c0 + s;
        if (r0 >= n) continue;
        const int64_t m = (n - r0) < s ? (n - r0) : s;
        const int64_t row0 = tr * 64, col0 = tc * 64;
        if (row0 >= m) continue;
        const int64_t kdim = phase == 0 ? s : m;
        const double* A = phase == 0 ? L + r0 * ldl + c0 : X + r0 * ldx + r0;
        const int64_t lda = phase == 0 ? ldl : ldx;
        const double* Bm = phase == 0 ? X + c0 * ldx + c0 : W + r0 * n + c0;
        const int64_t ldb = phase == 0 ? ldx : n;
        double* Cm = phase == 0 ? W + r0 * n + c0 : X + r0 * ldx + c0;
        const int64_t ldc = phase == 0 ? n : ldx;
        const double sgn = phase == 0 ? 1.0 : -1.0;
        const int tx = tid & 15, ty = tid >> 4;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
        // triangular operands: X11 lower (phase 0: B[k][j] = 0 for k < j), X22 lower (phase 1: A[i][k] = 0 for k > i)
        const int64_t k_lo = phase == 0 ? (col0 / 16) * 16 : 0;
        const int64_t k_hi = phase == 0 ? kdim : ((row0 + 64) < kdim ? (row0 + 64) : kdim);
        for (int64_t kk0 = k_lo; kk0 < k_hi; kk0 += 16) {
          __syncthreads();
          for (int t = tid; t < 64 * 16; t += FT) {
            {
              const int kk = t % 16, i = t / 16;
              const int64_t gi = row0 + i, gk = kk0 + kk;
              tb.a[kk][i] = (gi < m && gk < kdim) ? __ldcg(A + gi * lda + gk) : 0.0;
            }
            {
              const int j = t % 64, kk = t / 64;
              const int64_t gj = col0 + j, gk = kk0 + kk;
              tb.b[kk][j] = (gj < s && gk < kdim) ? __ldcg(Bm + gk * ldb + gj) : 0.0;
            }
          }
          __syncthreads();
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) av[r] = tb.a[kk][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) bv[c] = tb.b[kk][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int64_t gi = row0 + ty * 4 + r;
          if (gi >= m) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int64_t gj = col0 + tx * 4 + c;
            if (gj < s) __stcg(Cm + gi * ldc + gj, sgn * acc[r][c]);
          }
        }
      }
      fac_cluster_sync();
    }
  }
  // alpha = X^T z ; tr K^-1 = |X|_F^2.  A thread owns one column of one row segment (consecutive threads =
  // consecutive columns: coalesced), loads go out eight at a time.  Partial sums go to the merge workspace
  // (free by now) and are added in a FIXED order afterwards: results are bitwise reproducible.
  {
    int64_t segs = gstride / n > 0 ? gstride / n : 1;
    if (segs > n) segs = n;                // the partials must fit the n x n workspace
    const int64_t seg_rows = (n + segs - 1) / segs;
    double* part = W;                      // [segs][n] partial alphas, then [C] partial traces
    double tr = 0.0;
    for (int64_t w = gt; w < segs * n; w += gstride) {
      const int64_t j = w % n, sg = w / n;
      int64_t i = sg * seg_rows;
      const int64_t i_end = (i + seg_rows < n) ? i + seg_rows : n;
      if (i < j) i = j;
      double acc = 0.0;
      for (; i + 8 <= i_end; i += 8) {
        double xv[8], zv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xv[u] = __ldcg(X + (i + u) * ldx + j); zv[u] = __ldcg(Q.z + i + u); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc = fma(xv[u], zv[u], acc); tr = fma(xv[u], xv[u], tr); }
      }
      for (; i < i_end; ++i) {
        const double x = __ldcg(X + i * ldx + j);
        acc = fma(x, __ldcg(Q.z + i), acc);
        tr = fma(x, x, tr);
      }
      __stcg(part + sg * n + j, acc);
    }
    const double t1 = block_sum(tr, red);
    if (tid == 0) __stcg(part + segs * n + rank, t1);
    fac_cluster_sync();
    double aa = 0.0;
    for (int64_t j = gt; j < n; j += gstride) {
      double a = 0.0;
      for (int64_t sg = 0; sg < segs; ++sg) a += __ldcg(part + sg * n + j);
      if (Q.alpha) Q.alpha[j] = a;
      __stcg(Q.alpha_acc + j, a);
    }
    fac_cluster_sync();
    if (rank == 0) {
      for (int64_t j = tid; j < n; j += FT) { const double a = __ldcg(Q.alpha_acc + j); aa = fma(a, a, aa); }
      const double t2 = block_sum(aa, red);
      if (tid == 0) {
        double trs = 0.0;
        for (uint32_t r = 0; r < C; ++r) trs += __ldcg(part + segs * n + r);
        Q.stats[2] = trs;
        Q.stats[3] = t2;
      }
    }
  }
}

int cluster_size_for(int64_t n) {
  if (n <= 64) return 1;
  if (n <= 128) return 2;
  if (n <= 256) return 4;
  if (n <= 1024) return 8;
  return 16;
}

int launch_factor(nbw_ctx* ctx, const FactorBatch& B, int n_problems, cudaStream_t st) {
  int C = cluster_size_for(B.n);
  if (const char* v = getenv("NBW_FACTOR_CLUSTER")) C = atoi(v);
  NBW_REQUIRE(ctx, C == 1 || C == 2 || C == 4 || C == 8 || C == 16, "cluster size must be 1, 2, 4, 8 or 16");
  if (C == 16 && !(ctx->attr_set & NBW_ATTR_FACTOR_NP)) {
    if (cudaFuncSetAttribute(gp_factor_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      C = 8;
    } else {
      ctx->attr_set |= NBW_ATTR_FACTOR_NP;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_problems * C));
  cfg.blockDim = dim3(FT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)C;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  NBW_CUDA(ctx, cudaLaunchKernelEx(&cfg, gp_factor_kernel, B));
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

}  // namespace

// Is the one-launch factorisation used for this size?  NBW_FACTOR_FUSED_MAX overrides the crossover (0 = never:
// the per-panel launches of linalg_f64.cu, which spread the trailing update over the whole GPU).
int64_t nbw_factor_fused_max() {
  if (const char* v = getenv("NBW_FACTOR_FUSED_MAX")) return atoll(v);
  return 256;     // measured (profiles/r02_factor_bench.json): on a par at n = 150..300, the per-panel launches win beyond
}
bool nbw_factor_fused_ok(int64_t n) { return n <= nbw_factor_fused_max(); }

// One full / likelihood-only factorisation of an fp64 matrix (the body of nbw_gp_factor).  Device results only.
// W: [n x n] workspace and alpha_acc: [n] when Linv is requested.
int nbw_factor_fused(nbw_ctx* ctx, const double* K, int64_t ldk, int64_t n, double lik, const double* y, double* L,
                     int64_t ldl, double* Linv, int64_t ldi, double* W, double* alpha, double* z, double* stats,
                     int* info, cudaStream_t st) {
  FactorBatch B;
  memset(&B, 0, sizeof(B));
  B.y = y;
  B.n = n;
  FactorProblem& Q = B.p[0];
  Q.K = K; Q.ldk = ldk; Q.shift = lik;
  Q.L = L; Q.ldl = ldl; Q.Linv = Linv; Q.ldi = ldi; Q.alpha = alpha; Q.z = z; Q.stats = stats; Q.info = info;
  if (Linv) {
    Q.alpha_acc = W;            // the first n doubles of the workspace; the merge buffer follows
    Q.W = W + ((n + 31) / 32) * 32;
  }
  return launch_factor(ctx, B, 1, st);
}

extern "C" int nbw_gp_grid(nbw_ctx* ctx, const int32_t* level_grams, int64_t ldg, int64_t level_stride, int64_t n,
                           const int* n_levels_h, int n_cands, double lik, const double* y, double* stats_h,
                           int* info_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_gp_grid");
  NBW_REQUIRE(ctx, n > 0 && ldg >= n && level_stride >= n * ldg, "bad shape");
  NBW_REQUIRE(ctx, n_cands >= 1 && n_cands <= kMaxProblems, "1..%d candidates per call", kMaxProblems);
  NBW_REQUIRE(ctx, level_grams && n_levels_h && y && stats_h && info_h, "null pointer");
  NBW_REQUIRE(ctx, n <= 8192, "n too large for the one-launch factorisation");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t per = Carver::pad(sizeof(double) * (size_t)n * (size_t)n) + Carver::pad(sizeof(double) * (size_t)n) + 1024;
  int rc = nbw_scratch_ensure(ctx, per * (size_t)n_cands + 4096, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  double* d_stats = cv.take<double>(4 * kMaxProblems);
  int* d_info = cv.take<int>(kMaxProblems);
  FactorBatch B;
  memset(&B, 0, sizeof(B));
  B.y = y;
  B.n = n;
  for (int c = 0; c < n_cands; ++c) {
    NBW_REQUIRE(ctx, n_levels_h[c] >= 1 && n_levels_h[c] <= NBW_MAX_LEVELS, "candidate %d: bad level count", c);
    FactorProblem& Q = B.p[c];
    Q.G = level_grams; Q.ldg = ldg; Q.g_stride = level_stride; Q.n_levels = n_levels_h[c];
    Q.shift = lik;
    Q.L = cv.take<double>((size_t)n * (size_t)n);
    Q.ldl = n;
    Q.z = cv.take<double>((size_t)n);
    Q.stats = d_stats + 4 * c;
    Q.info = d_info + c;
  }
  rc = launch_factor(ctx, B, n_cands, st);
  if (rc) return rc;
  double hs[4 * kMaxProblems];
  NBW_CUDA(ctx, cudaMemcpyAsync(hs, d_stats, sizeof(double) * 4 * n_cands, cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaMemcpyAsync(info_h, d_info, sizeof(int) * n_cands, cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  for (int c = 0; c < n_cands; ++c) {
    stats_h[2 * c] = hs[4 * c];
    stats_h[2 * c + 1] = hs[4 * c + 1];
  }
  return NBW_OK;
}
