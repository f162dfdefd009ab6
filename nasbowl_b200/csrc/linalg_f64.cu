// fp64 GP linear algebra on device: weighted Gram accumulation + cosine normalisation
// (kernels/combine_kernels.py:36-56), blocked Cholesky with pivot check + log-determinant,
// triangular inverse (bayesopt/utils.py:120-140, compute_pd_inverse) and the GEMMs of the GP
// posterior (bayesopt/gp.py:271-276).  Training sets are n <= a few thousand, so these are
// small, latency-bound problems: plain shared-memory tiling, no tensor cores (the reference
// evaluates them in fp32; fp64 here is what lets the posterior meet the 1e-6 parity target).
#include "nbw_common.cuh"

#include <math.h>

namespace {

constexpr int NB = 32;   // Cholesky / trtri block size

// ------------------------------------------------------------------ element-wise kernels
__global__ void gram_accumulate_kernel(const int32_t* __restrict__ K, int64_t ldk, int64_t m, int64_t n,
                                       double w, double beta, double* __restrict__ out, int64_t ldo) {
  const int64_t total = m * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / n, j = t % n;
    const double v = w * (double)K[i * ldk + j];
    out[i * ldo + j] = (beta == 0.0) ? v : beta * out[i * ldo + j] + v;
  }
}

__global__ void axpby_kernel(const double* __restrict__ X, int64_t ldx, int64_t m, int64_t n, double a, double b,
                             double* __restrict__ Y, int64_t ldy) {
  const int64_t total = m * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / n, j = t % n;
    const double v = a * X[i * ldx + j];
    Y[i * ldy + j] = (b == 0.0) ? v : b * Y[i * ldy + j] + v;
  }
}

__global__ void inv_sqrt_diag_kernel(const double* __restrict__ K, int64_t ld, int64_t n, double* __restrict__ s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) s[i] = 1.0 / sqrt(K[i * ld + i]);
}

__global__ void scale_sym_kernel(double* __restrict__ K, int64_t ld, int64_t n, const double* __restrict__ s) {
  const int64_t total = n * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / n, j = t % n;
    K[i * ld + j] = K[i * ld + j] * s[i] * s[j];
  }
}

__global__ void scale_features_kernel(const int8_t* __restrict__ phi, int64_t ld_phi, int64_t n, int64_t k0,
                                      int64_t width, const double* __restrict__ s, double sqrt_w,
                                      double* __restrict__ B, int64_t ldb) {
  const int64_t total = n * width, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / width, j = t % width;
    B[i * ldb + j] = s[i] * sqrt_w * (double)phi[i * ld_phi + k0 + j];
  }
}

__global__ void posterior_cov_kernel(const double* __restrict__ Kss, int64_t ldk, const double* __restrict__ Q,
                                     int64_t ldq, int64_t m, double lik, double y_std,
                                     double* __restrict__ cov, int64_t ldc) {
  const int64_t total = m * m, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / m, j = t % m;
    double c = Kss[i * ldk + j] + (i == j ? lik : 0.0) - Q[i * ldq + j];
    c = fmax(c, lik);                     // torch.clamp(cov_s, likelihood, inf), gp.py:276
    const double sd = sqrt(c) * y_std;    // gp.py:278-279
    cov[i * ldc + j] = sd * sd;           // gp.py:280
  }
}

// sum of squares of an m x n block -> *out (double, atomically accumulated; caller zeroes)
__global__ void sumsq_kernel(const double* __restrict__ X, int64_t ld, int64_t m, int64_t n, double* out) {
  const int64_t total = m * n, stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const double v = X[(t / n) * ld + (t % n)];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) ws[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}

// ------------------------------------------------------------------ GEMM
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f64_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double* __restrict__ A, int64_t lda,
                const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C, int64_t ldc) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ double As[BK][BM + 2];
  __shared__ double Bs[BK][BN + 2];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;

  for (int64_t k0 = 0; k0 < k; k0 += BK) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int idx = threadIdx.x + s * 256;
      {
        int i, kk;
        if (TA) { i = idx % BM; kk = idx / BM; } else { kk = idx % BK; i = idx / BK; }
        const int64_t gi = row0 + i, gk = k0 + kk;
        double v = 0.0;
        if (gi < m && gk < k) v = TA ? A[gk * lda + gi] : A[gi * lda + gk];
        As[kk][i] = v;
      }
      {
        int j, kk;
        if (TB) { kk = idx % BK; j = idx / BK; } else { j = idx % BN; kk = idx / BN; }
        const int64_t gj = col0 + j, gk = k0 + kk;
        double v = 0.0;
        if (gj < n && gk < k) v = TB ? B[gj * ldb + gk] : B[gk * ldb + gj];
        Bs[kk][j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = row0 + ty * 4 + r;
    if (gi >= m) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t gj = col0 + tx * 4 + c;
      if (gj >= n) continue;
      const double v = alpha * acc[r][c];
      C[gi * ldc + gj] = (beta == 0.0) ? v : beta * C[gi * ldc + gj] + v;
    }
  }
}

int gemm_enqueue(nbw_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0) return NBW_OK;
  dim3 grid((unsigned)((n + 63) / 64), (unsigned)((m + 63) / 64));
  NBW_REQUIRE(ctx, grid.y <= 65535, "gemm: m too large");
  if (!ta && !tb) gemm_f64_kernel<false, false><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (ta && !tb) gemm_f64_kernel<true, false><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!ta && tb) gemm_f64_kernel<false, true><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else gemm_f64_kernel<true, true><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

// ------------------------------------------------------------------ Cholesky (lower, blocked)
__global__ void chol_init_kernel(const double* __restrict__ A, int64_t lda, int64_t n, double shift,
                                 double* __restrict__ L, int64_t ldl) {
  const int64_t total = n * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / n, j = t % n;
    L[i * ldl + j] = (j > i) ? 0.0 : A[i * lda + j] + (i == j ? shift : 0.0);
  }
}

// One panel step: factor the nb x nb diagonal block at (k0, k0) and solve the rows below it,
// X = A21 L11^-T.  Measured on B200, the three-kernel form of this step cost ~40 us per panel whatever
// n was: the time went into serial chains (a 32-step block-synchronous diagonal factorisation and a
// 32 x 16 deep dependent FMA chain per row), not into launches or flops.  Here
//   - warp 0 of EVERY CTA factors the diagonal block warp-synchronously: lane r keeps row r in
//     registers, the loops are fully unrolled, the column a(., j) travels by shuffle — no block barrier
//     inside the 32 steps; CTA 0 writes the factor back;
//   - the row solve is right-looking (x_j known -> all later x_t updated independently), which
//     turns the long dependent chain into 32 short ones.
__global__ void __launch_bounds__(128)
panel_kernel(double* __restrict__ L, int64_t ldl, int64_t k0, int nb, int64_t n, int* __restrict__ info) {
  __shared__ double l11[NB][NB + 1];
  __shared__ double inv_diag[NB];   // 1 / L11[j][j]: the chains below multiply instead of dividing
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    // right-looking in registers, the scaled pivot column broadcast through shared memory (see gp_factor.cu)
    __shared__ double colbuf[NB];
    double a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = (lane < nb && c <= lane && c < nb) ? L[(k0 + lane) * ldl + k0 + c] : 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j < nb) {                                            // uniform
        if (lane == j) {
          const double ajj = a[j];
          if (!(ajj > 0.0) && blockIdx.x == 0) atomicCAS(info, 0, (int)(k0 + j + 1));
          const double rinv = rsqrt(ajj);          // fp64 sqrt and divide are long dependent sequences
          a[j] = ajj * rinv;
          inv_diag[j] = rinv;
        }
        __syncwarp();
        const double rinv = inv_diag[j];
        if (lane > j) a[j] *= rinv;
        colbuf[lane] = (lane >= j) ? a[j] : 0.0;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < NB; ++c)
          if (lane >= c) a[c] = fma(-a[j], colbuf[c], a[c]);
        __syncwarp();
      }
    }
    if (lane >= nb) inv_diag[lane] = 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      l11[lane][c] = (c <= lane) ? a[c] : 0.0;
      if (blockIdx.x == 0 && lane < nb && c <= lane && c < nb) L[(k0 + lane) * ldl + k0 + c] = a[c];
    }
  }
  __syncthreads();
  const int64_t i = k0 + nb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[NB];
  double* row = L + i * ldl + k0;
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? row[j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (j < nb) {
      const double xj = x[j] * inv_diag[j];
      x[j] = xj;
#pragma unroll
      for (int t = j + 1; t < NB; ++t) x[t] = fma(-xj, l11[t][j], x[t]);
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) row[j] = x[j];
}

// Trailing update A22 -= P P^T on the lower triangle, P = L[k0+nb:, k0:k0+nb].
__global__ void __launch_bounds__(256)
syrk_update_kernel(double* __restrict__ L, int64_t ldl, int64_t k0, int nb, int64_t n) {
  const int64_t base = k0 + nb;
  const int64_t i0 = base + (int64_t)blockIdx.y * 64, j0 = base + (int64_t)blockIdx.x * 64;
  if (j0 > i0 + 63) return;   // tile strictly above the diagonal
  __shared__ double Pi[NB][64 + 2];
  __shared__ double Pj[NB][64 + 2];
  for (int t = threadIdx.x; t < 64 * NB; t += 256) {
    const int kk = t % NB, r = t / NB;
    const int64_t gi = i0 + r, gj = j0 + r;
    Pi[kk][r] = (gi < n && kk < nb) ? L[gi * ldl + k0 + kk] : 0.0;
    Pj[kk][r] = (gj < n && kk < nb) ? L[gj * ldl + k0 + kk] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll 8
  for (int kk = 0; kk < NB; ++kk) {
    double a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = Pi[kk][ty * 4 + r];
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = Pj[kk][tx * 4 + c];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = i0 + ty * 4 + r;
    if (gi >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t gj = j0 + tx * 4 + c;
      if (gj <= gi) L[gi * ldl + gj] -= acc[r][c];
    }
  }
}

__global__ void logdet_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += log(L[i * ldl + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) ws[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) *out = 2.0 * acc;
  }
}

int chol_enqueue(nbw_ctx* ctx, const double* A, int64_t lda, int64_t n, double shift, double* L, int64_t ldl,
                 double* d_logdet, int* d_info, cudaStream_t st) {
  NBW_CUDA(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), st));
  chol_init_kernel<<<nbw_grid_for(ctx, n * n, 256), 256, 0, st>>>(A, lda, n, shift, L, ldl);
  NBW_CHECK_LAUNCH(ctx);
  for (int64_t k0 = 0; k0 < n; k0 += NB) {
    const int nb = (int)((n - k0) < NB ? (n - k0) : NB);
    const int64_t rest = n - k0 - nb;
    panel_kernel<<<(unsigned)(rest > 0 ? (rest + 127) / 128 : 1), 128, 0, st>>>(L, ldl, k0, nb, n, d_info);
    NBW_CHECK_LAUNCH(ctx);
    if (rest > 0) {
      const unsigned tiles = (unsigned)((rest + 63) / 64);
      syrk_update_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(L, ldl, k0, nb, n);
      NBW_CHECK_LAUNCH(ctx);
    }
  }
  logdet_kernel<<<1, 256, 0, st>>>(L, ldl, n, d_logdet);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

// ------------------------------------------------------------------ triangular inverse
// X = L^-1 by recursive doubling: invert the 32 x 32 diagonal blocks (one warp each, right-looking
// substitution out of shared memory), then for s = 32, 64, 128, ... merge neighbouring s-blocks,
//     [ L11  0  ]^-1   [ X11            0  ]
//     [ L21 L22 ]    = [ -X22 L21 X11  X22 ],
// as two batched block products per level (W = L21 X11, X21 = -X22 W; all block pairs of a level in one
// launch).  1 + 2 log2(n / 32) launches of GEMM-shaped work replace the block-column sweep this file
// used before, whose n / 32 sequential steps per column made it the longest piece of the factorisation
// (1.4 ms of 3.2 ms at n = 1000).
static inline size_t trtri_front_bytes(int64_t n) { return Carver::pad(sizeof(double) * (size_t)n) + 8192; }
static inline size_t trtri_scratch_bytes(int64_t n) {
  // (+ n + 32 doubles: the one-launch factorisation keeps its alpha accumulator in front of the merge buffer)
  return trtri_front_bytes(n) + Carver::pad(sizeof(double) * ((size_t)n * (size_t)n + (size_t)n + 128));
}

__global__ void __launch_bounds__(32)
trtri_diag_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, double* __restrict__ X, int64_t ldx) {
  __shared__ double l[NB][NB + 1];
  const int lane = threadIdx.x;
  const int64_t k0 = (int64_t)blockIdx.x * NB;
  const int nb = (int)((n - k0) < NB ? (n - k0) : NB);
  for (int r = 0; r < NB; ++r) l[r][lane] = (r < nb && lane <= r) ? L[(k0 + r) * ldl + k0 + lane] : (r == lane ? 1.0 : 0.0);
  __syncwarp();
  // lane j solves L x = e_j
  double x[NB];
#pragma unroll
  for (int p = 0; p < NB; ++p) x[p] = (p == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int p = 0; p < NB; ++p) {
    const double xp = x[p] / l[p][p];
    x[p] = xp;
#pragma unroll
    for (int t = p + 1; t < NB; ++t) x[t] = fma(-xp, l[t][p], x[t]);
  }
  if (lane < nb) {
#pragma unroll
    for (int p = 0; p < NB; ++p)
      if (p < nb) X[(k0 + p) * ldx + k0 + lane] = (p >= lane) ? x[p] : 0.0;
  }
}

// phase 0: W21 = L21 * X11 ; phase 1: X21 = -X22 * W21, for every pair of neighbouring s-blocks (blockIdx.z)
__global__ void __launch_bounds__(256)
trtri_merge_kernel(int phase, int64_t s, int64_t n, const double* __restrict__ L, int64_t ldl, double* X, int64_t ldx,
                   double* W, int64_t ldw) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ double As[BK][BM + 2];
  __shared__ double Bs[BK][BN + 2];
  const int64_t c0 = (int64_t)blockIdx.z * 2 * s, r0 = c0 + s;
  if (r0 >= n) return;
  const int64_t m = (n - r0) < s ? (n - r0) : s;       // rows of the lower-left block
  const int64_t kdim = phase == 0 ? s : m;
  const double* A = phase == 0 ? L + r0 * ldl + c0 : X + r0 * ldx + r0;
  const int64_t lda = phase == 0 ? ldl : ldx;
  const double* B = phase == 0 ? X + c0 * ldx + c0 : W + r0 * ldw + c0;
  const int64_t ldb = phase == 0 ? ldx : ldw;
  double* C = phase == 0 ? W + r0 * ldw + c0 : X + r0 * ldx + c0;
  const int64_t ldc = phase == 0 ? ldw : ldx;
  const double alpha = phase == 0 ? 1.0 : -1.0;
  const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
  if (row0 >= m) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  // triangular operands: X11 is lower (phase 0: B[k][j] = 0 for k < j), X22 is lower (phase 1: A[i][k] = 0 for k > i)
  const int64_t k_lo = phase == 0 ? (col0 / BK) * BK : 0;
  const int64_t k_hi = phase == 0 ? kdim : ((row0 + BM) < kdim ? (row0 + BM) : kdim);
  for (int64_t k0 = k_lo; k0 < k_hi; k0 += BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + q * 256;
      {
        const int kk = idx % BK, i = idx / BK;
        const int64_t gi = row0 + i, gk = k0 + kk;
        As[kk][i] = (gi < m && gk < kdim) ? A[gi * lda + gk] : 0.0;
      }
      {
        const int j = idx % BN, kk = idx / BN;
        const int64_t gj = col0 + j, gk = k0 + kk;
        Bs[kk][j] = (gj < s && gk < kdim) ? B[gk * ldb + gj] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bs[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = row0 + ty * 4 + r;
    if (gi >= m) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t gj = col0 + tx * 4 + c;
      if (gj < s) C[gi * ldc + gj] = alpha * acc[r][c];
    }
  }
}

int trtri_enqueue(nbw_ctx* ctx, const double* L, int64_t ldl, int64_t n, double* X, int64_t ldx, cudaStream_t st) {
  int rc = nbw_scratch_ensure(ctx, trtri_scratch_bytes(n), st);
  if (rc) return rc;
  // W lives behind the front region, where nbw_gp_factor keeps its statistics and one n-vector
  double* W = reinterpret_cast<double*>(static_cast<char*>(ctx->scratch) + trtri_front_bytes(n));
  NBW_CUDA(ctx, cudaMemset2DAsync(X, ldx * sizeof(double), 0, n * sizeof(double), n, st));
  trtri_diag_kernel<<<(unsigned)((n + NB - 1) / NB), 32, 0, st>>>(L, ldl, n, X, ldx);
  NBW_CHECK_LAUNCH(ctx);
  for (int64_t s = NB; s < n; s *= 2) {
    const unsigned pairs = (unsigned)((n + 2 * s - 1) / (2 * s));
    const unsigned tiles = (unsigned)((s + 63) / 64);
    for (int phase = 0; phase < 2; ++phase) {
      trtri_merge_kernel<<<dim3(tiles, tiles, pairs), 256, 0, st>>>(phase, s, n, L, ldl, X, ldx, W, n);
      NBW_CHECK_LAUNCH(ctx);
    }
  }
  return NBW_OK;
}

// z = L^-1 y for ONE right-hand side (the grid search over h needs y^T K^-1 y = |z|^2 and log|K| only,
// not the inverse): one CTA walks the 32-row blocks; eight warps form the products with the part of z
// already known, warp 0 finishes the block by forward substitution.
__global__ void __launch_bounds__(256)
trsv_lower_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ y, double* z) {
  extern __shared__ double zs[];            // the part of z already known (n doubles)
  __shared__ double rhs[NB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r0 = 0; r0 < n; r0 += NB) {
    for (int i = warp; i < NB; i += 8) {
      const int64_t row = r0 + i;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // four loads in flight per lane
      if (row < n) {
        const double* lr = L + row * ldl;
        int64_t k = lane;
        for (; k + 96 < r0; k += 128) {
          a0 = fma(lr[k], zs[k], a0);
          a1 = fma(lr[k + 32], zs[k + 32], a1);
          a2 = fma(lr[k + 64], zs[k + 64], a2);
          a3 = fma(lr[k + 96], zs[k + 96], a3);
        }
        for (; k < r0; k += 32) a0 = fma(lr[k], zs[k], a0);
      }
      double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) rhs[i] = row < n ? y[row] - acc : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      const int64_t row = r0 + lane;
      // this lane's row of the diagonal block, fetched before the serial chain starts
      double lrow[NB];
#pragma unroll
      for (int p = 0; p < NB; ++p) lrow[p] = (row < n && p <= lane && r0 + p < n) ? L[row * ldl + r0 + p] : 0.0;
      double dinv = 0.0;
#pragma unroll
      for (int p = 0; p < NB; ++p)
        if (p == lane) dinv = (row < n) ? 1.0 / lrow[p] : 1.0;
      double v = rhs[lane];
#pragma unroll
      for (int p = 0; p < NB; ++p) {
        const double zp = __shfl_sync(0xffffffffu, v * dinv, p);
        if (lane == p) v = zp;
        else if (lane > p) v = fma(-lrow[p], zp, v);
      }
      if (row < n) { z[row] = v; zs[row] = v; }
    }
    __syncthreads();
  }
}

}  // namespace

// ====================================================================== C-ABI
extern "C" int nbw_gram_accumulate(nbw_ctx* ctx, const int32_t* K, int64_t ldk, int64_t m, int64_t n,
                                   double w, double beta, double* out, int64_t ldo, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, m >= 0 && n >= 0 && ldk >= n && ldo >= n, "bad shape");
  if (m == 0 || n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, K && out, "null pointer");
  gram_accumulate_kernel<<<nbw_grid_for(ctx, m * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      K, ldk, m, n, w, beta, out, ldo);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_axpby_f64(nbw_ctx* ctx, const double* X, int64_t ldx, int64_t m, int64_t n, double a, double b,
                             double* Y, int64_t ldy, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, m >= 0 && n >= 0 && ldx >= n && ldy >= n, "bad shape");
  if (m == 0 || n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, X && Y, "null pointer");
  axpby_kernel<<<nbw_grid_for(ctx, m * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, m, n, a, b, Y, ldy);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_gram_normalize(nbw_ctx* ctx, double* K, int64_t ld, int64_t n, double* inv_sqrt_diag,
                                  void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n >= 0 && ld >= n, "bad shape");
  if (n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, K && inv_sqrt_diag, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  inv_sqrt_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(K, ld, n, inv_sqrt_diag);
  NBW_CHECK_LAUNCH(ctx);
  scale_sym_kernel<<<nbw_grid_for(ctx, n * n, 256), 256, 0, st>>>(K, ld, n, inv_sqrt_diag);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_scale_features(nbw_ctx* ctx, const int8_t* phi, int64_t ld_phi, int64_t n, int64_t k0,
                                  int64_t k1, const double* s, double sqrt_w, double* B, int64_t ldb,
                                  void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n >= 0 && k1 >= k0 && k0 >= 0 && ld_phi >= k1 && ldb >= k1 - k0, "bad shape");
  if (n == 0 || k1 == k0) return NBW_OK;
  NBW_REQUIRE(ctx, phi && s && B, "null pointer");
  scale_features_kernel<<<nbw_grid_for(ctx, n * (k1 - k0), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      phi, ld_phi, n, k0, k1 - k0, s, sqrt_w, B, ldb);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_posterior_cov_finish(nbw_ctx* ctx, const double* Kss, int64_t ldk, const double* Q,
                                        int64_t ldq, int64_t m, double lik, double y_std, double* cov,
                                        int64_t ldc, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, m >= 0 && ldk >= m && ldq >= m && ldc >= m, "bad shape");
  if (m == 0) return NBW_OK;
  NBW_REQUIRE(ctx, Kss && Q && cov, "null pointer");
  posterior_cov_kernel<<<nbw_grid_for(ctx, m * m, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      Kss, ldk, Q, ldq, m, lik, y_std, cov, ldc);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_gemm_f64(nbw_ctx* ctx, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
                            double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                            double beta, double* C, int64_t ldc, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0 && ldc >= n, "bad shape");
  NBW_REQUIRE(ctx, lda >= (trans_a ? m : k) && ldb >= (trans_b ? k : n), "bad leading dimension");
  NBW_REQUIRE(ctx, (m == 0 || n == 0) || (A && B && C) || k == 0, "null pointer");
  return gemm_enqueue(ctx, trans_a, trans_b, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
                      static_cast<cudaStream_t>(stream));
}

extern "C" int nbw_chol_f64(nbw_ctx* ctx, const double* A, int64_t lda, int64_t n, double shift, double* L,
                            int64_t ldl, double* logdet_h, int* info_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_chol_f64");
  NBW_REQUIRE(ctx, n > 0 && lda >= n && ldl >= n && A && L && info_h, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = nbw_scratch_ensure(ctx, 1024, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  double* d_logdet = cv.take<double>(1);
  int* d_info = cv.take<int>(1);
  rc = chol_enqueue(ctx, A, lda, n, shift, L, ldl, d_logdet, d_info, st);
  if (rc) return rc;
  double ld = 0.0;
  NBW_CUDA(ctx, cudaMemcpyAsync(&ld, d_logdet, sizeof(double), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaMemcpyAsync(info_h, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  if (logdet_h) *logdet_h = ld;
  if (*info_h != 0) NBW_FAIL(ctx, NBW_ERR_NOT_PD, "matrix not positive definite at pivot %d", *info_h - 1);
  return NBW_OK;
}

extern "C" int nbw_trtri_lower_f64(nbw_ctx* ctx, const double* L, int64_t ldl, int64_t n, double* Linv,
                                   int64_t ldi, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n > 0 && ldl >= n && ldi >= n && L && Linv, "bad argument");
  return trtri_enqueue(ctx, L, ldl, n, Linv, ldi, static_cast<cudaStream_t>(stream));
}

extern "C" int nbw_gp_factor(nbw_ctx* ctx, const double* K, int64_t ldk, int64_t n, double lik,
                             const double* y, double* L, int64_t ldl, double* Linv, int64_t ldi,
                             double* alpha, double* stats_h, int* info_h, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_gp_factor");
  NBW_REQUIRE(ctx, n > 0 && ldk >= n && ldl >= n && (Linv == nullptr || ldi >= n), "bad shape");
  NBW_REQUIRE(ctx, K && y && L && stats_h && info_h, "null pointer");
  NBW_REQUIRE(ctx, (Linv == nullptr) == (alpha == nullptr), "Linv and alpha are requested together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // everything this call carves out of the arena is reserved up front: a later growth would move it
  int rc = nbw_scratch_ensure(ctx, Linv ? trtri_scratch_bytes(n) : Carver::pad(sizeof(double) * n) + 4096, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  double* d_stats = cv.take<double>(4);   // logdet, yKy, trKinv, |alpha|^2
  int* d_info = cv.take<int>(1);
  double* z = cv.take<double>(n);
  if (nbw_factor_fused_ok(n)) {
    // one launch: a cluster of CTAs walks the panels, the inverse merges and the reductions (gp_factor.cu)
    double* W = Linv ? reinterpret_cast<double*>(static_cast<char*>(ctx->scratch) + trtri_front_bytes(n)) : nullptr;
    rc = nbw_factor_fused(ctx, K, ldk, n, lik, y, L, ldl, Linv, ldi, W, alpha, z, d_stats, d_info, st);
    if (rc) return rc;
    double hs4[4];
    NBW_CUDA(ctx, cudaMemcpyAsync(hs4, d_stats, sizeof(hs4), cudaMemcpyDeviceToHost, st));
    NBW_CUDA(ctx, cudaMemcpyAsync(info_h, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    NBW_CUDA(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < 4; ++i) stats_h[i] = hs4[i];
    if (*info_h != 0) NBW_FAIL(ctx, NBW_ERR_NOT_PD, "matrix not positive definite at pivot %d", *info_h - 1);
    return NBW_OK;
  }
  NBW_CUDA(ctx, cudaMemsetAsync(d_stats, 0, sizeof(double) * 4, st));
  rc = chol_enqueue(ctx, K, ldk, n, lik, L, ldl, d_stats + 0, d_info, st);
  if (rc) return rc;
  if (Linv) {
    rc = trtri_enqueue(ctx, L, ldl, n, Linv, ldi, st);
    if (rc) return rc;
    // z = Linv y ; alpha = Linv^T z
    rc = gemm_enqueue(ctx, 0, 0, n, 1, n, 1.0, Linv, ldi, y, 1, 0.0, z, 1, st);
    if (rc) return rc;
    rc = gemm_enqueue(ctx, 1, 0, n, 1, n, 1.0, Linv, ldi, z, 1, 0.0, alpha, 1, st);
    if (rc) return rc;
  } else {
    // likelihood only: no inverse.  z is kept in shared memory, which bounds n for this path.
    NBW_REQUIRE(ctx, (size_t)n * sizeof(double) <= 200 * 1024, "likelihood-only factorisation handles n <= 25600");
    if (!(ctx->attr_set & NBW_ATTR_TRSV)) {
      NBW_CUDA(ctx, cudaFuncSetAttribute(trsv_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      ctx->attr_set |= NBW_ATTR_TRSV;
    }
    trsv_lower_kernel<<<1, 256, (size_t)n * sizeof(double), st>>>(L, ldl, n, y, z);
    NBW_CHECK_LAUNCH(ctx);
  }
  sumsq_kernel<<<1, 256, 0, st>>>(z, 1, n, 1, d_stats + 1);
  NBW_CHECK_LAUNCH(ctx);
  if (Linv) {
    sumsq_kernel<<<nbw_grid_for(ctx, n * n, 256, 2), 256, 0, st>>>(Linv, ldi, n, n, d_stats + 2);
    NBW_CHECK_LAUNCH(ctx);
    sumsq_kernel<<<1, 256, 0, st>>>(alpha, 1, n, 1, d_stats + 3);
    NBW_CHECK_LAUNCH(ctx);
  }
  double hs[4];
  NBW_CUDA(ctx, cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaMemcpyAsync(info_h, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  stats_h[0] = hs[0];
  stats_h[1] = hs[1];
  stats_h[2] = hs[2];
  stats_h[3] = hs[3];
  if (*info_h != 0) NBW_FAIL(ctx, NBW_ERR_NOT_PD, "matrix not positive definite at pivot %d", *info_h - 1);
  return NBW_OK;
}

// ====================================================================== gradient of the marginal likelihood w.r.t. kernel weights
// GraphGP.fit trains the weights of a sum of kernels by autograd through
//   K = normalise(sum_i w_i R_i) + lik I,  nlml' = (0.5 y^T K^-1 y - 0.5 log|K| + c) / n
// (bayesopt/gp.py:141-210, kernels/combine_kernels.py:36-56, utils.py:102-140 incl. the sign of the
// log-det term).  Closed form, with M = alpha alpha^T + K^-1, S = sum_i w_i R_i, s_j = S_jj^-1/2 and
// N = normalise(S):
//   d nlml' / d w_i = -1/(2n) [ sum_jk M_jk R_i,jk s_j s_k  -  sum_j R_i,jj s_j^2 (sum_k M_jk N_jk) ]
namespace {

// rowdot[j] = sum_k (alpha_j alpha_k + Kinv_jk) N_jk ; one block per row
__global__ void __launch_bounds__(256)
weight_grad_rowdot_kernel(const double* __restrict__ Kinv, int64_t ldk, const double* __restrict__ alpha,
                          const double* __restrict__ N, int64_t ldn, int64_t n, double* __restrict__ rowdot) {
  __shared__ double red[256];
  const int64_t j = blockIdx.x;
  const double aj = alpha[j];
  double acc = 0.0;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) acc += (aj * alpha[k] + Kinv[j * ldk + k]) * N[j * ldn + k];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowdot[j] = red[0];
}

// out += sum_k M_jk R_jk s_j s_k - R_jj s_j^2 rowdot_j for row j = blockIdx.x
__global__ void __launch_bounds__(256)
weight_grad_kernel(const double* __restrict__ Kinv, int64_t ldk, const double* __restrict__ alpha,
                   const double* __restrict__ R, int64_t ldr, const double* __restrict__ s,
                   const double* __restrict__ rowdot, int64_t n, double* out) {
  __shared__ double red[256];
  const int64_t j = blockIdx.x;
  const double aj = alpha[j], sj = s[j];
  double acc = 0.0;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x)
    acc += (aj * alpha[k] + Kinv[j * ldk + k]) * R[j * ldr + k] * sj * s[k];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, red[0] - R[j * ldr + j] * sj * sj * rowdot[j]);
}

}  // namespace

extern "C" int nbw_kernel_weight_grad(nbw_ctx* ctx, const double* Kinv, int64_t ldk, const double* alpha,
                                      const double* N, int64_t ldn, const double* inv_sqrt_diag, int64_t n,
                                      const double* const* R_h, int64_t ldr, int n_kernels, double* grad_h,
                                      void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_kernel_weight_grad");
  NBW_REQUIRE(ctx, n > 0 && ldk >= n && ldn >= n && ldr >= n && n_kernels >= 1 && n_kernels <= 64, "bad shape");
  NBW_REQUIRE(ctx, n < 65536, "n too large for one block per row");
  NBW_REQUIRE(ctx, Kinv && alpha && N && inv_sqrt_diag && R_h && grad_h, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = nbw_scratch_ensure(ctx, Carver::pad(sizeof(double) * n) + 4096, st);
  if (rc) return rc;
  Carver cv(ctx->scratch);
  double* d_out = cv.take<double>(64);
  double* rowdot = cv.take<double>(n);
  NBW_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(double) * 64, st));
  weight_grad_rowdot_kernel<<<(unsigned)n, 256, 0, st>>>(Kinv, ldk, alpha, N, ldn, n, rowdot);
  NBW_CHECK_LAUNCH(ctx);
  for (int i = 0; i < n_kernels; ++i) {
    NBW_REQUIRE(ctx, R_h[i] != nullptr, "R_h[%d] is NULL", i);
    weight_grad_kernel<<<(unsigned)n, 256, 0, st>>>(Kinv, ldk, alpha, R_h[i], ldr, inv_sqrt_diag, rowdot, n, d_out + i);
    NBW_CHECK_LAUNCH(ctx);
  }
  double hs[64];
  NBW_CUDA(ctx, cudaMemcpyAsync(hs, d_out, sizeof(hs), cudaMemcpyDeviceToHost, st));
  NBW_CUDA(ctx, cudaStreamSynchronize(st));
  for (int i = 0; i < n_kernels; ++i) grad_h[i] = -hs[i] / (2.0 * (double)n);
  return NBW_OK;
}

// ====================================================================== prefix form of the pool model
namespace {

__global__ void label_parents_kernel(const uint32_t* __restrict__ ids_prev, const uint32_t* __restrict__ ids_cur,
                                     int64_t n_keys, uint32_t* __restrict__ parent) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += stride) {
    const uint32_t c = ids_cur[i];
    if (c != NBW_INVALID_ID) parent[c] = ids_prev[i];   // every writer of a slot writes the same value
  }
}

__global__ void prefix_level_kernel(const double* __restrict__ B, int64_t ldb, int64_t n, int64_t col0,
                                    int64_t lab0, int64_t lab_prev0, int64_t width,
                                    const uint32_t* __restrict__ parent, double* Bp, int64_t ldp, int level) {
  const int64_t total = n * width, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / width, z = t % width;
    double v = B[i * ldb + col0 + z];
    if (level > 0) v += Bp[i * ldp + lab_prev0 + parent[lab0 + z]];
    Bp[i * ldp + lab0 + z] = v;
  }
}

}  // namespace

extern "C" int nbw_label_parents(nbw_ctx* ctx, const uint32_t* ids_prev, const uint32_t* ids_cur, int64_t n_keys,
                                 uint32_t* parent, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n_keys >= 0, "n_keys negative");
  if (n_keys == 0) return NBW_OK;
  NBW_REQUIRE(ctx, ids_prev && ids_cur && parent, "null pointer");
  label_parents_kernel<<<nbw_grid_for(ctx, n_keys, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids_prev, ids_cur, n_keys, parent);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_prefix_features(nbw_ctx* ctx, const double* B, int64_t ldb, int64_t n, int n_levels,
                                   const int64_t* col_off_h, const int64_t* label_off_h, const uint32_t* parent,
                                   double* Bp, int64_t ldp, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n_levels >= 1 && n_levels <= NBW_MAX_LEVELS && col_off_h && label_off_h, "bad levels");
  NBW_REQUIRE(ctx, n >= 0 && ldp >= label_off_h[n_levels] && ldb >= col_off_h[n_levels - 1], "bad shape");
  if (n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, B && Bp && (n_levels == 1 || parent), "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int l = 0; l < n_levels; ++l) {
    const int64_t width = label_off_h[l + 1] - label_off_h[l];
    if (width <= 0) continue;
    prefix_level_kernel<<<nbw_grid_for(ctx, n * width, 256), 256, 0, st>>>(
        B, ldb, n, col_off_h[l], label_off_h[l], l > 0 ? label_off_h[l - 1] : 0, width, parent, Bp, ldp, l);
    NBW_CHECK_LAUNCH(ctx);
  }
  return NBW_OK;
}

// ====================================================================== item features of the packed OA path
namespace {
__global__ void combine_columns_kernel(const double* __restrict__ B, int64_t n, int64_t ldb, const int32_t* __restrict__ idx,
                                       int64_t J, int K, double* __restrict__ out, int64_t ldo) {
  const int64_t total = n * J, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / J, j = t % J;
    const double* __restrict__ row = B + i * ldb;
    double v = 0.0;
    for (int k = 0; k < K; ++k) {
      const int32_t c = idx[j * K + k];
      if (c == 0x7FFFFFFF) continue;
      v += c >= 0 ? row[c] : -row[~c];
    }
    out[i * ldo + j] = v;
  }
}
}  // namespace

extern "C" int nbw_combine_columns(nbw_ctx* ctx, const double* B, int64_t n, int64_t ldb, const int32_t* idx, int64_t J,
                                   int K, double* out, int64_t ldo, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_REQUIRE(ctx, n >= 0 && J >= 0 && K >= 1 && K <= 64 && ldo >= J, "bad shape");
  if (n == 0 || J == 0) return NBW_OK;
  NBW_REQUIRE(ctx, B && idx && out, "null pointer");
  combine_columns_kernel<<<nbw_grid_for(ctx, n * J, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(B, n, ldb, idx, J, K,
                                                                                                      out, ldo);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}
