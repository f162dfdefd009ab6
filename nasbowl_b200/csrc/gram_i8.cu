// Exact integer Gram matrix  C[i,j] = sum_k A[i,k] * B[j,k]  (int8 x int8 -> int32).
// Replaces VertexHistogram._calculate_kernel_matrix (grakel_replace/vertex_histogram.py:211-259,
// a float64 BLAS dgemm over integer-valued histograms) and the per-level sum of
// weisfeiler_lehman.py:316-327: WL levels are contiguous column blocks of Phi, so the sum over
// levels 0..h is ONE GEMM over the K-range [0, col_off[h+1]).
//
// Product path (impl 0): tcgen05.mma kind::i8, cta_group::2 — a cluster of two CTAs (one TPC) owns a
// 256 x 256 output tile, accumulators in TMEM (128 lanes x 256 columns of s32 per CTA, two buffers),
// both operands K-major, staged by TMA into 128B-swizzled shared memory (one 128-byte K-block per
// row per stage; each CTA stages its 128 rows of A and its half of B), 6-stage mbarrier pipeline,
// persistent CTA pairs walking the tiles in a grouped order:
//   warp 0 lane 0 : TMA producer          (cp.async.bulk.tensor.2d -> the leader's full barrier)
//   warp 1 lane 0 : MMA issuer (leader)   (4 x UMMA_K=32 per stage, multicast tcgen05.commit)
//   warps 2..5    : epilogue              (tcgen05.ld 32x32b.x32 -> 128-byte row segments to HBM)
// int32 accumulation of int8 products is exact.  The single-CTA persistent kernel (cta_group::1) it
// replaced is kept as NBW_GRAM_VARIANT=4 for comparison.
//
// Cross-check path (impl 1): plain dp4a kernel, used by the tests to validate impl 0.
#include "nbw_common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128, BK = 128;   // BK in bytes == int8 elements == one swizzle row
constexpr int UMMA_K = 32;
constexpr int MAX_STAGES = 8;
constexpr int A_TILE_BYTES = BM * BK;          // 16 KB of A per stage
constexpr uint32_t SPIN_LIMIT = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > SPIN_LIMIT) __trap();   // a broken pipeline must fault, not hang the GPU
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int32_t x, int32_t y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);   // start address       [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset [16,30) (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version  [46,48) = 1 on sm_100
  d |= (uint64_t)2 << 61;                        // layout type         [61,64) = SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_load_32cols(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, M = 128, N = BN
// Grouped tile order: consecutive tile indices sweep GROUP_M row-tiles before moving to the next
// column-tile, so CTAs that run at the same time share both their A and their B tiles in L2
// (with a plain row-major order every row-tile re-streams all of B from HBM).
constexpr int GROUP_M = 16;
__device__ __forceinline__ void tile_coords(int64_t t, int32_t m_tiles, int32_t n_tiles, int32_t& mt, int32_t& nt) {
  const int64_t per_group = (int64_t)GROUP_M * n_tiles;
  const int32_t first_m = (int32_t)(t / per_group) * GROUP_M;
  const int32_t gsize = (m_tiles - first_m) < GROUP_M ? (m_tiles - first_m) : GROUP_M;
  const int64_t in_group = t % per_group;
  mt = first_m + (int32_t)(in_group % gsize);
  nt = (int32_t)(in_group / gsize);
}

template <int BN>
struct IdescI8 {
  static constexpr uint32_t value =
      (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// ------------------------------------------------------------------ persistent variant
// One CTA per SM walks the output tiles (128 x BN) in row-major tile order.  Six warps:
//   warp 0      : TMA producer over (tile, k-block)
//   warp 1      : MMA issuer; owns the TMEM allocation: two accumulator buffers of BN columns
//   warps 2..5  : epilogue of tile i (tcgen05.ld -> HBM) while warp 1 already accumulates tile i+1
// so neither the epilogue nor the TMEM/barrier set-up sits on the tensor pipe's critical path.
struct __align__(8) PersistBarriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};
template <int BN, int STAGES>
constexpr size_t gram_persist_smem_bytes() {
  return 1024 + (size_t)STAGES * (A_TILE_BYTES + BN * BK) + sizeof(PersistBarriers);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gram_i8_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          int32_t* __restrict__ C, int64_t ldc, int64_t m, int64_t n, int32_t kblk0, int32_t k_blocks,
                          int32_t m_tiles, int32_t n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_TILE_BYTES = BN * BK;
  constexpr uint32_t kIdesc = IdescI8<BN>::value;
  static_assert(2 * BN <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  uint8_t* tile_a = smem;
  uint8_t* tile_b = smem + (size_t)STAGES * A_TILE_BYTES;
  PersistBarriers* bars = reinterpret_cast<PersistBarriers*>(smem + (size_t)STAGES * (A_TILE_BYTES + B_TILE_BYTES));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total_tiles = (int64_t)m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int32_t mt, nt;
        tile_coords(t, m_tiles, n_tiles, mt, nt);
        const int32_t row0 = mt * BM, col0 = nt * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          mbar_expect_tx(&bars->full[stage], (uint32_t)(A_TILE_BYTES + B_TILE_BYTES));
          const int32_t kx = (kblk0 + kb) * BK;
          tma_load_2d(&map_a, &bars->full[stage], tile_a + (size_t)stage * A_TILE_BYTES, kx, row0);
          tma_load_2d(&map_b, &bars->full[stage], tile_b + (size_t)stage * B_TILE_BYTES, kx, col0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);        // epilogue has drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(tile_a + (size_t)stage * A_TILE_BYTES);
          const uint32_t b_addr = smem_u32(tile_b + (size_t)stage * B_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8(tmem_acc, umma_smem_desc(a_addr + k * UMMA_K), umma_smem_desc(b_addr + k * UMMA_K), kIdesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&bars->tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // epilogue warps 2..5: warp w may touch TMEM lanes [32 (w % 4), 32 (w % 4) + 32)
    const int quarter = warp & 3;
    const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int32_t mt, nt;
      tile_coords(t, m_tiles, n_tiles, mt, nt);
      const int64_t row0 = (int64_t)mt * BM, col0 = (int64_t)nt * BN;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t r = row0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_load_32cols(tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
        if (r < m) {
          const int64_t cbase = col0 + c * 32;
          int32_t* dst = C + r * ldc + cbase;
          if (vec_ok && cbase + 32 <= n) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<int4*>(dst + 4 * q) =
                  make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (cbase + q < n) dst[q] = (int32_t)v[q];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                 : "memory");
  }
}

// ------------------------------------------------------------------ epilogue forms
// One thread owns 32 consecutive columns of one output row (tcgen05.ld 32x32b.x32).  The accumulator leaves
// TMEM as s32; what reaches HBM is chosen per launch (nbw_gram_out): s32 (128 B per segment), u8 / s16
// with saturation (32 / 64 B: Gram entries of NAS cells are <= (h+1) n_max^2, 196 for NB101 at h = 3) or the
// cosine-normalised fp32 kernel value K_ij rsqrt(K_ii) rsqrt(K_jj) of kernels/combine_kernels.py:54-56.
struct GramOut {
  void* C;
  int64_t ldc;              // in elements of the output type
  const float* scale_a;     // F32_NORM: 1/sqrt(K_ii) of the rows of A / of B
  const float* scale_b;
  uint32_t* overflow;       // set to 1 when a narrow store had to saturate
  int64_t row_offset;       // global row index of A's first row (symmetric schedule)
  int upper_only;           // skip tiles strictly below the diagonal of the global matrix
};

template <int MODE>
__device__ __forceinline__ void store_segment(const GramOut& O, int64_t r, int64_t cbase, const uint32_t (&v)[32], int64_t n,
                                              bool vec_ok) {
  if (MODE == NBW_GRAM_S32) {
    int32_t* dst = static_cast<int32_t*>(O.C) + r * O.ldc + cbase;
    if (vec_ok && cbase + 32 <= n) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<int4*>(dst + 4 * q) = make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (cbase + q < n) dst[q] = (int32_t)v[q];
    }
  } else if (MODE == NBW_GRAM_U8) {
    uint8_t* dst = static_cast<uint8_t*>(O.C) + r * O.ldc + cbase;
    uint32_t w[8];
    bool sat = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int x = (int)v[4 * q + t];
        sat = sat || x > 255 || x < 0;
        b[t] = (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
      }
      w[q] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    }
    if (sat && O.overflow) atomicOr(O.overflow, 1u);
    if (vec_ok && cbase + 32 <= n) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (cbase + q < n) dst[q] = (uint8_t)(w[q >> 2] >> (8 * (q & 3)));
    }
  } else if (MODE == NBW_GRAM_S16) {
    int16_t* dst = static_cast<int16_t*>(O.C) + r * O.ldc + cbase;
    uint32_t w[16];
    bool sat = false;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      int a = (int)v[2 * q], b = (int)v[2 * q + 1];
      sat = sat || a > 32767 || a < -32768 || b > 32767 || b < -32768;
      a = a > 32767 ? 32767 : (a < -32768 ? -32768 : a);
      b = b > 32767 ? 32767 : (b < -32768 ? -32768 : b);
      w[q] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
    }
    if (sat && O.overflow) atomicOr(O.overflow, 1u);
    if (vec_ok && cbase + 32 <= n) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (cbase + q < n) dst[q] = (int16_t)(w[q >> 1] >> (16 * (q & 1)));
    }
  } else {
    float* dst = static_cast<float*>(O.C) + r * O.ldc + cbase;
    const float sa = __ldg(O.scale_a + r);
    if (vec_ok && cbase + 32 <= n) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 sb = __ldg(reinterpret_cast<const float4*>(O.scale_b + cbase) + q);
        *reinterpret_cast<float4*>(dst + 4 * q) =
            make_float4((float)(int)v[4 * q] * sa * sb.x, (float)(int)v[4 * q + 1] * sa * sb.y,
                        (float)(int)v[4 * q + 2] * sa * sb.z, (float)(int)v[4 * q + 3] * sa * sb.w);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (cbase + q < n) dst[q] = (float)(int)v[q] * sa * __ldg(O.scale_b + cbase + q);
    }
  }
}

// Symmetric schedule (A and B are row ranges of the same matrix): a 256 x 256 tile whose last column lies
// left of its first global row holds only entries K_ij with i > j — the mirror image of a tile that is
// computed — and is skipped by producer, MMA issuer and epilogue alike.
__device__ __forceinline__ bool tile_skipped(const GramOut& O, int32_t mt, int32_t nt, int tile_rows, int tile_cols) {
  return O.upper_only && ((int64_t)nt * tile_cols + (tile_cols - 1) < O.row_offset + (int64_t)mt * tile_rows);
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  Each CTA stages its own 128 rows of A
// and its own 128 rows (= half of N) of B per K-block: 32 KB per stage instead of 48 KB, and every
// operand byte is read from shared memory by one SM only — the single-CTA kernel is bound by exactly
// that bandwidth.  The leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256; the
// accumulator rows 128 r .. 128 r + 127 land in the TMEM of CTA r.  Barrier protocol:
//   full[s]      lives in the leader: one arrive.expect_tx (64 KB) by the leader's producer, TMA of BOTH
//                CTAs completes its bytes there (mbarrier address mapped to rank 0, .cta_group::2)
//   empty[s]     one per CTA: tcgen05.commit ... multicast::cluster (mask 0b11) releases both producers
//   tmem_full[b] one per CTA, same multicast commit
//   tmem_empty[b] in the leader: 4 epilogue warps x 2 CTAs arrive (the peer's remotely)
struct __align__(8) PairBarriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int PAIR_BN = 256;                       // N of the pair tile
constexpr int PAIR_B_ROWS = PAIR_BN / 2;           // rows of B each CTA stages
constexpr int PAIR_STAGE_BYTES = A_TILE_BYTES + PAIR_B_ROWS * BK;
constexpr int PAIR_STAGES = 6;                     // 4..7 measured within 1 % of each other (profiles/r01_gram_variants.txt)
template <int STAGES>
constexpr size_t gram_pair_smem_bytes() {
  return 1024 + (size_t)STAGES * PAIR_STAGE_BYTES + sizeof(PairBarriers);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, void* dst, int32_t x, int32_t y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int STAGES, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gram_i8_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GramOut O, int64_t m, int64_t n, int32_t kblk0, int32_t k_blocks,
                    int32_t m_tiles, int32_t n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_HALF_BYTES = PAIR_B_ROWS * BK;
  constexpr int PM = 2 * BM;                                                 // rows of the pair tile
  constexpr uint32_t kIdesc =
      (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PAIR_BN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
  uint8_t* tile_a = smem;
  uint8_t* tile_b = smem + (size_t)STAGES * A_TILE_BYTES;
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem + (size_t)STAGES * PAIR_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int64_t total_tiles = (int64_t)m_tiles * n_tiles;
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], 8);   // 4 epilogue warps of each CTA (used in the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();          // the peer's barriers exist before anything remote touches them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t t = pair_id; t < total_tiles; t += n_pairs) {
        int32_t mt, nt;
        tile_coords(t, m_tiles, n_tiles, mt, nt);
        if (tile_skipped(O, mt, nt, PM, PAIR_BN)) continue;
        const int32_t row0 = mt * PM + (int32_t)rank * BM, col0 = nt * PAIR_BN + (int32_t)rank * PAIR_B_ROWS;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          if (leader) mbar_expect_tx(&bars->full[stage], (uint32_t)(2 * PAIR_STAGE_BYTES));
          const uint32_t leader_full = map_to_rank(smem_u32(&bars->full[stage]), 0);
          const int32_t kx = (kblk0 + kb) * BK;
          tma_load_2d_pair(&map_a, leader_full, tile_a + (size_t)stage * A_TILE_BYTES, kx, row0);
          tma_load_2d_pair(&map_b, leader_full, tile_b + (size_t)stage * B_HALF_BYTES, kx, col0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t t = pair_id; t < total_tiles; t += n_pairs) {
        {
          int32_t mt, nt;
          tile_coords(t, m_tiles, n_tiles, mt, nt);
          if (tile_skipped(O, mt, nt, PM, PAIR_BN)) continue;
        }
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);        // both CTAs' epilogues have drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + acc * PAIR_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(tile_a + (size_t)stage * A_TILE_BYTES);
          const uint32_t b_addr = smem_u32(tile_b + (size_t)stage * B_HALF_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_i8_pair(tmem_acc, umma_smem_desc(a_addr + k * UMMA_K), umma_smem_desc(b_addr + k * UMMA_K), kIdesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&bars->empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(&bars->tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    constexpr int kElem = MODE == NBW_GRAM_U8 ? 1 : (MODE == NBW_GRAM_S16 ? 2 : 4);
    const bool vec_ok = ((O.ldc * kElem) % 16 == 0) && ((reinterpret_cast<uintptr_t>(O.C) & 15) == 0) &&
                        (MODE != NBW_GRAM_F32_NORM || (reinterpret_cast<uintptr_t>(O.scale_b) & 15) == 0);
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t t = pair_id; t < total_tiles; t += n_pairs) {
      int32_t mt, nt;
      tile_coords(t, m_tiles, n_tiles, mt, nt);
      if (tile_skipped(O, mt, nt, PM, PAIR_BN)) continue;
      const int64_t row0 = (int64_t)mt * PM + (int64_t)rank * BM, col0 = (int64_t)nt * PAIR_BN;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t r = row0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < PAIR_BN / 32; ++c) {
        uint32_t v[32];
        tmem_load_32cols(tmem_base + acc * PAIR_BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
        if (r < m) store_segment<MODE>(O, r, col0 + c * 32, v, n, vec_ok);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_rank(smem_u32(&bars->tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();          // nobody frees TMEM or leaves while the pair still computes / signals
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ dp4a cross-check kernel
__global__ void __launch_bounds__(256)
gram_i8_simt_kernel(const int8_t* __restrict__ A, int64_t m, int64_t lda, const int8_t* __restrict__ B, int64_t n,
                    int64_t ldb, int64_t k0, int64_t k1, int32_t* __restrict__ C, int64_t ldc) {
  __shared__ uint32_t As[64][17];
  __shared__ uint32_t Bs[64][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * 64, col0 = (int64_t)blockIdx.x * 64;
  int acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0;
  for (int64_t kk = k0; kk < k1; kk += 64) {
    {
      const int r = threadIdx.x >> 2, q = threadIdx.x & 3;   // 64 rows x 4 uint4
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (row0 + r < m) va = *reinterpret_cast<const uint4*>(A + (row0 + r) * lda + kk + q * 16);
      if (col0 + r < n) vb = *reinterpret_cast<const uint4*>(B + (col0 + r) * ldb + kk + q * 16);
      As[r][q * 4 + 0] = va.x; As[r][q * 4 + 1] = va.y; As[r][q * 4 + 2] = va.z; As[r][q * 4 + 3] = va.w;
      Bs[r][q * 4 + 0] = vb.x; Bs[r][q * 4 + 1] = vb.y; Bs[r][q * 4 + 2] = vb.z; Bs[r][q * 4 + 3] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      int a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = (int)As[ty * 4 + r][w];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = (int)Bs[tx * 4 + c][w];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = __dp4a(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t gi = row0 + ty * 4 + r, gj = col0 + tx * 4 + c;
      if (gi < m && gj < n) C[gi * ldc + gj] = acc[r][c];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(nbw_ctx* ctx, CUtensorMap* map, const int8_t* base, int64_t rows, int64_t ld, int box_rows) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    NBW_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) NBW_FAIL(ctx, NBW_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    ctx->encode_tiled = fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld};
  const cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) NBW_FAIL(ctx, NBW_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NBW_OK;
}

}  // namespace

namespace {

template <int MODE>
int launch_pair(nbw_ctx* ctx, const CUtensorMap& map_a, const CUtensorMap& map_b, const GramOut& O, int64_t m, int64_t n,
                int32_t kb0, int32_t nkb, cudaStream_t st) {
  constexpr size_t kBytes = gram_pair_smem_bytes<PAIR_STAGES>();
  auto kern = gram_i8_pair_kernel<PAIR_STAGES, MODE>;
  if (!(ctx->attr_set & (NBW_ATTR_GRAM_PAIR << MODE))) {
    NBW_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBytes));
    ctx->attr_set |= (NBW_ATTR_GRAM_PAIR << MODE);
  }
  int& max_pairs = ctx->gram_max_pairs;
  if (!max_pairs) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->sm_count & ~1));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = kBytes;
    int clusters = 0;
    NBW_CUDA(ctx, cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg));
    NBW_REQUIRE(ctx, clusters >= 1, "no CTA pair of the Gram kernel fits this device");
    max_pairs = clusters;
  }
  const int32_t mt2 = (int32_t)((m + 2 * BM - 1) / (2 * BM)), nt2 = (int32_t)((n + PAIR_BN - 1) / PAIR_BN);
  const int64_t tiles = (int64_t)mt2 * nt2;
  NBW_REQUIRE(ctx, tiles < ((int64_t)1 << 31), "gram: too many tiles for one launch");
  const unsigned pairs = (unsigned)(tiles < max_pairs ? tiles : max_pairs);
  kern<<<2 * pairs, 192, kBytes, st>>>(map_a, map_b, O, m, n, kb0, nkb, mt2, nt2);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

int check_operands(nbw_ctx* ctx, const int8_t* A, int64_t m, int64_t lda, const int8_t* B, int64_t n, int64_t ldb, int64_t k0,
                   int64_t k1) {
  NBW_REQUIRE(ctx, k0 >= 0 && k1 > k0 && k0 % 128 == 0 && k1 % 128 == 0, "k0, k1 must be multiples of 128 with k1 > k0");
  NBW_REQUIRE(ctx, lda % 128 == 0 && ldb % 128 == 0 && lda >= k1 && ldb >= k1, "lda/ldb must be multiples of 128 and >= k1");
  NBW_REQUIRE(ctx, A && B, "null pointer");
  NBW_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 127) == 0,
              "A and B must be 128-byte aligned");
  NBW_REQUIRE(ctx, m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31) && lda < ((int64_t)1 << 31) && ldb < ((int64_t)1 << 31),
              "dimension too large");
  return NBW_OK;
}

// C[i][j] = C[j][i] for i > j: completes a matrix of which only the upper 256 x 256 tiles were computed.
// 32 x 32 tiles through shared memory: coalesced reads of the upper tile, coalesced writes of its mirror.
template <typename T>
__global__ void __launch_bounds__(256)
sym_mirror_kernel(T* __restrict__ C, int64_t n, int64_t ldc) {
  __shared__ T tile[32][33];
  // triangular enumeration of tile pairs (bi < bj) plus the diagonal tiles (bi == bj)
  const int64_t nb = (n + 31) / 32;
  int64_t t = blockIdx.x;
  // row bi has nb - bi tiles; invert the triangular index with a float estimate and a fix-up
  int64_t bi = (int64_t)((2.0 * nb + 1.0 - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t)) * 0.5);
  while (bi > 0 && bi * nb - bi * (bi - 1) / 2 > t) --bi;
  while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
  const int64_t bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? C[i * ldc + j] : T(0);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bj * 32 + r, j = bi * 32 + tx;    // the mirrored tile: rows of bj, columns of bi
    if (i < n && j < n && i > j) C[i * ldc + j] = tile[tx][r];
  }
}

}  // namespace

extern "C" int nbw_gram_i8_ex(nbw_ctx* ctx, const int8_t* A, int64_t m, int64_t lda, const int8_t* B, int64_t n,
                              int64_t ldb, int64_t k0, int64_t k1, void* C, int64_t ldc, int out_mode, int upper_only,
                              int64_t row_offset, const float* scale_a, const float* scale_b, uint32_t* overflow,
                              void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_gram_i8_ex");
  NBW_REQUIRE(ctx, m >= 0 && n >= 0 && ldc >= n, "bad shape");
  NBW_REQUIRE(ctx, out_mode >= NBW_GRAM_S32 && out_mode <= NBW_GRAM_F32_NORM, "unknown output mode %d", out_mode);
  NBW_REQUIRE(ctx, out_mode != NBW_GRAM_F32_NORM || (scale_a && scale_b), "the normalised output needs both scale vectors");
  NBW_REQUIRE(ctx, row_offset >= 0, "row_offset negative");
  if (m == 0 || n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, C != nullptr, "C is NULL");
  int rc = check_operands(ctx, A, m, lda, B, n, ldb, k0, k1);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap map_a, map_b;
  rc = make_operand_map(ctx, &map_a, A, m, lda, BM);
  if (rc) return rc;
  rc = make_operand_map(ctx, &map_b, B, n, ldb, PAIR_B_ROWS);
  if (rc) return rc;
  GramOut O;
  O.C = C;
  O.ldc = ldc;
  O.scale_a = scale_a;
  O.scale_b = scale_b;
  O.overflow = overflow;
  O.row_offset = row_offset;
  O.upper_only = upper_only;
  const int32_t kb0 = (int32_t)(k0 / BK), nkb = (int32_t)((k1 - k0) / BK);
  switch (out_mode) {
    case NBW_GRAM_S32: return launch_pair<NBW_GRAM_S32>(ctx, map_a, map_b, O, m, n, kb0, nkb, st);
    case NBW_GRAM_U8: return launch_pair<NBW_GRAM_U8>(ctx, map_a, map_b, O, m, n, kb0, nkb, st);
    case NBW_GRAM_S16: return launch_pair<NBW_GRAM_S16>(ctx, map_a, map_b, O, m, n, kb0, nkb, st);
    default: return launch_pair<NBW_GRAM_F32_NORM>(ctx, map_a, map_b, O, m, n, kb0, nkb, st);
  }
}

extern "C" int nbw_sym_mirror(nbw_ctx* ctx, void* C, int64_t n, int64_t ldc, int elem_bytes, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_sym_mirror");
  NBW_REQUIRE(ctx, n >= 0 && ldc >= n && (elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4), "bad argument");
  if (n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, C != nullptr, "C is NULL");
  const int64_t nb = (n + 31) / 32;
  const int64_t tiles = nb * (nb + 1) / 2;
  NBW_REQUIRE(ctx, tiles < ((int64_t)1 << 31), "matrix too large for one mirror launch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 1) sym_mirror_kernel<uint8_t><<<(unsigned)tiles, 256, 0, st>>>(static_cast<uint8_t*>(C), n, ldc);
  else if (elem_bytes == 2) sym_mirror_kernel<uint16_t><<<(unsigned)tiles, 256, 0, st>>>(static_cast<uint16_t*>(C), n, ldc);
  else sym_mirror_kernel<uint32_t><<<(unsigned)tiles, 256, 0, st>>>(static_cast<uint32_t*>(C), n, ldc);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

extern "C" int nbw_gram_i8(nbw_ctx* ctx, const int8_t* A, int64_t m, int64_t lda, const int8_t* B, int64_t n,
                           int64_t ldb, int64_t k0, int64_t k1, int32_t* C, int64_t ldc, int impl, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_gram_i8");
  NBW_REQUIRE(ctx, m >= 0 && n >= 0 && ldc >= n, "bad shape");
  if (m == 0 || n == 0) return NBW_OK;
  NBW_REQUIRE(ctx, C != nullptr, "null pointer");
  int rc = check_operands(ctx, A, m, lda, B, n, ldb, k0, k1);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl == 1) {
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((m + 63) / 64));
    NBW_REQUIRE(ctx, grid.y <= 65535, "simt gram: m too large");
    gram_i8_simt_kernel<<<grid, 256, 0, st>>>(A, m, lda, B, n, ldb, k0, k1, C, ldc);
    NBW_CHECK_LAUNCH(ctx);
    return NBW_OK;
  }
  NBW_REQUIRE(ctx, impl == 0, "unknown impl %d", impl);
  // Default: the CTA-pair kernel: cta_group::2, 256 x 256 tiles, epilogue of tile i overlapped with the
  // main loop of tile i+1 through two TMEM accumulator buffers, grouped tile order for L2 reuse.
  // Measured int8 throughput: 3.20 Pop/s at 16384^2 x 8192, 3.9-4.0 Pop/s on the 200k x 200k x 33792
  // stress (profiles/).  NBW_GRAM_VARIANT=4 selects the single-CTA persistent kernel it replaced
  // (128 x 256 tiles, cta_group::1: 2.89 / 3.44 Pop/s), kept for comparison.
  int variant = 5;
  if (const char* v = getenv("NBW_GRAM_VARIANT")) variant = atoi(v);
  NBW_REQUIRE(ctx, variant == 4 || variant == 5, "NBW_GRAM_VARIANT must be 4 (single CTA) or 5 (CTA pair)");
  if (variant == 5)
    return nbw_gram_i8_ex(ctx, A, m, lda, B, n, ldb, k0, k1, C, ldc, NBW_GRAM_S32, 0, 0, nullptr, nullptr, nullptr, stream);
  CUtensorMap map_a, map_b;
  rc = make_operand_map(ctx, &map_a, A, m, lda, BM);
  if (rc) return rc;
  rc = make_operand_map(ctx, &map_b, B, n, ldb, 256);
  if (rc) return rc;
  const int32_t kb0 = (int32_t)(k0 / BK), nkb = (int32_t)((k1 - k0) / BK);
  constexpr size_t kBytes = gram_persist_smem_bytes<256, 4>();
  if (!(ctx->attr_set & NBW_ATTR_GRAM_SINGLE)) {
    NBW_CUDA(ctx, cudaFuncSetAttribute(gram_i8_persistent_kernel<256, 4>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBytes));
    ctx->attr_set |= NBW_ATTR_GRAM_SINGLE;
  }
  const int32_t mt = (int32_t)((m + BM - 1) / BM), nt = (int32_t)((n + 255) / 256);
  const int64_t tiles = (int64_t)mt * nt;
  NBW_REQUIRE(ctx, tiles < ((int64_t)1 << 31), "gram: too many tiles for one launch");
  const unsigned ctas = (unsigned)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  gram_i8_persistent_kernel<256, 4><<<ctas, 192, kBytes, st>>>(map_a, map_b, C, ldc, m, n, kb0, nkb, mt, nt);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}

// ====================================================================== per-level Grams of a small training set
// The grid search over the WL depth (bayesopt/gp.py:493-538) needs the Gram of every single level of the training
// set (vertex_histogram.py:243, one level at a time).  For a few hundred cells that is microseconds of arithmetic:
// one SIMT launch (dp4a, exact like the tensor-core path) computes all levels — the upper triangle of 32 x 32 tiles,
// mirrored on store — instead of one TMA descriptor + persistent tensor-core launch per level.
namespace {
constexpr int kLgTile = 32;
__global__ void __launch_bounds__(256)
level_grams_small_kernel(const int8_t* __restrict__ phi, int64_t ld_phi, int n, const int32_t* __restrict__ col_off,
                         int32_t* __restrict__ out, int64_t ld_out, int64_t level_stride, int tiles) {
  __shared__ uint32_t As[kLgTile][36], Bs[kLgTile][36];      // 128 bytes of K per row (+ padding against bank conflicts)
  const int level = blockIdx.y;
  // upper-triangle tile index -> (ti, tj), ti <= tj
  int t = blockIdx.x, ti = 0;
  while (t >= tiles - ti) { t -= tiles - ti; ++ti; }
  const int tj = ti + t;
  const int tid = threadIdx.x;
  const int row = tid >> 3, cg = tid & 7;                    // this thread: output row `row`, columns 4 cg .. 4 cg + 3
  const int k0 = col_off[level], k1 = col_off[level + 1];    // multiples of 128
  int acc[4] = {0, 0, 0, 0};
  const int lr = tid >> 3, lw = (tid & 7) * 4;               // loader: row lr, words lw .. lw + 3 (16 bytes)
  for (int k = k0; k < k1; k += 128) {
    const int ra = ti * kLgTile + lr, rb = tj * kLgTile + lr;
    uint4 va = make_uint4(0u, 0u, 0u, 0u), vb = va;
    if (ra < n) va = *reinterpret_cast<const uint4*>(phi + (int64_t)ra * ld_phi + k + 4 * lw);
    if (rb < n) vb = *reinterpret_cast<const uint4*>(phi + (int64_t)rb * ld_phi + k + 4 * lw);
    As[lr][lw] = va.x; As[lr][lw + 1] = va.y; As[lr][lw + 2] = va.z; As[lr][lw + 3] = va.w;
    Bs[lr][lw] = vb.x; Bs[lr][lw + 1] = vb.y; Bs[lr][lw + 2] = vb.z; Bs[lr][lw + 3] = vb.w;
    __syncthreads();
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
      const int a = (int)As[row][w];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = __dp4a(a, (int)Bs[4 * cg + c][w], acc[c]);
    }
    __syncthreads();
  }
  int32_t* o = out + (int64_t)level * level_stride;
  const int gi = ti * kLgTile + row;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int gj = tj * kLgTile + 4 * cg + c;
    if (gi < n && gj < n) {
      o[(int64_t)gi * ld_out + gj] = acc[c];
      o[(int64_t)gj * ld_out + gi] = acc[c];
    }
  }
}
}  // namespace

extern "C" int nbw_level_grams_i32(nbw_ctx* ctx, const int8_t* phi, int64_t n, int64_t ld_phi, const int32_t* col_off,
                                   int n_levels, int32_t* out, int64_t ld_out, int64_t level_stride, void* stream) {
  if (!ctx) return NBW_ERR_INVALID;
  NbwRange range("nbw_level_grams_i32");
  NBW_REQUIRE(ctx, n >= 1 && n <= 4096 && n_levels >= 1 && n_levels <= NBW_MAX_LEVELS, "bad n / n_levels");
  NBW_REQUIRE(ctx, phi && col_off && out && ld_out >= n && level_stride >= n * ld_out && ld_phi % 16 == 0, "bad argument");
  const int tiles = (int)((n + kLgTile - 1) / kLgTile);
  dim3 grid((unsigned)(tiles * (tiles + 1) / 2), (unsigned)n_levels);
  level_grams_small_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(phi, ld_phi, (int)n, col_off, out, ld_out,
                                                                              level_stride, tiles);
  NBW_CHECK_LAUNCH(ctx);
  return NBW_OK;
}
