// Second-generation integer-slice contraction (see gemm_i8.cuh for the scheme): the same 21 exact
// int8 digit products, re-tiled so that every tcgen05.mma is M128 N128 K32 -- the N = 64 shape of the
// first kernel issues at 55 clk/MMA (61 % of the int8 peak, tools/ubench_i8.cu), N = 128 at the full
// rate (66.7 clk for twice the work).
//
// Tensor memory holds 512 columns = four 128-column int32 accumulators, but the scheme needs six digit
// groups (s + t = 2..7).  So a CTA makes TWO passes over its k-range for one 128 x 128 output tile:
//   pass A  groups 4..7  (18 products, all six digits of both operands, 48 KB per K = 32 block)
//           -> accumulators 0..3 (all 512 columns); the epilogue warps drain them into registers as the
//           fp64 partial v1 = sum_{d=4..7} 2^(-7d) G_d (64 values per thread), columns 0..255 first so
//           that
//   pass B  groups 2, 3  (3 products of digits 1, 2 only: one digit-pair plane per operand, so a stage
//           carries three K = 32 blocks in the same 48 KB) -> accumulators 0..1
//           can start while columns 256..511 are still being read; the final epilogue adds
//           2^-14 G_2 + 2^-21 G_3, applies the row / column scales, squares and reduces per candidate.
// Putting the byte-hungry products (18 per 48 KB) in one pass and the three leading products in a
// pass that moves a third of the bytes keeps both passes near the ~45 B/clk/SM that L2 delivers.
// Operands stream through one 4-stage TMA ring that runs uninterrupted across the two passes.  Digit
// planes are stored pair-interleaved at 32-k granularity: each 64-byte row segment holds 32 k-values of
// digit 2p followed by the same 32 k-values of digit 2p+1 (TMA box inner = 64 B, SWIZZLE_64B, UMMA
// descriptors at byte offsets 0 / 32).  k-blocks past the row block's triangular range are harmless:
// W is zero there (and beyond the matrix TMA zero-fills).
// Warp roles as in gemm_i8.cuh: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-9
// epilogue (two per TMEM lane quarter, 64 candidate columns each).
//
// The kernel is PERSISTENT: one CTA per SM walks a static list of tiles (row blocks heaviest first,
// dealt to the CTAs in serpentine order so every CTA gets the same k-depth to within ~0.5 %).  Tensor
// memory is allocated once, and the TMA ring runs on across tile boundaries, so the operands of tile
// j + 1 are already in shared memory while the epilogue warps finish tile j; a one-tile-per-CTA launch
// paid ~11 us of allocation / pipeline fill / epilogue per tile (15 % of the kernel at N = 5000), and the
// hardware's in-order block scheduler left a 7 % tail.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_i8.cuh"

namespace dfb {

constexpr int X2_BM = 128, X2_BN = 128, X2_BK = 32;
constexpr int X2_STAGES = 4;
constexpr int X2_A_PAIR = X2_BM * 2 * X2_BK;            // 8192 B: 128 rows x (32 B digit 2p | 32 B digit 2p+1)
constexpr int X2_B_PAIR = X2_BN * 2 * X2_BK;            // 8192 B
constexpr int X2_STAGE_BYTES = 3 * X2_A_PAIR + 3 * X2_B_PAIR;   // 49152: 3 planes (pass A) or 3 k-blocks (pass B) per operand
constexpr int X2_THREADS = 320;
constexpr size_t X2_SMEM_BYTES = (size_t)X2_STAGES * X2_STAGE_BYTES + 1024 + 2 * 4 * X2_BN * sizeof(double) +
                                 (2 * X2_STAGES + 4) * 8 + 64;
// M=128, N=128, A/B = signed int8 K-major, D = int32
constexpr uint32_t X2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(X2_BN >> 3) << 17) |
                              ((uint32_t)(X2_BM >> 4) << 24);

__device__ __forceinline__ void umma_i8_n128(unsigned tmem_d, uint64_t da, uint64_t db, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(X2_IDESC), "r"(accumulate)
      : "memory");
}

// Tile j of CTA p: serpentine deal of the list (rb descending, cb ascending); false when the list is exhausted.
__device__ __forceinline__ bool x2_tile(const ScoreI8Args& g, int j, int& rb, int& cb, int& nk) {
  const int P = (int)gridDim.x, p = (int)blockIdx.x;
  const int t = j * P + ((j & 1) ? P - 1 - p : p);
  if (t >= g.n_rb * g.n_cb) return false;
  rb = g.n_rb - 1 - t / g.n_cb;
  cb = t % g.n_cb;
  nk = min(g.K, (rb + 1) * TILE) / X2_BK;         // K = 32 blocks inside the triangular range of this row block
  return true;
}

__global__ void __launch_bounds__(X2_THREADS, 1)
score_i8x2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA3,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB3,
                  const ScoreI8Args g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* colsum = reinterpret_cast<double*>(tiles + (size_t)X2_STAGES * X2_STAGE_BYTES);   // [2][4][128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(colsum + 2 * 4 * X2_BN);
  uint64_t* empty_bar = full_bar + X2_STAGES;
  uint64_t* acc1_bar = empty_bar + X2_STAGES;     // pass-A accumulators complete
  uint64_t* drain_bar = acc1_bar + 1;             // columns 0..255 drained by all 8 epilogue warps
  uint64_t* acc2_bar = drain_bar + 1;             // pass-B accumulators complete
  uint64_t* epi_bar = acc2_bar + 1;               // all tensor-memory reads of the tile done (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < X2_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc1_bar, 1);
    mbar_init(drain_bar, 8);
    mbar_init(acc2_bar, 1);
    mbar_init(epi_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ---------------- TMA producer: one ring across passes and tiles ------------------------------------
    if (lane == 0) {
      unsigned git = 0;                                  // ring position, runs on across tiles
      int rb, cb, nk;
      for (int j = 0; x2_tile(g, j, rb, cb, nk); j++) {
        const int n_it = nk + (nk + 2) / 3;              // pass A: nk stages; pass B: three K = 32 blocks per stage
        for (int it = 0; it < n_it; it++, git++) {
          const unsigned s = git % X2_STAGES, n = git / X2_STAGES;
          mbar_wait(&empty_bar[s], (n & 1u) ^ 1u);
          mbar_expect_tx(&full_bar[s], (unsigned)X2_STAGE_BYTES);
          unsigned char* dst = tiles + (size_t)s * X2_STAGE_BYTES;
          // operand A sub-tiles at [0, 24 KB), operand B sub-tiles at [24 KB, 48 KB) in both passes
          if (it < nk) {
            tma_load_3d(dst, &tmA3, it * 2 * X2_BK, rb * X2_BM, 0, &full_bar[s]);
            tma_load_3d(dst + 3 * X2_A_PAIR, &tmB3, it * 2 * X2_BK, cb * X2_BN, 0, &full_bar[s]);
          } else {
            const int kb0 = (it - nk) * 3;
#pragma unroll
            for (int u = 0; u < 3; u++) {
              tma_load_3d(dst + u * X2_A_PAIR, &tmA1, (kb0 + u) * 2 * X2_BK, rb * X2_BM, 0, &full_bar[s]);
              tma_load_3d(dst + 3 * X2_A_PAIR + u * X2_B_PAIR, &tmB1, (kb0 + u) * 2 * X2_BK, cb * X2_BN, 0,
                          &full_bar[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ---------------------------------------------------------------------------
    if (lane == 0) {
      // K-major SWIZZLE_64B descriptor: SBO = 512 B (8 rows x 64 B), version 1, layout type 4
      constexpr uint64_t DESC_HI = ((uint64_t)(32u | (1u << 14) | (4u << 29))) << 32;
      unsigned git = 0;
      int rb, cb, nk;
      for (int j = 0; x2_tile(g, j, rb, cb, nk); j++) {
        const int n_it = nk + (nk + 2) / 3;
        const unsigned tpar = (unsigned)(j & 1);
        if (j > 0) {
          // the previous tile's accumulators must have been read out of tensor memory
          mbar_wait(epi_bar, tpar ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        for (int it = 0; it < n_it; it++, git++) {
          const unsigned s = git % X2_STAGES, n = git / X2_STAGES;
          const bool pb = it >= nk;
          if (it == nk) {
            // pass B reuses TMEM columns 0..255: wait until the epilogue has drained them
            mbar_wait(drain_bar, tpar);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          }
          mbar_wait(&full_bar[s], n & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const unsigned a0 = smem_u32(tiles + (size_t)s * X2_STAGE_BYTES);
          const unsigned a_lo = ((a0 & 0x3FFFFu) >> 4) | 0x10000u;
          const unsigned b_lo = (((a0 + 3 * X2_A_PAIR) & 0x3FFFFu) >> 4) | 0x10000u;
          const bool first = (it == 0) || (it == nk);
          if (!pb) {
            // groups 4..7 -> accumulators 0..3; sub-tile index = digit-pair plane
#pragma unroll
            for (int d = 4; d <= 7; d++) {
              const unsigned acc = (unsigned)((d - 4) * X2_BN);      // literal TMEM columns (base 0)
              bool lead = true;                                       // first product of this group
#pragma unroll
              for (int sa = 1; sa <= I8_S; sa++) {
                const int tb = d - sa;
                if (tb < 1 || tb > I8_S) continue;
                const unsigned aoff = ((sa - 1) >> 1) * X2_A_PAIR + ((sa - 1) & 1) * X2_BK;
                const unsigned boff = ((tb - 1) >> 1) * X2_B_PAIR + ((tb - 1) & 1) * X2_BK;
                umma_i8_n128(acc, DESC_HI | (uint64_t)(a_lo + (aoff >> 4)), DESC_HI | (uint64_t)(b_lo + (boff >> 4)),
                             (first && lead) ? 0u : 1u);
                lead = false;
              }
            }
          } else {
            // groups 2, 3 -> accumulators 0..1; sub-tile index = K = 32 block, digits 1 and 2 only
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const unsigned au = (unsigned)(u * X2_A_PAIR), bu = (unsigned)(u * X2_B_PAIR);
              const uint64_t a1 = DESC_HI | (uint64_t)(a_lo + (au >> 4));
              const uint64_t a2 = DESC_HI | (uint64_t)(a_lo + ((au + X2_BK) >> 4));
              const uint64_t b1 = DESC_HI | (uint64_t)(b_lo + (bu >> 4));
              const uint64_t b2 = DESC_HI | (uint64_t)(b_lo + ((bu + X2_BK) >> 4));
              umma_i8_n128(0u, a1, b1, (first && u == 0) ? 0u : 1u);                 // group 2: (1, 1)
              umma_i8_n128((unsigned)X2_BN, a1, b2, (first && u == 0) ? 0u : 1u);     // group 3: (1, 2)
              umma_i8_n128((unsigned)X2_BN, a2, b1, 1u);                              //          (2, 1)
            }
          }
          umma_commit(&empty_bar[s]);
          if (it == nk - 1) umma_commit(acc1_bar);
        }
        umma_commit(acc2_bar);
      }
    }
  } else {
    // ---------------- epilogue warps 2..9 ---------------------------------------------------------------------
    const int q = warp & 3;                  // TMEM lane quarter
    const int half = (warp - 2) >> 2;        // candidate columns [64 half, 64 half + 64)
    const int row = q * 32 + lane;
    const unsigned lane_addr = tmem_base + ((unsigned)(q * 32) << 16);
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    const int et = tid - 64;
    int rb, cb, nk;
    for (int j = 0; x2_tile(g, j, rb, cb, nk); j++) {
      const unsigned tpar = (unsigned)(j & 1);
      double* cs = colsum + (size_t)tpar * 4 * X2_BN;      // double-buffered by tile parity
      // tmem_base != 0 would mean the literal accumulator addresses were wrong: poison the result
      const double rs = (tmem_base == 0u) ? g.rowscale[(int64_t)rb * X2_BM + row] * g.colscale
                                          : __longlong_as_double(0x7ff8000000000000ll);
      double v1[64];
      mbar_wait(acc1_bar, tpar);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // drain groups 4, 5 (columns 0..255) first, release them to pass B, then groups 6, 7
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        int r2[8], r3[8];
        tmem_ld8(lane_addr + (unsigned)(0 * X2_BN + half * 64 + c0), r2);
        tmem_ld8(lane_addr + (unsigned)(1 * X2_BN + half * 64 + c0), r3);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++) v1[c0 + j2] = fma((double)r2[j2], 0x1p-28, (double)r3[j2] * 0x1p-35);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      if (lane == 0) mbar_arrive(drain_bar);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        int r4[8], r5[8];
        tmem_ld8(lane_addr + (unsigned)(2 * X2_BN + half * 64 + c0), r4);
        tmem_ld8(lane_addr + (unsigned)(3 * X2_BN + half * 64 + c0), r5);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++)
          v1[c0 + j2] += fma((double)r4[j2], 0x1p-42, (double)r5[j2] * 0x1p-49);
      }
      mbar_wait(acc2_bar, tpar);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        int r6[8], r7[8];
        tmem_ld8(lane_addr + (unsigned)(0 * X2_BN + half * 64 + c0), r6);
        tmem_ld8(lane_addr + (unsigned)(1 * X2_BN + half * 64 + c0), r7);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c0 == 56) {
          // last tensor-memory read of this tile: the issuer may start the next tile's pass A
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          if (lane == 0) mbar_arrive(epi_bar);
        }
        double sq[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++) {
          double v = v1[c0 + j2] + fma((double)r6[j2], 0x1p-14, (double)r7[j2] * 0x1p-21);
          v *= rs;
          sq[j2] = v * v;
        }
        // halving butterfly over the warp's 32 rows (see gemm_i8.cuh)
        double w4[4], w2[2], w1;
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) {
          const double send = h16 ? sq[j2] : sq[j2 + 4];
          const double keep = h16 ? sq[j2 + 4] : sq[j2];
          w4[j2] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int j2 = 0; j2 < 2; j2++) {
          const double send = h8 ? w4[j2] : w4[j2 + 2];
          const double keep = h8 ? w4[j2 + 2] : w4[j2];
          w2[j2] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const double send = h4 ? w2[0] : w2[1];
          const double keep = h4 ? w2[1] : w2[0];
          w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
        if ((lane & 3) == 0)
          cs[q * X2_BN + half * 64 + c0 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = w1;
      }
      asm volatile("bar.sync 1, 256;\n" ::: "memory");     // the eight epilogue warps
      if (et < X2_BN)
        g.partial[(int64_t)rb * g.ld_partial + (int64_t)cb * X2_BN + et] =
            ((cs[et] + cs[X2_BN + et]) + cs[2 * X2_BN + et]) + cs[3 * X2_BN + et];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

}  // namespace dfb
