// CTA-pair version of the two-pass integer-slice contraction (gemm_i8x2.cuh): the same passes, digit
// groups, ring and persistent schedule, but every MMA is a tcgen05.mma.cta_group::2 of shape M256 N128
// K32 issued by the leader CTA of a 2-CTA cluster (the two SMs of a TPC).
//
// Why: a single-CTA M128 N128 K32 int8 MMA reads 8 KB of operands from shared memory per 66.7 clk --
// 123 of the SM's 128 B/clk -- so the TMA ring's own writes (40 B/clk) cannot fit beside it and the
// one-CTA kernel tops out at ~75 % of the tensor issue rate.  In pair mode each CTA holds its own 128
// rows of W's digits and only HALF of the K_* tile (64 candidate rows); the hardware feeds both halves
// of the B operand to both SMs' tensor cores, so each SM reads 6 KB per MMA and fetches 36 KB instead
// of 48 KB per stage from L2.
//
// Pair geometry: cluster c works on row-block pair rp (rows [256 rp, 256 rp + 256)) x candidate tile cb;
// CTA rank r owns row block 2 rp + r (its A tile, its 128 TMEM lanes, its epilogue and its output row)
// and loads candidate rows [128 cb + 64 r, + 64) as its half of B.  Both CTAs run a TMA producer; the
// transaction bytes of both land on the LEADER's full barrier (cta_group::2 TMA form).  The leader's
// single MMA thread issues for the pair and commits to both CTAs' barriers (multicast commit); the
// epilogue warps of both CTAs report "columns drained" on the leader's barriers with remote arrives.
// Digits: with R256 the kernel uses FIVE radix-256 digits per operand (top digit 7 bits, four balanced bytes:
// digits_radix256 in kernels.cu) instead of six radix-128 digits: x ~ sum_s a_s 2^-(8s-1).  The products
// a_s b_t with s + t <= 6 are kept (15 MMAs per K = 32 block instead of 21; the dropped s + t = 7 terms
// are ~2^-40 of the operands' scale, the same order as the digit truncation), grouped by d = s + t with
// weight 2^-(8d-2).  Pass A accumulates groups 3..6 (14 products, all digits) in the four accumulators,
// pass B group 2 (the single leading product) in accumulator 0, streaming a compact copy of the leading
// digit (plain row-major int8 rows, SWIZZLE_128B, 128 k-values per stage: a sixth of pass A's bytes).  int32 headroom: 5 K 2^14 < 2^31 needs
// K <= 26214 (api.cu refuses the path beyond npad = 24576).  The tensor pipe -- and at the 1 kW power cap
// the whole step -- scales with the MMA count, so this is a straight 29 % cut of the dominant cost for
// an error that stays ~30x inside the 1e-8 contract at N = 5000.  R256 = false keeps the six radix-128 digits
// and 21 products of gemm_i8x2.cuh (12x more accurate): api.cu picks per posterior from the a-priori bound.
// The k-range is the triangular range of the lower row block of the pair (the upper one's digits are
// zero there).  A phantom row block (odd n_rb) reads zeros through TMA's out-of-bounds fill.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_i8.cuh"
#include "gemm_i8x2.cuh"

namespace dfb {

constexpr int C2_STAGES = 6;
constexpr int C2_DIGITS = 5;                           // radix-256 digits per operand
constexpr int C2_A_SUB = X2_BM * 2 * X2_BK;             // 8192 B: 128 rows x 64 B
constexpr int C2_B_SUB = (X2_BN / 2) * 2 * X2_BK;       // 4096 B:  64 rows x 64 B
constexpr int C2_STAGE_BYTES = 3 * C2_A_SUB + 3 * C2_B_SUB;          // 36864 per CTA
constexpr int C2_PB_A_BYTES = X2_BM * 128;              // pass B, radix 256: 128 rows x 128 k-values of the leading digit
constexpr int C2_PB_BYTES = C2_PB_A_BYTES + (X2_BN / 2) * 128;       // + 64 candidate rows: 24576 per CTA
// Warp roles: 0..7 epilogue (two per TMEM lane quarter), 8 TMA producer, 9 MMA issuer (+ TMEM allocation), 10..11 idle.
// Register budget: the kernel is LAUNCHED with C2_LAUNCH_REGS = 136 registers per thread (52224 per CTA, 13056 per SM
// sub-partition: the register file is partitioned 4 x 16384 and warp w lives on partition w % 4).  Right after
// set-up the light warpgroup (8..11) drops to C2_LIGHT_REGS = 56 (setmaxnreg.dec: 10240 registers into the CTA's
// pool) and the two epilogue warpgroups rise to C2_EPI_REGS = 168 (setmaxnreg.inc: 8192 out of it) -- deallocated
// registers only ever return to the CTA's own pool, so what a co-resident kernel can use is fixed by the launch
// count: 16384 - 13056 = 3328 registers per partition, room for one 88-register warp of the K_* kernel
// (kernels.cu: kstar_seg_kernel, four warps per CTA, one per partition).
constexpr int C2_THREADS = 384;
constexpr int C2_LAUNCH_REGS = 136;
constexpr int C2_LIGHT_REGS = 56;
constexpr int C2_EPI_REGS = 168;
constexpr size_t C2_SMEM_BYTES = (size_t)C2_STAGES * C2_STAGE_BYTES + 1024 + 2 * 4 * X2_BN * sizeof(double) +
                                 (2 * C2_STAGES + 4) * 8 + 64;
// M=256 (pair), N=128, A/B = signed int8 K-major, D = int32
constexpr uint32_t C2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(X2_BN >> 3) << 17) |
                              ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ unsigned c2_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Waiting without occupying issue slots: mbarrier.try_wait with a suspend-time hint parks the warp in hardware until the
// phase completes (or the hint expires) instead of returning after the short default limit.  The per-CTA trace
// (tools/trace_overlap.py) showed the co-resident K_* warps running at a sixth of their stand-alone rate while this
// kernel's five or six waiting warps per CTA polled with the default limit.
__device__ __forceinline__ void c2_wait(void* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000000u)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void c2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in the leader CTA (rank 0)
__device__ __forceinline__ void c2_arrive_leader(void* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// TMA load into this CTA's shared memory, transaction bytes credited to the leader CTA's barrier; `policy` is an L2
// eviction-priority descriptor (the fixed encodings CUTLASS uses: normal / evict-first / evict-last)
constexpr uint64_t C2_L2_NORMAL = 0x1000000000000000ull;
constexpr uint64_t C2_L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t C2_L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void c2_tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2,
                                               void* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void c2_umma(unsigned tmem_d, uint64_t da, uint64_t db, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(C2_IDESC), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void c2_commit(void* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}

// stages of pass B for nk K = 32 blocks: radix 256 carries 128 k-values of the compact leading-digit plane per
// stage, radix 128 three K = 32 blocks of the first digit-pair plane
template <bool R256>
__device__ __forceinline__ int c2_pass_b_stages(int nk) { return R256 ? (nk + 3) / 4 : (nk + 2) / 3; }

// Tile j of cluster c: serpentine deal of the tile list.  List order (g.cb_group = G candidate tiles per group):
// group of G candidate tiles ascending -> row-block pair descending (heaviest first) -> candidate tile within the
// group.  The clusters work on ~74 consecutive list entries at any time, so with G x 20 row-block pairs per group a
// group's K_* digit tiles (G x 4.6 MB at N = 5000) are consumed by ALL row-block pairs while they sit in the L2, and
// only W's digits are re-streamed, once per group: ~(n_cb / G) x |W| + |K_*| bytes from HBM per launch instead of
// ~n_rp / 2 x |K_*| for the plain row-pair-major order (G = 0 or >= n_cb).
__device__ __forceinline__ bool c2_tile(const ScoreI8Args& g, int j, int& rp, int& cb, int& nk) {
  const int P = (int)gridDim.x >> 1, c = (int)blockIdx.x >> 1;
  const int n_rp = (g.n_rb + 1) >> 1;
  const int t = j * P + ((j & 1) ? P - 1 - c : c);
  if (t >= n_rp * g.n_cb) return false;
  const int G = (g.cb_group > 0 && g.cb_group < g.n_cb) ? g.cb_group : g.n_cb;
  const int per_group = n_rp * G;
  const int grp = t / per_group;
  const int cb0 = grp * G;
  const int width = min(G, g.n_cb - cb0);                 // the last group may be narrower
  const int u = t - grp * per_group;                      // index inside the group: (rp descending, cb ascending)
  // groups before the last one have exactly per_group entries; the last one n_rp * width
  rp = n_rp - 1 - u / width;
  cb = cb0 + u % width;
  nk = min(g.K, (2 * rp + 2) * TILE) / X2_BK;
  return true;
}

template <bool R256>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(C2_LAUNCH_REGS)
score_i8c2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA3,
                  const __grid_constant__ CUtensorMap tmA1c, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmB1c,
                  const ScoreI8Args g) {
  // uniform over the grid: the counter is only written by earlier kernels of the same stream
  if (g.abort_count != nullptr && *g.abort_count > g.abort_cap) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* colsum = reinterpret_cast<double*>(tiles + (size_t)C2_STAGES * C2_STAGE_BYTES);   // [2][4][128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(colsum + 2 * 4 * X2_BN);   // used in the leader only
  uint64_t* empty_bar = full_bar + C2_STAGES;     // per CTA, multicast commit
  uint64_t* acc1_bar = empty_bar + C2_STAGES;     // per CTA, multicast commit: pass-A accumulators complete
  uint64_t* drain_bar = acc1_bar + 1;             // leader: columns 0..255 drained by the 16 epilogue warps
  uint64_t* acc2_bar = drain_bar + 1;             // per CTA, multicast commit: pass-B accumulators complete
  uint64_t* epi_bar = acc2_bar + 1;               // leader: all tensor-memory reads of the tile done (16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = c2_cta_rank();
  const unsigned long long t_begin = (tid == 0 && g_trace != nullptr) ? trace_now() : 0ull;

  if (tid == 0) {
    for (int s = 0; s < C2_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc1_bar, 1);
    mbar_init(drain_bar, 16);
    mbar_init(acc2_bar, 1);
    mbar_init(epi_bar, 16);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  c2_cluster_sync();                              // both CTAs' barriers exist before any remote arrival
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp >= 8) {
  // the light warpgroup: nothing below this point of the branch may need more than C2_LIGHT_REGS registers (the
  // epilogue sits on the other side of the branch so that ptxas does not apply the cap to it)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C2_LIGHT_REGS));
  if (warp == 8) {
    // ---------------- TMA producer (both CTAs): own A tile + own half of B --------------------------------
    if (lane == 0) {
      unsigned git = 0;
      int rp, cb, nk;
      for (int j = 0; c2_tile(g, j, rp, cb, nk); j++) {
        const int rb = 2 * rp + (int)rank;
        const int brow = cb * X2_BN + (int)rank * (X2_BN / 2);
        const int n_it = nk + c2_pass_b_stages<R256>(nk);
        for (int it = 0; it < n_it; it++, git++) {
          const unsigned s = git % C2_STAGES, n = git / C2_STAGES;
          c2_wait(&empty_bar[s], (n & 1u) ^ 1u);
          unsigned char* dst = tiles + (size_t)s * C2_STAGE_BYTES;
          if (R256 && it >= nk) {
            // pass B, radix 256: 128 k-values of the compact leading-digit planes (A 16 KB, B half 8 KB)
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * (unsigned)C2_PB_BYTES);
            c2_tma_load_3d(dst, &tmA1c, (it - nk) * 128, rb * X2_BM, 0, &full_bar[s], g.l2_a);
            c2_tma_load_3d(dst + C2_PB_A_BYTES, &tmB1c, (it - nk) * 128, brow, 0, &full_bar[s], g.l2_b);
            continue;
          }
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * (unsigned)C2_STAGE_BYTES);   // both CTAs' bytes
          if (it < nk) {
            c2_tma_load_3d(dst, &tmA3, it * 2 * X2_BK, rb * X2_BM, 0, &full_bar[s], g.l2_a);
            c2_tma_load_3d(dst + 3 * C2_A_SUB, &tmB3, it * 2 * X2_BK, brow, 0, &full_bar[s], g.l2_b);
          } else {
            const int kb0 = (it - nk) * 3;
#pragma unroll
            for (int u = 0; u < 3; u++) {
              c2_tma_load_3d(dst + u * C2_A_SUB, &tmA1, (kb0 + u) * 2 * X2_BK, rb * X2_BM, 0, &full_bar[s], g.l2_a);
              c2_tma_load_3d(dst + 3 * C2_A_SUB + u * C2_B_SUB, &tmB1, (kb0 + u) * 2 * X2_BK, brow, 0, &full_bar[s], g.l2_b);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer: leader CTA only, for the pair -----------------------------------------
    // The whole warp walks the loops in uniform control flow and ONE elected lane issues: descriptors,
    // stage indices and barrier addresses then live in uniform registers, which is what UTCIMMA / UTCBAR
    // take -- under a divergent `if (lane == 0)` every operand needs an R2UR first and the issue path
    // (~76 clk per MMA) is slower than the tensor pipe (66.7 clk).
    if (rank == 0) {
      const bool el = elect_one();
      constexpr uint64_t DESC_HI = ((uint64_t)(32u | (1u << 14) | (4u << 29))) << 32;   // SWIZZLE_64B, SBO 512
      unsigned git = 0;
      int rp, cb, nk;
      const bool timed = g.timing != nullptr;
      long long t_full = 0, t_drain = 0, t_epi = 0, t0 = 0;
      const long long t_start = clock64();
      for (int j = 0; c2_tile(g, j, rp, cb, nk); j++) {
        const int n_it = nk + c2_pass_b_stages<R256>(nk);
        const unsigned tpar = (unsigned)(j & 1);
        if (j > 0) {
          if (timed) t0 = clock64();
          c2_wait(epi_bar, tpar ^ 1u);
          if (timed) t_epi += clock64() - t0;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        for (int it = 0; it < n_it; it++, git++) {
          const unsigned s = git % C2_STAGES, n = git / C2_STAGES;
          const bool pb = it >= nk;
          if (it == nk) {
            if (timed) t0 = clock64();
            c2_wait(drain_bar, tpar);
            if (timed) t_drain += clock64() - t0;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          }
          if (timed) t0 = clock64();
          c2_wait(&full_bar[s], n & 1u);
          if (timed) t_full += clock64() - t0;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const unsigned a0 = smem_u32(tiles + (size_t)s * C2_STAGE_BYTES);
          const unsigned a_lo = ((a0 & 0x3FFFFu) >> 4) | 0x10000u;
          const unsigned b_lo = (((a0 + 3 * C2_A_SUB) & 0x3FFFFu) >> 4) | 0x10000u;
          const bool first = (it == 0) || (it == nk);
          if constexpr (R256) {
            if (!pb) {
              // groups 3..6 -> accumulators 0..3; sub-tile index = digit-pair plane
#pragma unroll
              for (int d = 3; d <= 6; d++) {
                const unsigned acc = (unsigned)((d - 3) * X2_BN);
                bool lead = true;
#pragma unroll
                for (int sa = 1; sa <= C2_DIGITS; sa++) {
                  const int tb = d - sa;
                  if (tb < 1 || tb > C2_DIGITS) continue;
                  const unsigned aoff = ((sa - 1) >> 1) * C2_A_SUB + ((sa - 1) & 1) * X2_BK;
                  const unsigned boff = ((tb - 1) >> 1) * C2_B_SUB + ((tb - 1) & 1) * X2_BK;
                  if (el)
                    c2_umma(acc, DESC_HI | (uint64_t)(a_lo + (aoff >> 4)), DESC_HI | (uint64_t)(b_lo + (boff >> 4)),
                            (first && lead) ? 0u : 1u);
                  lead = false;
                }
              }
            } else {
              // group 2 = the leading product (1, 1) -> accumulator 0: four K = 32 steps along the 128-byte
              // SWIZZLE_128B rows of the compact leading-digit tiles (SBO = 1024 B, layout type 2)
              constexpr uint64_t DESC_HI128 = ((uint64_t)(64u | (1u << 14) | (2u << 29))) << 32;
              const unsigned pb_b_lo = (((a0 + C2_PB_A_BYTES) & 0x3FFFFu) >> 4) | 0x10000u;
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const uint64_t a1 = DESC_HI128 | (uint64_t)(a_lo + ((unsigned)(u * X2_BK) >> 4));
                const uint64_t b1 = DESC_HI128 | (uint64_t)(pb_b_lo + ((unsigned)(u * X2_BK) >> 4));
                if (el) c2_umma(0u, a1, b1, (first && u == 0) ? 0u : 1u);
              }
            }
          } else {
            // six radix-128 digits: groups 4..7 -> accumulators 0..3 in pass A, groups 2, 3 -> 0..1 in pass B
            if (!pb) {
#pragma unroll
              for (int d = 4; d <= 7; d++) {
                const unsigned acc = (unsigned)((d - 4) * X2_BN);
                bool lead = true;
#pragma unroll
                for (int sa = 1; sa <= I8_S; sa++) {
                  const int tb = d - sa;
                  if (tb < 1 || tb > I8_S) continue;
                  const unsigned aoff = ((sa - 1) >> 1) * C2_A_SUB + ((sa - 1) & 1) * X2_BK;
                  const unsigned boff = ((tb - 1) >> 1) * C2_B_SUB + ((tb - 1) & 1) * X2_BK;
                  if (el) c2_umma(acc, DESC_HI | (uint64_t)(a_lo + (aoff >> 4)), DESC_HI | (uint64_t)(b_lo + (boff >> 4)),
                          (first && lead) ? 0u : 1u);
                  lead = false;
                }
              }
            } else {
#pragma unroll
              for (int u = 0; u < 3; u++) {
                const unsigned au = (unsigned)(u * C2_A_SUB), bu = (unsigned)(u * C2_B_SUB);
                const uint64_t a1 = DESC_HI | (uint64_t)(a_lo + (au >> 4));
                const uint64_t a2 = DESC_HI | (uint64_t)(a_lo + ((au + X2_BK) >> 4));
                const uint64_t b1 = DESC_HI | (uint64_t)(b_lo + (bu >> 4));
                const uint64_t b2 = DESC_HI | (uint64_t)(b_lo + ((bu + X2_BK) >> 4));
                if (el) c2_umma(0u, a1, b1, (first && u == 0) ? 0u : 1u);
                if (el) c2_umma((unsigned)X2_BN, a1, b2, (first && u == 0) ? 0u : 1u);
                if (el) c2_umma((unsigned)X2_BN, a2, b1, 1u);
              }
            }
          }
          if (el) c2_commit(&empty_bar[s]);
          if (el && it == nk - 1) c2_commit(acc1_bar);
        }
        if (el) c2_commit(acc2_bar);
      }
      if (timed && el) {
        unsigned long long* o = g.timing + 4 * (blockIdx.x >> 1);
        o[0] = (unsigned long long)t_full; o[1] = (unsigned long long)t_drain; o[2] = (unsigned long long)t_epi;
        o[3] = (unsigned long long)(clock64() - t_start);
      }
    }
  }
  } else {
    // ---------------- epilogue warps 0..7 (both CTAs, own row block) ---------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C2_EPI_REGS));
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const unsigned lane_addr = tmem_base + ((unsigned)(q * 32) << 16);
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    const int et = tid;
    int rp, cb, nk;
    for (int j = 0; c2_tile(g, j, rp, cb, nk); j++) {
      const int rb = 2 * rp + (int)rank;
      const bool real = rb < g.n_rb;                       // false for the phantom row block of an odd n_rb
      const unsigned tpar = (unsigned)(j & 1);
      double* cs = colsum + (size_t)tpar * 4 * X2_BN;
      const double rs = (tmem_base == 0u) ? (real ? g.rowscale[(int64_t)rb * X2_BM + row] * g.colscale : 0.0)
                                          : __longlong_as_double(0x7ff8000000000000ll);
      double v1[64];
      c2_wait(acc1_bar, tpar);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // tensor-memory loads run one step ahead of the conversions (tcgen05.wait::ld covers all loads
      // issued so far, so the next pair is issued right after the wait and lands during the math)
      int ra[2][8], rc[2][8];
      tmem_ld8(lane_addr + (unsigned)(0 * X2_BN + half * 64), ra[0]);
      tmem_ld8(lane_addr + (unsigned)(1 * X2_BN + half * 64), rc[0]);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        const int cur = (c0 >> 3) & 1;
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c0 + 8 < 64) {
          tmem_ld8(lane_addr + (unsigned)(0 * X2_BN + half * 64 + c0 + 8), ra[cur ^ 1]);
          tmem_ld8(lane_addr + (unsigned)(1 * X2_BN + half * 64 + c0 + 8), rc[cur ^ 1]);
        } else {
          tmem_ld8(lane_addr + (unsigned)(2 * X2_BN + half * 64), ra[cur ^ 1]);        // first pair of the second drain
          tmem_ld8(lane_addr + (unsigned)(3 * X2_BN + half * 64), rc[cur ^ 1]);
        }
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++)
          v1[c0 + j2] = R256 ? fma((double)ra[cur][j2], 0x1p-22, (double)rc[cur][j2] * 0x1p-30)
                             : fma((double)ra[cur][j2], 0x1p-28, (double)rc[cur][j2] * 0x1p-35);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      if (lane == 0) c2_arrive_leader(drain_bar);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        const int cur = (c0 >> 3) & 1;          // (64 >> 3) & 1 == 0: the prefetched pair sits in buffer 0
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c0 + 8 < 64) {
          tmem_ld8(lane_addr + (unsigned)(2 * X2_BN + half * 64 + c0 + 8), ra[cur ^ 1]);
          tmem_ld8(lane_addr + (unsigned)(3 * X2_BN + half * 64 + c0 + 8), rc[cur ^ 1]);
        }
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++)
          v1[c0 + j2] += R256 ? fma((double)ra[cur][j2], 0x1p-38, (double)rc[cur][j2] * 0x1p-46)
                              : fma((double)ra[cur][j2], 0x1p-42, (double)rc[cur][j2] * 0x1p-49);
      }
      c2_wait(acc2_bar, tpar);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // radix 256: group 2 sits in accumulator 0 alone; radix 128: groups 2, 3 in accumulators 0, 1
      tmem_ld8(lane_addr + (unsigned)(half * 64), ra[0]);
      if (!R256) tmem_ld8(lane_addr + (unsigned)(X2_BN + half * 64), rc[0]);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        const int cur = (c0 >> 3) & 1;
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c0 + 8 < 64) {
          tmem_ld8(lane_addr + (unsigned)(half * 64 + c0 + 8), ra[cur ^ 1]);
          if (!R256) tmem_ld8(lane_addr + (unsigned)(X2_BN + half * 64 + c0 + 8), rc[cur ^ 1]);
        } else {
          // last tensor-memory read of this tile has landed: the issuer may start the next tile's pass A
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          if (lane == 0) c2_arrive_leader(epi_bar);
        }
        double sq[8];
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++) {
          double v = R256 ? fma((double)ra[cur][j2], 0x1p-14, v1[c0 + j2])
                          : v1[c0 + j2] + fma((double)ra[cur][j2], 0x1p-14, (double)rc[cur][j2] * 0x1p-21);
          v *= rs;
          sq[j2] = v * v;
        }
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
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      if (et < X2_BN && real)
        g.partial[(int64_t)rb * g.ld_partial + (int64_t)cb * X2_BN + et] =
            ((cs[et] + cs[X2_BN + et]) + cs[2 * X2_BN + et]) + cs[3 * X2_BN + et];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) trace_emit(2u, t_begin);
  c2_cluster_sync();                              // the peer may still be reading / being written to
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

}  // namespace dfb
