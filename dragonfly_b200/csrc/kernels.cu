// CUDA kernels of libdfb200.so other than the DMMA GEMM core (gemm.cuh): kernel-matrix builds,
// the diagonal-block Cholesky/inverse, small vector kernels and the acquisition + arg-max.
// sm_100a only.  Reference paths are relative to the reference tree (dragonfly-opt 0.1.7).
#include "kernels.cuh"
#include "exp_nonpos.h"

// ---- diagnostics: per-CTA (kind, SM id, start, end) records of the two kernels of the overlapped scoring pipeline ------
// dfb_debug_trace(buffer, capacity) arms it (buffer[0] = record counter, 4 words per record); NULL disarms.  Used by
// tools/trace_overlap.py to see whether the K_* CTAs really run beside the persistent contraction CTAs.
namespace dfb {
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned long long g_trace_cap = 0;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_emit(unsigned kind, unsigned long long t0) {
  unsigned long long* buf = g_trace;
  if (buf == nullptr) return;
  const unsigned long long idx = atomicAdd(buf, 1ull);
  if (idx >= g_trace_cap) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  buf[1 + 4 * idx + 0] = ((unsigned long long)kind << 32) | smid;
  buf[1 + 4 * idx + 1] = t0;
  buf[1 + 4 * idx + 2] = trace_now();
  buf[1 + 4 * idx + 3] = blockIdx.x;
}
int debug_set_trace(void* buf, long long cap_records) {
  unsigned long long* p = static_cast<unsigned long long*>(buf);
  unsigned long long c = (unsigned long long)(cap_records < 0 ? 0 : cap_records);
  DFB_CUDA_OK(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  DFB_CUDA_OK(cudaMemcpyToSymbol(g_trace_cap, &c, sizeof(c)));
  return 0;
}
}  // namespace dfb
#include "gemm_tma.cuh"
#include "gemm_i8.cuh"
#include "gemm_i8x2.cuh"
#include "gemm_i8c2.cuh"

namespace dfb {

#define DFB_TRY_RET(expr) do { int _r = (expr); if (_r != 0) return _r; } while (0)

// ================================================================================================
// Kernel evaluation in the reference's operation order (include/dfb200.h, "kernel descriptor").
// ================================================================================================
__device__ __forceinline__ double base_kernel_value(const dfb_factor_desc& f, double d2) {
  if (f.kind == DFB_BASE_SE) {
    // scale * np.exp(-dist_sq / 2)                                        kernel.py:176
    return __dmul_rn(f.scale, dfb_exp_nonpos(__dmul_rn(d2, -0.5)));
  }
  // Matern: dist = sqrt(D2); kernel.py:259-270, 292-299
  const double dist = sqrt(d2);
  const double mm = __dmul_rn(f.s8, dist);
  double u = 0.0;
  const int p = f.p;
  for (int i = 0; i <= p; i++) {
    const int e = p - i;
    double pw;
    if (e == 0) pw = 1.0;
    else if (e == 1) pw = mm;
    else if (e == 2) pw = __dmul_rn(mm, mm);
    else pw = pow(mm, (double)e);
    u = __dadd_rn(u, __dmul_rn(f.coeffs[i], pw));
  }
  const double w = __dmul_rn(f.gamma_ratio, dfb_exp_nonpos(__dmul_rn(-f.s2, dist)));
  u = __dmul_rn(u, w);
  return __dmul_rn(f.scale, u);
}

// (X**2).sum(axis=1) in NumPy's own association order (general_utils.py:66-67): add.reduce starts
// from the identity 0 and adds pairwise_sum(row): sequential for fewer than 8 elements, else eight
// interleaved accumulators combined as ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) plus a sequential
// tail (n <= 128: no recursive split).  Matching it keeps the rounding noise of
// D2(x, x) = (|x|^2 + |x|^2) - 2 x.x -- which sqrt() amplifies to ~1e-8 for Matern-1/2 -- identical
// to the reference's.
template <typename F>
__device__ __forceinline__ double numpy_sumsq(int n, F get) {
  double res;
  if (n < 8) {
    res = 0.0;
    for (int i = 0; i < n; i++) { const double v = get(i); res = __dadd_rn(res, __dmul_rn(v, v)); }
  } else {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { const double v = get(q); r[q] = __dmul_rn(v, v); }
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int q = 0; q < 8; q++) { const double v = get(i + q); r[q] = __dadd_rn(r[q], __dmul_rn(v, v)); }
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) { const double v = get(i); res = __dadd_rn(res, __dmul_rn(v, v)); }
  }
  return res;
}

// ---- scaled training set: x~ = x / bw (SoA, j contiguous) and per-factor squared norms ----------
// SEKernel.get_scaled_repr (kernel.py:179-181) + the (X2**2).sum(axis=1) of dist_squared
// (general_utils.py:66).
__global__ void prep_scaled_kernel(const dfb_kernel_desc* __restrict__ desc, int use_train_coords,
                                   const double* __restrict__ X, int64_t n, int d, double* xs,
                                   double* nrm, int64_t npad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  const int nf = desc->n_factors;
  for (int f = 0; f < nf; f++) {
    const dfb_factor_desc& fd = desc->factors[f];
    for (int q = 0; q < fd.n_dims; q++) {
      const int slot = fd.slot_off + q;
      double v = 0.0;
      if (j < n) {
        const int coord = use_train_coords ? desc->slot_train_coord[slot] : desc->slot_cand_coord[slot];
        v = X[j * d + coord] / desc->slot_bandwidth[slot];
      }
      xs[(int64_t)slot * npad + j] = v;
    }
    nrm[(int64_t)f * npad + j] = numpy_sumsq(fd.n_dims, [&](int q) {
      return xs[(int64_t)(fd.slot_off + q) * npad + j];
    });
  }
}

// ---- K_* rows for a block of candidates, fused with mu = mean + K_* alpha -----------------------
// Kernel.__call__(X_test, X) (kernel.py:72-83) + K_tetr.dot(alpha) (gp_core.py:173-174).
// One warp owns KSTAR_R candidate rows; its lanes stride over the training points so that the
// stores of each K_* row are 256 B coalesced and the training coordinates are read once per
// KSTAR_R rows.  The alpha-weighted row sum is reduced with warp shuffles.
constexpr int KSTAR_R = 2;
constexpr int KSTAR_WARPS = 8;
constexpr int KSTAR_CANDS = KSTAR_R * KSTAR_WARPS;

__global__ void __launch_bounds__(KSTAR_WARPS * 32)
kstar_kernel(const dfb_kernel_desc* __restrict__ desc_g, int cand_uses_train_coords,
             const double* __restrict__ xsT, const double* __restrict__ nrmT, int64_t npad_tr,
             const double* __restrict__ alpha, const double* __restrict__ Xc, int64_t m, int dc,
             int64_t m_rows, double* __restrict__ Ks, int64_t ldk, int64_t n_valid, int64_t n_write,
             double mean_const, double* __restrict__ mu, double* __restrict__ kss_out) {
  extern __shared__ __align__(16) unsigned char kraw[];
  dfb_kernel_desc* desc = reinterpret_cast<dfb_kernel_desc*>(kraw);
  {
    const int nwords = sizeof(dfb_kernel_desc) / 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(desc_g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(kraw);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int ns = desc->n_slots, nf = desc->n_factors, nt = desc->n_terms;
  double* xc = reinterpret_cast<double*>(kraw + ((sizeof(dfb_kernel_desc) + 15) / 16) * 16);
  double* nc = xc + KSTAR_CANDS * ns;
  const int64_t base = (int64_t)blockIdx.x * KSTAR_CANDS;

  for (int idx = threadIdx.x; idx < KSTAR_CANDS * ns; idx += blockDim.x) {
    const int r = idx / ns, s = idx - r * ns;
    const int64_t cand = base + r;
    double v = 0.0;
    if (cand < m) {
      const int coord = cand_uses_train_coords ? desc->slot_train_coord[s] : desc->slot_cand_coord[s];
      v = Xc[cand * dc + coord] / desc->slot_bandwidth[s];
    }
    xc[idx] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < KSTAR_CANDS * nf; idx += blockDim.x) {
    const int r = idx / nf, f = idx - r * nf;
    const dfb_factor_desc& fd = desc->factors[f];
    nc[idx] = numpy_sumsq(fd.n_dims, [&](int q) { return xc[r * ns + fd.slot_off + q]; });
  }
  __syncthreads();
  // k(x*, x*) the way the reference gets it: the diagonal of kernel(X_test, X_test)
  // (gp_core.py:179), i.e. through D2(x, x) = (|x|^2 + |x|^2) - 2 x.x with its rounding noise.
  if (kss_out != nullptr && threadIdx.x < KSTAR_CANDS) {
    const int r = threadIdx.x;
    const int64_t cand = base + r;
    if (cand < m) {
      double sum = 0.0;
      for (int t = 0; t < nt; t++) {
        double prod = desc->term_pre_scale[t];
        for (int f = desc->term_first_factor[t]; f < desc->term_first_factor[t + 1]; f++) {
          const dfb_factor_desc& fd = desc->factors[f];
          double dot = 0.0;
          for (int q = 0; q < fd.n_dims; q++) {
            const double v = xc[r * ns + fd.slot_off + q];
            dot = fma(v, v, dot);
          }
          const double nn = nc[r * nf + f];
          double d2 = __dadd_rn(__dadd_rn(nn, nn), -2.0 * dot);
          d2 = fmax(d2, 0.0);
          prod = __dmul_rn(prod, base_kernel_value(fd, d2));
        }
        sum = __dadd_rn(sum, prod);
      }
      kss_out[cand] = __dmul_rn(desc->post_scale, sum);
    }
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * KSTAR_R;
  const int64_t cand0 = base + r0;
  if (cand0 >= m_rows) return;
  double mu_acc[KSTAR_R];
#pragma unroll
  for (int r = 0; r < KSTAR_R; r++) mu_acc[r] = 0.0;
  const double post = desc->post_scale;

  for (int64_t j = lane; j < n_write; j += 32) {
    double kv[KSTAR_R];
#pragma unroll
    for (int r = 0; r < KSTAR_R; r++) kv[r] = 0.0;
    if (j < n_valid) {
      double sum[KSTAR_R];
#pragma unroll
      for (int r = 0; r < KSTAR_R; r++) sum[r] = 0.0;
      for (int t = 0; t < nt; t++) {
        double prod[KSTAR_R];
#pragma unroll
        for (int r = 0; r < KSTAR_R; r++) prod[r] = desc->term_pre_scale[t];
        for (int f = desc->term_first_factor[t]; f < desc->term_first_factor[t + 1]; f++) {
          const dfb_factor_desc& fd = desc->factors[f];
          double dot[KSTAR_R];
#pragma unroll
          for (int r = 0; r < KSTAR_R; r++) dot[r] = 0.0;
          for (int q = 0; q < fd.n_dims; q++) {
            const int s = fd.slot_off + q;
            const double xt = xsT[(int64_t)s * npad_tr + j];
#pragma unroll
            for (int r = 0; r < KSTAR_R; r++) dot[r] = fma(xc[(r0 + r) * ns + s], xt, dot[r]);
          }
          const double nt2 = nrmT[(int64_t)f * npad_tr + j];
#pragma unroll
          for (int r = 0; r < KSTAR_R; r++) {
            // (|y|^2 + |x|^2) - 2 x.y, clipped at 0                      general_utils.py:66-69
            double d2 = __dadd_rn(__dadd_rn(nt2, nc[(r0 + r) * nf + f]), -2.0 * dot[r]);
            d2 = fmax(d2, 0.0);
            prod[r] = __dmul_rn(prod[r], base_kernel_value(fd, d2));
          }
        }
#pragma unroll
        for (int r = 0; r < KSTAR_R; r++) sum[r] = __dadd_rn(sum[r], prod[r]);
      }
#pragma unroll
      for (int r = 0; r < KSTAR_R; r++) kv[r] = __dmul_rn(post, sum[r]);
    }
    const double aj = (alpha != nullptr && j < n_valid) ? alpha[j] : 0.0;
#pragma unroll
    for (int r = 0; r < KSTAR_R; r++) {
      const int64_t cand = cand0 + r;
      if (cand < m_rows) {
        const double v = (cand < m) ? kv[r] : 0.0;
        Ks[cand * ldk + j] = v;
        mu_acc[r] = fma(v, aj, mu_acc[r]);
      }
    }
  }
  if (mu != nullptr) {
#pragma unroll
    for (int r = 0; r < KSTAR_R; r++) {
      double s = mu_acc[r];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int64_t cand = cand0 + r;
      if (lane == 0 && cand < m) mu[cand] = mean_const + s;
    }
  }
}

// ---- fast path of kstar_kernel for the plain SE / Matern kernels (1 term, 1 factor, d <= 8) ------------
// Same arithmetic, same operation order; what changes is the bookkeeping: kind, p and d are compile
// time, the candidate coordinates live in registers and each lane carries 4 x KF_R independent entries
// (4 consecutive training points x KF_R candidate rows) through the exp/sqrt dependency chains.  (The generic kernel spends ~220 of its ~270
// instructions per entry interpreting the descriptor: profiles/r01_kstar_ncu_summary.txt.)
constexpr int KF_R = 2;
constexpr int KF_WARPS = 4;
// Two warps share a candidate pair, each over half of the training points: 2.76 waves of half-length
// warp tasks per 6528-candidate chunk instead of 1.38 waves of full-length ones (the second wave of
// which left 60 % of the machine idle); the two halves of mu are added in shared memory, in fixed order.
constexpr int KF_SPLIT = 2;
constexpr int KF_CANDS = KF_R * KF_WARPS / KF_SPLIT;

template <int KIND, int P>
__device__ __forceinline__ double base_value_fast(const dfb_factor_desc& f, double d2) {
  if (KIND == DFB_BASE_SE) return __dmul_rn(f.scale, dfb_exp_nonpos(__dmul_rn(d2, -0.5)));
  const double dist = dfb_sqrt_nonneg(d2);
  const double mm = __dmul_rn(f.s8, dist);
  double u;
  if (P == 0) {
    u = __dadd_rn(0.0, __dmul_rn(f.coeffs[0], 1.0));
  } else if (P == 1) {
    u = __dadd_rn(0.0, __dmul_rn(f.coeffs[0], mm));
    u = __dadd_rn(u, __dmul_rn(f.coeffs[1], 1.0));
  } else {
    u = __dadd_rn(0.0, __dmul_rn(f.coeffs[0], __dmul_rn(mm, mm)));
    u = __dadd_rn(u, __dmul_rn(f.coeffs[1], mm));
    u = __dadd_rn(u, __dmul_rn(f.coeffs[2], 1.0));
  }
  const double w = __dmul_rn(f.gamma_ratio, dfb_exp_nonpos(__dmul_rn(-f.s2, dist)));
  return __dmul_rn(f.scale, __dmul_rn(u, w));
}

// I8OUT: instead of the fp64 K_* rows, emit their six signed 7-bit digit planes (pair-interleaved layout
// of gemm_i8.cuh) for the tcgen05 contraction -- the fp64 matrix is then never written.
// Radix-256 digits for the CTA-pair kernel (gemm_i8c2.cuh): x (|x| <= 1/2) ~ a0 2^-7 + a1 2^-15 + a2 2^-23 +
// a3 2^-31 + a4 2^-39, a0 in [-64, 64], a1..a4 balanced bytes in [-128, 127]; |x - sum| <= 2^-40.
// hi = rint(x 2^15) and lo = rint((x 2^15 - hi) 2^24) are read off the low mantissa word of
// (value + 1.5 * 2^52); the bytes then peel off with sign extension, carries rippling upwards.
__device__ __forceinline__ void digits_radix256(double x, int (&a)[5]) {
  const double MAGIC = 6755399441055744.0;
  const double t1 = fma(x, 0x1p15, MAGIC);
  int hi = __double2loint(t1);
  const double rem = fma(x, 0x1p15, -(t1 - MAGIC));          // exact, |rem| <= 1/2
  const int lo = __double2loint(fma(rem, 0x1p24, MAGIC));
  a[4] = (int)(signed char)lo;
  int r = (lo - a[4]) >> 8;
  a[3] = (int)(signed char)r;
  r = (r - a[3]) >> 8;
  a[2] = (int)(signed char)r;
  hi += (r - a[2]) >> 8;
  a[1] = (int)(signed char)hi;
  a[0] = (hi - a[1]) >> 8;
}
__device__ __forceinline__ uint32_t pack4_i8(int a, int b, int c, int d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

struct KstarI8Out {
  uint8_t* planes;        // Ki8
  int64_t plane_bytes;    // 2 * chunk * npad
  int64_t row_bytes;      // 2 * npad
  double inv_colscale;    // 2^-F
  int kb;                 // k-values per interleave block (64 or 32)
  int radix256;           // 1: five radix-256 digits (digits_radix256), 0: six radix-128 digits
  const int* abort_count; // the launch is a no-op once *abort_count > abort_cap (shortlist overflow); may be NULL
  int abort_cap;
};

template <int KIND, int P, int D, bool I8OUT>
__global__ void __launch_bounds__(KF_WARPS * 32)
kstar_fast_kernel(const dfb_kernel_desc* __restrict__ desc_g, int cand_uses_train_coords,
                  const double* __restrict__ xsT, const double* __restrict__ nrmT, int64_t npad_tr,
                  const double* __restrict__ alpha, const double* __restrict__ Xc, int64_t m, int dc,
                  int64_t m_rows, double* __restrict__ Ks, int64_t ldk, int64_t n_valid, int64_t n_write,
                  double mean_const, double* __restrict__ mu, double* __restrict__ kss_out,
                  const KstarI8Out i8o) {
  if (I8OUT && i8o.abort_count != nullptr && *i8o.abort_count > i8o.abort_cap) return;
  __shared__ dfb_factor_desc fsh;
  __shared__ int coord_sh[8];
  __shared__ double bw_sh[8];
  __shared__ double scal_sh[2];
  if (threadIdx.x == 0) {
    fsh = desc_g->factors[0];
    scal_sh[0] = desc_g->term_pre_scale[0];
    scal_sh[1] = desc_g->post_scale;
  }
  if (threadIdx.x < D) {
    coord_sh[threadIdx.x] = cand_uses_train_coords ? desc_g->slot_train_coord[threadIdx.x]
                                                   : desc_g->slot_cand_coord[threadIdx.x];
    bw_sh[threadIdx.x] = desc_g->slot_bandwidth[threadIdx.x];
  }
  __syncthreads();
  const dfb_factor_desc f = fsh;
  const double pre = scal_sh[0], post = scal_sh[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = warp % KF_SPLIT;
  const int64_t cand0 = (int64_t)blockIdx.x * KF_CANDS + (warp / KF_SPLIT) * KF_R;
  const bool active = cand0 < m_rows;
  const int64_t span = ((n_write + 128 * KF_SPLIT - 1) / (128 * KF_SPLIT)) * 128;
  const int64_t j_lo = active ? part * span : n_write;
  const int64_t j_hi = (j_lo + span < n_write) ? j_lo + span : n_write;
  __shared__ double mu_sh[KF_WARPS][KF_R];

  double xc[KF_R][D], nc[KF_R];
#pragma unroll
  for (int r = 0; r < KF_R; r++) {
    const int64_t cand = cand0 + r;
#pragma unroll
    for (int q = 0; q < D; q++) xc[r][q] = (cand < m) ? Xc[cand * dc + coord_sh[q]] / bw_sh[q] : 0.0;
    double s = 0.0;                    // numpy_sumsq for n < 8; n == 8 uses the 8-accumulator form
    if (D < 8) {
#pragma unroll
      for (int q = 0; q < D; q++) s = __dadd_rn(s, __dmul_rn(xc[r][q], xc[r][q]));
    } else {
      s = numpy_sumsq(D, [&](int q) { return xc[r][q]; });
    }
    nc[r] = s;
  }
  if (kss_out != nullptr && lane == 0 && part == 0 && active) {
#pragma unroll
    for (int r = 0; r < KF_R; r++) {
      const int64_t cand = cand0 + r;
      if (cand < m) {
        double dot = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) dot = fma(xc[r][q], xc[r][q], dot);
        double d2 = __dadd_rn(__dadd_rn(nc[r], nc[r]), -2.0 * dot);
        d2 = fmax(d2, 0.0);
        const double prod = __dmul_rn(pre, base_value_fast<KIND, P>(f, d2));
        kss_out[cand] = __dmul_rn(post, __dadd_rn(0.0, prod));
      }
    }
  }

  double mu_acc[KF_R];
#pragma unroll
  for (int r = 0; r < KF_R; r++) mu_acc[r] = 0.0;
  // each lane owns 4 consecutive training points per step (128 per warp step): 16-byte loads of the
  // SoA coordinates, 16-byte stores of the fp64 rows or one packed 32-bit store per digit plane, and
  // 4 x KF_R independent exp/sqrt chains in flight
  for (int64_t j0 = j_lo + 4 * lane; j0 < j_hi; j0 += 128) {
    double kv[KF_R][4];
    double aj[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < KF_R; r++)
#pragma unroll
      for (int e = 0; e < 4; e++) kv[r][e] = 0.0;
    if (j0 < n_valid) {
      double xt[D][4];
#pragma unroll
      for (int q = 0; q < D; q++) {
        const double2 lo = *reinterpret_cast<const double2*>(xsT + (int64_t)q * npad_tr + j0);
        const double2 hi = *reinterpret_cast<const double2*>(xsT + (int64_t)q * npad_tr + j0 + 2);
        xt[q][0] = lo.x; xt[q][1] = lo.y; xt[q][2] = hi.x; xt[q][3] = hi.y;
      }
      double nt2[4];
      {
        const double2 lo = *reinterpret_cast<const double2*>(nrmT + j0);
        const double2 hi = *reinterpret_cast<const double2*>(nrmT + j0 + 2);
        nt2[0] = lo.x; nt2[1] = lo.y; nt2[2] = hi.x; nt2[3] = hi.y;
      }
      if (alpha != nullptr) {
        const double2 lo = *reinterpret_cast<const double2*>(alpha + j0);
        const double2 hi = *reinterpret_cast<const double2*>(alpha + j0 + 2);
        aj[0] = lo.x; aj[1] = lo.y; aj[2] = hi.x; aj[3] = hi.y;
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const bool valid = (j0 + e) < n_valid;
        if (!valid) aj[e] = 0.0;
#pragma unroll
        for (int r = 0; r < KF_R; r++) {
          double dot = 0.0;
#pragma unroll
          for (int q = 0; q < D; q++) dot = fma(xc[r][q], xt[q][e], dot);
          double d2 = __dadd_rn(__dadd_rn(nt2[e], nc[r]), -2.0 * dot);
          d2 = fmax(d2, 0.0);
          const double prod = __dmul_rn(pre, base_value_fast<KIND, P>(f, d2));
          kv[r][e] = valid ? __dmul_rn(post, __dadd_rn(0.0, prod)) : 0.0;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < KF_R; r++) {
      const int64_t cand = cand0 + r;
      if (cand < m_rows) {
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (cand < m) ? kv[r][e] : 0.0;
        if (I8OUT) {
          // exact digit expansion of v * 2^-F (0 <= v * 2^-F < 1/2, so every digit fits int8 without
          // clamping; see slice_i8_kernel); four columns packed per 32-bit store
          // x = hi 2^-21 + lo 2^-42 with hi = rint(x 2^21), lo = rint((x 2^21 - hi) 2^21), both read off the
          // low mantissa word of (value + 1.5 * 2^52); each 21-bit half then splits into three balanced
          // 7-bit digits with 32-bit integer shifts.
          uint8_t* dst = i8o.planes + cand * i8o.row_bytes + (j0 / i8o.kb) * (2 * i8o.kb) + (j0 % i8o.kb);
          if (i8o.radix256) {
            int dg[4][5];
#pragma unroll
            for (int e = 0; e < 4; e++) digits_radix256(v[e] * i8o.inv_colscale, dg[e]);
#pragma unroll
            for (int sd = 0; sd < 5; sd++)
              *reinterpret_cast<uint32_t*>(dst + (int64_t)(sd >> 1) * i8o.plane_bytes + (sd & 1) * i8o.kb) =
                  pack4_i8(dg[0][sd], dg[1][sd], dg[2][sd], dg[3][sd]);
            // compact copy of the leading digit (plain row-major rows) for pass B of gemm_i8c2.cuh
            *reinterpret_cast<uint32_t*>(i8o.planes + 3 * i8o.plane_bytes + cand * (i8o.row_bytes >> 1) + j0) =
                pack4_i8(dg[0][0], dg[1][0], dg[2][0], dg[3][0]);
          } else {
          const double MAGIC = 6755399441055744.0;
          int hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const double x = v[e] * i8o.inv_colscale;
            const double t1 = fma(x, 0x1p21, MAGIC);
            hi[e] = __double2loint(t1);
            const double rem = fma(x, 0x1p21, -(t1 - MAGIC));       // exact, |rem| <= 1/2
            lo[e] = __double2loint(fma(rem, 0x1p21, MAGIC));
          }
          uint32_t pack[I8_S];
#pragma unroll
          for (int hsel = 0; hsel < 2; hsel++) {
            int a1[4], a2[4], a3[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int w = hsel ? lo[e] : hi[e];
              a1[e] = (w + 8192) >> 14;
              const int r1 = w - (a1[e] << 14);
              a2[e] = (r1 + 64) >> 7;
              a3[e] = r1 - (a2[e] << 7);
            }
            pack[3 * hsel + 0] = __byte_perm(__byte_perm(a1[0], a1[1], 0x0040), __byte_perm(a1[2], a1[3], 0x0040), 0x5410);
            pack[3 * hsel + 1] = __byte_perm(__byte_perm(a2[0], a2[1], 0x0040), __byte_perm(a2[2], a2[3], 0x0040), 0x5410);
            pack[3 * hsel + 2] = __byte_perm(__byte_perm(a3[0], a3[1], 0x0040), __byte_perm(a3[2], a3[3], 0x0040), 0x5410);
          }
#pragma unroll
          for (int sd = 0; sd < I8_S; sd++)
            *reinterpret_cast<uint32_t*>(dst + (int64_t)(sd >> 1) * i8o.plane_bytes + (sd & 1) * i8o.kb) = pack[sd];
          }
        } else {
          double2 lo, hi;
          lo.x = v[0]; lo.y = v[1]; hi.x = v[2]; hi.y = v[3];
          *reinterpret_cast<double2*>(Ks + cand * ldk + j0) = lo;
          *reinterpret_cast<double2*>(Ks + cand * ldk + j0 + 2) = hi;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) mu_acc[r] = fma(v[e], aj[e], mu_acc[r]);
      }
    }
  }
  if (mu != nullptr) {
#pragma unroll
    for (int r = 0; r < KF_R; r++) {
      double s = mu_acc[r];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) mu_sh[warp][r] = s;
    }
    __syncthreads();
    if (part == 0 && lane == 0 && active) {
#pragma unroll
      for (int r = 0; r < KF_R; r++) {
        double s = mu_sh[warp][r];
#pragma unroll
        for (int q = 1; q < KF_SPLIT; q++) s += mu_sh[warp + q][r];
        const int64_t cand = cand0 + r;
        if (cand < m) mu[cand] = mean_const + s;
      }
    }
  }
}

// ================================================================================================
// K_* digit planes for the CTA-pair contraction, second generation (radix 256 only): kstar_seg_kernel.
//
// What changed against kstar_fast_kernel<.., I8OUT = true> (which stays as the radix-128 / fallback path):
//  * TRAINING-STATIONARY loop nest: a warp owns 64 training points -- two per lane, their scaled coordinates, norms
//    and alpha held in registers for the whole kernel -- and streams candidate rows past them, two rows per iteration
//    (four independent dependency chains per lane).  The only loads in the loop are the warp-uniform candidate rows
//    (64 bytes each).  ncu on the first version of this kernel, which re-read its training slice from L1 every row,
//    showed the fp64 pipe at 47 % with half of all issue slots lost to long-scoreboard stalls
//    (profiles/r02_kstar_seg_ncu_summary.txt); with nothing left to wait for, a handful of warps per SM is enough;
//  * which is the point: at <= 104 registers and no shared memory, one 4-warp CTA of it (one warp per SM sub-partition)
//    co-resides with the persistent tcgen05 kernel (gemm_i8c2.cuh: launched at 136 registers x 12 warps, 225 KB of
//    shared memory), and api.cu issues chunk c+1's K_* on a second stream while chunk c is being contracted: fp64 pipe
//    and tensor pipe of the same SM at the same time;
//  * candidate-side work (x / bw, |x~|^2, k(x*,x*) in the reference's own operation order) is done once per row by
//    cand_prep_kernel instead of once per (row, training slice);
//  * the five digits of x = v 2^-F come from ONE fused multiply-add: t = x 2^39 + (1.5 2^52 + 0x80808080) holds
//    q + bias in its low mantissa bits, and adding 0x80 to every byte position and then flipping that bit IS the
//    balanced (signed-byte) base-256 expansion -- no integer carry chain;
//  * the kernel value is formed with fused constants and FMA contraction: v differs from the reference-order value
//    of kstar_kernel by a few ulp, which is 10^6 times below the int8 screen's own error allowance, and every
//    candidate that matters is re-scored by the exact-order fp64 path anyway (dfb_score_argmax).  Exception: the
//    squared distance of Matern-1/2 keeps the reference's rounding order (see the loop);
//  * padding needs no selects: training points beyond n carry the norm 1e200 (their kernel value underflows to an
//    exact 0 for SE and Matern alike), candidate rows beyond m likewise;
//  * mu leaves as per-block partial sums mu_part[block][row] (one packed butterfly per row pair), added in a fixed
//    order by mu_reduce_kernel.
// Algorithmic bytes per candidate: 6 npad written (five digit planes + the compact leading plane) + 8 (D+2) read.
// ================================================================================================
constexpr int KS_BLK = 64;         // training points per warp (two per lane), register-resident
constexpr int KS_ROWS = 64;        // candidate rows per CTA
constexpr int KS_WARPS = 4;        // one warp per SM sub-partition
// Register cap.  104 is what fits beside the persistent contraction kernel (3328 registers per sub-partition are left by its
// 12 warps x 136 registers); stand-alone -- the default, see kstar_overlap in DESIGN.md 5.1 -- 128 is fastest
// (0.200 ms per 6528 x 5120 chunk against 0.209 / 0.219 / 0.244 at 104 / 88 / 72: profiles/r02_kstar_variants.txt).
#ifndef DFB_KS_MAXREG
#define DFB_KS_MAXREG 128
#endif
constexpr int KS_MAXREG = DFB_KS_MAXREG;
constexpr double KS_FAR = 1e200;   // squared norm of padding points: exp(-sqrt(1e200) c) == 0, no overflow on the way

template <int KIND, int P, int D>
__global__ void cand_prep_kernel(const dfb_kernel_desc* __restrict__ desc_g, const double* __restrict__ Xc, int64_t m,
                                 int dc, int64_t m_rows, double* __restrict__ cprep, double* __restrict__ kss_out) {
  constexpr int CP = (D + 2) & ~1;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_rows) return;
  double xc[D];
  double nc = KS_FAR;
#pragma unroll
  for (int q = 0; q < D; q++) xc[q] = 0.0;
  if (r < m) {
#pragma unroll
    for (int q = 0; q < D; q++) xc[q] = Xc[r * dc + desc_g->slot_cand_coord[q]] / desc_g->slot_bandwidth[q];
    if (D < 8) {
      nc = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) nc = __dadd_rn(nc, __dmul_rn(xc[q], xc[q]));
    } else {
      nc = numpy_sumsq(D, [&](int q) { return xc[q]; });
    }
    if (kss_out != nullptr) {
      // k(x*, x*) through D2(x, x) = (|x|^2 + |x|^2) - 2 x.x with its rounding noise (gp_core.py:179)
      const dfb_factor_desc f = desc_g->factors[0];
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) dot = fma(xc[q], xc[q], dot);
      double d2 = __dadd_rn(__dadd_rn(nc, nc), -2.0 * dot);
      d2 = fmax(d2, 0.0);
      const double prod = __dmul_rn(desc_g->term_pre_scale[0], base_value_fast<KIND, P>(f, d2));
      kss_out[r] = __dmul_rn(desc_g->post_scale, __dadd_rn(0.0, prod));
    }
  }
#pragma unroll
  for (int q = 0; q < D; q++) cprep[r * CP + q] = xc[q];
  cprep[r * CP + D] = nc;
  if (CP > D + 1) cprep[r * CP + D + 1] = 0.0;
}

// constants of the segment kernel as constant-bank operands (literals would be re-materialised with two moves each)
__constant__ double ks_cd[6] = {1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
                                6755399441055744.0, 6755399441055744.0 + 2155905152.0, 0x1p-960};

// exp(x), x <= 0 (or a rounding residue above 0): dfb_exp_nonpos without the argument clamp -- whatever the polynomial
// makes of x < -707 is discarded by the final select (no traps on the device); NaN propagates through the arithmetic
__device__ __forceinline__ double ks_exp(double x) {
  const double t = fma(x, ks_cd[0], ks_cd[3]);
  const int n = __double2loint(t);
  const double tn = t - ks_cd[3];
  double r = fma(tn, ks_cd[1], x);
  r = fma(tn, ks_cd[2], r);
  double p = dfb_exp_cd[0];
#pragma unroll
  for (int i = 1; i < 14; i++) p = fma(p, r, dfb_exp_cd[i]);
  const double out = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return (x < -707.0) ? 0.0 : out;
}

// Four of them, stage by stage (see the staging note in kstar_seg_kernel).  Degree-13 polynomial by Horner's rule: with
// 16 warps per SM stand-alone the kernel is bound by instruction count, not by the depth of the dependency chain, and
// Horner is three multiplications shorter (0.171 ms per chunk against 0.189 for the Estrin form, which -DDFB_KS_ESTRIN
// keeps for the co-resident, latency-bound setting of kstar_overlap: depth 4 instead of 13).
__device__ __forceinline__ void ks_exp4(const double (&x)[4], double (&out)[4]) {
  double t[4], r[4], p[4];
  int n[4];
#pragma unroll
  for (int e = 0; e < 4; e++) { t[e] = fma(x[e], ks_cd[0], ks_cd[3]); n[e] = __double2loint(t[e]); t[e] -= ks_cd[3]; }
#pragma unroll
  for (int e = 0; e < 4; e++) r[e] = fma(t[e], ks_cd[2], fma(t[e], ks_cd[1], x[e]));
#ifndef DFB_KS_ESTRIN
#pragma unroll
  for (int e = 0; e < 4; e++) p[e] = dfb_exp_cd[0];
#pragma unroll
  for (int i = 1; i < 14; i++) {
#pragma unroll
    for (int e = 0; e < 4; e++) p[e] = fma(p[e], r[e], dfb_exp_cd[i]);
  }
#else
  double r2[4], r4[4], a[4][7];
#pragma unroll
  for (int e = 0; e < 4; e++) r2[e] = r[e] * r[e];
#pragma unroll
  for (int k = 0; k < 7; k++) {
#pragma unroll
    for (int e = 0; e < 4; e++) a[e][k] = fma(dfb_exp_cd[13 - (2 * k + 1)], r[e], dfb_exp_cd[13 - 2 * k]);
  }
#pragma unroll
  for (int e = 0; e < 4; e++) r4[e] = r2[e] * r2[e];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const double b0 = fma(a[e][1], r2[e], a[e][0]);
    const double b1 = fma(a[e][3], r2[e], a[e][2]);
    const double b2 = fma(a[e][5], r2[e], a[e][4]);
    const double e0 = fma(b1, r4[e], b0);
    const double e1 = fma(a[e][6], r4[e], b2);
    p[e] = fma(e1, r4[e] * r4[e], e0);
  }
#endif
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const double o = __hiloint2double(__double2hiint(p[e]) + (n[e] << 20), __double2loint(p[e]));
    out[e] = (x[e] < -707.0) ? 0.0 : o;
  }
}

struct KsegArgs {
  const double* xsT; const double* nrm; const double* alpha; int64_t npad_tr; int64_t n_valid;
  const double* cprep; int64_t m_rows; int64_t n_write;
  uint8_t* planes; int64_t plane_bytes; int64_t row_bytes;
  double cdig;       // 2^-F 2^39: kernel value -> q
  double cval;       // post * pre * scale (* Gamma(p+1)/Gamma(2p+1)): the constant factors of the kernel, fused
  double s8, ms2, c0, c1, c2;         // ms2 = -sqrt(2 nu)
  double* mu_part; int64_t ld_mu;       // mu_part may be NULL (no mu wanted)
  const int* abort_count; int abort_cap;
  double* rows64; int64_t ld64;         // ROWS64 variant
};

// ROWS64 = true: the same kernel writing the fp64 K_* rows (g.rows64, leading dimension g.ld64) instead of the digit
// planes -- the materialising build of the fp64 scoring path, dfb_eval and the Thompson-sampling blocks.
template <int KIND, int P, int D, bool ROWS64>
__global__ void __maxnreg__(KS_MAXREG) kstar_seg_kernel(const KsegArgs g) {
  if (g.abort_count != nullptr && *g.abort_count > g.abort_cap) return;
  constexpr int CP = (D + 2) & ~1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long t_begin = (threadIdx.x == 0 && g_trace != nullptr) ? trace_now() : 0ull;
  const int blk = blockIdx.x * KS_WARPS + warp;                   // this warp's block of 64 training points
  const int64_t j = (int64_t)blk * KS_BLK + 2 * lane;             // this lane's two points: j, j + 1
  if (j >= g.n_write) return;                                     // whole warps only (no barriers in this kernel)
  double xt[D][2], nt2[2], aj[2];
#pragma unroll
  for (int q = 0; q < D; q++) {
    const double2 v = *reinterpret_cast<const double2*>(g.xsT + (int64_t)q * g.npad_tr + j);
    xt[q][0] = v.x; xt[q][1] = v.y;
  }
  {
    const double2 v = *reinterpret_cast<const double2*>(g.nrm + j);
    nt2[0] = (j < g.n_valid) ? v.x : KS_FAR;
    nt2[1] = (j + 1 < g.n_valid) ? v.y : KS_FAR;
    aj[0] = 0.0; aj[1] = 0.0;
    if (g.alpha != nullptr) {
      const double2 a2 = *reinterpret_cast<const double2*>(g.alpha + j);
      aj[0] = (j < g.n_valid) ? a2.x : 0.0;
      aj[1] = (j + 1 < g.n_valid) ? a2.y : 0.0;
    }
  }
  const int64_t r_lo = (int64_t)blockIdx.y * KS_ROWS;
  const int64_t r_hi = (r_lo + KS_ROWS < g.m_rows) ? r_lo + KS_ROWS : g.m_rows;       // m_rows is a multiple of 128
  // byte offset of this lane's two digits inside a pair-interleaved row: ((j >> 5) << 6) + (j & 31)
  const unsigned doff = (unsigned)(((j >> 5) << 6) + (j & 31));
  const bool up = (lane & 16) != 0;
  // candidate rows r, r + 1 (warp-uniform, 64 bytes each): loaded one iteration ahead -- the coordinates are dead after the
  // dot-product stage, so the next pair is fetched into the same registers right there and has the rest of the iteration
  // (~250 instructions) to arrive (ncu: long-scoreboard was the kernel's top stall with the loads at the top of the loop)
  double xc[2][D], nc[2];
  auto load_rows = [&](int64_t r) {
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const double2* cp = reinterpret_cast<const double2*>(g.cprep + (r + rr) * CP);
#pragma unroll
      for (int q = 0; q < D; q += 2) {
        const double2 w2 = cp[q >> 1];
        xc[rr][q] = w2.x;
        if (q + 1 < D) xc[rr][q + 1] = w2.y;
      }
      nc[rr] = g.cprep[(r + rr) * CP + D];
    }
  };
  if (r_lo < r_hi) load_rows(r_lo);
  for (int64_t r = r_lo; r < r_hi; r += 2) {
    // Rows r and r + 1 against the lane's two points: four independent dependency chains c = 2 * row + point, advanced
    // STAGE BY STAGE (every stage an unrolled loop over c) so that all four stay in flight.
    double d2[4], v[4];
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        if (KIND == DFB_BASE_MATERN && P == 0) {
          // Matern-1/2 is not smooth at 0: for a candidate that coincides with a training point d2 is pure rounding
          // residue and sqrt() turns 1e-16 of it into 1e-8 of the kernel value, so d2 is formed exactly as the
          // reference-order kernels form it: sequential FMA chain, (|y|^2 + |x|^2) - 2 x.y (general_utils.py:66-69)
          double dot = 0.0;
#pragma unroll
          for (int q = 0; q < D; q++) dot = fma(xc[rr][q], xt[q][e], dot);
          d2[2 * rr + e] = __dadd_rn(__dadd_rn(nt2[e], nc[rr]), -2.0 * dot);
        } else {
          // smooth at 0 (SE, Matern-3/2, -5/2: value = 1 - O(d2)): the residue is harmless, so the dot product runs
          // as two half-length chains and d2 by one fused multiply-add (shorter dependency chain)
          double p0 = xc[rr][0] * xt[0][e], p1 = (D > 1) ? xc[rr][1] * xt[1][e] : 0.0;
#pragma unroll
          for (int q = 2; q < D; q += 2) {
            p0 = fma(xc[rr][q], xt[q][e], p0);
            if (q + 1 < D) p1 = fma(xc[rr][q + 1], xt[q + 1][e], p1);
          }
          d2[2 * rr + e] = fma(-2.0, p0 + p1, nt2[e] + nc[rr]);
        }
      }
    }
    if (r + 2 < r_hi) load_rows(r + 2);
    if (KIND == DFB_BASE_SE) {
      // d2 < 0 (rounding residue of coincident points) -> exp(+1e-16) = 1: no clip needed
      double x[4];
#pragma unroll
      for (int c = 0; c < 4; c++) x[c] = d2[c] * -0.5;
      ks_exp4(x, v);
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] *= g.cval;
    } else {
      // dist = sqrt(max(d2, 0)); d2 below 2^-960 (including the negative rounding residues) gives dist = 0
      double y[4], ee[4], dist[4];
#pragma unroll
      for (int c = 0; c < 4; c++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[c]) : "d"(d2[c]));
#pragma unroll
      for (int c = 0; c < 4; c++) ee[c] = fma(d2[c], -(y[c] * y[c]), 1.0);
#pragma unroll
      for (int c = 0; c < 4; c++) y[c] = fma(fma(ee[c], 0.375, 0.5), y[c] * ee[c], y[c]);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        // y ~ d2^-1/2 to 2^-58 after the third-order refinement: d2 y is sqrt(d2) to an ulp (the Markstein
        // correction of dfb_sqrt_nonneg would add three dependent operations for the last half ulp)
        const double ss = d2[c] * y[c];
        dist[c] = (d2[c] < ks_cd[5]) ? 0.0 : ss;             // NaN d2: comparison false, ss = NaN propagates
      }
      double x[4], u[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        x[c] = g.ms2 * dist[c];
        const double mm = g.s8 * dist[c];
        if (P == 0) u[c] = g.c0;
        else if (P == 1) u[c] = fma(g.c0, mm, g.c1);
        else u[c] = fma(fma(g.c0, mm, g.c1), mm, g.c2);
        u[c] *= g.cval;
      }
      ks_exp4(x, v);
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] *= u[c];
    }
    // mu partials of the two rows over this warp's 64 points: one packed butterfly (the half-warps swap the row they
    // do not keep in the first step), lane 0 ends up with row r, lane 16 with row r + 1
    {
      const double m0 = fma(v[1], aj[1], v[0] * aj[0]);
      const double m1 = fma(v[3], aj[1], v[2] * aj[0]);
      double keep = up ? m1 : m0;
      const double send = up ? m0 : m1;
      keep += __shfl_xor_sync(0xffffffffu, send, 16);
      keep += __shfl_xor_sync(0xffffffffu, keep, 8);
      keep += __shfl_xor_sync(0xffffffffu, keep, 4);
      keep += __shfl_xor_sync(0xffffffffu, keep, 2);
      keep += __shfl_xor_sync(0xffffffffu, keep, 1);
      if ((lane & 15) == 0 && g.mu_part != nullptr) g.mu_part[(int64_t)blk * g.ld_mu + r + (lane >> 4)] = keep;
    }
    if (ROWS64) {
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        double2 o;
        o.x = v[2 * rr]; o.y = v[2 * rr + 1];
        *reinterpret_cast<double2*>(g.rows64 + (r + rr) * g.ld64 + j) = o;
      }
      continue;
    }
    // digits: word of chain c = bytes (a4, a3, a2, a1), a0 in the low byte of the high word; two points per 16-bit store
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const double t0 = fma(v[2 * rr], g.cdig, ks_cd[4]), t1 = fma(v[2 * rr + 1], g.cdig, ks_cd[4]);
      const unsigned w0 = (unsigned)__double2loint(t0) ^ 0x80808080u, w1 = (unsigned)__double2loint(t1) ^ 0x80808080u;
      const unsigned short d0 = (unsigned short)__byte_perm((unsigned)__double2hiint(t0), (unsigned)__double2hiint(t1), 0x0040);
      const unsigned short d1 = (unsigned short)__byte_perm(w0, w1, 0x0073);
      const unsigned short d2w = (unsigned short)__byte_perm(w0, w1, 0x0062);
      const unsigned short d3 = (unsigned short)__byte_perm(w0, w1, 0x0051);
      const unsigned short d4 = (unsigned short)__byte_perm(w0, w1, 0x0040);
      // pair-interleaved planes: digits (2p, 2p+1) side by side in 32-byte k segments (gemm_i8c2.cuh)
      uint8_t* dst = g.planes + (r + rr) * g.row_bytes + doff;
      *reinterpret_cast<unsigned short*>(dst) = d0;
      *reinterpret_cast<unsigned short*>(dst + 32) = d1;
      *reinterpret_cast<unsigned short*>(dst + g.plane_bytes) = d2w;
      *reinterpret_cast<unsigned short*>(dst + g.plane_bytes + 32) = d3;
      *reinterpret_cast<unsigned short*>(dst + 2 * g.plane_bytes) = d4;
      *reinterpret_cast<unsigned short*>(g.planes + 3 * g.plane_bytes + (r + rr) * (g.row_bytes >> 1) + j) = d0;
    }
  }
  if (threadIdx.x == 0) trace_emit(1u, t_begin);       // warp 0 of the CTA: representative (all warps do equal work)
}

__global__ void mu_reduce_kernel(const double* __restrict__ part, int n_seg, int64_t ld, int64_t m, double mean_const,
                                 double* __restrict__ mu) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double s = 0.0;
  for (int k = 0; k < n_seg; k++) s += part[(int64_t)k * ld + r];
  mu[r] = mean_const + s;
}

// ---- tall factorisation matrix set-up -------------------------------------------------------------
// T = [ K + (noise + jitter) I (padded with identity) ; I ; y_c^T (row 0 of the last block) ].
__global__ void init_tall_kernel(double* T, int64_t n, int64_t npad, double diag_add,
                                 const double* __restrict__ yc, int with_bottom) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  if (i < n) T[i * npad + i] += diag_add;       // K_trtr_wo_noise + noise_var * np.eye (gp_core.py:843)
  else T[i * npad + i] = 1.0;
  if (with_bottom) T[(npad + i) * npad + i] = 1.0;
  T[2 * npad * npad + i] = (i < n) ? yc[i] : 0.0;
}

// ---- diagonal block: Cholesky factor AND its inverse in one pass ---------------------------------------
// Factorises the 128 x 128 block T[step] in place (lower) and writes W_kk = L_kk^-1 (lower) to
// Dinv.  The inverse costs no extra storage: the same right-looking elimination is applied to the
// virtual tall block [A_kk ; I]; the strictly-upper triangle of the shared array holds the evolving
// L_kk^-T while the lower triangle holds L_kk.  A non-positive (or NaN) pivot reports
// info = global index + 1, the LAPACK dpotrf convention behind np.linalg.LinAlgError
// (general_utils.py:176-180).
// Register-resident version: thread (ty, tx) of a 16 x 16 grid owns the 8 x 8 block-cyclic elements
// (ty + 16a, tx + 16b); per column only the scaled column (128 values) goes through shared memory.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* T, int64_t ld, int step,
                                                         double* Dinv, int* info) {
  __shared__ double colbuf[TILE];
  __shared__ double dLs[TILE];
  __shared__ double piv_sh;
  if (*info != 0) return;
  double* blk = T + (int64_t)step * TILE * ld + (int64_t)step * TILE;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double e[8][8];
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int i = ty + 16 * a, c = tx + 16 * b;
      e[a][b] = (c <= i) ? blk[(int64_t)i * ld + c] : 0.0;
    }
  for (int j = 0; j < TILE; j++) {
    const int jb = j >> 4, jt = j & 15;
    if (ty == jt && tx == jt) {
#pragma unroll
      for (int a = 0; a < 8; a++)
        if (a == jb) piv_sh = e[a][a];
    }
    __syncthreads();
    const double piv = piv_sh;
    if (!(piv > 0.0)) {                 // uniform across the block
      if (tid == 0) atomicCAS(info, 0, step * TILE + j + 1);
      return;
    }
    // d_j = sqrt(pivot) and 1 / d_j from ONE reciprocal-square-root seed: the two software sequences of sqrt() and of the
    // division (~19 dependent fp64 operations, paid 128 times in a row by a single CTA) share their refinement -- 9
    // dependent operations.  d_j is the correctly rounded root (the Markstein step of CUDA's own sqrt); 1 / d_j is one
    // Newton step of y ~ pivot^-1/2 against the ROUNDED d_j, i.e. the reciprocal dpotf2 scales by, to well below an ulp.
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
    double dj, rj;
    if (piv >= 0x1p-900 && piv <= 0x1p900) {
      const double e0 = fma(piv, -(y * y), 1.0);
      y = fma(fma(e0, 0.375, 0.5), y * e0, y);                    // pivot^-1/2 to ~2^-58
      const double g0 = piv * y;
      dj = fma(fma(-g0, g0, piv), 0.5 * y, g0);
      rj = fma(y, fma(-dj, y, 1.0), y);
    } else {                                                      // out of the seed's comfortable range: the library pair
      dj = sqrt(piv);
      rj = 1.0 / dj;
    }
    if (tx == jt) {                     // owners of column j scale it and publish it
#pragma unroll
      for (int b = 0; b < 8; b++) {
        if (b == jb) {
#pragma unroll
          for (int a = 0; a < 8; a++) {
            const int i = ty + 16 * a;
            const double nv = (i == j) ? rj : e[a][b] * rj;   // dpotf2 scales by the reciprocal too
            e[a][b] = nv;
            colbuf[i] = nv;
          }
        }
      }
      if (ty == jt) dLs[j] = dj;
    }
    __syncthreads();
    // e(i,c) -= m_i * col_c for c > j and (i <= j [inverse rows] or c <= i [Cholesky rows]);
    // m_i = colbuf[i] (1/d_j for i == j), col_c = colbuf[c].  Branch-free: an element is statically
    // lower (c <= i: active iff c > j) or upper (c > i: active iff i <= j and c > j), so masking the
    // column factor by (c > j) and, for upper elements, the row factor by (i <= j) leaves 64 plain FMAs.
    double mrow[8], mrow_inv[8], mcol[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int i = ty + 16 * a;
      mrow[a] = colbuf[i];
      mrow_inv[a] = (i <= j) ? mrow[a] : 0.0;
    }
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int c = tx + 16 * b;
      mcol[b] = (c > j) ? colbuf[c] : 0.0;
    }
    const bool diag_lower = (tx <= ty);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const double mdiag = diag_lower ? mrow[a] : mrow_inv[a];
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const double mr = (b < a) ? mrow[a] : ((b == a) ? mdiag : mrow_inv[a]);
        e[a][b] = fma(-mr, mcol[b], e[a][b]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int i = ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int c = tx + 16 * b;
      if (c < i) {
        blk[(int64_t)i * ld + c] = e[a][b];                 // L
      } else if (c == i) {
        blk[(int64_t)i * ld + c] = dLs[i];
        Dinv[i * TILE + c] = e[a][b];                        // 1 / L_ii
      } else {
        blk[(int64_t)i * ld + c] = 0.0;
        Dinv[i * TILE + c] = 0.0;                            // L^-1 is lower triangular
        Dinv[c * TILE + i] = e[a][b];                        // (L^-T)[i][c] = (L^-1)[c][i]
      }
    }
  }
}

// ---- W = (L^-T)^T : 32 x 32 tile transpose ----------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n) {
  __shared__ double tile[32][33];
  const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    tile[r][threadIdx.x] = src[(by + r) * n + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    dst[(bx + r) * n + by + threadIdx.x] = tile[threadIdx.x][r];
}

// ---- alpha = L^-T (L^-1 y) = rows of L^-T dotted with v = L^-1 y  (gp_core.py:162-163) ---------
__global__ void alpha_kernel(const double* __restrict__ Wt, const double* __restrict__ v,
                             double* alpha, int64_t n, int64_t npad) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= npad) return;
  double s = 0.0;
  if (j < n) {
    const double* row = Wt + j * npad;
    for (int64_t i = (j & ~(int64_t)31) + lane; i < npad; i += 32) s = fma(row[i], v[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (lane == 0) alpha[j] = s;
}

// ---- LML pieces: sum log L_ii, y_c . alpha, |L^-1 y_c|^2  (gp_core.py:222-227) --------------------
__global__ void lml_reduce_kernel(const double* __restrict__ T, const double* __restrict__ yc,
                                  const double* __restrict__ alpha, const double* __restrict__ v,
                                  int64_t n, int64_t npad, double* out) {
  __shared__ double sh[3][32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    a += log(T[i * npad + i]);
    if (alpha != nullptr) b = fma(yc[i], alpha[i], b);
    c = fma(v[i], v[i], c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sh[2][warp] = c; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? sh[0][lane] : 0.0;
    b = lane < nw ? sh[1][lane] : 0.0;
    c = lane < nw ? sh[2][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { out[0] = a; out[1] = b; out[2] = c; }
  }
}

// ---- copy-outs for gp.L / generic strided copies ----------------------------------------------------
__global__ void extract_lower_kernel(const double* __restrict__ T, int64_t npad, double* L, int64_t n) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t i = idx / n, c = idx - i * n;
  L[idx] = (c <= i) ? T[i * npad + c] : 0.0;
}

__global__ void copy_pad_kernel(const double* __restrict__ src, int64_t n_src, double* dst, int64_t n_dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_dst) dst[i] = (i < n_src) ? src[i] : 0.0;
}

__global__ void copy_rows_kernel(const double* __restrict__ src, int64_t ld_src, double* dst,
                                 int64_t ld_dst, int64_t rows, int64_t cols) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols, c = idx - r * cols;
  dst[r * ld_dst + c] = src[r * ld_src + c];
}

// ================================================================================================
// Acquisition + arg-max                                       dragonfly/opt/gpb_acquisitions.py
// ================================================================================================
// scipy.stats.norm.cdf == scipy.special.ndtr (cephes): 0.5 erfc(-z / sqrt 2) split at |x| < sqrt(1/2).
__device__ __forceinline__ double norm_cdf_ref(double z) {
  if (isnan(z)) return z;
  const double x = z * 0.70710678118654752440;
  const double ax = fabs(x);
  double y;
  if (ax < 0.70710678118654752440) {
    y = 0.5 + 0.5 * erf(x);
  } else {
    y = 0.5 * erfc(ax);
    if (x > 0) y = 1.0 - y;
  }
  return y;
}
// scipy.stats.norm.pdf: exp(-x**2 / 2) / sqrt(2 pi)
__device__ __forceinline__ double norm_pdf_ref(double z) {
  return exp(__dmul_rn(__dmul_rn(z, z), -0.5)) / 2.50662827463100050242;
}
__device__ __forceinline__ double ei_for_norm_diff(double z) {     // gpb_acquisitions.py:247-249
  return __dadd_rn(__dmul_rn(z, norm_cdf_ref(z)), norm_pdf_ref(z));
}

// np.argmax order: NaN beats everything, ties go to the lower index (oper_utils.py:73).
__device__ __forceinline__ bool better(double sa, int64_t ia, double sb, int64_t ib) {
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  const bool na = isnan(sa), nb = isnan(sb);
  if (na || nb) {
    if (na && nb) return ia < ib;
    return na;
  }
  if (sa > sb) return true;
  if (sa < sb) return false;
  return ia < ib;
}

// Error model of the int8-slice scoring pass (api.cu: i8_sigma2_bound): |sigma^2_int8 - sigma^2_fp64| <= b2 for
// every candidate.  mu is computed in fp64 by both passes, so a score differs only through sigma:
//   |sd_64 - sd_i8| <= min(b2 / sd_i8, sqrt(b2))          (|sqrt a - sqrt b| <= |a - b| / (sqrt a + sqrt b) <= sqrt|a - b|)
//   UCB:  |d score| <= |beta| e;   EI, TTEI: d/d sigma = phi(z) [x sigma / sigma_c] <= 0.4, so <= 0.4 e;
//   PI:   d/d sigma = -z phi(z) / sigma, |z phi(z)| <= 0.242: <= 0.25 e / (sd_i8 - e), capped at 1.
// `sens` carries |beta| / 0.4 / 0.25.  Returns < 0 for a candidate whose fp64 variance may be negative (NaN score,
// which np.argmax treats as the maximum): sd_i8 <= sqrt(b2) or NaN -- such candidates are always re-scored.
__device__ __forceinline__ double i8_score_err(const I8ErrModel& em, double sd) {
  const double root = sqrt(em.b2);
  if (!(sd > root)) return -1.0;
  const double e = em.b2 / sd;
  if (em.kind == DFB_ACQ_PI) {
    const double lo = sd - e;
    return (lo > 0.0) ? fmin(1.0, em.sens * e / lo) : 1.0;
  }
  return em.sens * e;
}

__global__ void __launch_bounds__(256)
acq_kernel(const dfb_acq_desc acq, const double* __restrict__ mu, const double* __restrict__ partial,
           int64_t ld_partial, int nrb, const double* __restrict__ kss, int64_t m, int64_t idx_base,
           int want_std, double* __restrict__ sd_out, double* __restrict__ score_out, double* blk_score,
           int64_t* blk_index, const int64_t* __restrict__ idx_map, const I8ErrModel em, double* blk_lb,
           const int* __restrict__ abort_count, int abort_cap) {
  if (abort_count != nullptr && *abort_count > abort_cap) return;     // shortlist overflowed: this pass is void
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double score = 0.0;
  int64_t index = -1;
  double lb = -__longlong_as_double(0x7ff0000000000000ll);     // -inf
  if (i < m) {
    const double mean = mu[i];
    double sd = 0.0;
    if (want_std) {
      double vn = 0.0;
      for (int rb = 0; rb < nrb; rb++) vn += partial[(int64_t)rb * ld_partial + i];
      sd = sqrt(__dadd_rn(kss[i], -vn));    // np.sqrt(np.diag(K_tete - V.T.dot(V))): no clamp
      if (sd_out != nullptr) sd_out[i] = sd;
    }
    switch (acq.kind) {
      case DFB_ACQ_UCB:                       // mu + beta_th * sigma            :219-222
        score = __dadd_rn(mean, __dmul_rn(acq.beta, sd));
        break;
      case DFB_ACQ_EI: {                      // sigma * EI((mu - best) / sigma)  :255-260
        const double z = __dadd_rn(mean, -acq.best) / sd;
        score = __dmul_rn(sd, ei_for_norm_diff(z));
        break;
      }
      case DFB_ACQ_PI:                        // Phi((mu - best) / sigma)         :235-238
        score = norm_cdf_ref(__dadd_rn(mean, -acq.best) / sd);
        break;
      case DFB_ACQ_TTEI: {                    // :274-279
        const double comb = sqrt(__dadd_rn(__dmul_rn(acq.ref_std, acq.ref_std), __dmul_rn(sd, sd)));
        const double z = __dadd_rn(mean, -acq.ref_mean) / comb;
        score = __dmul_rn(comb, ei_for_norm_diff(z));
        break;
      }
      default:
        score = mean;
    }
    if (score_out != nullptr) score_out[i] = score;
    index = (idx_map != nullptr) ? idx_map[i] : idx_base + i;
    if (blk_lb != nullptr) {
      // a certain lower bound of this candidate's fp64 score (none for suspects / NaN)
      const double e = i8_score_err(em, sd);
      if (e >= 0.0 && !isnan(score)) lb = score - e;
    }
  }
  if (blk_score == nullptr) return;
  // block arg-max
  for (int o = 16; o > 0; o >>= 1) {
    const double so = __shfl_xor_sync(0xffffffffu, score, o);
    const int64_t io = __shfl_xor_sync(0xffffffffu, index, o);
    if (better(so, io, score, index)) { score = so; index = io; }
    lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
  }
  __shared__ double ss[8];
  __shared__ int64_t si[8];
  __shared__ double sl[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = score; si[warp] = index; sl[warp] = lb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      if (better(ss[w], si[w], score, index)) { score = ss[w]; index = si[w]; }
      lb = fmax(lb, sl[w]);
    }
    blk_score[blockIdx.x] = score;
    blk_index[blockIdx.x] = index;
    if (blk_lb != nullptr) blk_lb[blockIdx.x] = lb;
  }
}

// Folds the per-block winners of one chunk into the running (score, index) of the whole call.
// blk_lb / best_lb (optional): the running maximum of the candidates' certain lower bounds (int8 pass).
__global__ void __launch_bounds__(256)
argmax_merge_kernel(const double* __restrict__ blk_score, const int64_t* __restrict__ blk_index,
                    int nblk, double* best_score, int64_t* best_index, const double* __restrict__ blk_lb,
                    double* best_lb, const int* __restrict__ abort_count, int abort_cap) {
  if (abort_count != nullptr && *abort_count > abort_cap) return;
  double score = 0.0;
  int64_t index = -1;
  double lb = -__longlong_as_double(0x7ff0000000000000ll);
  if (threadIdx.x == 0) { score = *best_score; index = *best_index; if (best_lb != nullptr) lb = *best_lb; }
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    if (better(blk_score[b], blk_index[b], score, index)) { score = blk_score[b]; index = blk_index[b]; }
    if (blk_lb != nullptr) lb = fmax(lb, blk_lb[b]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double so = __shfl_xor_sync(0xffffffffu, score, o);
    const int64_t io = __shfl_xor_sync(0xffffffffu, index, o);
    if (better(so, io, score, index)) { score = so; index = io; }
    lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
  }
  __shared__ double ss[8];
  __shared__ int64_t si[8];
  __shared__ double sl[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = score; si[warp] = index; sl[warp] = lb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      if (better(ss[w], si[w], score, index)) { score = ss[w]; index = si[w]; }
      lb = fmax(lb, sl[w]);
    }
    *best_score = score;
    *best_index = index;
    if (best_lb != nullptr) *best_lb = lb;
  }
}

// ---- multi-objective scalarisations + arg-max        dragonfly/opt/multiobjective_gpb_acquisitions.py ----
// One thread per candidate combines the objectives' (mu, sd) -- or posterior samples -- in the reference's
// own operation order (no FMA contraction), then the same block arg-max as acq_kernel.
struct MooArgs {
  dfb_moo_desc d;
  const double* a[DFB_MOO_MAX_OBJ];    // mu_k, or the sampled values v_k
  const double* b[DFB_MOO_MAX_OBJ];    // sd_k (UCB kinds)
};
__device__ __forceinline__ double np_minimum(double x, double y) {   // np.minimum: NaN propagates
  if (isnan(x)) return x;
  if (isnan(y)) return y;
  return x < y ? x : y;
}
__global__ void __launch_bounds__(256)
moo_kernel(const MooArgs g, int64_t m, int64_t idx_base, double* __restrict__ score_out, double* blk_score,
           int64_t* blk_index) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double score = 0.0;
  int64_t index = -1;
  if (i < m) {
    const int K = g.d.n_obj;
    if (g.d.kind == DFB_MOO_LIN_UCB) {              // :79-91
      double mu_tot = 0.0, s2_tot = 0.0;
      for (int k = 0; k < K; k++) {
        const double w = g.d.weight[k], sd = g.b[k][i];
        mu_tot = __dadd_rn(mu_tot, __dmul_rn(g.a[k][i], w));
        s2_tot = __dadd_rn(s2_tot, __dmul_rn(__dmul_rn(sd, sd), __dmul_rn(w, w)));
      }
      score = __dadd_rn(mu_tot, __dmul_rn(g.d.beta, sqrt(s2_tot)));
    } else if (g.d.kind == DFB_MOO_TCH_UCB) {       // :94-107 (takes the square root of the std, as written there)
      score = __longlong_as_double(0x7ff0000000000000ll);
      for (int k = 0; k < K; k++) {
        const double ucb = __dadd_rn(__dadd_rn(g.a[k][i], __dmul_rn(g.d.beta, sqrt(g.b[k][i]))), -g.d.ref[k]);
        score = np_minimum(score, ucb / g.d.weight[k]);
      }
    } else if (g.d.kind == DFB_MOO_LIN_VAL) {       // :31-39
      for (int k = 0; k < K; k++) score = __dadd_rn(score, __dmul_rn(g.a[k][i], g.d.weight[k]));
    } else {                                        // DFB_MOO_TCH_VAL :56-65
      score = __longlong_as_double(0x7ff0000000000000ll);
      for (int k = 0; k < K; k++)
        score = np_minimum(score, __dadd_rn(g.a[k][i], -g.d.ref[k]) / g.d.weight[k]);
    }
    if (score_out != nullptr) score_out[i] = score;
    index = idx_base + i;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double so = __shfl_xor_sync(0xffffffffu, score, o);
    const int64_t io = __shfl_xor_sync(0xffffffffu, index, o);
    if (better(so, io, score, index)) { score = so; index = io; }
  }
  __shared__ double ss[8];
  __shared__ int64_t si[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = score; si[warp] = index; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      if (better(ss[w], si[w], score, index)) { score = ss[w]; index = si[w]; }
    blk_score[blockIdx.x] = score;
    blk_index[blockIdx.x] = index;
  }
}

// ---- counter-based normals for Thompson sampling at scale ----------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter = (column index lo, hi, draw index, stream), key = seed.  Element (s, a)
// of the S x m matrix depends only on (seed, col0 + a, s), so a candidate gets the same normals whatever the block,
// chunk or rank layout.  Box-Muller in fp64 on two 53-bit uniforms in (0, 1).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {      // (0, 1): (53-bit integer + 1/2) 2^-53
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)v + 0.5) * 0x1p-53;
}
__global__ void fill_rng_kernel(uint64_t seed, int64_t col0, int S, int64_t m, int what, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)S * m) return;
  const int64_t s = idx / m, a = idx - s * m;
  const uint64_t col = (uint64_t)(col0 + a);
  uint32_t r[4];
  philox4x32_10((uint32_t)col, (uint32_t)(col >> 32), (uint32_t)s, (uint32_t)what, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  const double u1 = u53(r[0], r[1]);
  if (what == DFB_RNG_UNIFORM) { out[idx] = u1; return; }
  const double u2 = u53(r[2], r[3]);
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  out[idx] = sqrt(-2.0 * log(u1)) * cs;
}

// Candidate generation on the device: random_sample + map_to_bounds (oper_utils.py:59-67, general_utils.py:25-27) with
// the same counter-based uniforms as fill_rng_kernel -- coordinate s of candidate row a is
// lo[s] + u(seed, row0 + a, s) * (hi[s] - lo[s]), u = the DFB_RNG_UNIFORM element (s, row0 + a): it depends on the GLOBAL
// row index only, so any rank can generate exactly its shard, and the winning row can be regenerated on its own.
// Output row-major m x d (what dfb_score_argmax takes).
struct CandBounds {
  double lo[DFB_MAX_SLOTS];
  double width[DFB_MAX_SLOTS];     // hi - lo, formed on the host like map_to_bounds does
};
__global__ void fill_candidates_kernel(uint64_t seed, int64_t row0, int64_t m, int d, const CandBounds b,
                                       double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * d) return;
  const int64_t a = idx / d;
  const int s = (int)(idx - a * d);
  const uint64_t col = (uint64_t)(row0 + a);
  uint32_t r[4];
  philox4x32_10((uint32_t)col, (uint32_t)(col >> 32), (uint32_t)s, (uint32_t)DFB_RNG_UNIFORM, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  out[idx] = __dadd_rn(__dmul_rn(u53(r[0], r[1]), b.width[s]), b.lo[s]);      // pts * (hi - lo) + lo
}

// running arg-max per draw: row s of samples (S x ld) over columns [0, m) -> (best[s], index[s]) in np.argmax order
__global__ void __launch_bounds__(256)
ts_argmax_kernel(const double* __restrict__ samples, int64_t ld, int64_t m, int64_t idx_base, int reset,
                 double* best, int64_t* index) {
  const int s = blockIdx.x;
  double score = 0.0;
  int64_t idx = -1;
  if (!reset && threadIdx.x == 0) { score = best[s]; idx = index[s]; }
  for (int64_t a = threadIdx.x; a < m; a += blockDim.x) {
    const double v = samples[(int64_t)s * ld + a];
    if (better(v, idx_base + a, score, idx)) { score = v; idx = idx_base + a; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double so = __shfl_xor_sync(0xffffffffu, score, o);
    const int64_t io = __shfl_xor_sync(0xffffffffu, idx, o);
    if (better(so, io, score, idx)) { score = so; idx = io; }
  }
  __shared__ double ss[8];
  __shared__ int64_t si[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = score; si[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      if (better(ss[w], si[w], score, idx)) { score = ss[w]; idx = si[w]; }
    best[s] = score;
    index[s] = idx;
  }
}

__global__ void reset_best_kernel(double* best_score, int64_t* best_index, double* best_lb) {
  *best_score = 0.0;
  *best_index = -1;
  *best_lb = -__longlong_as_double(0x7ff0000000000000ll);
}

// samples[s][a] += mu[a]   (draw_gaussian_samples: L.dot(U).T + mu, general_utils.py:231)
__global__ void add_row_vector_kernel(double* M, int64_t ld, int64_t rows, int64_t cols,
                                      const double* __restrict__ v) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols, c = idx - r * cols;
  M[r * ld + c] += v[c];
}

// covar[a][b] = kss_or_K**[a][b] handled by the caller; this zero-fills a padded matrix
__global__ void fill_kernel(double* p, int64_t n, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void set_diag_kernel(double* M, int64_t ld, int64_t from, int64_t to, double v, int add) {
  const int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) M[i * ld + i] = add ? (M[i * ld + i] + v) : v;
}

// Shortlist for the exact re-score that follows a fast (int8-slice) scoring pass.  With the error model above,
// candidate i's fp64 score lies in [s_i - E_i, s_i + E_i]; LB = max_j (s_j - E_j) over the candidates seen so far
// (best_lb: includes this chunk, folded by argmax_merge_kernel) is a certain lower bound of the final fp64 maximum.
// Kept: every candidate with s_i + E_i >= LB - pad (a superset of what the final LB would keep, since LB only
// grows) -- the fp64 arg-max and all its exact ties are among them --, every NaN, and every suspect (fp64 variance
// possibly negative).  Rows are gathered so that host-staged chunks can be re-scored later; the int8 score and its
// error allowance are kept for the self-check of the error model after the exact pass (selfcheck_kernel).
__global__ void collect_shortlist_kernel(const double* __restrict__ score, const double* __restrict__ sd,
                                         int64_t mc, int64_t idx_base, const double* best_lb,
                                         const I8ErrModel em, double pad,
                                         const double* __restrict__ Xc, int dc, int64_t* list_idx,
                                         double* list_X, double* list_s8, double* list_err, int* list_count, int cap) {
  if (*list_count > cap) return;                       // already overflowed: the whole pass is void
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mc) return;
  const double s = score[i];
  const double e = i8_score_err(em, sd[i]);
  const bool keep = isnan(s) || e < 0.0 || s + e >= *best_lb - pad;
  if (!keep) return;
  const int pos = atomicAdd(list_count, 1);
  if (pos >= cap) return;
  list_idx[pos] = idx_base + i;
  list_s8[pos] = s;
  list_err[pos] = isnan(s) ? -1.0 : e;
  for (int q = 0; q < dc; q++) list_X[(int64_t)pos * dc + q] = Xc[i * dc + q];
}

// After the exact fp64 re-score of the shortlist: every listed candidate with a defined allowance must satisfy
// |s_int8 - s_fp64| <= E_i -- a live check of the a-priori error model on exactly the candidates that matter.
// out[0] = number of violations, out[1] = max over the list of |s_int8 - s_fp64| / E_i scaled by 1e6 (diagnostic).
__global__ void selfcheck_kernel(const double* __restrict__ s8, const double* __restrict__ err,
                                 const double* __restrict__ s64, int count, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double e = err[i];
  if (e < 0.0) return;
  const double d = fabs(s8[i] - s64[i]);
  if (isnan(s64[i]) || d > e) atomicAdd(out, 1);
  if (e > 0.0) {
    const double r = fmin(d / e, 1000.0) * 1e6;
    atomicMax(out + 1, (int)r);
  }
}

__global__ void vec_max_kernel(const double* __restrict__ v, int64_t n, double* out) {
  __shared__ double sh[32];
  double m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = lane < (blockDim.x >> 5) ? sh[lane] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[0] = m;
  }
}

// max over the first n diagonal entries (stable_cholesky's np.diag(M).max(), general_utils.py:184)
__global__ void diag_max_kernel(const double* __restrict__ M, int64_t ld, int64_t n, double* out) {
  __shared__ double sh[32];
  double v = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v = fmax(v, M[i * ld + i]);
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? sh[lane] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) out[0] = v;
  }
}

// ---- integer digit planes for the tcgen05 path (gemm_i8.cuh) ------------------------------------------
// rowscale[i] = 2^E_i with |M[i,:]| * 2^-E_i < 1/2  (rowinv = 2^-E_i); one warp per row.
__global__ void row_exponent_kernel(const double* __restrict__ M, int64_t ld, int64_t rows, int64_t cols,
                                    double* rowscale, double* rowinv) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  double mx = 0.0;
  for (int64_t c = lane; c < cols; c += 32) mx = fmax(mx, fabs(M[row * ld + c]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    int e = 0;
    if (mx > 0.0 && isfinite(mx)) frexp(mx, &e);
    rowscale[row] = ldexp(1.0, e + 1);
    rowinv[row] = ldexp(1.0, -(e + 1));
  }
}

// Exact expansion of x = M * 2^-E (|x| < 1/2) into I8_S signed 7-bit digits: y = 128 x, a = rint(y),
// x <- y - a (all exact in fp64); four consecutive columns per thread, one 32-bit store per digit.
// Output layout (pair-interleaved planes, see gemm_i8.cuh): byte offset of (digit s, row, k) =
//   (s/2) * plane_bytes + row * 2*cols + (k/64) * 128 + (s%2) * 64 + (k%64).
__global__ void slice_i8_kernel(const double* __restrict__ M, int64_t ld, int64_t rows, int64_t cols4,
                                const double* __restrict__ rowinv, double inv_const,
                                uint32_t* __restrict__ out, int64_t plane_words, int kb, int radix256) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols4) return;
  const int64_t row = idx / cols4, c4 = idx - row * cols4;
  const double inv = (rowinv != nullptr) ? rowinv[row] : inv_const;
  const double4 in = *reinterpret_cast<const double4*>(M + row * ld + 4 * c4);
  double x[4] = {in.x * inv, in.y * inv, in.z * inv, in.w * inv};
  const int64_t k = 4 * c4;
  const int64_t row_words = 2 * cols4;                       // 2 * cols bytes
  // kb = k-values per interleave block (64: gemm_i8.cuh, 32: gemm_i8x2.cuh)
  const int64_t base = row * row_words + (k / kb) * (kb >> 1) + ((k % kb) >> 2);
  if (radix256) {
    int dg[4][5];
#pragma unroll
    for (int q = 0; q < 4; q++) digits_radix256(x[q], dg[q]);
#pragma unroll
    for (int s = 0; s < 5; s++)
      out[(int64_t)(s >> 1) * plane_words + base + (s & 1) * (kb >> 2)] = pack4_i8(dg[0][s], dg[1][s], dg[2][s], dg[3][s]);
    out[3 * plane_words + row * cols4 + c4] = pack4_i8(dg[0][0], dg[1][0], dg[2][0], dg[3][0]);   // compact leading digit
    return;
  }
#pragma unroll
  for (int s = 0; s < I8_S; s++) {
    uint32_t pack = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double y = x[q] * 128.0;
      double a = rint(y);
      x[q] = y - a;
      a = fmin(fmax(a, -127.0), 127.0);
      pack |= ((uint32_t)((int)a) & 0xffu) << (8 * q);
    }
    out[(int64_t)(s >> 1) * plane_words + base + (s & 1) * (kb >> 2)] = pack;
  }
}

// ================================================================================================
// Host launchers
// ================================================================================================
static bool g_attr_done[2] = {false, false};

int launch_gemm(dfb_handle* h, const GemmArgs& g, int epi, int n_blocks) {
  if (n_blocks <= 0) return 0;
  if (epi == EPI_SUMSQ) {
    if (!g_attr_done[1]) {
      DFB_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<EPI_SUMSQ>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GEMM_SMEM_BYTES));
      g_attr_done[1] = true;
    }
    gemm_tn_kernel<EPI_SUMSQ><<<n_blocks, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(g);
  } else {
    if (!g_attr_done[0]) {
      DFB_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<EPI_STORE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GEMM_SMEM_BYTES));
      g_attr_done[0] = true;
    }
    gemm_tn_kernel<EPI_STORE><<<n_blocks, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(g);
  }
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// D[r][c] = (C ? C[r][c] : 0) + part_0[r][c] + part_1[r][c] + ...  (fixed order: deterministic)
__global__ void splitk_reduce_kernel(const double* __restrict__ part, int ksplit, int64_t rows, int64_t cols,
                                     const double* __restrict__ C, int64_t ldc, double* __restrict__ D, int64_t ldd) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols, c = idx - r * cols;
  double s = (C != nullptr) ? C[r * ldc + c] : 0.0;
  for (int k = 0; k < ksplit; k++) s += part[(int64_t)k * rows * cols + idx];
  D[r * ldd + c] = s;
}

int launch_gemm_splitk(dfb_handle* h, GemmArgs g, int ksplit, double* part) {
  const int tiles = g.n_rb * g.n_cb;
  if (tiles <= 0) return 0;
  if (g.mode != MODE_GENERIC || ksplit < 2 || g.lower_only) { set_error("launch_gemm_splitk: bad arguments"); return -1; }
  const double* C = g.C; const int64_t ldc = g.ldc;
  double* D = g.D; const int64_t ldd = g.ldd;
  g.C = nullptr; g.ksplit = ksplit; g.part = part;
  { const int r = launch_gemm(h, g, EPI_STORE, tiles * ksplit); if (r != 0) return r; }
  const int64_t rows = (int64_t)g.n_rb * TILE, cols = (int64_t)g.n_cb * TILE;
  splitk_reduce_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, h->stream>>>(part, ksplit, rows, cols, C, ldc,
                                                                                  D, ldd);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- |L^-1 k_*|^2 for a handful of candidates (GP.eval of 1..16 points: the sequential maximisers' one-point
// objective, TTEI's reference arm, BOCA's fidelity scan) ---------------------------------------------------------
// The tile kernels spend a 128-wide candidate tile (and ~0.7 ms at N = 5000: 40 CTAs with k-depths up to N) on
// a single point.  Here one warp owns one row i of W = L^-1 at a time: v_i = sum_{k <= i} W[i][k] k_*[k] with the
// row streamed once, coalesced (the read of W's lower triangle -- 105 MB at N = 5000 -- is the whole cost), for MC
// candidates at once; squares accumulate per warp and acq_kernel adds the per-warp partials in index order
// (deterministic).
template <int MC>
__global__ void __launch_bounds__(256)
small_sumsq_kernel(const double* __restrict__ W, int64_t ldw, const double* __restrict__ Ks, int64_t ldk,
                   int64_t n_rows, int c0, double* __restrict__ part, int64_t ld_part) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t warps_total = (int64_t)gridDim.x * 8;
  double vs[MC];
#pragma unroll
  for (int c = 0; c < MC; c++) vs[c] = 0.0;
  for (int64_t i = warp_global; i < n_rows; i += warps_total) {
    const double* wr = W + i * ldw;
    double acc[MC];
#pragma unroll
    for (int c = 0; c < MC; c++) acc[c] = 0.0;
    int64_t k = lane;
    for (; k + 96 <= i; k += 128) {          // four independent 256-byte row segments in flight per warp
      const double w0 = wr[k], w1 = wr[k + 32], w2 = wr[k + 64], w3 = wr[k + 96];
#pragma unroll
      for (int c = 0; c < MC; c++) {
        const double* kr = Ks + (int64_t)(c0 + c) * ldk + k;
        const double k0 = kr[0], k1 = kr[32], k2 = kr[64], k3 = kr[96];
        acc[c] = fma(w3, k3, fma(w2, k2, fma(w1, k1, fma(w0, k0, acc[c]))));
      }
    }
    for (; k <= i; k += 32) {
      const double w = wr[k];
#pragma unroll
      for (int c = 0; c < MC; c++) acc[c] = fma(w, Ks[(int64_t)(c0 + c) * ldk + k], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < MC; c++) {
      double a = acc[c];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      vs[c] = fma(a, a, vs[c]);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < MC; c++) part[warp_global * ld_part + c0 + c] = vs[c];
  }
}

int launch_small_sumsq(dfb_handle* h, const double* W, int64_t ldw, const double* Ks, int64_t ldk, int64_t n_rows,
                       int m, double* part, int64_t ld_part, int* n_warps_out) {
  const unsigned blocks = (unsigned)((n_rows + 7) / 8);
  *n_warps_out = (int)blocks * 8;
  for (int c0 = 0; c0 < m; c0 += 8) {
    const int mc = (m - c0 < 8) ? (m - c0) : 8;
    if (mc == 8) small_sumsq_kernel<8><<<blocks, 256, 0, h->stream>>>(W, ldw, Ks, ldk, n_rows, c0, part, ld_part);
    else if (mc >= 5) {       // 5..7: an 8-wide pass over rows c0 .. c0+7 (rows beyond m are zero K_* rows)
      small_sumsq_kernel<8><<<blocks, 256, 0, h->stream>>>(W, ldw, Ks, ldk, n_rows, c0, part, ld_part);
    } else if (mc >= 3) small_sumsq_kernel<4><<<blocks, 256, 0, h->stream>>>(W, ldw, Ks, ldk, n_rows, c0, part, ld_part);
    else if (mc == 2) small_sumsq_kernel<2><<<blocks, 256, 0, h->stream>>>(W, ldw, Ks, ldk, n_rows, c0, part, ld_part);
    else small_sumsq_kernel<1><<<blocks, 256, 0, h->stream>>>(W, ldw, Ks, ldk, n_rows, c0, part, ld_part);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// ---- TMA tensor maps + the v2 scoring kernel ---------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int make_tensor_map_2d_f64(CUtensorMap* out, const double* base, int64_t rows, int64_t cols_ld,
                           int64_t cols) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DFB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return -2;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols_ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)TILE};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return -2;
  }
  return 0;
}

int make_tensor_map_3d_u8(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t planes,
                          int64_t row_ld_bytes, int64_t plane_stride_bytes, int box_cols, int box_rows,
                          int box_planes) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DFB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return -2;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  const cuuint64_t gstride[2] = {(cuuint64_t)row_ld_bytes, (cuuint64_t)plane_stride_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3d u8) failed with CUresult %d", (int)r);
    return -2;
  }
  return 0;
}

static bool g_i8_attr = false;
int launch_score_i8(dfb_handle* h, const CUtensorMap& tmA, const CUtensorMap& tmB, const ScoreI8Args& g) {
  ScoreI8Args ga = g;
  if (ga.cb_group > ga.n_cb) ga.cb_group = ga.n_cb;
  const int n_groups = (ga.n_cb + ga.cb_group - 1) / ga.cb_group;
  const int n_blocks = ga.n_rb * ga.cb_group * n_groups;
  if (n_blocks <= 0) return 0;
  if (!g_i8_attr) {
    DFB_CUDA_OK(cudaFuncSetAttribute(score_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)I8_SMEM_BYTES));
    DFB_CUDA_OK(cudaFuncSetAttribute(score_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)I8_SMEM_BYTES));
    g_i8_attr = true;
  }
  if (h->i8_ts)
    score_i8_kernel<true><<<n_blocks, I8_THREADS, I8_SMEM_BYTES, h->stream>>>(tmA, tmB, ga);
  else
    score_i8_kernel<false><<<n_blocks, I8_THREADS, I8_SMEM_BYTES, h->stream>>>(tmA, tmB, ga);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_score_i8_args(dfb_handle* h, const CUtensorMap& tmA, const CUtensorMap& tmB, int n_rb, int n_cb,
                         int K, double* partial, int64_t ld_partial, const double* rowscale, double colscale) {
  ScoreI8Args g;
  memset(&g, 0, sizeof(g));
  g.n_rb = n_rb; g.n_cb = n_cb; g.K = K; g.partial = partial; g.ld_partial = ld_partial;
  g.rowscale = rowscale; g.colscale = colscale;
  g.cb_group = h->i8_cb_group;
  g.timing = nullptr;
  return launch_score_i8(h, tmA, tmB, g);
}

static bool g_i8x2_attr = false;
int launch_score_i8x2_args(dfb_handle* h, const CUtensorMap& tmA2, const CUtensorMap& tmA3,
                           const CUtensorMap& tmB2, const CUtensorMap& tmB3, int n_rb, int n_cb, int K,
                           double* partial, int64_t ld_partial, const double* rowscale, double colscale) {
  ScoreI8Args g;
  memset(&g, 0, sizeof(g));
  g.n_rb = n_rb; g.n_cb = n_cb; g.K = K; g.partial = partial; g.ld_partial = ld_partial;
  g.rowscale = rowscale; g.colscale = colscale;
  g.cb_group = 0;                                   // unused: the persistent kernel deals tiles itself
  g.timing = nullptr;
  const int n_tiles = n_rb * n_cb;
  if (n_tiles <= 0) return 0;
  static int n_sm[64] = {0};                        // one CTA per SM (201 KB of shared memory each)
  if (n_sm[h->device & 63] == 0)
    DFB_CUDA_OK(cudaDeviceGetAttribute(&n_sm[h->device & 63], cudaDevAttrMultiProcessorCount, h->device));
  const int n_blocks = n_tiles < n_sm[h->device & 63] ? n_tiles : n_sm[h->device & 63];
  if (!g_i8x2_attr) {
    DFB_CUDA_OK(cudaFuncSetAttribute(score_i8x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)X2_SMEM_BYTES));
    g_i8x2_attr = true;
  }
  score_i8x2_kernel<<<n_blocks, X2_THREADS, X2_SMEM_BYTES, h->stream>>>(tmA2, tmA3, tmB2, tmB3, g);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

static bool g_i8c2_attr = false;
// Candidate tiles per group for the pair kernel's tile order (gemm_i8c2.cuh: c2_tile).  The static serpentine deal is only
// as balanced as the list order lets it be, so the group width is chosen by simulating the deal: among the widths whose
// K_* digit group stays L2-resident (<= 64 MB), the one with the least HBM traffic estimate whose heaviest cluster is
// within 0.5 % of the best balance found (no grouping included).
static int choose_cb_group(int n_rb, int n_cb, int K, int P) {
  const int n_rp = (n_rb + 1) / 2;
  const long long group_bytes_per_cb = 7ll * TILE * K;
  double best_bal = 1e30;
  int cand[64], n_cand = 0;
  double bal_of[64];
  for (int G = 4; G <= n_cb; G++) {
    if (G < n_cb && group_bytes_per_cb * G > (64ll << 20)) continue;
    if (n_cand == 64) break;
    double load[256];
    for (int c = 0; c < P && c < 256; c++) load[c] = 0.0;
    int t = 0;
    for (int cb0 = 0; cb0 < n_cb; cb0 += G) {
      const int width = (n_cb - cb0 < G) ? n_cb - cb0 : G;
      for (int u = 0; u < n_rp * width; u++, t++) {
        const int rp = n_rp - 1 - u / width;
        int nk = (2 * rp + 2) * TILE;
        if (nk > K) nk = K;
        nk /= 32;
        const int j = t / P, cc = t % P;
        load[(j & 1) ? P - 1 - cc : cc] += nk + (nk + 3) / 4;
      }
    }
    double mx = 0.0, sum = 0.0;
    for (int c = 0; c < P && c < 256; c++) { sum += load[c]; if (load[c] > mx) mx = load[c]; }
    const double bal = mx / (sum / P);
    cand[n_cand] = G; bal_of[n_cand] = bal; n_cand++;
    if (bal < best_bal) best_bal = bal;
  }
  int best = n_cb;
  for (int i = 0; i < n_cand; i++)          // widths ascending: the first acceptable width beyond which traffic only grows
    if (bal_of[i] <= best_bal * 1.005) {    // least traffic = widest acceptable group below the residency limit
      if (cand[i] < n_cb) best = cand[i];
    }
  return best;
}

int launch_score_i8c2_args(dfb_handle* h, const CUtensorMap& tmA1, const CUtensorMap& tmA3, const CUtensorMap& tmA1c,
                           const CUtensorMap& tmB1h, const CUtensorMap& tmB3h, const CUtensorMap& tmB1c, int n_rb, int n_cb, int K,
                           double* partial, int64_t ld_partial, const double* rowscale, double colscale,
                           const int* abort_count) {
  ScoreI8Args g;
  memset(&g, 0, sizeof(g));
  g.n_rb = n_rb; g.n_cb = n_cb; g.K = K; g.partial = partial; g.ld_partial = ld_partial;
  g.rowscale = rowscale; g.colscale = colscale;
  g.abort_count = abort_count; g.abort_cap = SHORTLIST_CAP;
  const int n_tiles = ((n_rb + 1) / 2) * n_cb;      // row-block pairs x candidate tiles
  if (n_tiles <= 0) return 0;
  static int n_sm[64] = {0};
  if (n_sm[h->device & 63] == 0)
    DFB_CUDA_OK(cudaDeviceGetAttribute(&n_sm[h->device & 63], cudaDevAttrMultiProcessorCount, h->device));
  const int max_clusters = n_sm[h->device & 63] / 2;                 // one CTA per SM, two SMs per cluster
  const int n_clusters = n_tiles < max_clusters ? n_tiles : max_clusters;
  {
    // group width of the tile order: option i8_c2_group > 0 forces it, 0 = chosen by simulation (cached per geometry)
    static int memo_key[4] = {0, 0, 0, 0}, memo_val = 0;
    if (h->i8_c2_group > 0) g.cb_group = h->i8_c2_group;
    else {
      if (memo_key[0] != n_rb || memo_key[1] != n_cb || memo_key[2] != K || memo_key[3] != n_clusters) {
        memo_val = choose_cb_group(n_rb, n_cb, K, n_clusters);
        memo_key[0] = n_rb; memo_key[1] = n_cb; memo_key[2] = K; memo_key[3] = n_clusters;
      }
      g.cb_group = memo_val;
    }
    h->last_c2_group = g.cb_group;
  }
  // L2 priorities (option i8_l2_hint): 0 = normal / normal, 1 = K_* digits evict-last (they are re-read by every row-block
  // pair of the group), 2 = additionally W digits evict-first (streamed once per group)
  g.l2_a = (h->i8_l2_hint == 2) ? C2_L2_EVICT_FIRST : C2_L2_NORMAL;
  g.l2_b = (h->i8_l2_hint >= 1) ? C2_L2_EVICT_LAST : C2_L2_NORMAL;
  if (!g_i8c2_attr) {
    DFB_CUDA_OK(cudaFuncSetAttribute(score_i8c2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)C2_SMEM_BYTES));
    DFB_CUDA_OK(cudaFuncSetAttribute(score_i8c2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)C2_SMEM_BYTES));
    g_i8c2_attr = true;
  }
  // DFB200_I8_TIMING=1: where the MMA-issuing thread waits (diagnostics; synchronises after every launch)
  static const bool timing = getenv("DFB200_I8_TIMING") != nullptr;
  static unsigned long long* tbuf = nullptr;
  g.timing = nullptr;
  if (timing) {
    if (tbuf == nullptr) DFB_CUDA_OK(cudaMalloc(&tbuf, sizeof(unsigned long long) * 4 * 128));
    g.timing = tbuf;
  }
  if (h->i8_radix256)
    score_i8c2_kernel<true><<<2 * n_clusters, C2_THREADS, C2_SMEM_BYTES, h->stream>>>(tmA1, tmA3, tmA1c, tmB1h, tmB3h, tmB1c, g);
  else
    score_i8c2_kernel<false><<<2 * n_clusters, C2_THREADS, C2_SMEM_BYTES, h->stream>>>(tmA1, tmA3, tmA1c, tmB1h, tmB3h, tmB1c, g);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  if (timing) {
    static int printed = 0;
    unsigned long long host[4 * 128];
    DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
    DFB_CUDA_OK(cudaMemcpy(host, tbuf, sizeof(unsigned long long) * 4 * n_clusters, cudaMemcpyDeviceToHost));
    double a[4] = {0, 0, 0, 0};
    for (int c = 0; c < n_clusters; c++) for (int q = 0; q < 4; q++) a[q] += (double)host[4 * c + q] / n_clusters;
    if (printed++ < 6)
      fprintf(stderr, "[i8c2 timing] clusters %d  MMA thread clocks: total %.0f  wait full %.0f (%.1f%%)  drain %.0f (%.1f%%)  "
              "epilogue %.0f (%.1f%%)\n", n_clusters, a[3], a[0], 100 * a[0] / a[3], a[1], 100 * a[1] / a[3], a[2],
              100 * a[2] / a[3]);
  }
  return 0;
}

int launch_row_exponent(dfb_handle* h, const double* M, int64_t ld, int64_t rows, int64_t cols,
                        double* rowscale, double* rowinv) {
  row_exponent_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, h->stream>>>(M, ld, rows, cols, rowscale, rowinv);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_slice_i8(dfb_handle* h, const double* M, int64_t ld, int64_t rows, int64_t cols,
                    const double* rowinv, double inv_const, void* out, int64_t plane_bytes, int64_t out_ld_bytes) {
  (void)out_ld_bytes;
  const int64_t total = rows * (cols / 4);
  if (total <= 0) return 0;
  slice_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(
      M, ld, rows, cols / 4, rowinv, inv_const, reinterpret_cast<uint32_t*>(out), plane_bytes / 4,
      h->i8_impl >= 1 ? 32 : 64, h->i8_radix256);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

static bool g_tma_attr = false;
int launch_score_tma(dfb_handle* h, const CUtensorMap& tmW, const CUtensorMap& tmK,
                     const ScoreTmaArgs& g) {
  ScoreTmaArgs ga = g;
  if (ga.cb_group > ga.n_cb) ga.cb_group = ga.n_cb;
  const int n_groups = (ga.n_cb + ga.cb_group - 1) / ga.cb_group;
  const int n_blocks = ga.n_rb * ga.cb_group * n_groups;
  if (n_blocks <= 0) return 0;
  if (!g_tma_attr) {
    DFB_CUDA_OK(cudaFuncSetAttribute(score_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TMA_SMEM_BYTES));
    g_tma_attr = true;
  }
  score_tma_kernel<<<n_blocks, TMA_THREADS, TMA_SMEM_BYTES, h->stream>>>(tmW, tmK, ga);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_prep_scaled(dfb_handle* h, const dfb_kernel_desc* d_desc, int use_train_coords,
                       const double* X, int64_t n, int d, double* xs, double* nrm, int64_t npad) {
  const int threads = 128;
  prep_scaled_kernel<<<(unsigned)((npad + threads - 1) / threads), threads, 0, h->stream>>>(
      d_desc, use_train_coords, X, n, d, xs, nrm, npad);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int KIND, int P>
static bool launch_kstar_fast_d(dfb_handle* h, int d, unsigned blocks, const dfb_kernel_desc* d_desc,
                                int ctc, const double* xsT, const double* nrmT, int64_t npad_tr,
                                const double* alpha, const double* Xc, int64_t m, int dc, int64_t m_rows,
                                double* Ks, int64_t ldk, int64_t n_valid, int64_t n_write,
                                double mean_const, double* mu, double* kss_out, const KstarI8Out* i8o) {
  KstarI8Out none;
  memset(&none, 0, sizeof(none));
#define DFB_KF_CASE(DD)                                                                              \
  case DD:                                                                                           \
    if (i8o != nullptr)                                                                              \
      kstar_fast_kernel<KIND, P, DD, true><<<blocks, KF_WARPS * 32, 0, h->stream>>>(                 \
          d_desc, ctc, xsT, nrmT, npad_tr, alpha, Xc, m, dc, m_rows, Ks, ldk, n_valid, n_write,      \
          mean_const, mu, kss_out, *i8o);                                                            \
    else                                                                                             \
      kstar_fast_kernel<KIND, P, DD, false><<<blocks, KF_WARPS * 32, 0, h->stream>>>(                \
          d_desc, ctc, xsT, nrmT, npad_tr, alpha, Xc, m, dc, m_rows, Ks, ldk, n_valid, n_write,      \
          mean_const, mu, kss_out, none);                                                            \
    return true;
  switch (d) {
    DFB_KF_CASE(1) DFB_KF_CASE(2) DFB_KF_CASE(3) DFB_KF_CASE(4)
    DFB_KF_CASE(5) DFB_KF_CASE(6) DFB_KF_CASE(7) DFB_KF_CASE(8)
    default: return false;
  }
#undef DFB_KF_CASE
}

// Returns 1 in *emitted_i8 when the digit planes were written by the K_* kernel itself (fused path).
int launch_kstar_i8(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc,
                    const double* xsT, const double* nrmT, int64_t npad_tr, const double* alpha,
                    const double* Xc, int64_t m, int dc, int64_t m_rows, int64_t n_valid, int64_t n_write,
                    double mean_const, double* mu, double* kss_out, void* planes, int64_t plane_bytes,
                    int64_t row_bytes, double inv_colscale, int* emitted_i8, const int* abort_count) {
  *emitted_i8 = 0;
  if (m_rows <= 0) return 0;
  if (!(h->kstar_fast && desc.n_terms == 1 && desc.n_factors == 1 && desc.factors[0].n_dims <= 8 &&
        desc.factors[0].slot_off == 0 && (desc.factors[0].kind == DFB_BASE_SE || desc.factors[0].p <= 2) &&
        n_write % 4 == 0 && npad_tr % 4 == 0))
    return 0;
  KstarI8Out o;
  o.planes = reinterpret_cast<uint8_t*>(planes); o.plane_bytes = plane_bytes; o.row_bytes = row_bytes;
  o.inv_colscale = inv_colscale;
  o.kb = (h->i8_impl >= 1) ? 32 : 64;
  o.radix256 = h->i8_radix256;
  o.abort_count = abort_count; o.abort_cap = SHORTLIST_CAP;
  const unsigned fblocks = (unsigned)((m_rows + KF_CANDS - 1) / KF_CANDS);
  const int d = desc.factors[0].n_dims;
  bool ok = false;
#define DFB_KF_ARGS h, d, fblocks, d_desc, 0, xsT, nrmT, npad_tr, alpha, Xc, m, dc, m_rows, nullptr, 0,  \
                    n_valid, n_write, mean_const, mu, kss_out, &o
  if (desc.factors[0].kind == DFB_BASE_SE) ok = launch_kstar_fast_d<DFB_BASE_SE, 0>(DFB_KF_ARGS);
  else if (desc.factors[0].p == 0) ok = launch_kstar_fast_d<DFB_BASE_MATERN, 0>(DFB_KF_ARGS);
  else if (desc.factors[0].p == 1) ok = launch_kstar_fast_d<DFB_BASE_MATERN, 1>(DFB_KF_ARGS);
  else ok = launch_kstar_fast_d<DFB_BASE_MATERN, 2>(DFB_KF_ARGS);
#undef DFB_KF_ARGS
  if (ok) {
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
    *emitted_i8 = 1;
  }
  return 0;
}

// Second-generation K_* digit path (kstar_seg_kernel): cand_prep -> segments -> mu.  Returns 1 in *emitted when it ran.
template <int KIND, int P, bool ROWS64>
static bool launch_kseg_d(dfb_handle* h, int d, const dfb_kernel_desc* d_desc, const double* Xc, int64_t m, int dc,
                          int64_t m_rows, double* cprep, double* kss_out, const KsegArgs& a, int n_seg) {
  const dim3 grid((unsigned)((n_seg + KS_WARPS - 1) / KS_WARPS), (unsigned)((m_rows + KS_ROWS - 1) / KS_ROWS));
  const unsigned pblocks = (unsigned)((m_rows + 127) / 128);
  // All kernels of the K stage ask for the maximum shared-memory carve-out -- the configuration the persistent tcgen05
  // kernel puts the SMs in -- so that their CTAs can be placed beside it instead of waiting for an SM to be re-configured.
#define DFB_KS_CASE(DD)                                                                                              \
  case DD: {                                                                                                         \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      cudaFuncSetAttribute(cand_prep_kernel<KIND, P, DD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);      \
      cudaFuncSetAttribute(kstar_seg_kernel<KIND, P, DD, ROWS64>, cudaFuncAttributePreferredSharedMemoryCarveout, ROWS64 ? -1 : 100); \
      cudaFuncSetAttribute(mu_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);                   \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    cand_prep_kernel<KIND, P, DD><<<pblocks, 128, 0, h->stream>>>(d_desc, Xc, m, dc, m_rows, cprep, kss_out);        \
    kstar_seg_kernel<KIND, P, DD, ROWS64><<<grid, KS_WARPS * 32, 0, h->stream>>>(a);                                 \
    return true;                                                                                                     \
  }
  switch (d) {
    DFB_KS_CASE(1) DFB_KS_CASE(2) DFB_KS_CASE(3) DFB_KS_CASE(4)
    DFB_KS_CASE(5) DFB_KS_CASE(6) DFB_KS_CASE(7) DFB_KS_CASE(8)
    default: return false;
  }
#undef DFB_KS_CASE
}

int launch_kstar_seg(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc, const double* xsT,
                     const double* nrm, int64_t npad_tr, const double* alpha, int64_t n_valid, const double* Xc, int64_t m, int dc,
                     int64_t m_rows, int64_t n_write, double mean_const, double* mu, double* kss_out, void* planes,
                     int64_t plane_bytes, int64_t row_bytes, double inv_colscale, double* cprep, double* mu_part,
                     int64_t ld_mu, int* emitted, const int* abort_count) {
  *emitted = 0;
  if (m_rows <= 0) return 0;
  const dfb_factor_desc& f = desc.factors[0];
  if (!(h->kstar_fast && h->kstar_seg && h->i8_impl == 2 && h->i8_radix256 && desc.n_terms == 1 && desc.n_factors == 1 &&
        f.n_dims <= 8 && f.slot_off == 0 && (f.kind == DFB_BASE_SE || f.p <= 2) && n_write % 128 == 0 && npad_tr % 2 == 0 &&
        m_rows % 2 == 0))
    return 0;
  KsegArgs a;
  memset(&a, 0, sizeof(a));
  a.xsT = xsT; a.nrm = nrm; a.alpha = alpha; a.npad_tr = npad_tr; a.n_valid = n_valid; a.cprep = cprep; a.m_rows = m_rows;
  a.n_write = n_write; a.planes = reinterpret_cast<uint8_t*>(planes); a.plane_bytes = plane_bytes; a.row_bytes = row_bytes;
  a.cdig = inv_colscale * 0x1p39;
  a.cval = desc.post_scale * desc.term_pre_scale[0] * f.scale * (f.kind == DFB_BASE_MATERN ? f.gamma_ratio : 1.0);
  a.s8 = f.s8; a.ms2 = -f.s2; a.c0 = f.coeffs[0]; a.c1 = f.coeffs[1]; a.c2 = f.coeffs[2];
  a.mu_part = mu_part; a.ld_mu = ld_mu; a.abort_count = abort_count; a.abort_cap = SHORTLIST_CAP;
  const int n_seg = (int)(n_write / KS_BLK);             // 64-point blocks, one per warp
  bool ok = false;
#define DFB_KS_ARGS h, f.n_dims, d_desc, Xc, m, dc, m_rows, cprep, kss_out, a, n_seg
  if (f.kind == DFB_BASE_SE) ok = launch_kseg_d<DFB_BASE_SE, 0, false>(DFB_KS_ARGS);
  else if (f.p == 0) ok = launch_kseg_d<DFB_BASE_MATERN, 0, false>(DFB_KS_ARGS);
  else if (f.p == 1) ok = launch_kseg_d<DFB_BASE_MATERN, 1, false>(DFB_KS_ARGS);
  else ok = launch_kseg_d<DFB_BASE_MATERN, 2, false>(DFB_KS_ARGS);
#undef DFB_KS_ARGS
  if (!ok) return 0;
  h->launches += 2;
  DFB_CUDA_OK(cudaGetLastError());
  if (mu != nullptr) {
    mu_reduce_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(mu_part, n_seg, ld_mu, m, mean_const, mu);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
  }
  *emitted = 1;
  return 0;
}

// fp64-row form of the segment kernel: K_* rows of m candidates (rows m .. m_rows-1 zero) + mu + k(x*,x*).  Served when the
// candidate side uses candidate coordinates and the shapes are the padded ones of the scoring paths; returns 0 in *done
// otherwise (the caller falls back to kstar_fast_kernel / kstar_kernel).
int launch_kstar_rows64(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc, const double* xsT,
                        const double* nrm, int64_t npad_tr, const double* alpha, const double* Xc, int64_t m, int dc,
                        int64_t m_rows, double* Ks, int64_t ldk, int64_t n_valid, int64_t n_write, double mean_const,
                        double* mu, double* kss_out, int* done) {
  *done = 0;
  const dfb_factor_desc& f = desc.factors[0];
  if (!(h->kstar_fast && h->kstar_seg && desc.n_terms == 1 && desc.n_factors == 1 && f.n_dims <= 8 && f.slot_off == 0 &&
        (f.kind == DFB_BASE_SE || f.p <= 2) && n_write % KS_BLK == 0 && npad_tr % 2 == 0 && m_rows % 2 == 0 && ldk % 2 == 0 &&
        m_rows <= h->chunk && (reinterpret_cast<uintptr_t>(Ks) & 15) == 0 && h->cprep != nullptr &&
        (alpha == nullptr || (reinterpret_cast<uintptr_t>(alpha) & 15) == 0)))
    return 0;
  const int n_seg = (int)(n_write / KS_BLK);
  const bool want_mu = (mu != nullptr) && n_seg <= (int)(h->npad_max / KS_BLK) + 2;
  if (mu != nullptr && !want_mu) return 0;
  KsegArgs a;
  memset(&a, 0, sizeof(a));
  a.xsT = xsT; a.nrm = nrm; a.alpha = alpha; a.npad_tr = npad_tr; a.n_valid = n_valid; a.cprep = h->cprep; a.m_rows = m_rows;
  a.n_write = n_write;
  a.cval = desc.post_scale * desc.term_pre_scale[0] * f.scale * (f.kind == DFB_BASE_MATERN ? f.gamma_ratio : 1.0);
  a.s8 = f.s8; a.ms2 = -f.s2; a.c0 = f.coeffs[0]; a.c1 = f.coeffs[1]; a.c2 = f.coeffs[2];
  a.mu_part = want_mu ? h->mu_part : nullptr; a.ld_mu = h->chunk;
  a.rows64 = Ks; a.ld64 = ldk;
  bool ok = false;
#define DFB_KS_ARGS h, f.n_dims, d_desc, Xc, m, dc, m_rows, h->cprep, kss_out, a, n_seg
  if (f.kind == DFB_BASE_SE) ok = launch_kseg_d<DFB_BASE_SE, 0, true>(DFB_KS_ARGS);
  else if (f.p == 0) ok = launch_kseg_d<DFB_BASE_MATERN, 0, true>(DFB_KS_ARGS);
  else if (f.p == 1) ok = launch_kseg_d<DFB_BASE_MATERN, 1, true>(DFB_KS_ARGS);
  else ok = launch_kseg_d<DFB_BASE_MATERN, 2, true>(DFB_KS_ARGS);
#undef DFB_KS_ARGS
  if (!ok) return 0;
  h->launches += 2;
  DFB_CUDA_OK(cudaGetLastError());
  if (want_mu) {
    mu_reduce_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(h->mu_part, n_seg, h->chunk, m, mean_const, mu);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
  }
  *done = 1;
  return 0;
}

int launch_kstar(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc,
                 int cand_uses_train_coords, const double* xsT, const double* nrmT, int64_t npad_tr,
                 const double* alpha, const double* Xc, int64_t m, int dc, int64_t m_rows, double* Ks,
                 int64_t ldk, int64_t n_valid, int64_t n_write, double mean_const, double* mu,
                 double* kss_out) {
  if (m_rows <= 0) return 0;
  if (!cand_uses_train_coords && h->kstar_rows64) {
    int done = 0;
    DFB_TRY_RET(launch_kstar_rows64(h, d_desc, desc, xsT, nrmT, npad_tr, alpha, Xc, m, dc, m_rows, Ks, ldk, n_valid, n_write,
                                    mean_const, mu, kss_out, &done));
    if (done) return 0;
  }
  // fast path: plain SE / Matern(p <= 2) on <= 8 coordinates
  if (h->kstar_fast && desc.n_terms == 1 && desc.n_factors == 1 && desc.factors[0].n_dims <= 8 &&
      desc.factors[0].slot_off == 0 && (desc.factors[0].kind == DFB_BASE_SE || desc.factors[0].p <= 2) &&
      n_write % 4 == 0 && npad_tr % 4 == 0 && ldk % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(Ks) & 15) == 0 && (alpha == nullptr || (reinterpret_cast<uintptr_t>(alpha) & 15) == 0)) {
    const unsigned fblocks = (unsigned)((m_rows + KF_CANDS - 1) / KF_CANDS);
    const int d = desc.factors[0].n_dims;
    bool ok = false;
#define DFB_KF_ARGS h, d, fblocks, d_desc, cand_uses_train_coords, xsT, nrmT, npad_tr, alpha, Xc, m, dc, \
                    m_rows, Ks, ldk, n_valid, n_write, mean_const, mu, kss_out, nullptr
    if (desc.factors[0].kind == DFB_BASE_SE) ok = launch_kstar_fast_d<DFB_BASE_SE, 0>(DFB_KF_ARGS);
    else if (desc.factors[0].p == 0) ok = launch_kstar_fast_d<DFB_BASE_MATERN, 0>(DFB_KF_ARGS);
    else if (desc.factors[0].p == 1) ok = launch_kstar_fast_d<DFB_BASE_MATERN, 1>(DFB_KF_ARGS);
    else ok = launch_kstar_fast_d<DFB_BASE_MATERN, 2>(DFB_KF_ARGS);
#undef DFB_KF_ARGS
    if (ok) {
      h->launches++;
      DFB_CUDA_OK(cudaGetLastError());
      return 0;
    }
  }
  const size_t smem = ((sizeof(dfb_kernel_desc) + 15) / 16) * 16 +
                      sizeof(double) * KSTAR_CANDS * (size_t)(desc.n_slots + desc.n_factors);
  const unsigned blocks = (unsigned)((m_rows + KSTAR_CANDS - 1) / KSTAR_CANDS);
  kstar_kernel<<<blocks, KSTAR_WARPS * 32, smem, h->stream>>>(
      d_desc, cand_uses_train_coords, xsT, nrmT, npad_tr, alpha, Xc, m, dc, m_rows, Ks, ldk, n_valid,
      n_write, mean_const, mu, kss_out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_init_tall(dfb_handle* h, double* T, int64_t n, int64_t npad, double diag_add,
                     const double* yc, int with_bottom) {
  init_tall_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, h->stream>>>(T, n, npad, diag_add, yc,
                                                                         with_bottom);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

static bool g_diag_attr = false;
int launch_chol_diag(dfb_handle* h, double* T, int64_t ld, int step, double* Dinv, int* info) {
  (void)g_diag_attr;
  chol_diag_kernel<<<1, 256, 0, h->stream>>>(T, ld, step, Dinv, info);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_transpose(dfb_handle* h, const double* src, double* dst, int64_t n) {
  dim3 grid((unsigned)(n / 32), (unsigned)(n / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, h->stream>>>(src, dst, n);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_alpha(dfb_handle* h, const double* Wt, const double* v, double* alpha, int64_t n,
                 int64_t npad) {
  alpha_kernel<<<(unsigned)((npad + 7) / 8), 256, 0, h->stream>>>(Wt, v, alpha, n, npad);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_lml_reduce(dfb_handle* h, const double* T, const double* yc, const double* alpha,
                      const double* v, int64_t n, int64_t npad, double* out) {
  lml_reduce_kernel<<<1, 1024, 0, h->stream>>>(T, yc, alpha, v, n, npad, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_extract_lower(dfb_handle* h, const double* T, int64_t npad, double* L, int64_t n) {
  extract_lower_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, h->stream>>>(T, npad, L, n);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_copy_pad(dfb_handle* h, const double* src, int64_t n_src, double* dst, int64_t n_dst) {
  copy_pad_kernel<<<(unsigned)((n_dst + 255) / 256), 256, 0, h->stream>>>(src, n_src, dst, n_dst);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_copy_rows(dfb_handle* h, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                     int64_t rows, int64_t cols) {
  if (rows * cols <= 0) return 0;
  copy_rows_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, h->stream>>>(src, ld_src, dst,
                                                                               ld_dst, rows, cols);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_acq(dfb_handle* h, const dfb_acq_desc& acq, const double* mu, const double* partial,
               int64_t ld_partial, int nrb, const double* kss, int64_t m, int64_t idx_base, int want_std,
               double* sd_out, double* score_out, bool do_argmax, const int64_t* idx_map,
               const I8ErrModel* em) {
  if (m <= 0) return 0;
  const unsigned blocks = (unsigned)((m + 255) / 256);
  I8ErrModel none;
  none.b2 = 0.0; none.sens = 0.0; none.kind = 0;
  const bool track = do_argmax && em != nullptr && em->b2 > 0.0;
  const int* abort_count = track ? h->list_count : nullptr;     // int8 pass: void once the shortlist overflowed
  acq_kernel<<<blocks, 256, 0, h->stream>>>(acq, mu, partial, ld_partial, nrb, kss, m, idx_base,
                                            want_std, sd_out, score_out,
                                            do_argmax ? h->blk_score : nullptr,
                                            do_argmax ? h->blk_index : nullptr, idx_map,
                                            track ? *em : none, track ? h->blk_lb : nullptr, abort_count,
                                            SHORTLIST_CAP);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  if (do_argmax) {
    argmax_merge_kernel<<<1, 256, 0, h->stream>>>(h->blk_score, h->blk_index, (int)blocks,
                                                  h->best_score, h->best_index, track ? h->blk_lb : nullptr,
                                                  track ? h->best_lb : nullptr, abort_count, SHORTLIST_CAP);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// scores m candidates in slices of the handle's chunk (the block arg-max scratch is sized for one chunk)
int launch_moo(dfb_handle* h, const dfb_moo_desc& d, const double* const* a, const double* const* b, int64_t m,
               double* scores) {
  for (int64_t c0 = 0; c0 < m; c0 += h->chunk) {
    const int64_t mc = (m - c0 < h->chunk) ? (m - c0) : h->chunk;
    MooArgs g;
    memset(&g, 0, sizeof(g));
    g.d = d;
    for (int k = 0; k < d.n_obj; k++) {
      g.a[k] = a[k] + c0;
      g.b[k] = (b != nullptr && b[k] != nullptr) ? b[k] + c0 : nullptr;
    }
    const unsigned blocks = (unsigned)((mc + 255) / 256);
    moo_kernel<<<blocks, 256, 0, h->stream>>>(g, mc, c0, scores ? scores + c0 : nullptr, h->blk_score, h->blk_index);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
    argmax_merge_kernel<<<1, 256, 0, h->stream>>>(h->blk_score, h->blk_index, (int)blocks, h->best_score,
                                                  h->best_index, nullptr, nullptr, nullptr, 0);
    h->launches++;
    DFB_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int launch_fill_rng(dfb_handle* h, uint64_t seed, int64_t col0, int S, int64_t m, int what, double* out) {
  const int64_t total = (int64_t)S * m;
  if (total <= 0) return 0;
  fill_rng_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(seed, col0, S, m, what, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fill_candidates(dfb_handle* h, uint64_t seed, int64_t row0, int64_t m, int d, const double* lo,
                           const double* hi, double* out) {
  if (m * d <= 0) return 0;
  CandBounds b;
  memset(&b, 0, sizeof(b));
  for (int s = 0; s < d; s++) { b.lo[s] = lo[s]; b.width[s] = hi[s] - lo[s]; }
  fill_candidates_kernel<<<(unsigned)((m * d + 255) / 256), 256, 0, h->stream>>>(seed, row0, m, d, b, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_ts_argmax(dfb_handle* h, const double* samples, int64_t ld, int S, int64_t m, int64_t idx_base, int reset,
                     double* best, int64_t* index) {
  if (S <= 0) return 0;
  ts_argmax_kernel<<<(unsigned)S, 256, 0, h->stream>>>(samples, ld, m, idx_base, reset, best, index);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_reset_best(dfb_handle* h) {
  reset_best_kernel<<<1, 1, 0, h->stream>>>(h->best_score, h->best_index, h->best_lb);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_add_row_vector(dfb_handle* h, double* M, int64_t ld, int64_t rows, int64_t cols,
                          const double* v) {
  if (rows * cols <= 0) return 0;
  add_row_vector_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, h->stream>>>(M, ld, rows,
                                                                                    cols, v);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_collect_shortlist(dfb_handle* h, const double* score, const double* sd, int64_t mc,
                             int64_t idx_base, const I8ErrModel& em, double pad, const double* Xc, int dc) {
  if (mc <= 0) return 0;
  collect_shortlist_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, h->stream>>>(
      score, sd, mc, idx_base, h->best_lb, em, pad, Xc, dc, h->list_idx, h->list_X, h->list_s8, h->list_err,
      h->list_count, SHORTLIST_CAP);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// compares the int8 scores of the shortlist with the exact ones (s64: the re-score pass's output, same order)
int launch_selfcheck(dfb_handle* h, const double* s64, int count) {
  if (count <= 0) return 0;
  selfcheck_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->list_s8, h->list_err, s64, count,
                                                                         h->list_count + 1);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_vec_max(dfb_handle* h, const double* v, int64_t n, double* out) {
  vec_max_kernel<<<1, 1024, 0, h->stream>>>(v, n, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_diag_max(dfb_handle* h, const double* M, int64_t ld, int64_t n, double* out) {
  diag_max_kernel<<<1, 1024, 0, h->stream>>>(M, ld, n, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fill(dfb_handle* h, double* p, int64_t n, double v) {
  if (n <= 0) return 0;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(p, n, v);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_set_diag(dfb_handle* h, double* M, int64_t ld, int64_t from, int64_t to, double v, int add) {
  if (to <= from) return 0;
  set_diag_kernel<<<(unsigned)((to - from + 255) / 256), 256, 0, h->stream>>>(M, ld, from, to, v, add);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ================================================================================================
// Log-marginal-likelihood gradients.
//   GP.compute_grad_log_marginal_likelihood (gp_core.py:229-240):  1/2 tr((alpha alpha^T - K^-1) G),
//   G = Kernel.gradient(param, X, X) (kernel.py:202-217 SE, 301-322 Matern).
// K^-1 = W^T W comes from one triangular DMMA product of L^-T with itself (gemm.cuh, tri = 3); this kernel then
// walks the lower 128 x 128 tiles, re-derives every G entry from the scaled coordinates (dK/dparam is never
// materialised) and reduces M_ij G_ij for ALL parameters of the kernel in one pass:
//   slot 0 'scale'                G = K itself
//   slot 1 trace part of 'noise_var'  sum_i M_ii          (host multiplies by noise_var)
//   slot 3 'same_dim_bandwidths'  SE: K D2 / bw_0          Matern: T1 (-r / bw_0)
//   slot 4+q 'dim_bandwidths', q  SE: K d2_q / bw_q        Matern: T1 (-(d2_q / bw_q) / r), 0 on the diagonal
// with D2 the scaled squared distance, d2_q its one-coordinate version (dist_squared on a column, as the
// reference forms it), r = sqrt(D2) and T1 = scale c w (u' - s2 u) in the notation of kernel.py:272-290.
// Symmetric: off-diagonal entries of the lower triangle count twice.  Per-warp slots in shared memory and a
// fixed-order final sum keep the result deterministic.
// ================================================================================================
constexpr int GRAD_FIXED = 4;

__global__ void __launch_bounds__(256)
lml_grad_tile_kernel(const dfb_kernel_desc* __restrict__ desc_g, const double* __restrict__ xs,
                     const double* __restrict__ nrm, int64_t npad, const double* __restrict__ alpha,
                     const double* __restrict__ Kinv, int64_t ldk, int rb0, int nb, int64_t n, int pstride,
                     double* __restrict__ partial) {
  const int rb = rb0 + (int)blockIdx.x / nb, cb = (int)blockIdx.x % nb;
  if (cb > rb) return;
  __shared__ dfb_factor_desc fsh;
  __shared__ double bw_sh[DFB_MAX_SLOTS];
  __shared__ double red[GRAD_FIXED + DFB_MAX_SLOTS][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = desc_g->factors[0].n_dims;
  if (tid == 0) fsh = desc_g->factors[0];
  for (int q = tid; q < D; q += 256) bw_sh[q] = desc_g->slot_bandwidth[q];
  for (int idx = tid; idx < (GRAD_FIXED + D) * 8; idx += 256) (&red[0][0])[idx] = 0.0;
  __syncthreads();
  const dfb_factor_desc f = fsh;
  const bool se = (f.kind == DFB_BASE_SE);
  const double bw0 = bw_sh[0];
  const int64_t j = (int64_t)cb * TILE + (tid & 127);
  const int half = tid >> 7;
  const bool jvalid = j < n;
  const double aj = alpha[j], nj = nrm[j];
  double acc_scale = 0.0, acc_tr = 0.0, acc_same = 0.0;
  for (int pass = 0; pass < 4; pass++) {
    const int64_t i0 = (int64_t)rb * TILE + half * 64 + pass * 16;
    double t1m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int64_t i = i0 + k;
      double wgt = 0.0;
      if (jvalid && i < n) wgt = (rb != cb || j < i) ? 2.0 : (j == i ? 1.0 : 0.0);
      double dot = 0.0;
      for (int q = 0; q < D; q++) dot = fma(xs[(int64_t)q * npad + i], xs[(int64_t)q * npad + j], dot);
      double d2 = __dadd_rn(__dadd_rn(nj, nrm[i]), -2.0 * dot);
      d2 = fmax(d2, 0.0);
      const double M = wgt * (alpha[i] * aj - Kinv[(i - (int64_t)rb0 * TILE) * ldk + j]);
      if (i == j && jvalid) acc_tr += M;
      if (se) {
        const double base = f.scale * dfb_exp_nonpos(d2 * -0.5);
        acc_scale = fma(M, base, acc_scale);
        acc_same = fma(M, base * (d2 / bw0), acc_same);
        t1m[k] = M * base;
      } else {
        const double dist = sqrt(d2);
        const double mult = f.s8 * dist;
        double u = 0.0, up = 0.0;
        for (int t = 0; t <= f.p; t++) {
          const int e = f.p - t;
          double pw = 1.0, pw1 = 1.0;                      // mult^e, mult^(e-1)
          for (int r = 0; r < e; r++) { pw1 = pw; pw *= mult; }
          u += f.coeffs[t] * pw;
          if (e > 0) up += f.s8 * (double)e * f.coeffs[t] * pw1;
        }
        const double w = f.gamma_ratio * dfb_exp_nonpos(-f.s2 * dist);
        acc_scale = fma(M, f.scale * (u * w), acc_scale);
        const double T1 = f.scale * w * (up - f.s2 * u);
        acc_same = fma(M, T1 * (-(dist / bw0)), acc_same);
        // the reference zeroes the diagonal distances (np.fill_diagonal) before the per-dimension gradient
        t1m[k] = (i == j || wgt == 0.0) ? 0.0 : M * T1 * (-1.0 / dist);     // wgt 0: padding / upper half of a diagonal tile
      }
    }
    for (int q = 0; q < D; q++) {
      const double xj = xs[(int64_t)q * npad + j];
      const double xj2 = xj * xj, ibw = bw_sh[q];
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const double xi = xs[(int64_t)q * npad + i0 + k];
        double dsq = __dadd_rn(__dadd_rn(xj2, __dmul_rn(xi, xi)), -2.0 * __dmul_rn(xi, xj));
        dsq = fmax(dsq, 0.0);
        if (t1m[k] != 0.0) sacc = fma(t1m[k], dsq / ibw, sacc);
      }
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) red[GRAD_FIXED + q][warp] += sacc;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc_scale += __shfl_xor_sync(0xffffffffu, acc_scale, o);
    acc_tr += __shfl_xor_sync(0xffffffffu, acc_tr, o);
    acc_same += __shfl_xor_sync(0xffffffffu, acc_same, o);
  }
  if (lane == 0) { red[0][warp] = acc_scale; red[1][warp] = acc_tr; red[3][warp] = acc_same; }
  __syncthreads();
  if (tid < GRAD_FIXED + D) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; w8++) s += red[tid][w8];
    const int64_t t = (int64_t)rb * (rb + 1) / 2 + cb;
    partial[t * pstride + tid] = s;
  }
}

// out[p] = sum over the lower tiles, in tile order; out[2] = sum(alpha) ('noise_mean', gp_core.py:234-235)
__global__ void lml_grad_reduce_kernel(const double* __restrict__ partial, int64_t n_tiles, int pstride, int n_out,
                                       const double* __restrict__ alpha, int64_t n, double* __restrict__ out) {
  const int p = threadIdx.x;
  if (p >= n_out) return;
  double s = 0.0;
  if (p == 2) {
    for (int64_t i = 0; i < n; i++) s += alpha[i];
  } else {
    for (int64_t t = 0; t < n_tiles; t++) s += partial[t * pstride + p];
  }
  out[p] = s;
}

int launch_lml_grad_tiles(dfb_handle* h, const dfb_kernel_desc* d_desc, const double* xs, const double* nrm, int64_t npad,
                          const double* alpha, const double* Kinv, int64_t ldk, int rb0, int n_rb, int nb, int64_t n,
                          int pstride, double* partial) {
  lml_grad_tile_kernel<<<(unsigned)(n_rb * nb), 256, 0, h->stream>>>(d_desc, xs, nrm, npad, alpha, Kinv, ldk, rb0, nb, n,
                                                                    pstride, partial);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_lml_grad_reduce(dfb_handle* h, const double* partial, int64_t n_tiles, int pstride, int n_out,
                           const double* alpha, int64_t n, double* out) {
  lml_grad_reduce_kernel<<<1, 256, 0, h->stream>>>(partial, n_tiles, pstride, n_out, alpha, n, out);
  h->launches++;
  DFB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dfb
