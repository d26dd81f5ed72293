// C-ABI of libdfb200.so (include/dfb200.h): handle, workspace carve-up and the call sequences.
// No CPU fallback anywhere: every entry point drives CUDA kernels on the handle's stream.
#include <stdarg.h>
#include <new>
#include <stdlib.h>
#include "kernels.cuh"
#include "gemm_tma.h"

namespace dfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int64_t default_chunk(int64_t npad) {
  int64_t c = ((int64_t)32 << 20) / npad;   // ~256 MB of K_* rows per chunk
  c = c / TILE * TILE;
  if (c < TILE) c = TILE;
  if (c > 65536) c = 65536;
  return c;
}

struct Carver {
  char* base;
  size_t off;
  explicit Carver(char* b) : base(b), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

// One definition of the layout, used both to size the workspace and to carve it.
static size_t carve(dfb_handle* h, char* base, int64_t n_max, int64_t chunk) {
  const int64_t npad = round_up(n_max < 1 ? 1 : n_max, TILE);
  if (chunk <= 0) chunk = default_chunk(npad);
  chunk = round_up(chunk, TILE);
  const int64_t nb = npad / TILE;
  Carver c(base);
  double* T = c.take<double>((size_t)(2 * npad + TILE) * npad);
  double* W = c.take<double>((size_t)npad * npad);
  double* Dinv = c.take<double>((size_t)TILE * TILE);
  double* X = c.take<double>((size_t)npad * DFB_MAX_SLOTS);
  double* yc = c.take<double>((size_t)npad);
  double* alpha = c.take<double>((size_t)npad);
  double* tr_xs = c.take<double>((size_t)npad * DFB_MAX_SLOTS);
  double* tr_nrm = c.take<double>((size_t)npad * DFB_MAX_FACTORS);
  double* te_xs = c.take<double>((size_t)npad * DFB_MAX_SLOTS);
  double* te_nrm = c.take<double>((size_t)npad * DFB_MAX_FACTORS);
  double* Ks = c.take<double>((size_t)chunk * npad);
  // three pair-interleaved digit planes (2 bytes per entry each) + one compact plane of the leading digit
  int8_t* Wi8 = c.take<int8_t>((size_t)7 * npad * npad);
  int8_t* Ki8 = c.take<int8_t>((size_t)7 * chunk * npad);
  int8_t* Ki8b = c.take<int8_t>((size_t)7 * chunk * npad);
  double* mu_b = c.take<double>((size_t)chunk);
  double* kssv_b = c.take<double>((size_t)chunk);
  double* cprep = c.take<double>((size_t)chunk * 10);
  double* mu_part = c.take<double>((size_t)(npad / 64 + 2) * chunk);      // per 64-point block partial sums of mu
  double* rowscale = c.take<double>((size_t)npad);
  double* rowinv = c.take<double>((size_t)npad);
  int64_t* list_idx = c.take<int64_t>((size_t)SHORTLIST_CAP);
  double* list_X = c.take<double>((size_t)SHORTLIST_CAP * DFB_MAX_SLOTS);
  int* list_count = c.take<int>(4);
  double* list_s8 = c.take<double>((size_t)SHORTLIST_CAP);
  double* list_err = c.take<double>((size_t)SHORTLIST_CAP);
  double* blk_lb = c.take<double>((size_t)chunk / 128 + 16);
  double* best_lb = c.take<double>(1);
  double* partial = c.take<double>((size_t)nb * chunk);
  double* mu = c.take<double>((size_t)chunk);
  double* sd = c.take<double>((size_t)chunk);
  double* score = c.take<double>((size_t)chunk);
  double* kssv = c.take<double>((size_t)chunk);
  double* stage = c.take<double>((size_t)chunk * DFB_MAX_SLOTS);
  double* blk_score = c.take<double>((size_t)chunk / 128 + 16);
  int64_t* blk_index = c.take<int64_t>((size_t)chunk / 128 + 16);
  double* best_score = c.take<double>(1);
  int64_t* best_index = c.take<int64_t>(1);
  double* red = c.take<double>(8);
  int* info = c.take<int>(4);
  dfb_kernel_desc* d0 = c.take<dfb_kernel_desc>(1);
  dfb_kernel_desc* d1 = c.take<dfb_kernel_desc>(1);
  dfb_kernel_desc* d2 = c.take<dfb_kernel_desc>(1);
  // dfb_extend_posterior's snapshot of what it overwrites: the last row block of L, the last block
  // column of L^-T, that block of the y row, alpha
  double* ext_save = c.take<double>((size_t)(2 * TILE + 1) * npad + TILE);
  if (h != nullptr && base != nullptr) {
    h->ext_save = ext_save;
    h->Ki8b = Ki8b; h->mu_b = mu_b; h->kssv_b = kssv_b; h->cprep = cprep; h->mu_part = mu_part;
    h->T = T; h->W = W; h->Dinv = Dinv; h->X = X; h->yc = yc; h->alpha = alpha;
    h->tr.xs = tr_xs; h->tr.nrm = tr_nrm; h->te.xs = te_xs; h->te.nrm = te_nrm;
    h->Ks = Ks; h->Wi8 = Wi8; h->Ki8 = Ki8; h->rowscale = rowscale; h->rowinv = rowinv; h->list_idx = list_idx; h->list_X = list_X; h->list_count = list_count; h->list_s8 = list_s8; h->list_err = list_err; h->blk_lb = blk_lb; h->best_lb = best_lb; h->partial = partial; h->mu = mu; h->sd = sd; h->score = score; h->kssv = kssv; h->stage = stage;
    h->blk_score = blk_score; h->blk_index = blk_index; h->best_score = best_score;
    h->best_index = best_index; h->red = red; h->info = info;
    h->d_desc_tr = d0; h->d_desc_te = d1; h->d_desc_tmp = d2;
    h->n_max = n_max; h->npad_max = npad; h->chunk = chunk;
  }
  return c.off + 256;
}

static size_t carve_ts(dfb_handle* h, char* base, int64_t n_max, int64_t mb) {
  const int64_t npad = round_up(n_max < 1 ? 1 : n_max, TILE);
  const int64_t mbp = round_up(mb < 1 ? 1 : mb, TILE);
  Carver c(base);
  double* Vt = c.take<double>((size_t)mbp * npad);
  double* cxs = c.take<double>((size_t)mbp * DFB_MAX_SLOTS);
  double* cnrm = c.take<double>((size_t)mbp * DFB_MAX_FACTORS);
  double* Cov = c.take<double>((size_t)mbp * mbp);
  double* T2 = c.take<double>((size_t)(2 * mbp + TILE) * mbp);
  double* Ut = c.take<double>((size_t)256 * mbp);
  double* Sm = c.take<double>((size_t)256 * mbp);
  double* mu = c.take<double>((size_t)mbp);
  double* red = c.take<double>(8);
  int* info = c.take<int>(4);
  if (h != nullptr && base != nullptr) {
    h->ts_Vt = Vt; h->ts_cxs = cxs; h->ts_cnrm = cnrm; h->ts_Cov = Cov; h->ts_T = T2; h->ts_Ut = Ut;
    h->ts_Sm = Sm; h->ts_mu = mu; h->ts_red = red; h->ts_info = info; h->ts_mb = mbp;
  }
  return c.off + 256;
}

static int check_desc(const dfb_kernel_desc* d) {
  if (d == nullptr) { set_error("kernel descriptor is NULL"); return -1; }
  if (d->n_terms < 1 || d->n_terms > DFB_MAX_TERMS || d->n_factors < 1 ||
      d->n_factors > DFB_MAX_FACTORS || d->n_slots < 1 || d->n_slots > DFB_MAX_SLOTS) {
    set_error("kernel descriptor out of range (terms %d, factors %d, slots %d)", d->n_terms,
              d->n_factors, d->n_slots);
    return -1;
  }
  if (d->term_first_factor[0] != 0 || d->term_first_factor[d->n_terms] != d->n_factors) {
    set_error("kernel descriptor: term_first_factor does not cover the factors");
    return -1;
  }
  for (int f = 0; f < d->n_factors; f++) {
    const dfb_factor_desc& fd = d->factors[f];
    if ((fd.kind != DFB_BASE_SE && fd.kind != DFB_BASE_MATERN) || fd.n_dims < 1 ||
        fd.slot_off < 0 || fd.slot_off + fd.n_dims > d->n_slots || fd.p < 0 ||
        fd.p > DFB_MAX_MATERN_P) {
      set_error("kernel descriptor: bad factor %d", f);
      return -1;
    }
  }
  for (int s = 0; s < d->n_slots; s++) {
    if (d->slot_train_coord[s] < 0 || d->slot_train_coord[s] >= d->train_dim ||
        d->slot_cand_coord[s] < 0 || d->slot_cand_coord[s] >= d->cand_dim ||
        !(d->slot_bandwidth[s] > 0.0)) {
      set_error("kernel descriptor: bad slot %d", s);
      return -1;
    }
  }
  return 0;
}

#define DFB_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

static int need(dfb_handle* h, bool ws, bool kern, bool train, bool post, bool w) {
  if (h == nullptr) { set_error("handle is NULL"); return -1; }
  if (ws && h->ws == nullptr) { set_error("no workspace: call dfb_set_workspace first"); return -1; }
  if (kern && !h->have_kernel) { set_error("no kernel: call dfb_set_kernel first"); return -1; }
  if (train && !h->have_train) { set_error("no training data: call dfb_set_train first"); return -1; }
  if (post && !h->have_post) { set_error("no posterior: call dfb_build_posterior first"); return -1; }
  if (w && !h->have_w) { set_error("posterior was built LML-only: W = L^-1 is not available"); return -1; }
  return 0;
}

static int ensure_train_scaled(dfb_handle* h) {
  if (!h->tr_prepped) {
    DFB_TRY(launch_prep_scaled(h, h->d_desc_tr, 1, h->X, h->n, h->d, h->tr.xs, h->tr.nrm, h->npad));
    h->tr_prepped = true;
  }
  return 0;
}

static int ensure_test_scaled(dfb_handle* h) {
  if (h->have_test_kernel && !h->te_prepped) {
    DFB_TRY(launch_prep_scaled(h, h->d_desc_te, 1, h->X, h->n, h->d, h->te.xs, h->te.nrm, h->npad));
    h->te_prepped = true;
  }
  return 0;
}

// The blocked right-looking factorisation of the tall matrix [A ; I ; y^T] (see gemm.cuh):
// top -> L, bottom -> L^-T, y row -> (L^-1 y)^T.
//
// Schedule: chol_diag is a one-CTA, latency-bound kernel (~0.1 ms x npad/128 steps), so the plain
// step-after-step order leaves 147 SMs idle for a third of the build.  With look-ahead the trailing update of
// step k is split: the column of the NEXT panel (block k+1) is updated first on the critical-path stream, so
// chol_diag(k+1) and the panel solve of step k+1 run while the bulk of update k (column blocks >= k+2) is still
// in flight on a second stream.
//   hi:  chol(k) panel(k) [P_k] wait(R_k-1) next(k)  chol(k+1) panel(k+1) [P_k+1] wait(R_k) next(k+1) ...
//   lo:                   wait(P_k) rest(k) [R_k]                         wait(P_k+1) rest(k+1) [R_k+1]
// next(k) and rest(k-1) both accumulate into column block k+1, hence wait(R_k-1); rest(k) after rest(k-1) by
// stream order.  The arithmetic per tile is unchanged (same kernel, same k-order): results are bit-identical
// to the single-stream schedule.
struct StreamSwap {
  dfb_handle* h;
  cudaStream_t user;
  explicit StreamSwap(dfb_handle* hh) : h(hh), user(hh->stream) {}
  ~StreamSwap() { h->stream = user; }
};

static int ensure_factor_streams(dfb_handle* h) {
  if (h->fs_hi != nullptr) return 0;
  int lo = 0, hi = 0;
  DFB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least, hi = greatest priority
  DFB_CUDA_OK(cudaStreamCreateWithPriority(&h->fs_hi, cudaStreamNonBlocking, hi));
  DFB_CUDA_OK(cudaStreamCreateWithPriority(&h->fs_lo, cudaStreamNonBlocking, lo));
  cudaEvent_t* evs[5] = {&h->fe_fork, &h->fe_panel, &h->fe_rest, &h->fe_join_hi, &h->fe_join_lo};
  for (int i = 0; i < 5; i++) DFB_CUDA_OK(cudaEventCreateWithFlags(evs[i], cudaEventDisableTiming));
  return 0;
}

static int factorise_tall(dfb_handle* h, double* T, int64_t npad, double* Dinv, int* info,
                          bool with_bottom) {
  const int nb = (int)(npad / TILE);
  const bool la = h->lookahead != 0 && nb >= 4;
  StreamSwap guard(h);
  if (la) {
    DFB_TRY(ensure_factor_streams(h));
    DFB_CUDA_OK(cudaEventRecord(h->fe_fork, guard.user));
    DFB_CUDA_OK(cudaStreamWaitEvent(h->fs_hi, h->fe_fork, 0));
    DFB_CUDA_OK(cudaStreamWaitEvent(h->fs_lo, h->fe_fork, 0));
  }
  bool rest_pending = false;
  for (int step = 0; step < nb; step++) {
    if (la) h->stream = h->fs_hi;
    DFB_TRY(launch_chol_diag(h, T, npad, step, Dinv, info));
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = T; g.lda = npad; g.B = Dinv; g.ldb = TILE; g.D = T; g.ldd = npad;
    g.alpha = 1.0; g.mode = MODE_PANEL; g.K = TILE; g.step = step; g.nb = nb; g.info = info;
    g.skip_bottom = with_bottom ? 0 : 1;
    const int rows = 2 * nb + 1 - (step + 1);
    DFB_TRY(launch_gemm(h, g, EPI_STORE, rows));
    const int ncols = nb - step - 1;
    if (ncols <= 0) continue;
    g.mode = MODE_TRAIL; g.alpha = -1.0; g.B = nullptr; g.ldb = npad;
    if (!la) {
      DFB_TRY(launch_gemm(h, g, EPI_STORE, rows * ncols));
      continue;
    }
    DFB_CUDA_OK(cudaEventRecord(h->fe_panel, h->fs_hi));
    if (rest_pending) DFB_CUDA_OK(cudaStreamWaitEvent(h->fs_hi, h->fe_rest, 0));
    g.tr_j0 = 0; g.tr_nc = 1;                       // the next panel's column, on the critical path
    DFB_TRY(launch_gemm(h, g, EPI_STORE, rows));
    if (ncols > 1) {
      h->stream = h->fs_lo;
      DFB_CUDA_OK(cudaStreamWaitEvent(h->fs_lo, h->fe_panel, 0));
      g.tr_j0 = 1; g.tr_nc = ncols - 1;
      DFB_TRY(launch_gemm(h, g, EPI_STORE, rows * (ncols - 1)));
      DFB_CUDA_OK(cudaEventRecord(h->fe_rest, h->fs_lo));
      rest_pending = true;
    }
  }
  if (la) {
    DFB_CUDA_OK(cudaEventRecord(h->fe_join_hi, h->fs_hi));
    DFB_CUDA_OK(cudaEventRecord(h->fe_join_lo, h->fs_lo));
    DFB_CUDA_OK(cudaStreamWaitEvent(guard.user, h->fe_join_hi, 0));
    DFB_CUDA_OK(cudaStreamWaitEvent(guard.user, h->fe_join_lo, 0));
  }
  return 0;
}

// ---- optional per-class event timing -------------------------------------------------------------
static int prof_flush(dfb_handle* h, int cls) {
  ProfClass& pc = h->prof[cls];
  if (pc.n == 0) return 0;
  DFB_CUDA_OK(cudaEventSynchronize(pc.stop[pc.n - 1]));
  for (int i = 0; i < pc.n; i++) {
    float ms = 0.f;
    DFB_CUDA_OK(cudaEventElapsedTime(&ms, pc.start[i], pc.stop[i]));
    pc.acc_ms += ms;
    pc.acc_units += pc.units[i];
    pc.acc_launches += 1;
  }
  pc.n = 0;
  return 0;
}
static int prof_begin(dfb_handle* h, int cls) {
  if (!h->prof_on) return 0;
  ProfClass& pc = h->prof[cls];
  if (!pc.created) {
    for (int i = 0; i < PROF_RING; i++) {
      DFB_CUDA_OK(cudaEventCreate(&pc.start[i]));
      DFB_CUDA_OK(cudaEventCreate(&pc.stop[i]));
    }
    pc.created = true;
  }
  if (pc.n == PROF_RING) DFB_TRY(prof_flush(h, cls));
  DFB_CUDA_OK(cudaEventRecord(pc.start[pc.n], h->stream));
  return 0;
}
static int prof_end(dfb_handle* h, int cls, double units) {
  if (!h->prof_on) return 0;
  ProfClass& pc = h->prof[cls];
  DFB_CUDA_OK(cudaEventRecord(pc.stop[pc.n], h->stream));
  pc.units[pc.n] = units;
  pc.n++;
  return 0;
}

// A-priori bound on the int8-slice path's ABSOLUTE sigma^2 error for the active kernel.
//
// Error model.  One entry of v = L^-1 k_* is a sum over k <= i of digit-truncation and dropped-product terms, each
// bounded by c 2^-q rowscale_i colscale (q = 43 for six radix-128 digits, 40 for five radix-256 digits; the low
// digits of an operand are unrelated to its magnitude, so the terms do not shrink with |W_ik K_k|) and, being
// rounding residues of unrelated numbers, of effectively independent sign: |dv_i| grows like sqrt(n), exactly as
// the rounding error of the fp64 dot product it replaces (whose worst-case bound n eps is never approached either).
// d(sigma^2) = 2 sum_i v_i dv_i has standard deviation <= 2 |v| max_i sd(dv_i) <= 2 sqrt(k(x,x)) max_i sd(dv_i)
// for independent dv_i (|v|^2 <= k(x,x) - sigma^2 <= k(x,x)).  A worst-case (n instead of sqrt(n), aligned signs
// over i) bound would be ~sqrt(n) n / 8 ~ 4 10^4 times larger at N = 5000 and is as unattainable as LAPACK's own.
//
// The constant 8 keeps the bound above every maximum measured over the validation sweep (tools/sweep_i8_bound.py ->
// profiles/r02_i8_bound_sweep.json: 480 configurations -- N 1024..5000, noise 1e-2..1e-8 of the scale, scale 1e-2..1e4,
// SE / Matern-5/2 / Matern-1/2 / additive / product kernels, both digit schemes, 13056 candidates each, guard off): the
// worst measured / bound ratio is 0.28 (Matern-1/2, N = 1024, radix 256), typically 0.02-0.15; the worst ABSOLUTE error
// among the 185 configurations the guard admits is 1.05e-9 against the 1e-8 contract.  It is NOT a worst-case bound;
// three things keep the arg-max exact in spite of that:
//   (1) the limit below is ABSOLUTE: the int8 pass is used only while the bound is <= 5e-9, half of the
//       north-star's 1e-8 contract on sigma^2, whatever the kernel scale;
//   (2) dfb_score_argmax re-scores in fp64 every candidate whose int8 score, widened by the bound, could reach the
//       fp64 maximum, and returns the fp64 arg-max of those;
//   (3) after that exact pass the int8 and fp64 scores of the shortlist are compared (selfcheck_kernel): a single
//       candidate outside its allowance voids the int8 pass and the whole call is repeated in fp64
//       (query "last_selfcheck_violations" / "last_selfcheck_ratio").
// score_impl = 0 (env DFB200_SCORE=fp64, option "score_impl") switches the int8 path off altogether.
static double i8_colscale(const dfb_kernel_desc& desc) {
  int e = 0;
  frexp(desc.kss * (1.0 + 1e-9), &e);
  return ldexp(1.0, e + 1);
}
static double i8_sigma2_bound(const dfb_handle* h, const dfb_kernel_desc& desc) {
  return 8.0 * h->i8_rowscale_max * sqrt((double)h->n) * i8_colscale(desc) *
         ldexp(1.0, h->i8_radix256 ? -40 : -43) * sqrt(desc.kss);
}
static const double I8_BOUND_LIMIT = 5e-9;     // absolute: half of the 1e-8 sigma^2 contract
static bool i8_usable(const dfb_handle* h, const dfb_kernel_desc& desc) {
  if (!h->i8_ready || !(desc.kss > 0.0)) return false;
  if (h->i8_unguarded) return true;                 // diagnostics only (tools/sweep_i8_bound.py)
  return i8_sigma2_bound(h, desc) <= I8_BOUND_LIMIT;
}

// Digit planes of W = L^-1 for the tcgen05 path + the tensor maps of both operands.
static int prepare_i8(dfb_handle* h) {
  const int64_t npad = h->npad;
  DFB_TRY(launch_row_exponent(h, h->W, npad, npad, npad, h->rowscale, h->rowinv));
  DFB_TRY(launch_vec_max(h, h->rowscale, h->n, h->red + 4));
  DFB_CUDA_OK(cudaMemcpyAsync(&h->i8_rowscale_max, h->red + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  // Digit scheme of the CTA-pair kernel: five radix-256 digits (15 products) when the a-priori bound of
  // that coarser expansion passes for the training kernel, else six radix-128 digits (21 products).  The
  // radix-256 groups also need int32 headroom: 5 K 2^14 < 2^31.
  h->i8_radix256 = 0;
  if (h->i8_impl == 2 && h->i8_radix_opt != 0 && npad <= 24576) {
    h->i8_radix256 = 1;
    const dfb_kernel_desc& dtr = h->desc_tr;
    if (h->i8_radix_opt < 0 && dtr.kss > 0.0 && i8_sigma2_bound(h, dtr) > I8_BOUND_LIMIT)
      h->i8_radix256 = 0;
  }
  // pair-interleaved digit planes: 3 planes of rows x (2 * npad) bytes
  DFB_TRY(launch_slice_i8(h, h->W, npad, npad, npad, h->rowinv, 0.0, h->Wi8, 2 * npad * npad, 2 * npad));
  if (h->i8_impl >= 1) {
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK2h, h->Ki8, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 64, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK3h, h->Ki8, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 64, 3));
    // compact plane of the leading digit (row-major rows of npad bytes, SWIZZLE_128B boxes of 128 k-values)
    DFB_TRY(make_tensor_map_3d_u8(&h->tmW1c, h->Wi8 + 6 * npad * npad, npad, npad, 1, npad, npad * npad, 128, 128, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK1c, h->Ki8 + 6 * h->chunk * npad, npad, h->chunk, 1, npad, h->chunk * npad, 128, 64, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK2h_b, h->Ki8b, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 64, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK3h_b, h->Ki8b, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 64, 3));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK1c_b, h->Ki8b + 6 * h->chunk * npad, npad, h->chunk, 1, npad, h->chunk * npad, 128, 64, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmW2, h->Wi8, 2 * npad, npad, 3, 2 * npad, 2 * npad * npad, 64, 128, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmW3, h->Wi8, 2 * npad, npad, 3, 2 * npad, 2 * npad * npad, 64, 128, 3));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK2, h->Ki8, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 128, 1));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmK3, h->Ki8, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 64, 128, 3));
  } else {
    DFB_TRY(make_tensor_map_3d_u8(&h->tmWi8, h->Wi8, 2 * npad, npad, 3, 2 * npad, 2 * npad * npad, 128, 128, 3));
    DFB_TRY(make_tensor_map_3d_u8(&h->tmKi8, h->Ki8, 2 * npad, h->chunk, 3, 2 * npad, 2 * h->chunk * npad, 128, 64, 3));
  }
  h->i8_ready = true;
  return 0;
}


// ---- incremental posterior update ------------------------------------------------------
// Appending q training points changes only the LAST row block of L (as long as n + q stays inside the same
// padded size): Cholesky row i depends on rows <= i alone.  Rather than re-running the N^3/3 right-looking
// factorisation, the pre-step state of the last block column of the tall matrix [A ; I ; y^T] is rebuilt by
// left-looking products against the finished factor and the last factorisation step is replayed:
//     P    = A[last, :m0] L00^-T = A[last, :m0] W00^T          top, row block nb-1, columns < m0
//     S    = A[last, last] - P P^T                             top, diagonal block
//     Wt_c = -(L00^-T) P^T                                     L^-T rows < m0, last block column
//     y_c  = y[last] - v[:m0] P^T                              y row, last block
// then chol_diag + the panel solve of step nb-1 turn (S, I, Wt_c, y_c) into (L_dd, L_dd^-T, L^-T's last
// block column, v[last]).  Cost 4 N^2 * 128 flops + one 128 x 128 Cholesky instead of 2 N^3 / 3.
static int replay_last_block(dfb_handle* h, int32_t flags, double* lml_out_host) {
  const int64_t n = h->n, npad = h->npad;
  const int nb = (int)(npad / TILE), step = nb - 1;
  const int64_t m0 = (int64_t)step * TILE;
  double* top_row = h->T + m0 * npad;                      // row block nb-1 of the top
  double* mid = h->T + npad * npad;                        // L^-T
  double* yrow = h->T + 2 * npad * npad;                   // y row block (row 0 holds the data)
  h->have_post = h->have_w = false;
  h->tr_prepped = h->te_prepped = false;
  DFB_TRY(ensure_train_scaled(h));
  DFB_CUDA_OK(cudaMemsetAsync(h->info, 0, sizeof(int) * 4, h->stream));
  // A[last, :] = K(X[last], X) + (noise + jitter) I, identity on the padding rows -> Ks scratch (128 x npad)
  DFB_TRY(launch_kstar(h, h->d_desc_tr, h->desc_tr, 1, h->tr.xs, h->tr.nrm, npad, nullptr, h->X + m0 * h->d,
                       n - m0, h->d, TILE, h->Ks, npad, n, npad, 0.0, nullptr, nullptr));
  DFB_TRY(launch_set_diag(h, h->Ks + m0, npad, 0, n - m0, h->noise_plus_jitter, 1));
  DFB_TRY(launch_set_diag(h, h->Ks + m0, npad, n - m0, TILE, 1.0, 0));
  // The four left-looking products have 1 .. nb-1 output tiles with k-depths up to m0 ~ N: each tile's k-range is
  // split over several CTAs (launch_gemm_splitk) so that they fill the GPU; scratch lives in the K_* chunk buffer
  // behind the 128 rows of A.
  const int KS = 8;
  double* scratch = h->Ks + (int64_t)TILE * npad;
  const bool split = step >= 2 && h->chunk >= (int64_t)TILE * (1 + KS);       // KS * 128 x npad doubles of scratch
  GemmArgs g;
  if (step > 0) {
    // P[a][i] = sum_{k <= i} A[a][k] W[i][k]
    memset(&g, 0, sizeof(g));
    g.A = h->Ks; g.lda = npad; g.B = h->W; g.ldb = npad; g.D = top_row; g.ldd = npad; g.alpha = 1.0;
    g.mode = MODE_GENERIC; g.n_rb = 1; g.n_cb = step; g.K = (int)m0; g.tri = 2;
    if (split) DFB_TRY(launch_gemm_splitk(h, g, 4, scratch));
    else DFB_TRY(launch_gemm(h, g, EPI_STORE, step));
    // Wt_c[j][a] = -sum_k Wt[j][k] P[a][k]
    memset(&g, 0, sizeof(g));
    g.A = mid; g.lda = npad; g.B = top_row; g.ldb = npad; g.D = mid + m0; g.ldd = npad; g.alpha = -1.0;
    g.mode = MODE_GENERIC; g.n_rb = step; g.n_cb = 1; g.K = (int)m0;
    if (split) DFB_TRY(launch_gemm_splitk(h, g, 4, scratch));
    else DFB_TRY(launch_gemm(h, g, EPI_STORE, step));
  }
  // S = A[last, last] - P P^T  (K = 0 degenerates to a copy)
  memset(&g, 0, sizeof(g));
  g.A = top_row; g.lda = npad; g.B = top_row; g.ldb = npad; g.C = h->Ks + m0; g.ldc = npad;
  g.D = top_row + m0; g.ldd = npad; g.alpha = -1.0; g.mode = MODE_GENERIC; g.n_rb = 1; g.n_cb = 1; g.K = (int)m0;
  if (split) DFB_TRY(launch_gemm_splitk(h, g, KS, scratch));
  else DFB_TRY(launch_gemm(h, g, EPI_STORE, 1));
  // identity in the diagonal block of L^-T
  DFB_CUDA_OK(cudaMemset2DAsync(mid + m0 * npad + m0, sizeof(double) * npad, 0, sizeof(double) * TILE, TILE, h->stream));
  DFB_TRY(launch_set_diag(h, mid, npad, m0, npad, 1.0, 0));
  // y_c = y[last] - v[:m0] P^T (rows 1..127 of the y block are zero and stay zero)
  DFB_TRY(launch_copy_pad(h, h->yc + m0, n - m0, yrow + m0, TILE));
  memset(&g, 0, sizeof(g));
  g.A = yrow; g.lda = npad; g.B = top_row; g.ldb = npad; g.C = yrow + m0; g.ldc = npad;
  g.D = yrow + m0; g.ldd = npad; g.alpha = -1.0; g.mode = MODE_GENERIC; g.n_rb = 1; g.n_cb = 1; g.K = (int)m0;
  if (split) DFB_TRY(launch_gemm_splitk(h, g, KS, scratch));
  else DFB_TRY(launch_gemm(h, g, EPI_STORE, 1));
  // replay of factorisation step nb-1
  DFB_TRY(launch_chol_diag(h, h->T, npad, step, h->Dinv, h->info));
  memset(&g, 0, sizeof(g));
  g.A = h->T; g.lda = npad; g.B = h->Dinv; g.ldb = TILE; g.D = h->T; g.ldd = npad;
  g.alpha = 1.0; g.mode = MODE_PANEL; g.K = TILE; g.step = step; g.nb = nb; g.info = h->info;
  DFB_TRY(launch_gemm(h, g, EPI_STORE, 2 * nb + 1 - (step + 1)));
  const double* Wt = mid;
  const double* v = yrow;
  DFB_TRY(launch_transpose(h, Wt, h->W, npad));
  if (flags == DFB_BUILD_FULL) DFB_TRY(launch_alpha(h, Wt, v, h->alpha, n, npad));
  DFB_TRY(launch_lml_reduce(h, h->T, h->yc, flags == DFB_BUILD_FULL ? h->alpha : nullptr, v, n, npad, h->red));
  double red[3];
  int info = 0;
  DFB_CUDA_OK(cudaMemcpyAsync(red, h->red, sizeof(red), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(&info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (info != 0) {
    set_error("extended matrix is not positive definite: non-positive pivot at index %d", info - 1);
    return info;
  }
  h->have_post = true;
  h->have_w = true;
  h->i8_ready = false;
  if (h->score_impl == 1 || (h->score_impl == 2 && h->n >= 1024)) DFB_TRY(prepare_i8(h));
  // the fp64 TMA maps of W and Ks describe (address, npad) only: still valid
  if (lml_out_host != nullptr) {
    const double quad = (flags == DFB_BUILD_FULL) ? red[1] : red[2];
    *lml_out_host = -0.5 * quad - red[0] - 0.5 * (double)n * log(2.0 * M_PI);
  }
  return 0;
}

struct ChunkOut {
  double* mu; double* sd; double* score;   // user pointers (space given), may be NULL
};

// Scores m candidates chunk by chunk: K_* rows + mu -> |L^-1 k_*|^2 -> sd / acquisition / arg-max.
struct ChunkMode {
  bool want_std, do_argmax;
  bool use_i8;                 // int8-slice tcgen05 contraction instead of fp64 DMMA
  bool collect;                // gather the shortlist for the exact re-score
  I8ErrModel em;               // int8 error model (collect): bound on |d sigma^2|, score sensitivity
  double pad;                  // extra slack of the shortlist test
  const int64_t* idx_map;      // global index of each row (re-score pass), NULL = c0 + i
  bool allow_small;            // dfb_eval of <= SMALL_EVAL_M points: row-streaming kernel instead of the tile GEMM
  bool keep_scores;            // leave the scores of a single-chunk pass in h->score (self-check of the shortlist)
};
static ChunkMode chunk_mode(bool want_std, bool do_argmax, bool use_i8, const int64_t* idx_map = nullptr) {
  ChunkMode md;
  memset(&md, 0, sizeof(md));
  md.want_std = want_std; md.do_argmax = do_argmax; md.use_i8 = use_i8; md.idx_map = idx_map;
  return md;
}
constexpr int64_t SMALL_EVAL_M = 32;      // up to four 8-wide passes over W's rows (105 MB each at N = 5000): still ~10x cheaper than one 128-wide tile pass

static int ensure_ks_stream(dfb_handle* h) {
  if (h->ks_stream != nullptr) return 0;
  int lo = 0, hi = 0;
  DFB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority: the contraction's CTAs go first
  if (getenv("DFB200_KS_PRIO") != nullptr) {                    // diagnostics: 0 = equal priorities, 1 = swapped
    if (atoi(getenv("DFB200_KS_PRIO")) == 0) hi = lo; else { const int t = lo; lo = hi; hi = t; }
  }
  DFB_CUDA_OK(cudaStreamCreateWithPriority(&h->ks_stream, cudaStreamNonBlocking, lo));
  DFB_CUDA_OK(cudaStreamCreateWithPriority(&h->gs_stream, cudaStreamNonBlocking, hi));
  DFB_CUDA_OK(cudaEventCreateWithFlags(&h->ks_fork, cudaEventDisableTiming));
  DFB_CUDA_OK(cudaEventCreateWithFlags(&h->ks_join, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    DFB_CUDA_OK(cudaEventCreateWithFlags(&h->ks_k[i], cudaEventDisableTiming));
    DFB_CUDA_OK(cudaEventCreateWithFlags(&h->ks_g[i], cudaEventDisableTiming));
  }
  return 0;
}

// Per chunk two stages:
//   K: (host candidates: staging copy) K_* rows / digit planes + mu + k(x*,x*)      fp64 pipe
//   G: the contraction |L^-1 k_*|^2 -> sd / acquisition / arg-max / shortlist       tensor pipe (int8) or DMMA
// With the CTA-pair int8 contraction the stages of consecutive chunks are software-pipelined over two streams: K(c+1)
// runs on h->ks_stream into the second digit buffer while G(c) runs on the caller's stream (the lean K_* kernel
// co-resides with the persistent tcgen05 kernel: 88 registers x 128 threads and < 1 KB of shared memory per CTA).
// Events carry the two dependencies per buffer: K(c) -> G(c) and G(c) -> K(c+2).  Everything else is unchanged: the
// arithmetic per chunk, the order of the arg-max folds (G stages stay in stream order) and therefore every result.
static int run_chunks(dfb_handle* h, const dfb_acq_desc& acq, const double* Xc, int64_t m, int32_t dc,
                      int32_t space, double mean_const, ChunkOut out, const ChunkMode& md) {
  const bool want_std = md.want_std, do_argmax = md.do_argmax;
  const dfb_kernel_desc& desc = h->have_test_kernel ? h->desc_te : h->desc_tr;
  const dfb_kernel_desc* d_desc = h->have_test_kernel ? h->d_desc_te : h->d_desc_tr;
  const ScaledSet& ss = h->have_test_kernel ? h->te : h->tr;
  if (dc != desc.cand_dim) {
    set_error("candidates have %d columns, the kernel descriptor expects %d", dc, desc.cand_dim);
    return -1;
  }
  if (space == DFB_HOST && dc > DFB_MAX_SLOTS) { set_error("host candidates: dc > %d", DFB_MAX_SLOTS); return -1; }
  DFB_TRY(ensure_train_scaled(h));
  DFB_TRY(ensure_test_scaled(h));
  const int64_t npad = h->npad, Mc = h->chunk;
  const int nb = (int)(npad / TILE);
  if (do_argmax) DFB_TRY(launch_reset_best(h));
  const bool i8 = want_std && md.use_i8;
  const bool seg_ok = i8 && h->i8_fuse && h->kstar_fast && h->kstar_seg && h->i8_impl == 2 && h->i8_radix256;
  const bool pipelined = i8 && h->i8_impl == 2 && h->kstar_overlap && m > Mc;
  StreamSwap guard(h);
  cudaStream_t s_user = guard.user, s_k = s_user, s_g = s_user;
  if (pipelined) {
    DFB_TRY(ensure_ks_stream(h));
    s_k = h->ks_stream;
    s_g = h->gs_stream;
    DFB_CUDA_OK(cudaEventRecord(h->ks_fork, s_user));
    DFB_CUDA_OK(cudaStreamWaitEvent(s_k, h->ks_fork, 0));
    DFB_CUDA_OK(cudaStreamWaitEvent(s_g, h->ks_fork, 0));
  }
  // host candidates are staged in batches of as many whole chunks as the staging buffer holds
  // (chunk x DFB_MAX_SLOTS doubles), so a 6-column candidate matrix needs 1 copy per ~21 chunks
  const int64_t stage_rows = (Mc * DFB_MAX_SLOTS / dc) / Mc * Mc;
  int64_t staged_lo = 0, staged_hi = 0;
  // Page-locked host candidates (what the streamed `rand` maximiser hands over): the staging buffer is used as two halves
  // and the copy of batch b+1 runs on a copy stream while batch b is scored (ev_cp: copy done, ev_free: every stage
  // that reads the half is done).  Pageable memory keeps the single-buffer copy on the compute stream: its
  // cudaMemcpyAsync would block the host on the half's ev_free and stall the launches of the batch in flight.
  const int64_t half_rows = (stage_rows / Mc / 2) * Mc;
  bool dbuf = false;
  if (space == DFB_HOST && half_rows >= Mc && m > half_rows) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, Xc) == cudaSuccess && pa.type == cudaMemoryTypeHost) dbuf = true;
    cudaGetLastError();                                   // an unregistered pointer may leave a sticky-free error behind
  }
  if (dbuf) {
    if (h->cp_stream == nullptr) {
      DFB_CUDA_OK(cudaStreamCreateWithFlags(&h->cp_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        DFB_CUDA_OK(cudaEventCreateWithFlags(&h->cp_done[i], cudaEventDisableTiming));
        DFB_CUDA_OK(cudaEventCreateWithFlags(&h->cp_free[i], cudaEventDisableTiming));
      }
      DFB_CUDA_OK(cudaEventCreateWithFlags(&h->cp_fork, cudaEventDisableTiming));
    }
    DFB_CUDA_OK(cudaEventRecord(h->cp_fork, s_user));     // the copy stream starts after everything already on the caller's stream
    DFB_CUDA_OK(cudaStreamWaitEvent(h->cp_stream, h->cp_fork, 0));
  }
  auto issue_copy = [&](int64_t bi) -> int {               // batch bi -> half bi % 2, on the copy stream
    const int64_t lo = bi * half_rows;
    const int64_t hi = (m - lo < half_rows) ? m : lo + half_rows;
    if (bi >= 2) DFB_CUDA_OK(cudaStreamWaitEvent(h->cp_stream, h->cp_free[bi & 1], 0));
    DFB_CUDA_OK(cudaMemcpyAsync(h->stage + (bi & 1) * half_rows * dc, Xc + lo * dc, sizeof(double) * (hi - lo) * dc,
                                cudaMemcpyHostToDevice, h->cp_stream));
    DFB_CUDA_OK(cudaEventRecord(h->cp_done[bi & 1], h->cp_stream));
    return 0;
  };
  const int64_t n_chunks = (m + Mc - 1) / Mc;
  const double* xc_of[2] = {nullptr, nullptr};

  auto stage_k = [&](int64_t ci) -> int {
    const int b = pipelined ? (int)(ci & 1) : 0;
    const int64_t c0 = ci * Mc;
    const int64_t mc = (m - c0 < Mc) ? (m - c0) : Mc;
    const int64_t m_rows = round_up(mc, TILE);
    h->stream = s_k;
    if (pipelined && ci >= 2) DFB_CUDA_OK(cudaStreamWaitEvent(s_k, h->ks_g[b], 0));     // G(ci-2) is done with buffer b
    const double* xc_dev;
    if (space == DFB_HOST && dbuf) {
      const int64_t bi = c0 / half_rows;
      if (c0 % half_rows == 0) {                           // first chunk of a batch
        if (bi == 0) DFB_TRY(issue_copy(0));
        if ((bi + 1) * half_rows < m) DFB_TRY(issue_copy(bi + 1));
        DFB_CUDA_OK(cudaStreamWaitEvent(s_k, h->cp_done[bi & 1], 0));
      }
      xc_dev = h->stage + (bi & 1) * half_rows * dc + (c0 - bi * half_rows) * dc;
    } else if (space == DFB_HOST) {
      if (c0 >= staged_hi) {
        // the shortlist collection of earlier chunks reads the staged rows: wait for the latest G stage
        if (pipelined && ci >= 1) DFB_CUDA_OK(cudaStreamWaitEvent(s_k, h->ks_g[(ci - 1) & 1], 0));
        staged_lo = c0;
        staged_hi = (m - c0 < stage_rows) ? m : c0 + stage_rows;
        DFB_CUDA_OK(cudaMemcpyAsync(h->stage, Xc + staged_lo * dc, sizeof(double) * (staged_hi - staged_lo) * dc,
                                    cudaMemcpyHostToDevice, s_k));
      }
      xc_dev = h->stage + (c0 - staged_lo) * dc;
    } else {
      xc_dev = Xc + c0 * dc;
    }
    xc_of[b] = xc_dev;
    double* mu_dev = (space == DFB_DEVICE && out.mu) ? out.mu + c0 : (b ? h->mu_b : h->mu);
    double* kss_dev = b ? h->kssv_b : h->kssv;
    int8_t* planes = b ? h->Ki8b : h->Ki8;
    const int* abort_count = md.collect ? h->list_count : nullptr;
    DFB_TRY(prof_begin(h, DFB_PROF_KSTAR));
    int fused_digits = 0;
    if (seg_ok)
      DFB_TRY(launch_kstar_seg(h, d_desc, desc, ss.xs, ss.nrm, npad, h->alpha, h->n, xc_dev, mc, dc, m_rows, npad, mean_const,
                               mu_dev, kss_dev, planes, 2 * h->chunk * npad, 2 * npad, 1.0 / i8_colscale(desc), h->cprep,
                               h->mu_part, Mc, &fused_digits, abort_count));
    if (!fused_digits && i8 && h->i8_fuse)
      DFB_TRY(launch_kstar_i8(h, d_desc, desc, ss.xs, ss.nrm, npad, h->alpha, xc_dev, mc, dc, m_rows, h->n, npad,
                              mean_const, mu_dev, kss_dev, planes, 2 * h->chunk * npad, 2 * npad,
                              1.0 / i8_colscale(desc), &fused_digits, abort_count));
    if (!fused_digits) {
      DFB_TRY(launch_kstar(h, d_desc, desc, 0, ss.xs, ss.nrm, npad, h->alpha, xc_dev, mc, dc, m_rows,
                           h->Ks, npad, h->n, npad, mean_const, mu_dev, want_std ? kss_dev : nullptr));
      if (i8)     // K_* = 2^F * digits: |K_*| <= k(x,x) for every supported (stationary, non-negative) kernel
        DFB_TRY(launch_slice_i8(h, h->Ks, npad, m_rows, npad, nullptr, 1.0 / i8_colscale(desc), planes,
                                2 * h->chunk * npad, 2 * npad));
    }
    DFB_TRY(prof_end(h, DFB_PROF_KSTAR, (double)mc));
    if (pipelined) DFB_CUDA_OK(cudaEventRecord(h->ks_k[b], s_k));
    return 0;
  };

  auto stage_g = [&](int64_t ci) -> int {
    const int b = pipelined ? (int)(ci & 1) : 0;
    const int64_t c0 = ci * Mc;
    const int64_t mc = (m - c0 < Mc) ? (m - c0) : Mc;
    const int64_t m_rows = round_up(mc, TILE);
    h->stream = s_g;
    if (pipelined) DFB_CUDA_OK(cudaStreamWaitEvent(s_g, h->ks_k[b], 0));
    const double* xc_dev = xc_of[b];
    double* mu_dev = (space == DFB_DEVICE && out.mu) ? out.mu + c0 : (b ? h->mu_b : h->mu);
    double* kss_dev = b ? h->kssv_b : h->kssv;
    double* sd_dev = (space == DFB_DEVICE && out.sd) ? out.sd + c0 : h->sd;
    double* sc_dev = (space == DFB_DEVICE && out.score) ? out.score + c0
                                                         : ((out.score || md.collect || md.keep_scores) ? h->score : nullptr);
    const int* abort_count = md.collect ? h->list_count : nullptr;
    int small_warps = 0;
    const bool small = want_std && md.allow_small && !md.use_i8 && m <= SMALL_EVAL_M &&
                       (int64_t)((h->n + 7) / 8 * 8) * SMALL_EVAL_M <= (int64_t)nb * Mc;
    if (small) {
      // the padded rows of K_* beyond mc are zero, so an 8-wide pass may run past mc (within the 128-row tile)
      DFB_TRY(prof_begin(h, DFB_PROF_GEMM));
      DFB_TRY(launch_small_sumsq(h, h->W, npad, h->Ks, npad, h->n, (int)mc, h->partial, SMALL_EVAL_M, &small_warps));
      DFB_TRY(prof_end(h, DFB_PROF_GEMM, (double)mc));
    } else if (want_std) {
      GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.A = h->W; g.lda = npad; g.B = h->Ks; g.ldb = npad; g.mode = MODE_SCORE;
      g.n_rb = nb; g.n_cb = (int)(m_rows / TILE); g.K = (int)npad;
      g.partial = h->partial; g.ld_partial = Mc;
      DFB_TRY(prof_begin(h, DFB_PROF_GEMM));
      if (md.use_i8) {
        const double colscale = i8_colscale(desc);
        if (h->i8_impl == 2)
          DFB_TRY(launch_score_i8c2_args(h, h->tmW2, h->tmW3, h->tmW1c, b ? h->tmK2h_b : h->tmK2h, b ? h->tmK3h_b : h->tmK3h,
                                         b ? h->tmK1c_b : h->tmK1c, nb, (int)(m_rows / TILE), (int)npad,
                                         h->partial, Mc, h->rowscale, colscale, abort_count));
        else if (h->i8_impl == 1)
          DFB_TRY(launch_score_i8x2_args(h, h->tmW2, h->tmW3, h->tmK2, h->tmK3, nb, (int)(m_rows / TILE), (int)npad,
                                         h->partial, Mc, h->rowscale, colscale));
        else
          DFB_TRY(launch_score_i8_args(h, h->tmWi8, h->tmKi8, nb, (int)(m_rows / 64), (int)npad, h->partial, Mc,
                                       h->rowscale, colscale));
      } else if (h->gemm_impl == 1 && h->tma_ready) {
        ScoreTmaArgs ta;
        ta.n_rb = g.n_rb; ta.n_cb = g.n_cb; ta.K = g.K; ta.partial = g.partial; ta.ld_partial = g.ld_partial;
        ta.cb_group = h->tma_cb_group;
        DFB_TRY(launch_score_tma(h, h->tmW, h->tmK, ta));
      } else {
        DFB_TRY(launch_gemm(h, g, EPI_SUMSQ, g.n_rb * g.n_cb));
      }
      DFB_TRY(prof_end(h, DFB_PROF_GEMM, (double)mc));
    }
    if (want_std || do_argmax || sc_dev != nullptr) {
      DFB_TRY(prof_begin(h, DFB_PROF_ACQ));
      DFB_TRY(launch_acq(h, acq, mu_dev, h->partial, small ? SMALL_EVAL_M : Mc, small ? small_warps : nb, kss_dev, mc, c0,
                         want_std ? 1 : 0, want_std ? sd_dev : nullptr, sc_dev, do_argmax, md.idx_map,
                         md.collect ? &md.em : nullptr));
      if (md.collect)
        DFB_TRY(launch_collect_shortlist(h, sc_dev, sd_dev, mc, c0, md.em, md.pad, xc_dev, dc));
      DFB_TRY(prof_end(h, DFB_PROF_ACQ, (double)mc));
    }
    if (space == DFB_HOST) {
      if (out.mu)
        DFB_CUDA_OK(cudaMemcpyAsync(out.mu + c0, mu_dev, sizeof(double) * mc, cudaMemcpyDeviceToHost, s_g));
      if (out.sd && want_std)
        DFB_CUDA_OK(cudaMemcpyAsync(out.sd + c0, sd_dev, sizeof(double) * mc, cudaMemcpyDeviceToHost, s_g));
      if (out.score)
        DFB_CUDA_OK(cudaMemcpyAsync(out.score + c0, sc_dev, sizeof(double) * mc, cudaMemcpyDeviceToHost, s_g));
    }
    if (pipelined) DFB_CUDA_OK(cudaEventRecord(h->ks_g[b], s_g));
    if (dbuf) {                                            // last chunk of its batch: the half may be overwritten
      const int64_t bi = c0 / half_rows;
      const int64_t b_hi = (m - bi * half_rows < half_rows) ? m : (bi + 1) * half_rows;
      if (c0 + Mc >= b_hi) DFB_CUDA_OK(cudaEventRecord(h->cp_free[bi & 1], s_g));
    }
    return 0;
  };

  if (!pipelined) {
    for (int64_t ci = 0; ci < n_chunks; ci++) {
      DFB_TRY(stage_k(ci));
      DFB_TRY(stage_g(ci));
    }
  } else {
    // issue order: the contraction of chunk c is enqueued BEFORE the K_* of chunk c+1, so that the block scheduler
    // places the persistent kernel's CTAs first and the K_* CTAs fill the registers / shared memory left over
    DFB_TRY(stage_k(0));
    for (int64_t ci = 0; ci < n_chunks; ci++) {
      DFB_TRY(stage_g(ci));
      if (ci + 1 < n_chunks) DFB_TRY(stage_k(ci + 1));
    }
    DFB_CUDA_OK(cudaEventRecord(h->ks_join, s_g));          // every K stage precedes a G stage: joining G joins both
    DFB_CUDA_OK(cudaStreamWaitEvent(s_user, h->ks_join, 0));
    h->last_overlapped = n_chunks;
  }
  return 0;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_version(void) { return DFB_VERSION; }

const char* dfb_last_error(void) { return g_err; }

int dfb_create(dfb_handle** out, int device) {
  if (out == nullptr) { set_error("out is NULL"); return -1; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device visible (%s): libdfb200 has no CPU fallback", cudaGetErrorString(e));
    return -2;
  }
  if (device < 0 || device >= count) { set_error("device %d out of range (0..%d)", device, count - 1); return -1; }
  DFB_CUDA_OK(cudaSetDevice(device));
  int cc_major = 0, cc_minor = 0;     // attribute queries: cudaGetDeviceProperties costs milliseconds per call
  DFB_CUDA_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  DFB_CUDA_OK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  if (cc_major < 10) {
    set_error("device %d is sm_%d%d; libdfb200 is built for sm_100a (B200) only", device, cc_major, cc_minor);
    return -2;
  }
  dfb_handle* h = new (std::nothrow) dfb_handle();
  if (h == nullptr) { set_error("out of host memory"); return -2; }
  h->device = device;
  const char* impl = getenv("DFB200_GEMM");       // "v1" = cp.async ring, "tma" = TMA + mbarrier ring
  h->gemm_impl = (impl != nullptr && strcmp(impl, "v1") == 0) ? 0 : 1;   // default: TMA ring
  const char* simpl = getenv("DFB200_SCORE");     // "i8" = int8-slice tcgen05 contraction
  h->score_impl = 2;                              // auto
  if (simpl != nullptr && strcmp(simpl, "i8") == 0) h->score_impl = 1;
  if (simpl != nullptr && strcmp(simpl, "fp64") == 0) h->score_impl = 0;
  *out = h;
  return 0;
}

void dfb_destroy(dfb_handle* h) {
  if (h == nullptr) return;
  if (h->fs_hi != nullptr) {
    cudaStreamDestroy(h->fs_hi); cudaStreamDestroy(h->fs_lo);
    cudaEventDestroy(h->fe_fork); cudaEventDestroy(h->fe_panel); cudaEventDestroy(h->fe_rest);
    cudaEventDestroy(h->fe_join_hi); cudaEventDestroy(h->fe_join_lo);
  }
  if (h->cp_stream != nullptr) {
    cudaStreamDestroy(h->cp_stream); cudaEventDestroy(h->cp_fork);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(h->cp_done[i]); cudaEventDestroy(h->cp_free[i]); }
  }
  if (h->ks_stream != nullptr) {
    cudaStreamDestroy(h->ks_stream); cudaStreamDestroy(h->gs_stream); cudaEventDestroy(h->ks_fork); cudaEventDestroy(h->ks_join);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(h->ks_k[i]); cudaEventDestroy(h->ks_g[i]); }
  }
  if (h->prof != nullptr) {
    for (int c = 0; c < PROF_CLASSES; c++)
      if (h->prof[c].created)
        for (int i = 0; i < PROF_RING; i++) { cudaEventDestroy(h->prof[c].start[i]); cudaEventDestroy(h->prof[c].stop[i]); }
    delete[] h->prof;
  }
  delete h;
}

int dfb_set_stream(dfb_handle* h, void* cuda_stream) {
  DFB_TRY(need(h, false, false, false, false, false));
  h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  return 0;
}

size_t dfb_workspace_bytes(int64_t n_max, int32_t n_slots, int64_t chunk) {
  (void)n_slots;
  return carve(nullptr, nullptr, n_max, chunk);
}

int dfb_set_workspace(dfb_handle* h, void* workspace_dev, size_t bytes, int64_t n_max, int64_t chunk) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (workspace_dev == nullptr || n_max < 1) { set_error("bad workspace arguments"); return -1; }
  const size_t want = carve(nullptr, nullptr, n_max, chunk);
  if (bytes < want) { set_error("workspace too small: %zu bytes given, %zu needed", bytes, want); return -1; }
  if ((reinterpret_cast<uintptr_t>(workspace_dev) & 255) != 0) { set_error("workspace must be 256-byte aligned"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  h->ws = static_cast<char*>(workspace_dev);
  h->ws_bytes = bytes;
  carve(h, h->ws, n_max, chunk);
  h->have_train = h->have_post = h->have_w = false;
  h->tr_prepped = h->te_prepped = false;
  if (h->have_kernel)
    DFB_CUDA_OK(cudaMemcpyAsync(h->d_desc_tr, &h->desc_tr, sizeof(dfb_kernel_desc), cudaMemcpyHostToDevice, h->stream));
  if (h->have_test_kernel)
    DFB_CUDA_OK(cudaMemcpyAsync(h->d_desc_te, &h->desc_te, sizeof(dfb_kernel_desc), cudaMemcpyHostToDevice, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dfb_set_kernel(dfb_handle* h, const dfb_kernel_desc* desc) {
  DFB_TRY(need(h, true, false, false, false, false));
  DFB_TRY(check_desc(desc));
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));   // the pageable host copy below must not race
  h->desc_tr = *desc;
  DFB_CUDA_OK(cudaMemcpyAsync(h->d_desc_tr, &h->desc_tr, sizeof(dfb_kernel_desc), cudaMemcpyHostToDevice, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->have_kernel = true;
  h->tr_prepped = false;
  h->have_post = h->have_w = false;
  return 0;
}

int dfb_set_test_kernel(dfb_handle* h, const dfb_kernel_desc* desc) {
  DFB_TRY(need(h, true, false, false, false, false));
  if (desc == nullptr) { h->have_test_kernel = false; h->te_prepped = false; return 0; }
  DFB_TRY(check_desc(desc));
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->desc_te = *desc;
  DFB_CUDA_OK(cudaMemcpyAsync(h->d_desc_te, &h->desc_te, sizeof(dfb_kernel_desc), cudaMemcpyHostToDevice, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->have_test_kernel = true;
  h->te_prepped = false;
  return 0;
}

int dfb_set_train(dfb_handle* h, const double* X_dev, int64_t n, int32_t d, const double* y_centred_dev) {
  DFB_TRY(need(h, true, false, false, false, false));
  if (n < 1 || n > h->n_max) { set_error("n = %lld outside [1, n_max = %lld]", (long long)n, (long long)h->n_max); return -1; }
  if (d < 1 || d > DFB_MAX_SLOTS) { set_error("d = %d outside [1, %d]", d, DFB_MAX_SLOTS); return -1; }
  if (X_dev == nullptr || y_centred_dev == nullptr) { set_error("X / y pointer is NULL"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  h->n = n; h->d = d; h->npad = round_up(n, TILE);
  DFB_CUDA_OK(cudaMemcpyAsync(h->X, X_dev, sizeof(double) * n * d, cudaMemcpyDeviceToDevice, h->stream));
  DFB_TRY(launch_copy_pad(h, y_centred_dev, n, h->yc, h->npad));
  DFB_TRY(launch_fill(h, h->alpha, h->npad, 0.0));
  h->have_train = true;
  h->tr_prepped = h->te_prepped = false;
  h->have_post = h->have_w = false;
  return 0;
}

int dfb_build_posterior(dfb_handle* h, double noise_var, double jitter, int32_t flags, double* lml_out_host) {
  DFB_TRY(need(h, true, true, true, false, false));
  if (h->desc_tr.train_dim != h->d) { set_error("kernel train_dim %d != data dim %d", h->desc_tr.train_dim, h->d); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  const int64_t n = h->n, npad = h->npad;
  const bool with_bottom = (flags != DFB_BUILD_LML_ONLY);
  h->have_post = h->have_w = false;
  DFB_TRY(prof_begin(h, DFB_PROF_BUILD));
  DFB_CUDA_OK(cudaMemsetAsync(h->info, 0, sizeof(int) * 4, h->stream));
  DFB_CUDA_OK(cudaMemsetAsync(h->T, 0, sizeof(double) * (size_t)(2 * npad + TILE) * npad, h->stream));
  DFB_TRY(ensure_train_scaled(h));
  // K(X, X): GP._get_training_kernel_matrix (gp_core.py:149-153)
  DFB_TRY(launch_kstar(h, h->d_desc_tr, h->desc_tr, 1, h->tr.xs, h->tr.nrm, npad, nullptr, h->X, n, h->d,
                       n, h->T, npad, n, npad, 0.0, nullptr, nullptr));
  DFB_TRY(launch_init_tall(h, h->T, n, npad, noise_var + jitter, h->yc, with_bottom ? 1 : 0));
  DFB_TRY(factorise_tall(h, h->T, npad, h->Dinv, h->info, with_bottom));
  const double* Wt = h->T + (size_t)npad * npad;
  const double* v = h->T + (size_t)2 * npad * npad;
  if (with_bottom) DFB_TRY(launch_transpose(h, Wt, h->W, npad));
  if (flags == DFB_BUILD_FULL) DFB_TRY(launch_alpha(h, Wt, v, h->alpha, n, npad));
  DFB_TRY(launch_lml_reduce(h, h->T, h->yc, flags == DFB_BUILD_FULL ? h->alpha : nullptr, v, n, npad, h->red));
  DFB_TRY(prof_end(h, DFB_PROF_BUILD, 1.0));
  double red[3];
  int info = 0;
  DFB_CUDA_OK(cudaMemcpyAsync(red, h->red, sizeof(red), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(&info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->max_diag = h->desc_tr.kss + noise_var;
  if (info != 0) {
    set_error("matrix is not positive definite: non-positive pivot at index %d", info - 1);
    return info;
  }
  h->noise_plus_jitter = noise_var + jitter;
  h->have_post = true;
  h->have_w = with_bottom;
  h->i8_ready = false;
  if (with_bottom && (h->score_impl == 1 || (h->score_impl == 2 && h->n >= 1024))) DFB_TRY(prepare_i8(h));
  h->tma_ready = false;
  if (with_bottom && h->gemm_impl == 1) {
    DFB_TRY(make_tensor_map_2d_f64(&h->tmW, h->W, npad, npad, npad));
    DFB_TRY(make_tensor_map_2d_f64(&h->tmK, h->Ks, h->chunk, npad, npad));
    h->tma_ready = true;
  }
  if (lml_out_host != nullptr) {
    const double quad = (flags == DFB_BUILD_FULL) ? red[1] : red[2];
    *lml_out_host = -0.5 * quad - red[0] - 0.5 * (double)n * log(2.0 * M_PI);
  }
  return 0;
}

int dfb_restore_posterior(dfb_handle* h);

int dfb_extend_posterior(dfb_handle* h, const double* X_new_dev, int64_t q, const double* y_centred_new_dev,
                         int32_t flags, double* lml_out_host) {
  DFB_TRY(need(h, true, true, true, true, true));
  const int32_t build_flags = flags & ~DFB_EXTEND_SAVE;
  if (build_flags != DFB_BUILD_FULL && build_flags != DFB_BUILD_NO_ALPHA) { set_error("dfb_extend_posterior: flags must be DFB_BUILD_FULL or DFB_BUILD_NO_ALPHA (| DFB_EXTEND_SAVE)"); return -1; }
  if (q < 1 || X_new_dev == nullptr || y_centred_new_dev == nullptr) { set_error("bad extend arguments (q = %lld)", (long long)q); return -1; }
  if (h->n + q > h->npad) {
    set_error("dfb_extend_posterior: %lld + %lld points do not fit the padded size %lld of this posterior: rebuild",
              (long long)h->n, (long long)q, (long long)h->npad);
    return -1;
  }
  if (h->ext_saved) { set_error("dfb_extend_posterior: a saved extension is active, call dfb_restore_posterior first"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  const int64_t n0 = h->n, npad = h->npad;
  const int64_t m0 = npad - TILE;
  if (flags & DFB_EXTEND_SAVE) {
    double* s = h->ext_save;
    DFB_CUDA_OK(cudaMemcpyAsync(s, h->T + m0 * npad, sizeof(double) * TILE * npad, cudaMemcpyDeviceToDevice, h->stream));
    DFB_CUDA_OK(cudaMemcpy2DAsync(s + TILE * npad, sizeof(double) * TILE, h->T + npad * npad + m0, sizeof(double) * npad,
                                  sizeof(double) * TILE, (size_t)npad, cudaMemcpyDeviceToDevice, h->stream));
    DFB_CUDA_OK(cudaMemcpyAsync(s + 2 * TILE * npad, h->T + 2 * npad * npad + m0, sizeof(double) * TILE,
                                cudaMemcpyDeviceToDevice, h->stream));
    DFB_CUDA_OK(cudaMemcpyAsync(s + 2 * TILE * npad + TILE, h->alpha, sizeof(double) * npad, cudaMemcpyDeviceToDevice, h->stream));
    h->ext_saved_n = n0;
  }
  DFB_CUDA_OK(cudaMemcpyAsync(h->X + n0 * h->d, X_new_dev, sizeof(double) * q * h->d, cudaMemcpyDeviceToDevice, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(h->yc + n0, y_centred_new_dev, sizeof(double) * q, cudaMemcpyDeviceToDevice, h->stream));
  h->n = n0 + q;
  if (h->n_max < h->n) h->n_max = h->n;
  const int r = replay_last_block(h, build_flags, lml_out_host);
  if (flags & DFB_EXTEND_SAVE) {
    h->ext_saved = true;
    if (r > 0) {                       // not positive definite: put the un-extended posterior back
      char msg[512];
      snprintf(msg, sizeof(msg), "%s", g_err);
      DFB_TRY(dfb_restore_posterior(h));
      set_error("%s", msg);
    }
  }
  return r;
}

int dfb_restore_posterior(dfb_handle* h) {
  DFB_TRY(need(h, true, true, true, false, false));
  if (!h->ext_saved) { set_error("dfb_restore_posterior: nothing saved"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  const int64_t npad = h->npad, m0 = npad - TILE, n0 = h->ext_saved_n;
  const double* s = h->ext_save;
  DFB_CUDA_OK(cudaMemcpyAsync(h->T + m0 * npad, s, sizeof(double) * TILE * npad, cudaMemcpyDeviceToDevice, h->stream));
  DFB_CUDA_OK(cudaMemcpy2DAsync(h->T + npad * npad + m0, sizeof(double) * npad, s + TILE * npad, sizeof(double) * TILE,
                                sizeof(double) * TILE, (size_t)npad, cudaMemcpyDeviceToDevice, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(h->T + 2 * npad * npad + m0, s + 2 * TILE * npad, sizeof(double) * TILE,
                              cudaMemcpyDeviceToDevice, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(h->alpha, s + 2 * TILE * npad + TILE, sizeof(double) * npad, cudaMemcpyDeviceToDevice, h->stream));
  DFB_TRY(launch_fill(h, h->yc + n0, npad - n0, 0.0));           // zero the appended targets
  h->n = n0;
  h->ext_saved = false;
  h->tr_prepped = h->te_prepped = false;
  DFB_TRY(launch_transpose(h, h->T + npad * npad, h->W, npad));
  h->have_post = h->have_w = true;
  h->i8_ready = false;
  if (h->score_impl == 1 || (h->score_impl == 2 && h->n >= 1024)) DFB_TRY(prepare_i8(h));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dfb_lml_gradients(dfb_handle* h, double* out_host, int32_t n_out) {
  DFB_TRY(need(h, true, true, true, true, true));
  const dfb_kernel_desc& desc = h->desc_tr;
  if (desc.n_terms != 1 || desc.n_factors != 1) {
    // the reference's composite kernels inherit Kernel._child_gradient, which raises (kernel.py:123-125)
    set_error("LML gradients are defined for plain SE / Matern kernels only (kernel has %d terms, %d factors)",
              desc.n_terms, desc.n_factors);
    return -3;
  }
  const int D = desc.factors[0].n_dims;
  const int P = 4 + D;
  if (out_host == nullptr || n_out < P) { set_error("lml_gradients: out needs 4 + d = %d entries", P); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  const int64_t npad = h->npad;
  const int nb = (int)(npad / TILE);
  const int64_t n_tiles = (int64_t)nb * (nb + 1) / 2;
  if (n_tiles * P > (int64_t)nb * h->chunk || P > h->chunk) {
    set_error("lml_gradients: scoring chunk %lld too small for %lld tile partials", (long long)h->chunk, (long long)n_tiles);
    return -1;
  }
  DFB_TRY(ensure_train_scaled(h));
  // K^-1 = W^T W row-block stripe by stripe into the K_* chunk buffer (chunk rows x npad), each followed by the
  // fused reduction against the kernel derivatives.  L^-T sits in the middle block of the tall matrix.
  const double* Wt = h->T + (size_t)npad * npad;
  const int stripe = (int)((h->chunk / TILE) < nb ? (h->chunk / TILE) : nb);
  for (int rb0 = 0; rb0 < nb; rb0 += stripe) {
    const int R = (nb - rb0 < stripe) ? nb - rb0 : stripe;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = Wt + (int64_t)rb0 * TILE * npad; g.lda = npad; g.B = Wt; g.ldb = npad; g.D = h->Ks; g.ldd = npad;
    g.alpha = 1.0; g.mode = MODE_GENERIC; g.n_rb = R; g.n_cb = nb; g.K = (int)npad; g.tri = 3; g.lower_only = 1;
    g.rb0 = rb0;
    DFB_TRY(launch_gemm(h, g, EPI_STORE, R * nb));
    DFB_TRY(launch_lml_grad_tiles(h, h->d_desc_tr, h->tr.xs, h->tr.nrm, npad, h->alpha, h->Ks, npad, rb0, R, nb, h->n, P,
                                  h->partial));
  }
  DFB_TRY(launch_lml_grad_reduce(h, h->partial, n_tiles, P, P, h->alpha, h->n, h->score));
  DFB_CUDA_OK(cudaMemcpyAsync(out_host, h->score, sizeof(double) * P, cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  // 1/2 tr(.) (gp_core.py:240); slot 1 still lacks the factor noise_var, slot 2 is sum(alpha) as is
  for (int p = 0; p < P; p++) if (p != 2) out_host[p] *= 0.5;
  return 0;
}

int dfb_get_max_diag(dfb_handle* h, double* out_host) {
  DFB_TRY(need(h, true, true, true, false, false));
  if (out_host == nullptr) { set_error("out is NULL"); return -1; }
  *out_host = h->max_diag;
  return 0;
}

int dfb_get_state(dfb_handle* h, double* L_dev, double* alpha_dev, double* K_dev) {
  DFB_TRY(need(h, true, true, true, true, false));
  DFB_CUDA_OK(cudaSetDevice(h->device));
  if (L_dev) DFB_TRY(launch_extract_lower(h, h->T, h->npad, L_dev, h->n));
  if (alpha_dev) DFB_TRY(launch_copy_pad(h, h->alpha, h->n, alpha_dev, h->n));
  if (K_dev)
    DFB_TRY(launch_kstar(h, h->d_desc_tr, h->desc_tr, 1, h->tr.xs, h->tr.nrm, h->npad, nullptr, h->X, h->n,
                         h->d, h->n, K_dev, h->n, h->n, h->n, 0.0, nullptr, nullptr));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dfb_set_alpha(dfb_handle* h, const double* alpha_dev, int64_t n) {
  DFB_TRY(need(h, true, true, true, false, false));
  if (alpha_dev == nullptr || n < 0 || n > h->n) { set_error("bad alpha arguments"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_TRY(launch_copy_pad(h, alpha_dev, n, h->alpha, h->npad));
  return 0;
}

int dfb_eval(dfb_handle* h, const double* Xc, int64_t m, int32_t dc, int32_t space, double mean_const,
             double* mu, double* sd) {
  DFB_TRY(need(h, true, true, true, true, sd != nullptr));
  if (m < 0 || (m > 0 && (Xc == nullptr || mu == nullptr))) { set_error("bad eval arguments"); return -1; }
  if (m == 0) return 0;
  DFB_CUDA_OK(cudaSetDevice(h->device));
  dfb_acq_desc acq;
  memset(&acq, 0, sizeof(acq));
  acq.kind = DFB_ACQ_MEAN;
  ChunkOut out = {mu, sd, nullptr};
  const dfb_kernel_desc& desc = h->have_test_kernel ? h->desc_te : h->desc_tr;
  ChunkMode md = chunk_mode(sd != nullptr, false, false);
  md.use_i8 = (sd != nullptr) && (h->score_impl == 1) && i8_usable(h, desc);
  md.allow_small = h->small_eval != 0;
  h->last_used_i8 = md.use_i8 ? 1 : 0;
  DFB_TRY(run_chunks(h, acq, Xc, m, dc, space, mean_const, out, md));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dfb_score_argmax(dfb_handle* h, const dfb_acq_desc* acq, const double* Xc, int64_t m, int32_t dc,
                     int32_t space, double mean_const, double* scores, double* best_score_host,
                     int64_t* best_index_host) {
  if (acq == nullptr) { set_error("acq is NULL"); return -1; }
  const bool want_std = (acq->kind != DFB_ACQ_MEAN);
  DFB_TRY(need(h, true, true, true, true, want_std));
  if (m < 1 || Xc == nullptr) { set_error("bad score arguments (m = %lld)", (long long)m); return -1; }
  if (acq->kind < DFB_ACQ_MEAN || acq->kind > DFB_ACQ_TTEI) { set_error("unknown acquisition kind %d", acq->kind); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  ChunkOut out = {nullptr, nullptr, scores};
  const dfb_kernel_desc& desc = h->have_test_kernel ? h->desc_te : h->desc_tr;
  // A caller that asks for the full score vector gets fp64 scores (parity use); the shortlist scheme
  // only guarantees the arg-max, so the int8 pass is reserved for arg-max-only calls unless forced.
  const bool fast = want_std && h->score_impl != 0 && i8_usable(h, desc) &&
                    (scores == nullptr || h->score_impl == 1);
  ChunkMode md = chunk_mode(want_std, true, fast);
  h->last_used_i8 = fast ? 1 : 0;
  h->last_shortlist = 0;
  h->last_selfcheck_violations = 0;
  h->last_selfcheck_ratio = 0.0;
  double bs = 0.0;
  int64_t bi = -1;
  bool need_exact_pass = !fast;
  if (fast) {
    // Pass 1: int8-slice scoring of everything, collecting the shortlist of candidates whose fp64 score could be
    // the maximum under the error model (kernels.cu: i8_score_err / collect_shortlist_kernel).  The pass is void
    // -- and its remaining launches return at once -- as soon as the shortlist overflows (masses of exact ties).
    const double sk = sqrt(desc.kss);
    double scale = sk;                    // natural score scale, for the slack only
    if (acq->kind == DFB_ACQ_UCB) scale = (1.0 + fabs(acq->beta)) * sk + fabs(mean_const);
    else if (acq->kind == DFB_ACQ_PI) scale = 1.0;
    md.collect = true;
    md.em.b2 = i8_sigma2_bound(h, desc);
    md.em.kind = acq->kind;
    md.em.sens = (acq->kind == DFB_ACQ_UCB) ? fabs(acq->beta) : (acq->kind == DFB_ACQ_PI ? 0.25 : 0.4);
    md.pad = 1e-9 * scale;
    DFB_CUDA_OK(cudaMemsetAsync(h->list_count, 0, sizeof(int) * 4, h->stream));
    DFB_TRY(run_chunks(h, *acq, Xc, m, dc, space, mean_const, out, md));
    int count = 0;
    DFB_CUDA_OK(cudaMemcpyAsync(&count, h->list_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (count > SHORTLIST_CAP || count > h->chunk) {
      h->last_shortlist = -1;          // too many candidates within reach of the maximum: exact pass over everything
      need_exact_pass = true;
    } else {
      // Pass 2: exact fp64 (DMMA) re-score of the shortlist; indices map back to the caller's rows.  Then the
      // self-check: int8 vs fp64 score of every listed candidate against its allowance.
      h->last_shortlist = count;
      ChunkMode ex = chunk_mode(want_std, true, false, h->list_idx);
      ex.keep_scores = true;
      ChunkOut none = {nullptr, nullptr, nullptr};
      DFB_TRY(run_chunks(h, *acq, h->list_X, count, dc, DFB_DEVICE, mean_const, none, ex));
      DFB_TRY(launch_selfcheck(h, h->score, count));
      int chk[2] = {0, 0};
      DFB_CUDA_OK(cudaMemcpyAsync(chk, h->list_count + 1, sizeof(chk), cudaMemcpyDeviceToHost, h->stream));
      DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
      h->last_selfcheck_violations = chk[0];
      h->last_selfcheck_ratio = (double)chk[1] * 1e-6;
      if (chk[0] > 0) need_exact_pass = true;      // the error model failed on a candidate that matters: fp64
    }
  }
  if (need_exact_pass) {
    ChunkMode ex = chunk_mode(want_std, true, false);
    DFB_TRY(run_chunks(h, *acq, Xc, m, dc, space, mean_const, out, ex));
  }
  DFB_CUDA_OK(cudaMemcpyAsync(&bs, h->best_score, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(&bi, h->best_index, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (best_score_host) *best_score_host = bs;
  if (best_index_host) *best_index_host = bi;
  return 0;
}

int dfb_moo_score_argmax(dfb_handle* h, const dfb_moo_desc* desc, const double* const* a_dev,
                         const double* const* b_dev, int64_t m, double* scores_dev,
                         double* best_score_host, int64_t* best_index_host) {
  DFB_TRY(need(h, true, false, false, false, false));
  if (desc == nullptr || a_dev == nullptr || m < 1) { set_error("bad moo arguments (m = %lld)", (long long)m); return -1; }
  if (desc->kind < DFB_MOO_LIN_UCB || desc->kind > DFB_MOO_TCH_VAL) { set_error("unknown scalarisation kind %d", desc->kind); return -1; }
  if (desc->n_obj < 1 || desc->n_obj > DFB_MOO_MAX_OBJ) { set_error("n_obj = %d outside [1, %d]", desc->n_obj, DFB_MOO_MAX_OBJ); return -1; }
  const bool ucb = desc->kind == DFB_MOO_LIN_UCB || desc->kind == DFB_MOO_TCH_UCB;
  for (int k = 0; k < desc->n_obj; k++)
    if (a_dev[k] == nullptr || (ucb && (b_dev == nullptr || b_dev[k] == nullptr))) { set_error("objective %d: NULL vector", k); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_TRY(launch_reset_best(h));
  DFB_TRY(launch_moo(h, *desc, a_dev, ucb ? b_dev : nullptr, m, scores_dev));
  double bs = 0.0;
  int64_t bi = -1;
  DFB_CUDA_OK(cudaMemcpyAsync(&bs, h->best_score, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(&bi, h->best_index, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (best_score_host) *best_score_host = bs;
  if (best_index_host) *best_index_host = bi;
  return 0;
}

int dfb_kernel_matrix(dfb_handle* h, const dfb_kernel_desc* desc, const double* X1_dev, int64_t n1,
                      int32_t d1, const double* X2_dev, int64_t n2, int32_t d2, double* K_dev) {
  DFB_TRY(need(h, true, false, false, false, false));
  DFB_TRY(check_desc(desc));
  if (n1 < 1 || n2 < 1 || X1_dev == nullptr || X2_dev == nullptr || K_dev == nullptr) { set_error("bad kernel_matrix arguments"); return -1; }
  if (d1 != d2 || d1 != desc->train_dim) { set_error("kernel_matrix: dims %d, %d vs kernel train_dim %d", d1, d2, desc->train_dim); return -1; }
  const int64_t np2 = round_up(n2, TILE);
  if (np2 > h->npad_max) { set_error("kernel_matrix: n2 = %lld exceeds the workspace (n_max = %lld)", (long long)n2, (long long)h->n_max); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->desc_tmp = *desc;
  DFB_CUDA_OK(cudaMemcpyAsync(h->d_desc_tmp, &h->desc_tmp, sizeof(dfb_kernel_desc), cudaMemcpyHostToDevice, h->stream));
  h->te_prepped = false;   // the test-kernel scaled set is used as scratch
  DFB_TRY(launch_prep_scaled(h, h->d_desc_tmp, 1, X2_dev, n2, d2, h->te.xs, h->te.nrm, np2));
  DFB_TRY(launch_kstar(h, h->d_desc_tmp, h->desc_tmp, 1, h->te.xs, h->te.nrm, np2, nullptr, X1_dev, n1, d1,
                       n1, K_dev, n2, n2, n2, 0.0, nullptr, nullptr));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

// mu and the padded posterior covariance (h->ts_Cov, ld = mbp) of one block of m candidates:
// K_* -> V^T = K_* W^T -> Cov = K** - V^T V    (gp_core.py:173-181)
static int posterior_covariance(dfb_handle* h, const double* Xc_dev, int64_t m, int32_t dc,
                                double mean_const, int64_t* mbp_out, bool lower_only) {
  const dfb_kernel_desc& desc = h->have_test_kernel ? h->desc_te : h->desc_tr;
  const dfb_kernel_desc* d_desc = h->have_test_kernel ? h->d_desc_te : h->d_desc_tr;
  const ScaledSet& ss = h->have_test_kernel ? h->te : h->tr;
  if (h->ts_ws == nullptr) { set_error("no Thompson-sampling workspace: call dfb_set_ts_workspace first"); return -1; }
  if (dc != desc.cand_dim) { set_error("candidates have %d columns, the kernel descriptor expects %d", dc, desc.cand_dim); return -1; }
  const int64_t mbp = round_up(m, TILE);
  if (m < 1 || mbp > h->ts_mb || mbp > h->chunk) { set_error("block of %lld candidates exceeds the TS workspace (%lld) / chunk (%lld)", (long long)m, (long long)h->ts_mb, (long long)h->chunk); return -1; }
  DFB_TRY(ensure_train_scaled(h));
  DFB_TRY(ensure_test_scaled(h));
  const int64_t npad = h->npad;
  const int nb = (int)(npad / TILE), mbb = (int)(mbp / TILE);
  // K_* rows (zero rows beyond m) and mu
  DFB_TRY(launch_kstar(h, d_desc, desc, 0, ss.xs, ss.nrm, npad, h->alpha, Xc_dev, m, dc, mbp, h->Ks, npad,
                       h->n, npad, mean_const, h->ts_mu, nullptr));
  // V^T[a][i] = sum_{k <= i} K_*[a][k] W[i][k]
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = h->Ks; g.lda = npad; g.B = h->W; g.ldb = npad; g.D = h->ts_Vt; g.ldd = npad; g.alpha = 1.0;
  g.mode = MODE_GENERIC; g.n_rb = mbb; g.n_cb = nb; g.K = (int)npad; g.tri = 2;
  DFB_TRY(launch_gemm(h, g, EPI_STORE, mbb * nb));
  // K** = kernel(X_test, X_test) (gp_core.py:179) with the candidates as both sides
  DFB_TRY(launch_prep_scaled(h, d_desc, 0, Xc_dev, m, dc, h->ts_cxs, h->ts_cnrm, mbp));
  DFB_TRY(launch_kstar(h, d_desc, desc, 0, h->ts_cxs, h->ts_cnrm, mbp, nullptr, Xc_dev, m, dc, mbp, h->ts_Cov,
                       mbp, m, mbp, 0.0, nullptr, nullptr));
  // Cov = K** - V^T V
  memset(&g, 0, sizeof(g));
  g.A = h->ts_Vt; g.lda = npad; g.B = h->ts_Vt; g.ldb = npad; g.C = h->ts_Cov; g.ldc = mbp;
  g.D = h->ts_Cov; g.ldd = mbp; g.alpha = -1.0; g.mode = MODE_GENERIC; g.n_rb = mbb; g.n_cb = mbb;
  g.K = (int)npad;
  g.lower_only = lower_only ? 1 : 0;     // the factorisation that follows reads the lower triangle only
  DFB_TRY(launch_gemm(h, g, EPI_STORE, mbb * mbb));
  *mbp_out = mbp;
  return 0;
}

size_t dfb_ts_workspace_bytes(int64_t n_max, int64_t mb) { return carve_ts(nullptr, nullptr, n_max, mb); }

int dfb_set_ts_workspace(dfb_handle* h, void* workspace_dev, size_t bytes, int64_t mb) {
  DFB_TRY(need(h, true, false, false, false, false));
  if (workspace_dev == nullptr || mb < 1) { set_error("bad TS workspace arguments"); return -1; }
  const size_t want = carve_ts(nullptr, nullptr, h->n_max, mb);
  if (bytes < want) { set_error("TS workspace too small: %zu bytes given, %zu needed", bytes, want); return -1; }
  if ((reinterpret_cast<uintptr_t>(workspace_dev) & 255) != 0) { set_error("TS workspace must be 256-byte aligned"); return -1; }
  h->ts_ws = static_cast<char*>(workspace_dev);
  carve_ts(h, h->ts_ws, h->n_max, mb);
  return 0;
}

int dfb_eval_covar(dfb_handle* h, const double* Xc_dev, int64_t m, int32_t dc, double mean_const,
                   double* mu_dev, double* covar_dev) {
  DFB_TRY(need(h, true, true, true, true, true));
  if (Xc_dev == nullptr || mu_dev == nullptr || covar_dev == nullptr) { set_error("bad eval_covar arguments"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  int64_t mbp = 0;
  DFB_TRY(posterior_covariance(h, Xc_dev, m, dc, mean_const, &mbp, false));
  DFB_TRY(launch_copy_pad(h, h->ts_mu, m, mu_dev, m));
  DFB_TRY(launch_copy_rows(h, h->ts_Cov, mbp, covar_dev, m, m, m));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dfb_ts_draws(dfb_handle* h, const double* Xc_dev, int64_t m, int32_t dc, double mean_const,
                 const double* Ut_dev, int32_t S, double jitter, double* samples_dev, double* mu_dev,
                 double* max_diag_host) {
  DFB_TRY(need(h, true, true, true, true, true));
  if (Xc_dev == nullptr || Ut_dev == nullptr || samples_dev == nullptr || S < 1 || S > 256) { set_error("bad ts_draws arguments (1 <= S <= 256)"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  int64_t mbp = 0;
  DFB_TRY(posterior_covariance(h, Xc_dev, m, dc, mean_const, &mbp, true));
  DFB_TRY(launch_diag_max(h, h->ts_Cov, mbp, m, h->ts_red));
  // stable_cholesky(K) (general_utils.py:224-229): factorise Cov + jitter I (identity padding)
  DFB_CUDA_OK(cudaMemsetAsync(h->ts_info, 0, sizeof(int) * 4, h->stream));
  DFB_CUDA_OK(cudaMemsetAsync(h->ts_T, 0, sizeof(double) * (size_t)(2 * mbp + TILE) * mbp, h->stream));
  DFB_TRY(launch_copy_rows(h, h->ts_Cov, mbp, h->ts_T, mbp, mbp, mbp));
  DFB_TRY(launch_set_diag(h, h->ts_T, mbp, 0, m, jitter, 1));
  DFB_TRY(launch_set_diag(h, h->ts_T, mbp, m, mbp, 1.0, 0));
  DFB_TRY(factorise_tall(h, h->ts_T, mbp, h->Dinv, h->ts_info, false));
  // samples^T = L_post U: samples[s][a] = sum_{b <= a} U^T[s][b] L[a][b]
  const int64_t Sp = round_up(S, TILE);
  DFB_CUDA_OK(cudaMemsetAsync(h->ts_Ut, 0, sizeof(double) * (size_t)Sp * mbp, h->stream));
  DFB_TRY(launch_copy_rows(h, Ut_dev, m, h->ts_Ut, mbp, S, m));
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = h->ts_Ut; g.lda = mbp; g.B = h->ts_T; g.ldb = mbp; g.D = h->ts_Sm; g.ldd = mbp; g.alpha = 1.0;
  g.mode = MODE_GENERIC; g.n_rb = (int)(Sp / TILE); g.n_cb = (int)(mbp / TILE); g.K = (int)mbp; g.tri = 2;
  g.info = h->ts_info;
  DFB_TRY(launch_gemm(h, g, EPI_STORE, g.n_rb * g.n_cb));
  DFB_TRY(launch_add_row_vector(h, h->ts_Sm, mbp, S, m, h->ts_mu));
  DFB_TRY(launch_copy_rows(h, h->ts_Sm, mbp, samples_dev, m, S, m));
  if (mu_dev != nullptr) DFB_TRY(launch_copy_pad(h, h->ts_mu, m, mu_dev, m));
  double mx = 0.0;
  int info = 0;
  DFB_CUDA_OK(cudaMemcpyAsync(&mx, h->ts_red, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaMemcpyAsync(&info, h->ts_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DFB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (max_diag_host != nullptr) *max_diag_host = mx;
  if (info != 0) { set_error("posterior covariance is not positive definite at jitter %g (pivot %d)", jitter, info - 1); return info; }
  return 0;
}

int dfb_fill_rng(dfb_handle* h, uint64_t seed, int64_t col0, int32_t S, int64_t m, int32_t what, double* out_dev) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (out_dev == nullptr || S < 1 || m < 1 || col0 < 0 || (what != DFB_RNG_NORMAL && what != DFB_RNG_UNIFORM)) { set_error("bad fill_rng arguments"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  return launch_fill_rng(h, seed, col0, S, m, what, out_dev);
}

int dfb_fill_candidates(dfb_handle* h, uint64_t seed, int64_t row0, int64_t m, int32_t d, const double* lo_host,
                        const double* hi_host, double* out_dev) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (out_dev == nullptr || lo_host == nullptr || hi_host == nullptr || m < 1 || row0 < 0 || d < 1 || d > DFB_MAX_SLOTS) {
    set_error("bad fill_candidates arguments (m = %lld, d = %d)", (long long)m, d);
    return -1;
  }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  return launch_fill_candidates(h, seed, row0, m, d, lo_host, hi_host, out_dev);
}

int dfb_ts_argmax(dfb_handle* h, const double* samples_dev, int64_t ld, int32_t S, int64_t m, int64_t idx_base,
                  int32_t reset, double* best_dev, int64_t* index_dev) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (samples_dev == nullptr || best_dev == nullptr || index_dev == nullptr || S < 1 || m < 1 || ld < m) { set_error("bad ts_argmax arguments"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  return launch_ts_argmax(h, samples_dev, ld, S, m, idx_base, reset, best_dev, index_dev);
}

int64_t dfb_launch_count(dfb_handle* h) { return h ? h->launches : 0; }

int dfb_debug_trace(void* buf_dev, int64_t cap_records) { return debug_set_trace(buf_dev, (long long)cap_records); }

int dfb_query(dfb_handle* h, const char* name, double* out) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (name == nullptr || out == nullptr) { set_error("bad query arguments"); return -1; }
  const dfb_kernel_desc& desc = h->have_test_kernel ? h->desc_te : h->desc_tr;
  if (strcmp(name, "i8_sigma2_bound") == 0) { *out = h->i8_ready ? i8_sigma2_bound(h, desc) : -1.0; return 0; }
  if (strcmp(name, "last_used_i8") == 0) { *out = (double)h->last_used_i8; return 0; }
  if (strcmp(name, "last_shortlist") == 0) { *out = (double)h->last_shortlist; return 0; }
  if (strcmp(name, "last_selfcheck_violations") == 0) { *out = (double)h->last_selfcheck_violations; return 0; }
  if (strcmp(name, "last_selfcheck_ratio") == 0) { *out = h->last_selfcheck_ratio; return 0; }
  if (strcmp(name, "chunk") == 0) { *out = (double)h->chunk; return 0; }
  if (strcmp(name, "last_c2_group") == 0) { *out = (double)h->last_c2_group; return 0; }
  if (strcmp(name, "last_overlapped") == 0) { *out = (double)h->last_overlapped; return 0; }     // chunks of the last PIPELINED pass
  if (strcmp(name, "i8_bound_limit") == 0) { *out = I8_BOUND_LIMIT; return 0; }
  if (strcmp(name, "score_impl") == 0) { *out = (double)h->score_impl; return 0; }
  if (strcmp(name, "i8_ready") == 0) { *out = h->i8_ready ? 1.0 : 0.0; return 0; }
  if (strcmp(name, "i8_impl") == 0) { *out = (double)h->i8_impl; return 0; }
  if (strcmp(name, "i8_radix256") == 0) { *out = (double)h->i8_radix256; return 0; }
  set_error("unknown query '%s'", name);
  return -1;
}

int dfb_set_option(dfb_handle* h, const char* name, int64_t value) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (name == nullptr) { set_error("option name is NULL"); return -1; }
  if (strcmp(name, "gemm_impl") == 0) {
    if (value != 0 && value != 1) { set_error("gemm_impl must be 0 (cp.async) or 1 (TMA)"); return -1; }
    h->gemm_impl = (int)value;
    h->tma_ready = false;
    if (value == 1 && h->have_post && h->have_w) {
      DFB_CUDA_OK(cudaSetDevice(h->device));
      DFB_TRY(make_tensor_map_2d_f64(&h->tmW, h->W, h->npad, h->npad, h->npad));
      DFB_TRY(make_tensor_map_2d_f64(&h->tmK, h->Ks, h->chunk, h->npad, h->npad));
      h->tma_ready = true;
    }
    return 0;
  }
  if (strcmp(name, "kstar_fast") == 0) { h->kstar_fast = value ? 1 : 0; return 0; }
  if (strcmp(name, "lookahead") == 0) { h->lookahead = value ? 1 : 0; return 0; }
  if (strcmp(name, "small_eval") == 0) { h->small_eval = value ? 1 : 0; return 0; }
  if (strcmp(name, "i8_ts") == 0) { h->i8_ts = value ? 1 : 0; return 0; }
  if (strcmp(name, "i8_fuse") == 0) { h->i8_fuse = value ? 1 : 0; return 0; }
  if (strcmp(name, "kstar_seg") == 0) { h->kstar_seg = value ? 1 : 0; return 0; }
  if (strcmp(name, "kstar_rows64") == 0) { h->kstar_rows64 = value ? 1 : 0; return 0; }
  if (strcmp(name, "kstar_overlap") == 0) { h->kstar_overlap = value ? 1 : 0; return 0; }
  if (strcmp(name, "i8_unguarded") == 0) { h->i8_unguarded = value ? 1 : 0; return 0; }
  if (strcmp(name, "i8_impl") == 0) {
    if (value < 0 || value > 2) { set_error("i8_impl must be 0 (N=64, one pass), 1 (N=128, two passes) or 2 (CTA pairs)"); return -1; }
    h->i8_impl = (int)value;
    if (h->i8_ready) { DFB_CUDA_OK(cudaSetDevice(h->device)); DFB_TRY(prepare_i8(h)); }   // re-slice W in the new layout
    return 0;
  }
  if (strcmp(name, "i8_radix") == 0) {
    if (value < -1 || value > 1) { set_error("i8_radix must be -1 (auto), 0 (radix 128) or 1 (radix 256)"); return -1; }
    h->i8_radix_opt = (int)value;
    if (h->i8_ready) { DFB_CUDA_OK(cudaSetDevice(h->device)); DFB_TRY(prepare_i8(h)); }
    return 0;
  }
  if (strcmp(name, "tma_cb_group") == 0 && value >= 1) { h->tma_cb_group = (int)value; return 0; }
  if (strcmp(name, "i8_cb_group") == 0 && value >= 1) { h->i8_cb_group = (int)value; return 0; }
  if (strcmp(name, "i8_c2_group") == 0 && value >= 0) { h->i8_c2_group = (int)value; return 0; }
  if (strcmp(name, "i8_l2_hint") == 0 && value >= 0 && value <= 2) { h->i8_l2_hint = (int)value; return 0; }
  if (strcmp(name, "score_impl") == 0) {
    if (value < 0 || value > 2) { set_error("score_impl must be 0 (fp64 DMMA), 1 (int8-slice tcgen05) or 2 (auto)"); return -1; }
    h->score_impl = (int)value;
    h->i8_ready = false;
    if ((value == 1 || (value == 2 && h->n >= 1024)) && h->have_post && h->have_w) {
      DFB_CUDA_OK(cudaSetDevice(h->device));
      DFB_TRY(prepare_i8(h));
    }
    return 0;
  }
  set_error("unknown option '%s'", name);
  return -1;
}

int dfb_profile_enable(dfb_handle* h, int on) {
  DFB_TRY(need(h, false, false, false, false, false));
  DFB_CUDA_OK(cudaSetDevice(h->device));
  if (on && h->prof == nullptr) {
    h->prof = new (std::nothrow) ProfClass[PROF_CLASSES];
    if (h->prof == nullptr) { set_error("out of host memory"); return -2; }
  }
  h->prof_on = (on != 0);
  return 0;
}

int dfb_profile_read(dfb_handle* h, int cls, double* ms_total, int64_t* launches, double* units) {
  DFB_TRY(need(h, false, false, false, false, false));
  if (cls < 0 || cls >= PROF_CLASSES || h->prof == nullptr) { set_error("profiling not enabled / bad class"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(h->device));
  DFB_TRY(prof_flush(h, cls));
  ProfClass& pc = h->prof[cls];
  if (ms_total) *ms_total = pc.acc_ms;
  if (launches) *launches = pc.acc_launches;
  if (units) *units = pc.acc_units;
  pc.acc_ms = 0.0; pc.acc_units = 0.0; pc.acc_launches = 0;
  return 0;
}

}  // extern "C"
