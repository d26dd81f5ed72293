// Host-side launchers of the CUDA kernels in kernels.cu (one stream: h->stream).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace dfb {

int launch_prep_scaled(dfb_handle* h, const dfb_kernel_desc* d_desc, int use_train_coords,
                       const double* X, int64_t n, int d, double* xs, double* nrm, int64_t npad);
int launch_kstar(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc,
                 int cand_uses_train_coords, const double* xsT, const double* nrmT, int64_t npad_tr,
                 const double* alpha, const double* Xc, int64_t m, int dc, int64_t m_rows, double* Ks,
                 int64_t ldk, int64_t n_valid, int64_t n_write, double mean_const, double* mu,
                 double* kss_out);
int launch_kstar_i8(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc,
                    const double* xsT, const double* nrmT, int64_t npad_tr, const double* alpha,
                    const double* Xc, int64_t m, int dc, int64_t m_rows, int64_t n_valid, int64_t n_write,
                    double mean_const, double* mu, double* kss_out, void* planes, int64_t plane_bytes,
                    int64_t row_bytes, double inv_colscale, int* emitted_i8, const int* abort_count = nullptr);
// second-generation digit path (kernels.cu: kstar_seg_kernel)
int launch_kstar_seg(dfb_handle* h, const dfb_kernel_desc* d_desc, const dfb_kernel_desc& desc, const double* xsT,
                     const double* nrm, int64_t npad_tr, const double* alpha, int64_t n_valid, const double* Xc, int64_t m,
                     int dc, int64_t m_rows, int64_t n_write, double mean_const, double* mu, double* kss_out, void* planes,
                     int64_t plane_bytes, int64_t row_bytes, double inv_colscale, double* cprep, double* mu_part,
                     int64_t ld_mu, int* emitted, const int* abort_count = nullptr);
int launch_init_tall(dfb_handle* h, double* T, int64_t n, int64_t npad, double diag_add,
                     const double* yc, int with_bottom);
int launch_chol_diag(dfb_handle* h, double* T, int64_t ld, int step, double* Dinv, int* info);
int launch_transpose(dfb_handle* h, const double* src, double* dst, int64_t n);
int launch_alpha(dfb_handle* h, const double* Wt, const double* v, double* alpha, int64_t n,
                 int64_t npad);
int launch_lml_reduce(dfb_handle* h, const double* T, const double* yc, const double* alpha,
                      const double* v, int64_t n, int64_t npad, double* out);
int launch_extract_lower(dfb_handle* h, const double* T, int64_t npad, double* L, int64_t n);
int launch_copy_pad(dfb_handle* h, const double* src, int64_t n_src, double* dst, int64_t n_dst);
int launch_copy_rows(dfb_handle* h, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                     int64_t rows, int64_t cols);
// Error model of the int8-slice scoring pass handed to the acquisition / shortlist kernels (kernels.cu:
// i8_score_err): |sigma^2_int8 - sigma^2_fp64| <= b2; sens = |beta| (UCB), 0.4 (EI, TTEI), 0.25 (PI).
struct I8ErrModel {
  double b2;       // a-priori bound on |d sigma^2|; 0 = no int8 pass (no lower-bound tracking)
  double sens;
  int kind;        // DFB_ACQ_*
};
constexpr int SHORTLIST_CAP = 4096;
int launch_acq(dfb_handle* h, const dfb_acq_desc& acq, const double* mu, const double* partial,
               int64_t ld_partial, int nrb, const double* kss, int64_t m, int64_t idx_base,
               int want_std, double* sd_out, double* score_out, bool do_argmax,
               const int64_t* idx_map = nullptr, const I8ErrModel* em = nullptr);
int launch_collect_shortlist(dfb_handle* h, const double* score, const double* sd, int64_t mc,
                             int64_t idx_base, const I8ErrModel& em, double pad, const double* Xc, int dc);
int launch_selfcheck(dfb_handle* h, const double* s64, int count);
int launch_vec_max(dfb_handle* h, const double* v, int64_t n, double* out);
int launch_reset_best(dfb_handle* h);
int launch_fill_rng(dfb_handle* h, uint64_t seed, int64_t col0, int S, int64_t m, int what, double* out);
int launch_fill_candidates(dfb_handle* h, uint64_t seed, int64_t row0, int64_t m, int d, const double* lo,
                           const double* hi, double* out);
int launch_ts_argmax(dfb_handle* h, const double* samples, int64_t ld, int S, int64_t m, int64_t idx_base, int reset,
                     double* best, int64_t* index);
int launch_small_sumsq(dfb_handle* h, const double* W, int64_t ldw, const double* Ks, int64_t ldk, int64_t n_rows,
                       int m, double* part, int64_t ld_part, int* n_warps_out);
int launch_moo(dfb_handle* h, const dfb_moo_desc& d, const double* const* a, const double* const* b, int64_t m,
               double* scores);
int launch_add_row_vector(dfb_handle* h, double* M, int64_t ld, int64_t rows, int64_t cols,
                          const double* v);
int launch_diag_max(dfb_handle* h, const double* M, int64_t ld, int64_t n, double* out);
int launch_fill(dfb_handle* h, double* p, int64_t n, double v);
int launch_set_diag(dfb_handle* h, double* M, int64_t ld, int64_t from, int64_t to, double v, int add);
int debug_set_trace(void* buf, long long cap_records);
// LML gradients: reduction of (alpha alpha^T - K^-1) o dK/dparam over lower tiles of row blocks [rb0, rb0 + n_rb)
int launch_lml_grad_tiles(dfb_handle* h, const dfb_kernel_desc* d_desc, const double* xs, const double* nrm, int64_t npad,
                          const double* alpha, const double* Kinv, int64_t ldk, int rb0, int n_rb, int nb, int64_t n,
                          int pstride, double* partial);
int launch_lml_grad_reduce(dfb_handle* h, const double* partial, int64_t n_tiles, int pstride, int n_out,
                           const double* alpha, int64_t n, double* out);

}  // namespace dfb
