/*
 * dfb200.h -- C-ABI of libdfb200.so: the B200 (sm_100a) GP-BO inner loop behind Dragonfly's
 * Kernel / GP / gpb_acquisitions surfaces.
 *
 * The reference (dragonfly-opt 0.1.7) is pure Python + NumPy/SciPy and has NO C/FFI boundary on this
 * path; the entry points below are what a ctypes binding placed at the reference's
 * three Python-level extension points would call.  Each one cites the reference function whose
 * numeric body it replaces (paths relative to the reference tree).  INTEGRATION.md shows the
 * reference-side ctypes stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every function returns int: 0 = ok; > 0 = LAPACK-style info (dfb_build_posterior: the 1-based
 *     index of the first non-positive pivot, i.e. np.linalg.LinAlgError in the reference -- the
 *     HOST then walks the reference's jitter ladder, general_utils.py:183-203); < 0 = argument or
 *     CUDA error, message in dfb_last_error().
 *   - all floating point data is IEEE fp64, row-major.
 *   - pointers named *_dev are device pointers on the handle's device (e.g. torch.Tensor.data_ptr());
 *     pointers with a `space` argument are host (DFB_HOST) or device (DFB_DEVICE) pointers; host
 *     buffers are copied inside the call on the handle's stream (pinned memory recommended).
 *   - the library allocates NO large device memory: the caller sizes a workspace with
 *     dfb_workspace_bytes() and hands it over with dfb_set_workspace().
 *   - a handle is not thread-safe; work is issued on the stream given to dfb_set_stream()
 *     (default: the legacy default stream) and calls that return host scalars synchronise it.
 *   - there is no CPU fallback: without a CUDA device dfb_create() fails.
 */
#ifndef DFB200_H_
#define DFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_VERSION 100

/* ---- memory spaces ------------------------------------------------------------------------- */
#define DFB_DEVICE 0
#define DFB_HOST   1

/* ---- kernel descriptor ------------------------------------------------------------------------
 * Canonical sum-of-products form of the kernels on the hot path (dragonfly/gp/kernel.py):
 *
 *     k(x, y) = post_scale * SUM_t ( ((pre_scale_t * b_{t,0}) * b_{t,1}) * ... )
 *
 * where every base factor b is an SE or half-integer Matern kernel on a subset of coordinates,
 * evaluated exactly in the reference's operation order:
 *     x~ = x[coords] / bandwidths                                   kernel.py:179-181, 255-257
 *     D2 = max(0, (|y~|^2 + |x~|^2) - 2 x~.y~)                      general_utils.py:58-70
 *     SE      b = scale * exp(-D2 / 2)                              kernel.py:171-177
 *     Matern  r = sqrt(D2); m = s8 * r; u = SUM_i coeffs[i] * m^(p-i);
 *             u *= gamma_ratio * exp(-s2 * r); b = scale * u        kernel.py:259-270, 292-299
 *             (scale here = hyperparams['scale'] * norm_constant, formed on the host)
 *   SEKernel / MaternKernel      : 1 term, 1 factor, pre_scale 1, post_scale 1
 *   AdditiveKernel               : G terms of 1 factor, post_scale = outer scale   kernel.py:484-494
 *   CoordinateProductKernel (MF) : 1 term of F factors, pre_scale = outer scale    kernel.py:573-584
 * A slot is one (coordinate, bandwidth) pair of one factor; the train and candidate matrices may
 * use different column indices for the same slot (Add-UCB scores d_j-column candidates against
 * columns g_j of the training matrix, gpb_acquisitions.py:160-168).
 */
#define DFB_MAX_FACTORS   48
#define DFB_MAX_TERMS     48
#define DFB_MAX_SLOTS     128
#define DFB_MAX_MATERN_P  3

#define DFB_BASE_SE      0
#define DFB_BASE_MATERN  1

typedef struct dfb_factor_desc {
  int32_t kind;          /* DFB_BASE_SE | DFB_BASE_MATERN */
  int32_t p;             /* Matern: nu = p + 1/2 (0 <= p <= DFB_MAX_MATERN_P) */
  int32_t n_dims;        /* number of slots of this factor */
  int32_t slot_off;      /* first slot */
  double  scale;         /* SE: scale;  Matern: scale * norm_constant */
  double  s8;            /* sqrt(8 nu) */
  double  s2;            /* sqrt(2 nu) */
  double  gamma_ratio;   /* Gamma(p+1) / Gamma(2p+1) */
  double  coeffs[DFB_MAX_MATERN_P + 1];  /* (p+i)! / (i! (p-i)!) */
} dfb_factor_desc;

typedef struct dfb_kernel_desc {
  int32_t n_terms;
  int32_t n_factors;
  int32_t n_slots;
  int32_t train_dim;     /* columns of the training matrix */
  int32_t cand_dim;      /* columns of the candidate matrix */
  int32_t reserved;
  double  post_scale;
  double  kss;           /* k(x, x) for any x (all supported kernels are stationary) */
  int32_t term_first_factor[DFB_MAX_TERMS + 1];
  double  term_pre_scale[DFB_MAX_TERMS];
  dfb_factor_desc factors[DFB_MAX_FACTORS];
  int32_t slot_train_coord[DFB_MAX_SLOTS];
  int32_t slot_cand_coord[DFB_MAX_SLOTS];
  double  slot_bandwidth[DFB_MAX_SLOTS];
} dfb_kernel_desc;

/* ---- acquisition descriptor --------------------------------- dragonfly/opt/gpb_acquisitions.py */
#define DFB_ACQ_MEAN  0   /* score = mu                                   (uncert_form 'none')    */
#define DFB_ACQ_UCB   1   /* mu + beta * sigma                            :215-222, add_ucb :176  */
#define DFB_ACQ_EI    2   /* sigma (z Phi(z) + phi(z)), z=(mu-best)/sigma :247-260                */
#define DFB_ACQ_PI    3   /* Phi((mu - best)/sigma)                       :230-238                */
#define DFB_ACQ_TTEI  4   /* EI against a reference arm (ref_mean, ref_std) :269-279              */

typedef struct dfb_acq_desc {
  int32_t kind;
  int32_t reserved;
  double  beta;        /* UCB: beta_th */
  double  best;        /* EI / PI: curr_max_val */
  double  ref_mean;    /* TTEI */
  double  ref_std;     /* TTEI */
} dfb_acq_desc;

/* ---- multi-objective scalarisation --------- dragonfly/opt/multiobjective_gpb_acquisitions.py */
#define DFB_MOO_MAX_OBJ 8
#define DFB_MOO_LIN_UCB 0   /* sum_k w_k mu_k + beta sqrt(sum_k w_k^2 sd_k^2)                     :79-91  */
#define DFB_MOO_TCH_UCB 1   /* min_k (mu_k + beta sqrt(sd_k) - ref_k) / w_k  (sqrt of the std, as written) :94-107 */
#define DFB_MOO_LIN_VAL 2   /* sum_k w_k v_k,  v = one posterior sample per objective (lin_ts)   :19-41  */
#define DFB_MOO_TCH_VAL 3   /* min_k (v_k - ref_k) / w_k                             (tch_ts)   :44-68  */

typedef struct dfb_moo_desc {
  int32_t kind;
  int32_t n_obj;                       /* 1 .. DFB_MOO_MAX_OBJ */
  double  beta;                        /* UCB kinds: beta_th = sqrt(0.2 d log(2 d t + 1)) (:73-75) */
  double  weight[DFB_MOO_MAX_OBJ];     /* anc_data.obj_weights */
  double  ref[DFB_MOO_MAX_OBJ];        /* anc_data.reference_point (Tchebychev kinds) */
} dfb_moo_desc;

/* ---- build flags ----------------------------------------------------------------------------- */
#define DFB_BUILD_FULL      0   /* L, W = L^-1, alpha, LML  (GP.build_posterior, gp_core.py:155-163) */
#define DFB_BUILD_LML_ONLY  1   /* L and LML only           (GPFitter._tuning_objective, :551-563)   */
#define DFB_BUILD_NO_ALPHA  2   /* L, W only; alpha is supplied with dfb_set_alpha (hallucinations)  */

typedef struct dfb_handle dfb_handle;

/* ---- life cycle ------------------------------------------------------------------------------ */
int         dfb_version(void);
const char* dfb_last_error(void);
int         dfb_create(dfb_handle** out, int device);
void        dfb_destroy(dfb_handle* h);
int         dfb_set_stream(dfb_handle* h, void* cuda_stream);
/* Bytes of device workspace needed for n_max training points, a kernel with n_slots slots and
 * scoring chunks of `chunk` candidates (chunk = 0: library default).  */
size_t      dfb_workspace_bytes(int64_t n_max, int32_t n_slots, int64_t chunk);
int         dfb_set_workspace(dfb_handle* h, void* workspace_dev, size_t bytes, int64_t n_max,
                              int64_t chunk);

/* ---- model ------------------------------------------------------------------------------------
 * dfb_set_kernel       : the GP's kernel (Kernel protocol, kernel.py:59-129), used for K(X, X).
 * dfb_set_test_kernel  : optional different descriptor for K(X*, X) and k(x*, x*) -- Add-UCB's
 *                        per-group kernel (gpb_acquisitions.py:160-176); NULL resets to the GP's.
 * dfb_set_train        : GP.set_data (gp_core.py:127-133): X (n x d) and y - mean_func(X).
 */
int dfb_set_kernel(dfb_handle* h, const dfb_kernel_desc* desc);
int dfb_set_test_kernel(dfb_handle* h, const dfb_kernel_desc* desc);
int dfb_set_train(dfb_handle* h, const double* X_dev, int64_t n, int32_t d,
                  const double* y_centred_dev);

/* GP.build_posterior (gp_core.py:155-163) + compute_log_marginal_likelihood (:222-227):
 * K = k(X,X); L = chol(K + (noise_var + jitter) I); alpha = L^-T L^-1 y_c; W = L^-1;
 * lml = -1/2 y_c^T alpha - sum log L_ii - n/2 log 2 pi.  Returns info > 0 when the matrix is not
 * positive definite (np.linalg.LinAlgError in stable_cholesky, general_utils.py:176-192).  */
int dfb_build_posterior(dfb_handle* h, double noise_var, double jitter, int32_t flags,
                        double* lml_out_host);
/* GP.add_data_multiple (gp_core.py:139-146) and the (N + q)-point factorisation of
 * eval_with_hallucinated_observations (gp_core.py:200-206) WITHOUT the reference's full rebuild: appends q
 * training points (X_new: q x d, y_centred_new: q) to a built posterior and re-derives only the last row block
 * of L, the last block column of L^-T / rows of W = L^-1, alpha and the LML (Cholesky row i depends on rows
 * <= i only).  Uses the kernel, noise_var and jitter of the last dfb_build_posterior.  Requires n + q to stay
 * within the posterior's padded size (multiple of 128), else returns < 0 and the caller rebuilds.
 * flags: DFB_BUILD_FULL or DFB_BUILD_NO_ALPHA, optionally | DFB_EXTEND_SAVE to snapshot what is overwritten so
 * that dfb_restore_posterior() puts the un-extended posterior back bit for bit (hallucinations are temporary).
 * Returns info > 0 if the extended matrix is not positive definite (with DFB_EXTEND_SAVE the old posterior is
 * restored first; without it the posterior is invalid and must be rebuilt).  */
#define DFB_EXTEND_SAVE 16
int dfb_extend_posterior(dfb_handle* h, const double* X_new_dev, int64_t q, const double* y_centred_new_dev,
                         int32_t flags, double* lml_out_host);
int dfb_restore_posterior(dfb_handle* h);
/* Gradients of the log marginal likelihood w.r.t. every hyper-parameter of a plain SE / Matern kernel, in one call:
 *     1/2 tr((alpha alpha^T - K^-1) dK/dparam)          GP.compute_grad_log_marginal_likelihood, gp_core.py:229-240
 * with dK/dparam = Kernel.gradient(param, X, X) (kernel.py:202-217 SE, 301-322 Matern) re-derived entry by entry
 * on the device (never materialised) and K^-1 = L^-T L^-1 from one triangular DMMA product.
 *   out_host[0]     'scale'
 *   out_host[1]     'noise_var' WITHOUT its factor noise_var: 1/2 (alpha.alpha - tr K^-1); the reference's value is
 *                   noise_var * out[1] (gp_core.py:232-233)
 *   out_host[2]     'noise_mean' = sum(alpha)                                                  gp_core.py:234-235
 *   out_host[3]     'same_dim_bandwidths'
 *   out_host[4 + j] 'dim_bandwidths', param_num = j  (j < d)
 * n_out >= 4 + d.  Needs a full posterior (DFB_BUILD_FULL).  Returns -3 for composite kernels (the reference raises
 * NotImplementedError there, kernel.py:123-125).  Uses the K_* chunk buffer as scratch. */
int dfb_lml_gradients(dfb_handle* h, double* out_host, int32_t n_out);

/* max(diag K) of the last build -- the jitter ladder's scale (general_utils.py:183-189). */
int dfb_get_max_diag(dfb_handle* h, double* out_host);
/* Copies of gp.L (n x n lower), gp.alpha (n), gp.K_trtr_wo_noise (n x n); any may be NULL.
 * Read directly by _add_ucb (gpb_acquisitions.py:169-171) and gp_core.py:203.  */
int dfb_get_state(dfb_handle* h, double* L_dev, double* alpha_dev, double* K_dev);
/* Overrides alpha (n values; shorter vectors are zero-extended): eval_with_hallucinated_observations
 * takes the mean from the un-augmented GP and the variance from the augmented one
 * (gp_core.py:192-220).  */
int dfb_set_alpha(dfb_handle* h, const double* alpha_dev, int64_t n);

/* ---- prediction -------------------------------------------------------------------------------
 * GP.eval(X_test, 'none' | 'std') (gp_core.py:165-190): mu = mean_const + K_* alpha;
 * sd = sqrt(k(x*,x*) - |L^-1 k_*|^2) (no clamp: NaN for negative variances, like np.sqrt).
 * Never materialises the M x M covariance.  sd may be NULL (uncert_form 'none').  */
int dfb_eval(dfb_handle* h, const double* Xc, int64_t m, int32_t dc, int32_t space,
             double mean_const, double* mu, double* sd);
/* Optional second workspace for the block-exact joint posterior (covariance / Thompson sampling)
 * of up to `mb` candidates at a time (mb <= the scoring chunk).  */
size_t dfb_ts_workspace_bytes(int64_t n_max, int64_t mb);
int    dfb_set_ts_workspace(dfb_handle* h, void* workspace_dev, size_t bytes, int64_t mb);

/* GP.eval(X_test, 'covar') (gp_core.py:165-187): mu (m) and the full m x m posterior covariance
 * K** - V^T V, V = L^-1 K_*^T (device pointers; m <= mb).  Used by draw_samples (gp_core.py:250-254). */
int dfb_eval_covar(dfb_handle* h, const double* Xc_dev, int64_t m, int32_t dc, double mean_const,
                   double* mu_dev, double* covar_dev);

/* The fused acquisition maximiser: the body of asy_ucb / asy_ei / asy_pi / _ttei / _add_ucb's
 * per-group objective + random_maximise's arg-max (oper_utils.py:70-80).  Scores every candidate
 * row and returns the FIRST index of the maximum with NaN counting as the maximum (np.argmax).
 * scores (m values, same space as Xc) may be NULL.  */
int dfb_score_argmax(dfb_handle* h, const dfb_acq_desc* acq, const double* Xc, int64_t m,
                     int32_t dc, int32_t space, double mean_const, double* scores,
                     double* best_score_host, int64_t* best_index_host);

/* The multi-objective acquisitions' scalarisation + random_maximise's arg-max (np.argmax order) over m
 * candidates that n_obj GPs have already scored with dfb_eval on the device: a_dev[k] = mu_k (UCB kinds) or the
 * sampled values v_k (VAL kinds), b_dev[k] = sd_k (UCB kinds; NULL otherwise).  a_dev / b_dev are HOST arrays of
 * n_obj device pointers (each m doubles); scores_dev (m) may be NULL.  The handle supplies stream and scratch
 * only (any handle with a workspace on the same device).  */
int dfb_moo_score_argmax(dfb_handle* h, const dfb_moo_desc* desc, const double* const* a_dev,
                         const double* const* b_dev, int64_t m, double* scores_dev,
                         double* best_score_host, int64_t* best_index_host);

/* Thompson sampling at scale (BASELINE config 5: 256 draws x 10^6 candidates).  The reference draws its normals
 * with np.random.normal on the host (general_utils.py:230), which dfb_ts_draws reproduces when the caller supplies
 * them; at 10^6 x 256 that is 2 GB of host RNG and copies per call.  dfb_fill_rng generates them on the device
 * with a counter-based generator (Philox4x32-10 + Box-Muller, fp64): element (s, a) of the S x m output depends
 * only on (seed, col0 + a, s) -- the same candidate gets the same normals whatever the block or rank layout.
 * what: DFB_RNG_NORMAL or DFB_RNG_UNIFORM (53-bit uniforms in (0, 1), e.g. random_sample's U, oper_utils.py:62).
 * dfb_ts_argmax folds one block of draws (S x m, row stride ld) into running per-draw (value, index) pairs in
 * np.argmax order (sample.argmax() of asy_ts, gpb_acquisitions.py:127); reset != 0 starts a new arg-max.  */
#define DFB_RNG_NORMAL  0
#define DFB_RNG_UNIFORM 1
int dfb_fill_rng(dfb_handle* h, uint64_t seed, int64_t col0, int32_t S, int64_t m, int32_t what, double* out_dev);
int dfb_ts_argmax(dfb_handle* h, const double* samples_dev, int64_t ld, int32_t S, int64_t m, int64_t idx_base,
                  int32_t reset, double* best_dev, int64_t* index_dev);
/* random_sample + map_to_bounds (oper_utils.py:59-67, general_utils.py:25-27) on the device: rows row0 .. row0+m-1 of
 * the candidate matrix, row-major m x d, coordinate s of global row a = lo[s] + u * (hi[s] - lo[s]) with u the
 * DFB_RNG_UNIFORM element (s, a) of dfb_fill_rng for the same seed.  A row depends on (seed, its global index) only:
 * every rank generates just its shard, nobody holds the M x d matrix on the host, and the winning row is
 * regenerated from its index.  (np.random's MT19937 stream cannot be reproduced this way: the operators keep the
 * reference's host generation for seeded parity and use this in their throughput mode.)  lo / hi: HOST arrays of d.  */
int dfb_fill_candidates(dfb_handle* h, uint64_t seed, int64_t row0, int64_t m, int32_t d, const double* lo_host,
                        const double* hi_host, double* out_dev);

/* Kernel.__call__(X1, X2) (kernel.py:72-83): the n1 x n2 Gram matrix, device pointers. */
int dfb_kernel_matrix(dfb_handle* h, const dfb_kernel_desc* desc, const double* X1_dev, int64_t n1,
                      int32_t d1, const double* X2_dev, int64_t n2, int32_t d2, double* K_dev);

/* Thompson sampling (asy_ts, gpb_acquisitions.py:119-127; GP.draw_samples, gp_core.py:250-254;
 * draw_gaussian_samples, general_utils.py:224-232): samples = (L_post U)^T + mu with
 * L_post = chol(K** - V^T V + jitter I) for ONE block of m <= mb candidates.  Ut_dev is U^T, the
 * S x m matrix of standard normals (drawn on the host with np.random.normal for parity),
 * samples_dev is S x m, mu_dev (m) may be NULL.  max_diag_host receives max(diag covariance), the
 * scale of stable_cholesky's jitter ladder.  Returns info > 0 if the covariance is not PD at this
 * jitter (the host then walks the ladder, general_utils.py:183-203).  S <= 256 per call.  */
int dfb_ts_draws(dfb_handle* h, const double* Xc_dev, int64_t m, int32_t dc, double mean_const,
                 const double* Ut_dev, int32_t S, double jitter, double* samples_dev, double* mu_dev,
                 double* max_diag_host);

/* Counters for bench.py: number of kernels this handle has launched. */
int64_t dfb_launch_count(dfb_handle* h);
/* Diagnostics: arm (buf_dev = device buffer of 1 + 4 * cap_records 64-bit words, zeroed) or disarm (NULL) the per-CTA
 * trace of the K_* and contraction kernels of the overlapped scoring pipeline: record = (kind << 32 | SM id, start ns,
 * end ns, CTA index), buf[0] = number of records wanted.  tools/trace_overlap.py reads it. */
int dfb_debug_trace(void* buf_dev, int64_t cap_records);

/* Tuning switches.
 *  "gemm_impl"  : 0 = cp.async-ring DMMA kernel, 1 = TMA + mbarrier warp-specialised DMMA kernel for the
 *                 fp64 scoring contraction (env DFB200_GEMM=v1|tma at dfb_create; default tma).
 *  "score_impl" : how |L^-1 k_*|^2 is contracted (env DFB200_SCORE=fp64|i8|auto at dfb_create; default auto;
 *                 readable back through dfb_query "score_impl").  ONE switch -- score_impl = 0 -- puts every call
 *                 on the pure fp64 DMMA path:
 *                 0 = fp64 DMMA everywhere (no reduced-precision arithmetic anywhere; 1.2 M candidates/s at N = 5000);
 *                 1 = int8-slice tcgen05 path everywhere, sigma vectors included (exact digit expansion of both fp64
 *                     operands, int32 accumulation in tensor memory; a-priori bound: query "i8_sigma2_bound");
 *                 2 = auto: dfb_eval stays fp64; dfb_score_argmax screens with the int8 path, re-scores in fp64
 *                     every candidate whose int8 score -- widened by the error allowance that follows from the
 *                     bound -- could reach the fp64 maximum, returns the fp64 arg-max (index and score) of those,
 *                     and finally checks the int8 scores of that shortlist against the fp64 ones: a candidate
 *                     outside its allowance voids the screen and the call is redone in fp64 (queries
 *                     "last_selfcheck_violations", "last_selfcheck_ratio").  The bound (api.cu: i8_sigma2_bound) is
 *                     a sqrt(n) rounding-error model validated by a committed sweep (profiles/r02_i8_bound_sweep.json:
 *                     480 configurations, worst measured / bound 0.28), not a worst-case bound.  The screen is skipped when the
 *                     bound exceeds 5e-9 ABSOLUTE (half the 1e-8 sigma^2 contract, whatever the kernel scale;
 *                     query "i8_bound_limit") or n < 1024.
 *  "i8_impl"    : which tcgen05 kernel the int8 path uses (switching re-slices W: layouts differ):
 *                 2 (default) = persistent CTA-pair kernel: tcgen05.mma.cta_group::2 M256 N128 K32, two passes
 *                     per 256 x 128 tile (gemm_i8c2.cuh);
 *                 1 = the same two passes from single CTAs, M128 N128 K32 (gemm_i8x2.cuh);
 *                 0 = one pass of M128 N64 K32 MMAs over 128 x 64 tiles, one CTA per tile (gemm_i8.cuh).
 *  "i8_radix"   : digit scheme of the CTA-pair kernel: 1 = five radix-256 digits, 15 products (measured
 *                 |d sigma^2| 3.6e-10 at N = 5000); 0 = six radix-128 digits, 21 products (3e-11), the scheme
 *                 of i8_impl 0 and 1; -1 (default) = radix 256 whenever its a-priori bound is below
 *                 5e-9 (absolute) for the training kernel, else radix 128.
 *  "i8_fuse"    : 1 (default) = the K_* kernel emits the int8 digit planes directly, 0 = via an fp64 K_* buffer.
 *  "i8_unguarded": diagnostics only (tools/sweep_i8_bound.py): 1 = run the int8 path even when its a-priori bound exceeds
 *                 the limit, so that the bound can be compared with the measured error where it would refuse.
 *  "i8_ts"      : i8_impl 0 only: 1 = stage W's digits in tensor memory (tcgen05.cp), default 0.
 *  "lookahead"  : 1 (default) = look-ahead schedule of the blocked factorisation (next panel's column updated first,
 *                 chol_diag + panel solve of step k+1 overlap the bulk trailing update of step k on a second stream;
 *                 bit-identical results), 0 = one stream, step after step.
 *  "small_eval" : 1 (default) = dfb_eval of <= 32 points computes |L^-1 k_*|^2 by streaming the rows of W once (one
 *                 warp per row, HBM-bound) instead of spending 128-wide DMMA tiles on them; 0 = tile kernels always.
 *  "kstar_seg"  : 1 (default) = second-generation K_* kernels (kstar_seg_kernel: training-stationary, digits from one FMA) for
 *                 plain SE / Matern on <= 8 dims; 0 = the round-1 kernels in the reference's operation order.
 *  "kstar_rows64": 1 (default) = the fp64 K_* rows of the fp64 scoring paths come from kstar_seg_kernel's row form too.
 *  "kstar_overlap": 1 = K_* of chunk c+1 on a second stream beside the contraction of chunk c (two-stream pipeline with
 *                 double-buffered digit planes); default 0 -- no net gain on B200.
 *  "i8_c2_group": candidate tiles per group of the CTA-pair kernel's tile order; 0 (default) = chosen per geometry by
 *                 simulating the static deal (balance first, then the least HBM traffic); query "last_c2_group".
 *  "i8_l2_hint" : L2 eviction priorities of that kernel's TMA loads: 0 (default) none, 1 = K_* digits evict-last,
 *                 2 = also W digits evict-first (measured neutral / slower).
 *  "kstar_fast", "tma_cb_group", "i8_cb_group": kernel-selection / scheduling knobs used by tools/. */
int dfb_set_option(dfb_handle* h, const char* name, int64_t value);
/* Diagnostics: "i8_sigma2_bound", "i8_bound_limit", "i8_ready", "i8_impl", "i8_radix256", "score_impl",
 * "last_used_i8", "last_shortlist" (-1 = overflow -> fp64 pass), "last_selfcheck_violations" (> 0: the int8 screen
 * was voided and the call redone in fp64), "last_selfcheck_ratio" (max |s_int8 - s_fp64| / allowance over the last
 * shortlist; the model's margin is its inverse), "chunk", "last_c2_group", "last_overlapped". */
int dfb_query(dfb_handle* h, const char* name, double* out);

/* Per-kernel-class device timing with CUDA events on the handle's stream (bench.py's roofline):
 * class 0 = K_* build (+mu), 1 = the DMMA contraction |L^-1 k_*|^2, 2 = acquisition + arg-max,
 * 3 = posterior build (whole dfb_build_posterior).  dfb_profile_read synchronises, returns the
 * accumulated milliseconds, launches and work units (candidates for 0-2, builds for 3) and resets. */
#define DFB_PROF_KSTAR 0
#define DFB_PROF_GEMM  1
#define DFB_PROF_ACQ   2
#define DFB_PROF_BUILD 3
int dfb_profile_enable(dfb_handle* h, int on);
int dfb_profile_read(dfb_handle* h, int cls, double* ms_total, int64_t* launches, double* units);

/* Live roofline denominators for bench.py, measured on `device` in the calling process (best of 3 short launches):
 * DFB_PEAK_TCGEN05_I8 -> issue rate of tcgen05.mma kind::i8 M128 N256 K32 in int8 TOP/s (2 per MAC), the peak the
 * int8-slice contraction is quoted against; DFB_PEAK_DMMA_F64 -> fp64 DMMA.8x8x4 issue rate in TFLOP/s.
 * MEASURED_PEAKS.json carries neither.  Not on the product path; replaces nothing in the reference.  */
#define DFB_PEAK_TCGEN05_I8 0
#define DFB_PEAK_DMMA_F64   1
int dfb_measure_peak(int device, int what, double* out_host);

#ifdef __cplusplus
}
#endif
#endif  /* DFB200_H_ */
