/* bsreg_cuda.h -- C ABI of the B200-native bsreg Gibbs hot path.
 *
 * The reference (nk027/bsreg v0.0.2) is pure R and has NO FFI boundary; the seam
 * this library replaces is the R6 protocol that R/50_execute.R drives:
 *
 *   bsreg_create      <- get_blm/get_bslx/get_bsar/get_bsem/get_bsdm/get_bsdem
 *                        (R/40_setup.R:11-123): class$new(priors) + mdl$setup(...)
 *                        i.e. Base$setup R/10_lm.R:30-44, setup_2SLX R/12_slx.R:51-68,
 *                        setup_4NG R/10_lm.R:142-153, setup_8SEM R/15_sem.R:53-110,
 *                        setup_9SAR R/15_sar.R:47-98, setup_HS / setup_SNG
 *                        R/11_shrinkage.R:112-122 / :35-49, starting_NG R/10_lm.R:155-164
 *   bsreg_run         <- sample() + tune() R/50_execute.R:12-83 (whole loop on device;
 *                        each iteration = Base$sample R/10_lm.R:46-63)
 *   bsreg_get_state   <- mdl$get_parameters, mdl$MH[[k]]$get_scale/get_accepted/...,
 *                        mdl$get_meta (R/10_lm.R:90,193; R/20_mh.R:85-98)
 *   bsreg_set_state   <- the mutable R6 model object that bm.bm continues
 *                        (R/60_interface.R:71-78); needed for saveRDS round trips
 *   bsreg_eval_*      <- single deterministic pieces, exposed for parity tests
 *
 * Conventions: plain C, no C++ or torch types.  Every function returns 0 on
 * success or a negative BSREG_E* code; the message is in bsreg_last_error()
 * (thread-local).  The caller owns every host pointer; create() copies inputs
 * to the device(s) and may be followed immediately by freeing them.  Dense
 * matrices are COLUMN-MAJOR (R layout).  CSR indices are 0-based int32.
 */
#ifndef BSREG_CUDA_H
#define BSREG_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSREG_ABI_VERSION 2

enum { BSREG_OK = 0, BSREG_EINVAL = -1, BSREG_ECUDA = -2, BSREG_ENOMEM = -3,
       BSREG_EABORT = -4, BSREG_ENOGPU = -5 };

/* model: SLX / SDM / SDEM are LM / SAR / SEM with k_slx > 0 (R/40_setup.R:34,95,116) */
enum { BSREG_MODEL_LM = 0, BSREG_MODEL_SAR = 1, BSREG_MODEL_SEM = 2 };
/* set_options(type = ...) R/41_set_options.R:21 */
enum { BSREG_PRIOR_INDEPENDENT = 0, BSREG_PRIOR_CONJUGATE = 1,
       BSREG_PRIOR_SHRINKAGE = 2, BSREG_PRIOR_HORSESHOE = 3 };
/* log|I - lambda W|: given nodes (eigenvalues, R/15_sar.R:91-96) or Chebyshev
 * moments estimated on the device from W (new at large N) */
enum { BSREG_LDET_NONE = 0, BSREG_LDET_NODES = 1, BSREG_LDET_CHEBYSHEV = 2,
       /* ldet_SAR / ldet_SEM = list(grid = TRUE): exact log|det(I - x W_sub)| * reps on a grid of x (dense LU on the
        * device, the algorithm behind R's determinant()), interpolated by the FMM cubic spline of stats::splinefun
        * (R/15_sar.R:79-87, R/15_sem.R:91-99, R/00_aux.R:105-107) */
       BSREG_LDET_SPLINE = 3 };
/* STREAM: W, y, X are swept every Gibbs iteration (fused SpMM + Gram reduction);
 * CACHED: the sweep runs once at create() -- same draws, see DESIGN.md */
enum { BSREG_STATS_STREAM = 0, BSREG_STATS_CACHED = 1 };

typedef struct bsreg_handle bsreg_handle;

typedef struct {
  int32_t n, k;              /* observations, columns of X                     */
  const double *y;           /* n                                              */
  const double *X;           /* n x k column-major                             */
  int32_t k_slx;             /* columns to lag (0 = none): X <- [X, Psi X_slx] */
  const double *X_slx;       /* n x k_slx column-major (R/12_slx.R:51-68)      */
  int64_t nnz;               /* CSR of Psi_SAR / Psi_SEM / Psi_SLX (NULL for LM) */
  const int32_t *row_ptr;    /* n + 1 */
  const int32_t *col_idx;    /* nnz   */
  const double *val;         /* nnz   */
  int64_t nnz_slx;           /* optional separate Psi_SLX (R/40_setup.R:99); 0 = use W */
  const int32_t *row_ptr_slx, *col_idx_slx;
  const double *val_slx;
  int32_t n_nodes;           /* BSREG_LDET_NODES: ldet(v) = sum_k w_k * 0.5*log((1-v re_k)^2 + (v im_k)^2) */
  const double *node_re, *node_im, *node_w;   /* eigenvalues and weight = reps (R/15_sar.R:68,95) */
} bsreg_problem;

typedef struct {
  int32_t model, prior, ldet, stats_mode;
  /* set_NG (R/41_set_options.R:50-61) */
  const double *ng_mu;        int32_t ng_mu_len;         /* length 1 or M */
  const double *ng_precision; int32_t ng_precision_len;  /* length 1 or M (diagonal) */
  double ng_shape, ng_rate;
  const double *ng_beta0;     /* starting beta (length M) or NULL = least squares */
  double ng_sigma0;           /* starting sigma^2, NaN = RSS/(N-M) (R/10_lm.R:161-163) */
  /* set_SNG (R/41_set_options.R:65-84) */
  double sng_lambda_a, sng_lambda_b, sng_theta_scale, sng_theta_a, sng_lambda, sng_tau, sng_theta;
  /* set_HS (R/41_set_options.R:88-100) */
  double hs_lambda, hs_tau, hs_zeta, hs_nu;
  /* set_SAR / set_SEM (R/41_set_options.R:115-145), lambda part */
  double sp_lambda_a, sp_lambda_b, sp_lambda, sp_lambda_scale, sp_lambda_min, sp_lambda_max;
  /* BSREG_LDET_CHEBYSHEV */
  int32_t cheb_order, cheb_probes;
  /* BSREG_LDET_SPLINE: i_lambda = c(from, to, length) and reps of ldet_SAR / ldet_SEM (R/15_sar.R:47); the grid
   * matrices are the leading (n / reps) x (n / reps) block of W (get_W, R/15_sar.R:72-76), at most 4096 wide */
  double ldet_grid_from, ldet_grid_to;
  int32_t ldet_grid_len, ldet_reps;
  /* Connectivity FUNCTION Psi_SLX(delta) (R/12_slx.R:69-74; SLX model only): slx_psi = 1 reads the values of the
   * Psi_SLX pattern (val_slx, or val if nnz_slx = 0) as distances d_ij > 0 and uses W(delta)_ij = d_ij ^ -delta
   * (the paper's inverse-distance decay, paper/0_data.R:31-35), rows divided by their sum if slx_psi_norm = 1.
   * set_SLX delta part (R/41_set_options.R:115-152): delta is sampled by random-walk MH iff slx_delta_scale > 0
   * (MH_SLX_delta, R/20_mh.R:229-278); the draws then carry a delta_SLX column after sigma (R/12_slx.R:117-121). */
  int32_t slx_psi, slx_psi_norm;
  double slx_delta, slx_delta_scale, slx_delta_a, slx_delta_b, slx_delta_min, slx_delta_max;
  /* chains: this handle runs global chains [chain_offset, chain_offset + n_chains) */
  int32_t n_chains, chain_offset;
  uint64_t seed;
  /* devices: n_devices = 0 -> current device only; chains are split in
   * contiguous blocks across the listed devices, no collective in the loop */
  int32_t n_devices;
  const int32_t *devices;
  void *stream;               /* cudaStream_t to run on (single-device only), NULL = own stream */
} bsreg_options;

/* set_mh (R/42_set_mh.R:13-26).  NOTE reference quirk: R reads acc_change from
 * adjust_burn; the host layer replicates that, the ABI takes the value to use. */
typedef struct { double adjust_burn, acc_lo, acc_hi, acc_change; } bsreg_mh;

typedef int (*bsreg_progress_fn)(int64_t done, int64_t total, void *user);  /* nonzero = abort */

int  bsreg_abi_version(void);
const char *bsreg_last_error(void);

int  bsreg_create(const bsreg_problem *prob, const bsreg_options *opt, bsreg_handle **out);
void bsreg_destroy(bsreg_handle *h);

/* n, M (after SLX augmentation), P = stored parameters per draw, chains, Gram dimension d */
int  bsreg_dims(const bsreg_handle *h, int32_t *n, int32_t *m, int32_t *p, int32_t *n_chains, int32_t *d);

/* Burn-in with MH tuning, then n_save stored draws (R/50_execute.R:12-83).
 * draws: caller-allocated, dims (n_save, P, n_chains) column-major, i.e. one
 * R-ready n_save x P matrix per chain; NULL keeps the draws on the device
 * (fetch the last block with bsreg_fetch_draws).  Pinned host memory makes the
 * device->host copies asynchronous.  Continues the chains if called again. */
int  bsreg_run(bsreg_handle *h, int32_t n_burn, int32_t n_save, const bsreg_mh *mh,
               double *draws, bsreg_progress_fn progress, void *user);
int  bsreg_fetch_draws(bsreg_handle *h, int32_t first, int32_t count, double *draws /* (count, P, n_chains) */);

/* Per-chain state.  params (P, C); p0 (M, C) prior precision diagonal; aux (A, C)
 * shrinkage state, A = bsreg_aux_len(); mh_* per MH object k (0 = spatial lambda,
 * 1 = SNG theta): value/scale (2, C), counters (4, 2, C) = accepted, proposed,
 * accepted_tune, proposed_tune; iterations (C). Any pointer may be NULL. */
int  bsreg_aux_len(const bsreg_handle *h);
int  bsreg_get_state(bsreg_handle *h, double *params, double *p0, double *aux, double *mh_value,
                     double *mh_scale, int64_t *mh_counters, int64_t *iterations);
int  bsreg_set_state(bsreg_handle *h, const double *params, const double *p0, const double *aux,
                     const double *mh_value, const double *mh_scale, const int64_t *mh_counters,
                     const int64_t *iterations);

/* State of the MH object SLX_delta per chain: value, proposal scale, counters (4, C) as above.  Pointers may be NULL. */
int  bsreg_get_delta_state(bsreg_handle *h, double *delta, double *scale, int64_t *counters);
int  bsreg_set_delta_state(bsreg_handle *h, const double *delta, const double *scale, const int64_t *counters);

/* ---- parity entry points: each runs the same device code the sampler uses ---- */
/* K1: out = W * B for a host matrix B (n x m column-major)      [R/15_sar.R:57, R/15_sem.R:63-64] */
int  bsreg_eval_spmm(bsreg_handle *h, const double *B, int32_t m, double *out);
/* K1+K2 fused sweep: G = Z'Z, Z = [y | Wy | X | WX] (d x d, symmetric)  [R/10_lm.R:36, R/15_sar.R:133-135, R/15_sem.R:34-35] */
int  bsreg_eval_gram(bsreg_handle *h, double *G);
/* K2 direct form: ss[e] = || (y - lam_e Wy) - (X - rho_e WX) beta_e ||^2  [R/10_lm.R:176,192; R/15_sar.R:29; R/15_sem.R:29-32] */
int  bsreg_eval_residual_ss(bsreg_handle *h, int32_t n_eval, const double *beta /* (M, n_eval) */,
                            const double *lam, const double *rho, double *ss);
/* K3: ldet[e] = log|I - v_e W| from the handle's nodes            [R/15_sar.R:94-96] */
int  bsreg_eval_logdet(bsreg_handle *h, int32_t n_eval, const double *v, double *ldet);
/* nodes in use (eigen nodes as given, or Chebyshev nodes built at create); arrays of bsreg_n_nodes() */
int  bsreg_n_nodes(const bsreg_handle *h);
int  bsreg_get_nodes(bsreg_handle *h, double *re, double *im, double *w);
/* BSREG_LDET_SPLINE: the grid and the spline in use: x, y = log-determinants at the grid points, FMM coefficients b, c, d
 * (s(u) = y_i + dx (b_i + dx (c_i + dx d_i)), dx = u - x_i); arrays of bsreg_n_spline() */
int  bsreg_n_spline(const bsreg_handle *h);
int  bsreg_get_spline(bsreg_handle *h, double *x, double *y, double *b, double *c, double *d);
/* K4: Chebyshev trace moments m_0..m_order of W with n_probes Rademacher probes */
int  bsreg_eval_trace_moments(bsreg_handle *h, int32_t order, int32_t n_probes, uint64_t seed, double *moments);
/* conditional pieces at fixed parameters, per evaluation e:
 *   mu1 (M, n_eval), rate1: posterior mean of beta and sigma^2 rate  [R/10_lm.R:168-169,176]
 *   logpost: log conditional posterior of the spatial parameter at v  [R/20_mh.R:192-199]
 *   rss: SAR rss(v) with (e0e0,e1e0,e1e1) [R/15_sar.R:104-116] or SEM OLS rss(v) [R/15_sem.R:116-120] */
int  bsreg_eval_conditionals(bsreg_handle *h, int32_t n_eval, const double *beta, const double *sigma2,
                             const double *lam, const double *p0 /* (M, n_eval) */, const double *v,
                             double *mu1, double *rate1, double *rss, double *logpost);
/* device RNG (Philox4x32-10 contract, see oracle/philox.py): kind 0 uniform, 1 normal,
 * 2 gamma(p0=shape,p1=rate), 3 exponential(p0=rate), 4 GIG(p0=lambda,p1=chi,p2=psi),
 * 5 truncated RW (p0=loc,p1=scale,p2=lower,p3=upper).  Draw i uses iteration = it0 + i. */
int  bsreg_eval_rng(int32_t kind, uint64_t seed, int32_t chain, int32_t slot, int32_t it0, int32_t n,
                    const double *params, double *out);

/* ---- measurement helpers (bench.py) ---- */
void *bsreg_stream(bsreg_handle *h);             /* cudaStream_t of device shard 0 */
int  bsreg_launch_sweep(bsreg_handle *h, int32_t reps);      /* enqueue reps fused sweeps, no sync */
int  bsreg_launch_chain_step(bsreg_handle *h, int32_t reps); /* enqueue reps chain steps (burn-in mode), no sync */
int  bsreg_sync(bsreg_handle *h);
int64_t bsreg_sweep_bytes(const bsreg_handle *h);            /* algorithmic bytes of one sweep, SURVEY.md 8(d) */
int64_t bsreg_kernel_launches(const bsreg_handle *h);        /* kernels launched so far by this handle */

/* Device buffers of destroyed handles are cached for the next bsreg_create (R sessions call bm() repeatedly;
 * cudaMalloc/cudaFree dominate a short fit otherwise).  This returns the cache to the driver. */
void bsreg_release_cache(void);

/* Host-side logic of bsreg_create that chooses the INTERNAL ordering of the observations for the tiled sweep (the
 * Gram matrix Z'Z -- everything the sampler needs from N-scale data, R/10_lm.R:36, R/15_sar.R:106-108 -- does not
 * depend on the order of the rows).  Returns 1 and fills perm[n] (internal row r = caller's observation perm[r],
 * breadth-first over the out-edges of W) if W has no locality or `force` is set, 0 if W is kept as it is, < 0 on
 * error.  Needs no device; exists for tests. */
int bsreg_debug_internal_order(int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t force, int32_t *perm);

#ifdef __cplusplus
}
#endif
#endif /* BSREG_CUDA_H */
