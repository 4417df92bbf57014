// Internal declarations shared by the kernel files and the C-ABI host code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsreg {

// ---------------------------------------------------------------------------
// Problem data resident in HBM (one replica per GPU).
//   X     row-major [n][m]   (m includes SLX-lagged columns)
//   y     [n]
//   W     CSR: rp int32 [n+1], ci int32 [nnz], va f64 [nnz]
//   G     symmetric Gram of Z = [y | Wy | X | WX], d = m + 2 (LM: Wy column is zero; SAR) or 2m + 2 (SEM),
//         held ONLY as int64 fixed-point accumulators Gfix[3][d][d] (upper triangle; integer
//         atomics make the sum independent of CTA order).  Sweep j accumulates into buffer
//         j % 3 and clears buffer (j + 1) % 3; the chain step of iteration j reads buffer j % 3,
//         so sweep j + 1 can overlap chain step j.  value = Gfix * Gscale (power of two).
//   Zt    ring sweep: [ntiles][m+1][R] tile-blocked copy of [X | y] (column-major inside a tile)
//   csr_blob / tile_off / tile_len: per tile {int rp_local[R+1], nu, 2 spare; double va[nz~2]; int ci[nz~4]; int ul[nu~4]},
//         16-byte aligned: ul = the distinct columns of the tile (ascending), ci indexes ul.  Rows and columns of the
//         tiles are in the handle's internal ordering (perm: a breadth-first relabelling of the observations made at
//         create when W has no locality; the Gram matrix does not depend on the order of the observations); yp = y in
//         that ordering.  Everything else (rp/ci/va, X, y, Wy, WX) stays in the caller's ordering.
// ---------------------------------------------------------------------------

constexpr int SWEEP_TILE_ROWS = 64;

// Row r of tile column c sits at position r ^ 4 (c & 3) of the column (an involution inside aligned groups of 16 rows): the
// FP64 mma fragments of the persistent sweep (lane = (column g, row t) of an 8 x 4 block) then fall into distinct banks,
// and lane = row readers just see a permutation of their 32 rows.
__host__ __device__ inline int tile_row_pos(int c, int r) { return r ^ ((c & 3) << 2); }

// A tile's blob: header {rp_local[R + 1], nu, wt, 1 spare}, values, column slots, ul[nu~4] (the distinct columns).
//   wt > 0 (rows of similar length, e.g. kNN): slot-major "ELL" layout va[wt][R], ci[wt][R] -- lane = row reads consecutive
//          addresses, no shared-memory bank conflicts; short rows are padded with (0.0, 0)
//   wt = 0 (ragged tiles): CSR layout va[nz~2], ci[nz~4]
// ci indexes ul.  ell_width() is the rule both the host (capacities) and the packing kernel use.
constexpr int SWEEP_PACK_CAP = 4096;      // nonzeros of a tile the packing kernel sorts (more: CSR layout, no de-duplication)
__host__ __device__ inline int ell_width(int nz, int max_len) {
  return (max_len > 0 && nz <= SWEEP_PACK_CAP && SWEEP_TILE_ROWS * max_len <= nz + nz / 4 + SWEEP_TILE_ROWS) ? max_len : 0;
}
__host__ __device__ inline int tile_blob_bytes(int nz, int nu, int wt) {
  const int body = wt > 0 ? wt * SWEEP_TILE_ROWS * 12 : ((nz + 1) & ~1) * 8 + ((nz + 3) & ~3) * 4;
  return ((SWEEP_TILE_ROWS + 4) * 4 + body + ((nu + 3) & ~3) * 4 + 15) & ~15;
}

struct DeviceData {
  int n, m, d, model;
  int64_t nnz;
  const int *rp, *ci;
  const double *va;
  const double *y, *X;
  const double *Zt;           // ring sweep: tile-blocked [X | y] (null: design not covered by the ring)
  const unsigned char *csr_blob;
  const long long *tile_off;  // ntiles + 1 byte offsets into csr_blob (capacity of each tile)
  const int *tile_len;        // ntiles: bytes of each tile's blob actually used (what the ring copies)
  const double *yp;           // y in the ordering of the tiles
  const int *tile_ul;         // SEM ring sweep: [ntiles][u_ints] foreign rows of every tile, last int = their number
  const double *Xr;           // SEM ring sweep (semsweep.cu): X row-major [ntiles * R][ldx] in the ordering of the tiles (null: not used)
  int tile_max_blob;          // largest used blob of one tile (bytes)
  int tile_max_nu;            // most distinct columns in one tile
  // W, X, y in the internal ordering for the sweeps that read them directly (sweep_gram_kernel: SEM, 24 < d <= 88);
  // null = the caller's ordering is local enough and is used as it is
  const int *sw_rp, *sw_ci;
  const double *sw_va, *sw_X, *sw_y;
  double *Wy, *WX;            // cached lags (R/15_sar.R:57, R/15_sem.R:63-64): direct-form kernel, generic sweep
  long long *Gfix;            // 3 * d*d
  int gcur, gzero;            // buffer this launch accumulates into / reads; buffer a sweep clears (-1: none)
  const double *Gscale;       // d*d: value = Gfix * Gscale  (power of two)
  const double *Ginv_scale;   // d*d: fixed = rint(value * Ginv_scale)
  // log-determinant nodes
  int n_nodes;
  const double *node_re, *node_im, *node_w;
  // grid + spline log-determinant (R/15_sar.R:79-87): spl = [x | y | b | c | d], spl_n knots each (0: nodes are used)
  int spl_n;
  const double *spl;
  // sampled SLX connectivity parameter (slxdelta.cu): every chain has its own Gram matrix [C][d][d] (its lag columns
  // depend on its delta); null otherwise
  const double *Gchain;
};

// Sampled connectivity parameter delta of the SLX model (slxdelta.cu): Psi(delta)_ij = dist_ij ^ -delta on a fixed pattern
struct DeltaData {
  int L, norm, sampled;         // lagged columns; 1 = rows of Psi(delta) divided by their sum; 0 = delta stays at its start
  const int *rp, *ci;           // pattern (CSR)
  const double *logd;           // log distances
  const double *Xl;             // X_SLX row-major [n][L]
  double *delta, *delta_prop, *scale;   // [C]
  long long *cnt;               // [C][4] accepted, proposed, accepted_tune, proposed_tune
  double *Gchain;               // [C][d][d]
  double prior_a, prior_b, lo, hi;
};

// entry (i, j) of the Gram matrix from the fixed-point buffer of this launch
__device__ __forceinline__ double g_load(const DeviceData &dd, int i, int j) {
  const int e = i <= j ? i * dd.d + j : j * dd.d + i;
  // through L2: inside the persistent kernel the buffer is rewritten by the sweep CTAs' atomics behind L1's back
  return (double)__ldcg(dd.Gfix + (size_t)dd.gcur * dd.d * dd.d + e) * dd.Gscale[e];
}

struct Priors {
  int prior;
  double shape0, rate0, shape1;
  double sng_lambda_a, sng_lambda_b, sng_theta_scale, sng_theta_a;
  double lam_a, lam_b, lam_lo, lam_hi, lbeta_ab;   // spatial lambda prior + bounds
};

// Per-chain state, structure of arrays over the C chains of this GPU.
struct ChainState {
  int n_chains, chain_offset;
  uint64_t seed;
  double *beta;        // [C][m]
  double *sigma2;      // [C]
  double *lam;         // [C]
  double *p0;          // [C][m]   prior precision diagonal
  double *aux;         // [C][A]   A = 2m + 4: vecA[m], vecB[m], s0, s1, ldet_cur, rss_cur
  double *mh_scale;    // [C][2]   0 = spatial lambda, 1 = SNG theta
  long long *mh_cnt;   // [C][2][4] accepted, proposed, accepted_tune, proposed_tune
  long long *iters;    // [C]      R/10_lm.R:75 meta$iterations
  int *step;           // [C]      iteration index inside the current bsreg_run
  const double *mu0;   // [m]      prior mean
};

struct RunCtl {        // lives in device memory; updated by the host between blocks
  int n_burn, tune_limit, n_save, save_base, save_cap, store;
  double acc_lo, acc_hi, acc_change;
};

constexpr int AUX_EXTRA = 4;
__host__ __device__ inline int aux_len(int m) { return 2 * m + AUX_EXTRA; }

// opt-in to more than 48 KB of dynamic shared memory.  cudaFuncSetAttribute is per device (context), so the largest size
// configured so far is tracked per (device, kernel); thread-safe (one host thread per device sets shards up concurrently)
cudaError_t smem_optin(const void *kernel, size_t bytes);

// launchers (sweep.cu)
void launch_transpose(const double *src_colmajor, double *dst_rowmajor, int n, int m, int ld_dst, int col0, cudaStream_t s);
void launch_spmm(int n, const int *rp, const int *ci, const double *va, const double *B, int m, int ldb,
                 double *out, int ldo, int col0, double alpha, const double *C, double beta, cudaStream_t s);
int  colnorm_parts();
void launch_colnorms(const DeviceData &dd, double *partial, cudaStream_t s);   // partial: colnorm_parts() * d
bool sweep_fast_supported(const DeviceData &dd);
bool sweep_ring_supported(int d, int model, int max_blob, int max_nu);   // does the TMA-fed ring sweep cover this design
void launch_pack_tiles(const double *X, const double *y, const int *perm, double *Zt, double *yp, int n, int m, cudaStream_t s);
void launch_gram_to_double(const DeviceData &dd, double *G, cudaStream_t s);
void launch_pack_csr_tiles(const int *rp, const int *ci, const double *va, const int *perm, const int *inv,
                           const long long *tile_off, int *tile_len, int *max_len_nu, unsigned char *blob, int n,
                           bool split_local, cudaStream_t s);
// launchers (semsweep.cu)
int  sem_ring_ldx(int m);
bool sem_ring_supported(int n, int m, int max_blob, int max_nuf);
void launch_pack_rows(const double *X, const double *y, const int *perm, double *Xr, double *yp, int n, int m, int ldx,
                      cudaStream_t s);
void launch_sweep_sem(const DeviceData &dd, int sm_count, cudaStream_t s);
int  sem_ul_ints(int max_nuf);
void launch_extract_ul(const unsigned char *blob, const long long *tile_off, int *tile_len, int *max_len, int *ul, int u_ints,
                       int n, cudaStream_t s);
void launch_permute_problem(const int *rp, const int *ci, const double *va, const double *X, const double *y, const int *perm,
                            const int *inv, const int *rp_new, int *ci_new, double *va_new, double *X_new, double *y_new,
                            int n, int m, cudaStream_t s);
void launch_trace_w2_rows(int n, const int *rp, const int *ci, const double *va, double *rowsum, cudaStream_t s);
void launch_sweep(const DeviceData &dd, int sm_count, cudaStream_t s);
int  sweep_launch_count();
void launch_residual_ss(const DeviceData &dd, int n_eval, const double *beta,
                        const double *lam, const double *rho, double *partial, double *out, cudaStream_t s);
void launch_probes(double *U, int n, int n_probes, uint64_t seed, cudaStream_t s, const int *perm = nullptr, int n_rows = 0);
int  spmm_dot_parts(int n, int m);   // per-CTA partials launch_spmm_dot writes
void launch_combine_rows(const double *partial, int nparts, int count, double *out, cudaStream_t s);
void launch_dot(const double *a, const double *b, int64_t len, double *partial, double *out, cudaStream_t s);
void launch_spmm_dot(int n, const int *rp, const int *ci, const double *va, const double *B, int m, double *out, double alpha,
                     const double *C, double beta, const double *U, double *partial, double *dot, cudaStream_t s);

// launchers (logdet.cu)
void launch_grid_logdets(int size, const int *rp, const int *ci, const double *va, int n_grid, const double *xs_dev,
                         double *A, double reps, double *out, cudaStream_t s);
void fmm_spline_coefficients(int n, const double *x, const double *y, double *b, double *c, double *d);

// launchers (slxdelta.cu)
void launch_slx_power_lag(const DeltaData &dl, double delta, double *X, int m, int k_base, int n, cudaStream_t s);
int  delta_sweep_ctas(int sm_count);
void launch_sweep_delta(const DeltaData &dl, const double *X, const double *y, int n, int m, int k_base, int n_chains,
                        const double *delta, double *partial, int sm_count, cudaStream_t s);
void launch_delta_propose(const DeltaData &dl, const ChainState &cs, cudaStream_t s);
void launch_delta_decide(const DeviceData &dd, const DeltaData &dl, const Priors &pr, const ChainState &cs, const RunCtl *ctl,
                         double *draws, const double *partial, int n_parts, int mode, cudaStream_t s);

// launchers (chain.cu)
size_t chain_smem_bytes(int m, int model, int prior);
int  chain_threads(int m);
struct ChainInit {
  const double *beta0;     // device [m] or null (least squares)
  double sigma0;           // NaN = RSS/(N-M)
  double lam0, scale0;
  double hs_lambda, hs_tau, hs_zeta, hs_nu, sng_lambda, sng_tau, sng_theta, sng_theta_scale;
  const double *p0_init;   // device [m]
  bool refresh_only;       // only recompute the cached log-det / rss at the current lambda
};
void launch_chain_init(const DeviceData &dd, const Priors &pr, const ChainState &cs, const ChainInit &ci, cudaStream_t s);
// nrep > 1: that many iterations in one launch (only with chain_step_loops(): fixed Gram matrix, i.e. cached statistics)
void launch_chain_step(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                       cudaStream_t s, int nrep = 1);
bool chain_step_loops(const DeviceData &dd, const Priors &pr);
// launchers (persist.cu)
bool persistent_supported(const DeviceData &dd, const Priors &pr, int n_chains, int sm_count, bool streamed);
size_t persistent_sync_bytes();
cudaError_t launch_persistent(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl &ctl, const RunCtl *ctl_dev,
                              double *draws, int K, bool cached, void *sync, int sm_count, cudaStream_t s);

void launch_eval_logdet(const DeviceData &dd, int n_eval, const double *v, double *out, cudaStream_t s);
void launch_eval_conditionals(const DeviceData &dd, const Priors &pr, const double *mu0, int n_eval, const double *beta,
                              const double *sigma2, const double *lam, const double *p0, const double *v, double *mu1,
                              double *rate1, double *rss, double *logpost, cudaStream_t s);
void launch_eval_rng(int kind, uint64_t seed, int chain, int slot, int it0, int n, const double *params, double *out,
                     cudaStream_t s);

}  // namespace bsreg
