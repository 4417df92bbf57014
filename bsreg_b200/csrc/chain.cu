// K5/K6/K7: everything of one Gibbs iteration that is per chain (one CTA per chain).
//
// Follows Base$sample, R/10_lm.R:46-63:
//   sample_sigma      R/10_lm.R:174-179   (Conjugate: R/10_lm.R:226-230)
//   sample_beta       R/10_lm.R:166-172 + rmvn R/00_aux.R:54-70 (Conjugate: R/10_lm.R:222-225)
//   sample_shrinkage  Horseshoe R/11_shrinkage.R:124-141, Normal-Gamma R/11_shrinkage.R:51-72 + R/20_mh.R:116-146
//   sample_latent     SAR lambda, R/15_sar.R:101-125 + MH_SAR_lambda R/20_mh.R:156-207
//   sample_volatility SEM lambda, R/15_sem.R:113-129 (MH_SEM_lambda alias, R/20_mh.R:218)
//   finalize + tuning R/10_lm.R:73-77, R/20_mh.R:73-98, R/50_execute.R:62-71
//   draw store        R/50_execute.R:25-29 with the layout of R/10_lm.R:193, R/15_sar.R:138-142
//
// All N-scale quantities enter through blocks of G = Z'Z (see sweep.cu); the K x K
// algebra runs in shared memory.  The Cholesky is "bordered": right-hand sides and an
// identity block are appended as extra rows, so forward solves and L^-1 come out of
// the same column sweep and no serial triangular solve is left on the critical path.
#include <cstdlib>

#include "internal.cuh"
#include "rng.cuh"

namespace bsreg {

enum { MODE_STEP = 0, MODE_INIT = 1, MODE_REFRESH = 2 };

__device__ __forceinline__ double warp_sum_c(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NT>
__device__ __forceinline__ void bsync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// fixed-order block sum, result in every thread
template <int NT>
__device__ __forceinline__ double bsum(double v, double *red) {
  v = warp_sum_c(v);
  if (NT == 32) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) s += red[w];
  return s;
}

// ---- views of the Gram matrix G (d x d, symmetric), Z = [y | Wy | X | WX] ----
// Small designs stage the whole matrix in shared memory at kernel start (Gs != null); wide ones
// (d*d beyond the budget) read the fixed-point buffer through L2 on demand.
struct GView {
  const double *Gs;
  const long long *Gfix;       // this launch's fixed-point buffer
  const double *Gscale;
  int d, m, model;
  __device__ __forceinline__ double at(int i, int j) const {
    if (Gs) return Gs[i * d + j];
    const int e = i <= j ? i * d + j : j * d + i;
    return (double)Gfix[e] * Gscale[e];
  }
  __device__ __forceinline__ double yy() const { return at(0, 0); }
  __device__ __forceinline__ double yWy() const { return at(0, 1); }
  __device__ __forceinline__ double WyWy() const { return at(1, 1); }
  __device__ __forceinline__ double Xy(int a) const { return at(0, 2 + a); }
  __device__ __forceinline__ double XWy(int a) const { return at(1, 2 + a); }
  __device__ __forceinline__ double XX(int a, int b) const { return at(2 + a, 2 + b); }
  __device__ __forceinline__ double WXy(int a) const { return at(0, 2 + m + a); }
  __device__ __forceinline__ double WXWy(int a) const { return at(1, 2 + m + a); }
  __device__ __forceinline__ double XWX(int a, int b) const { return at(2 + a, 2 + m + b); }
  __device__ __forceinline__ double WXWX(int a, int b) const { return at(2 + m + a, 2 + m + b); }
  // working quantities at lambda: R/15_sar.R:132-133 (z, X'z), R/15_sem.R:28-36 (filtered y, X)
  __device__ __forceinline__ double xx_eff(int a, int b, double l) const {
    if (model != 2) return XX(a, b);
    return XX(a, b) - l * (XWX(a, b) + XWX(b, a)) + l * l * WXWX(a, b);
  }
  __device__ __forceinline__ double xy_eff(int a, double l) const {
    if (model == 0) return Xy(a);
    if (model == 1) return Xy(a) - l * XWy(a);
    return Xy(a) - l * (XWy(a) + WXy(a)) + l * l * WXWy(a);
  }
  __device__ __forceinline__ double yy_eff(double l) const {
    if (model == 0) return yy();
    return yy() - 2.0 * l * yWy() + l * l * WyWy();
  }
};

// ---- bordered Cholesky: rows [0,m) = A (lower), rows [m,nrows) = extra rows B'.
// On exit rows [0,m) hold L (A = L L'), extra row r holds (L^-1 b_r)'. nrows <= 3*NT.
template <int NT>
__device__ void chol_bordered(double *T, double *dg, int m, int nrows, int ld, int tid = threadIdx.x) {
  for (int k = 0; k < m; ++k) {
    double s[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int r = tid + q * NT;
      s[q] = 0.0;
      if (r >= k && r < nrows) {
        const double *tr = T + r * ld, *tk = T + k * ld;
        double s0 = tr[k], s1 = 0.0;
        int j = 0;
        for (; j + 1 < k; j += 2) { s0 = fma(-tr[j], tk[j], s0); s1 = fma(-tr[j + 1], tk[j + 1], s1); }
        if (j < k) s0 = fma(-tr[j], tk[j], s0);
        s[q] = s0 + s1;
        if (r == k) dg[k] = s[q];
      }
    }
    bsync<NT>();
    const double piv = sqrt(dg[k]);
    const double dinv = 1.0 / piv;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int r = tid + q * NT;
      if (r >= k && r < nrows) T[r * ld + k] = (r == k) ? piv : s[q] * dinv;
    }
    bsync<NT>();
  }
}

// x[j] = sum_{k>=j} Y[j][k] w[k]  (Y = L^-T stored by rows, upper triangular)
__device__ __forceinline__ double upper_dot(const double *Yrow, const double *w, int j, int m) {
  double s0 = 0.0, s1 = 0.0;
  int k = j;
  for (; k + 1 < m; k += 2) { s0 = fma(Yrow[k], w[k], s0); s1 = fma(Yrow[k + 1], w[k + 1], s1); }
  if (k < m) s0 = fma(Yrow[k], w[k], s0);
  return s0 + s1;
}

template <int NT>
__device__ double logdet_nodes(const DeviceData &dd, double v, double *red) {
  double acc = 0.0;
  for (int k = threadIdx.x; k < dd.n_nodes; k += NT) {
    const double a = 1.0 - v * dd.node_re[k], b = v * dd.node_im[k];
    acc = fma(dd.node_w[k], 0.5 * log(a * a + b * b), acc);
  }
  return bsum<NT>(acc, red);
}

// dbeta((v+1)/2, a, b, log=TRUE) - log(2): R/20_mh.R:197-199
__device__ __forceinline__ double log_prior_lambda(double v, const Priors &pr) {
  const double x = 0.5 * (v + 1.0);
  return (pr.lam_a - 1.0) * log(x) + (pr.lam_b - 1.0) * log1p(-x) - pr.lbeta_ab - 0.6931471805599453;
}

// R/20_mh.R:192-195
__device__ __forceinline__ double logpost_lambda(double v, double ldet, double rss, int n, int m, const Priors &pr) {
  if (!(rss > 0.0)) return nan("");
  return ldet - 0.5 * (double)(n - m) * log(rss) + log_prior_lambda(v, pr);
}

// shared-memory workspace of one chain
struct Work {
  double *T, *dg, *p0, *mu0, *beta, *mu1, *b0, *b1, *eps, *t0, *t1, *red, *G;
};

constexpr int G_STAGE_MAX_D = 48;     // stage G in shared memory up to this Gram dimension (18 KB)
__host__ __device__ inline int gram_dim(int m, int model) { return model == 2 ? 2 * m + 2 : m + 2; }
__host__ __device__ inline bool g_staged(int m, int model) { return gram_dim(m, model) <= G_STAGE_MAX_D; }
__host__ __device__ inline int work_ld(int m) { return m | 1; }
__host__ __device__ inline size_t work_doubles(int m, int model) {
  const size_t d = gram_dim(m, model);
  return (size_t)(2 * m + 3) * work_ld(m) + 10 * (size_t)m + 16 + (g_staged(m, model) ? d * d : 0);
}

__device__ __forceinline__ Work carve(double *sm, int m, int model) {
  Work w;
  w.T = sm;
  w.dg = w.T + (size_t)(2 * m + 3) * work_ld(m);
  w.p0 = w.dg + m; w.mu0 = w.p0 + m; w.beta = w.mu0 + m; w.mu1 = w.beta + m;
  w.b0 = w.mu1 + m; w.b1 = w.b0 + m; w.eps = w.b1 + m; w.t0 = w.eps + m; w.t1 = w.t0 + m;
  w.red = w.t1 + m;
  w.G = g_staged(m, model) ? w.red + 16 : nullptr;
  return w;
}

// copy the Gram matrix of this launch's buffer into shared memory (both triangles)
template <int NT>
__device__ __forceinline__ void stage_gram(const DeviceData &dd, double *Gs) {
  if (Gs == nullptr) return;
  const int d = dd.d;
  for (int e = threadIdx.x; e < d * d; e += NT) Gs[e] = g_load(dd, e / d, e - (e / d) * d);
}

// || y* - X* b ||^2 = y*'y* - 2 b'X*'y* + b'X*'X* b   (R/10_lm.R:176 via R/10_lm.R:192)
template <int NT>
__device__ double resid_ss(const GView &g, const double *b, double lam, int m, double *red) {
  double part = 0.0;
  for (int a = threadIdx.x; a < m; a += NT) {
    double t = 0.0;
    for (int c = 0; c < m; ++c) t = fma(g.xx_eff(a, c, lam), b[c], t);
    part = fma(b[a], t - 2.0 * g.xy_eff(a, lam), part);
  }
  return g.yy_eff(lam) + bsum<NT>(part, red);
}

// Fill the identity block used to extract L^-T
template <int NT>
__device__ void fill_identity(double *T, int row0, int m, int ld) {
  for (int e = threadIdx.x; e < m * m; e += NT) {
    const int j = e / m, k = e - j * m;
    T[(row0 + j) * ld + k] = (j == k) ? 1.0 : 0.0;
  }
}

// P1 = X*'X*/s + P0 and its bordered factorisation with rhs rows; returns nothing, leaves
// L in rows [0,m), L^-1 rhs in rows [m, m+nrhs), L^-T in rows [m+nrhs, 2m+nrhs).
// rhs_kind: 0 = X*'y*/s + P0 mu0 (beta step); SAR extras 1 = P0 mu0 + X'y/s, 2 = P0 mu0 + X'Wy/s
template <int NT>
__device__ void factor_system(const GView &g, Work &w, double lam_eff, double s, int m, int nrhs, int first_rhs_kind) {
  const int ld = work_ld(m);
  for (int e = threadIdx.x; e < m * m; e += NT) {
    const int a = e / m, b = e - a * m;
    if (b <= a) w.T[a * ld + b] = g.xx_eff(a, b, lam_eff) / s + (a == b ? w.p0[a] : 0.0);
  }
  for (int e = threadIdx.x; e < nrhs * m; e += NT) {
    const int t = e / m, a = e - t * m;
    const int kind = first_rhs_kind + t;
    double v;
    if (kind == 0) v = g.xy_eff(a, lam_eff) / s + w.p0[a] * w.mu0[a];
    else if (kind == 1) v = w.p0[a] * w.mu0[a] + g.Xy(a) / s;
    else v = w.p0[a] * w.mu0[a] + g.XWy(a) / s;
    w.T[(m + t) * ld + a] = v;
  }
  fill_identity<NT>(w.T, m + nrhs, m, ld);
  bsync<NT>();
  chol_bordered<NT>(w.T, w.dg, m, 2 * m + nrhs, ld);
}

// SAR: coefficients of rss(v) = e0e0 - 2 v e1e0 + v^2 e1e1 from b0, b1 (R/15_sar.R:110-115)
template <int NT>
__device__ void sar_coeffs(const GView &g, Work &w, int m, double &e0e0, double &e1e0, double &e1e1) {
  for (int a = threadIdx.x; a < m; a += NT) {
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < m; ++c) { s0 = fma(g.XX(a, c), w.b0[c], s0); s1 = fma(g.XX(a, c), w.b1[c], s1); }
    w.t0[a] = s0; w.t1[a] = s1;
  }
  bsync<NT>();
  double p00 = 0.0, p10 = 0.0, p11 = 0.0;
  for (int a = threadIdx.x; a < m; a += NT) {
    p00 = fma(w.b0[a], w.t0[a] - 2.0 * g.Xy(a), p00);
    p10 += w.b1[a] * w.t0[a] - w.b0[a] * g.XWy(a) - w.b1[a] * g.Xy(a);
    p11 = fma(w.b1[a], w.t1[a] - 2.0 * g.XWy(a), p11);
  }
  e0e0 = g.yy() + bsum<NT>(p00, w.red);
  e1e0 = g.yWy() + bsum<NT>(p10, w.red);
  e1e1 = g.WyWy() + bsum<NT>(p11, w.red);
}

// SEM: OLS rss of (y - v Wy) on (X - v WX), R/15_sem.R:116-120, via Cholesky of X_s'X_s
template <int NT>
__device__ double sem_rss(const GView &g, Work &w, double v, int m) {
  const int ld = work_ld(m);
  for (int e = threadIdx.x; e < m * m; e += NT) {
    const int a = e / m, b = e - a * m;
    if (b <= a) w.T[a * ld + b] = g.xx_eff(a, b, v);
  }
  for (int a = threadIdx.x; a < m; a += NT) w.T[m * ld + a] = g.xy_eff(a, v);
  bsync<NT>();
  chol_bordered<NT>(w.T, w.dg, m, m + 1, ld);
  double part = 0.0;
  for (int a = threadIdx.x; a < m; a += NT) { const double q = w.T[m * ld + a]; part = fma(q, q, part); }
  return g.yy_eff(v) - bsum<NT>(part, w.red);
}

// MH bookkeeping of R/20_mh.R:73-81
__device__ __forceinline__ void mh_count(long long *cnt, bool accepted) {
  if (accepted) { cnt[0] += 1; cnt[2] += 1; }
  cnt[1] += 1; cnt[3] += 1;
}
// R/50_execute.R:62-71 with set_scale<- zeroing accepted/proposed (R/20_mh.R:88-93)
__device__ __forceinline__ void mh_tune(double &scale, long long *cnt, const RunCtl &ctl) {
  const double rate = (double)cnt[2] / (double)(cnt[3] > 1 ? cnt[3] : 1);
  if (rate < ctl.acc_lo) { scale = fmax(scale * (1.0 - ctl.acc_change), 1e-12); cnt[0] = 0; cnt[1] = 0; }
  else if (rate > ctl.acc_hi) { scale = scale * (1.0 + ctl.acc_change); cnt[0] = 0; cnt[1] = 0; }
}

struct InitArgs {
  const double *beta0;   // device [m] or null
  double sigma0;         // NaN = least squares
  double lam0, scale0;
  double hs_lambda, hs_tau, hs_zeta, hs_nu, sng_lambda, sng_tau, sng_theta, sng_theta_scale;
  const double *p0_init; // device [m] NG precision diagonal
};

template <int NT>
__global__ void __launch_bounds__(NT) chain_kernel(DeviceData dd, Priors pr, ChainState cs, const RunCtl *ctlp,
                                                   double *draws, int mode, InitArgs ia) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, m = dd.m, n = dd.n;
  const int A = aux_len(m), ld = work_ld(m);
  Work w = carve(sm, m, dd.model);
  stage_gram<NT>(dd, w.G);
  const GView g{w.G, dd.Gfix + (size_t)dd.gcur * dd.d * dd.d, dd.Gscale, dd.d, m, dd.model};
  const bool conj = pr.prior == 1, hs = pr.prior == 3, sng = pr.prior == 2;
  const bool sar = dd.model == 1, sem = dd.model == 2;
  double *aux = cs.aux + (size_t)c * A;
  double *vecA = aux, *vecB = aux + m;
  const Philox ph(cs.seed, (uint32_t)(cs.chain_offset + c));

  double sigma2, lam, scale_l, scale_t;
  long long cnt[2][4];
  uint32_t it;
  const int step0 = (mode == MODE_STEP) ? cs.step[c] : 0;
  long long iters = (mode == MODE_INIT) ? 0 : cs.iters[c];

  for (int a = tid; a < m; a += NT) w.mu0[a] = cs.mu0[a];

  if (mode == MODE_INIT) {
    // ---- setup_4NG / setup_HS / setup_SNG / starting_NG ----
    lam = ia.lam0; scale_l = ia.scale0; scale_t = ia.sng_theta_scale;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) cnt[k][q] = 0;
    it = 0;
    for (int a = tid; a < m; a += NT) {
      double p = ia.p0_init[a];
      if (hs) { vecA[a] = ia.hs_lambda; vecB[a] = ia.hs_nu; p = 1.0 / (ia.hs_lambda * ia.hs_tau); }
      else if (sng) { vecA[a] = ia.sng_tau; vecB[a] = 0.0; p = 1.0 / ia.sng_tau; }
      else { vecA[a] = 0.0; vecB[a] = 0.0; }
      w.p0[a] = p;
    }
    if (tid == 0) {
      aux[2 * m] = hs ? ia.hs_tau : (sng ? ia.sng_lambda : 0.0);
      aux[2 * m + 1] = hs ? ia.hs_zeta : (sng ? ia.sng_theta : 0.0);
    }
    bsync<NT>();
    sigma2 = 1.0;
    if (!conj) {
      // beta = solve(XX*, Xy*), sigma^2 = RSS/(N-M): R/10_lm.R:157-163
      if (ia.beta0 != nullptr) {
        for (int a = tid; a < m; a += NT) w.beta[a] = ia.beta0[a];
      } else {
        for (int e = tid; e < m * m; e += NT) {
          const int a = e / m, b = e - a * m;
          if (b <= a) w.T[a * ld + b] = g.xx_eff(a, b, lam);
        }
        for (int a = tid; a < m; a += NT) w.T[m * ld + a] = g.xy_eff(a, lam);
        fill_identity<NT>(w.T, m + 1, m, ld);
        bsync<NT>();
        chol_bordered<NT>(w.T, w.dg, m, 2 * m + 1, ld);
        for (int j = tid; j < m; j += NT) w.beta[j] = upper_dot(w.T + (m + 1 + j) * ld, w.T + m * ld, j, m);
      }
      bsync<NT>();
      sigma2 = isnan(ia.sigma0) ? resid_ss<NT>(g, w.beta, lam, m, w.red) / (double)(n - m) : ia.sigma0;
    }
  } else {
    for (int a = tid; a < m; a += NT) {
      w.p0[a] = cs.p0[(size_t)c * m + a];
      w.beta[a] = cs.beta[(size_t)c * m + a];
      if (conj) w.mu1[a] = vecA[a];
    }
    sigma2 = cs.sigma2[c]; lam = cs.lam[c];
    scale_l = cs.mh_scale[2 * c]; scale_t = cs.mh_scale[2 * c + 1];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) cnt[k][q] = cs.mh_cnt[(size_t)c * 8 + k * 4 + q];
    it = (uint32_t)(iters + 1);
    bsync<NT>();
  }

  const bool do_sweep = (mode == MODE_STEP) || (mode == MODE_INIT && conj);   // Conjugate starts with one beta, sigma draw
  const bool reuse = sar && pr.prior == 0;     // P0 and sigma^2 unchanged between the beta and lambda steps
  bool have_b = false;

  if (do_sweep) {
    // ---- sample_sigma (skipped at init; Conjugate draws beta first, R/10_lm.R:218-221) ----
    if (mode == MODE_STEP) {
      const double rss = resid_ss<NT>(g, conj ? w.mu1 : w.beta, lam, m, w.red);
      sigma2 = 1.0 / ph.gamma(it, SLOT_SIGMA, pr.shape1, pr.rate0 + 0.5 * rss);
    }
    // ---- sample_beta ----
    const int nrhs = reuse ? 3 : 1;
    factor_system<NT>(g, w, lam, conj ? 1.0 : sigma2, m, nrhs, 0);
    for (int a = tid; a < m; a += NT) w.eps[a] = ph.normal(it, SLOT_BETA, (uint32_t)a);
    bsync<NT>();
    for (int j = tid; j < m; j += NT) {
      const double *Y = w.T + (m + nrhs + j) * ld;
      const double mu = upper_dot(Y, w.T + m * ld, j, m);
      w.mu1[j] = mu;
      w.beta[j] = mu + upper_dot(Y, w.eps, j, m);
      if (reuse) {
        w.b0[j] = upper_dot(Y, w.T + (m + 1) * ld, j, m);
        w.b1[j] = upper_dot(Y, w.T + (m + 2) * ld, j, m);
      }
    }
    have_b = reuse;
    bsync<NT>();
    if (mode == MODE_INIT) {   // Conjugate starting sigma draw, R/10_lm.R:220
      const double rss = resid_ss<NT>(g, w.mu1, lam, m, w.red);
      sigma2 = 1.0 / ph.gamma(it, SLOT_SIGMA, pr.shape1, pr.rate0 + 0.5 * rss);
    }
  }

  if (mode == MODE_STEP) {
    // ---- sample_shrinkage ----
    if (hs) {
      const double tau_old = aux[2 * m], zeta_old = aux[2 * m + 1];
      double part = 0.0;
      for (int a = tid; a < m; a += NT) {
        const double b2 = w.beta[a] * w.beta[a];
        const double l = 1.0 / ph.exponential(it, SLOT_HS_LAMBDA, 1.0 / vecB[a] + b2 / (2.0 * tau_old * sigma2), (uint32_t)a);
        w.t0[a] = l;
        part += b2 / l;
      }
      const double ssum = bsum<NT>(part, w.red);
      const double tau = 1.0 / ph.gamma(it, SLOT_HS_TAU, 0.5 * (double)(m + 1), 1.0 / zeta_old + ssum / (2.0 * sigma2));
      const double zeta = 1.0 / ph.exponential(it, SLOT_HS_ZETA, 1.0 + 1.0 / tau);
      for (int a = tid; a < m; a += NT) {
        const double l = w.t0[a];
        vecA[a] = l;
        vecB[a] = 1.0 / ph.exponential(it, SLOT_HS_NU, 1.0 + 1.0 / l, (uint32_t)a);
        w.p0[a] = 1.0 / fmax(l * tau, 1e-12);
      }
      bsync<NT>();
      if (tid == 0) { aux[2 * m] = tau; aux[2 * m + 1] = zeta; }
    } else if (sng) {
      double theta = aux[2 * m + 1];
      double part = 0.0;
      for (int a = tid; a < m; a += NT) part += vecA[a];
      const double stau = bsum<NT>(part, w.red);
      const double lam_s = ph.gamma(it, SLOT_SNG_LAMBDA, pr.sng_lambda_a + (double)m * theta,
                                    pr.sng_lambda_b + theta * stau / 2.0);
      double p_tau = 0.0, p_log = 0.0;
      for (int a = tid; a < m; a += NT) {
        const double dlt = w.beta[a] - w.mu0[a];
        const double t = fmax(ph.gig(it, SLOT_SNG_TAU, theta - 0.5, dlt * dlt, lam_s * theta, (uint32_t)a), 1e-12);
        vecA[a] = t;
        w.p0[a] = 1.0 / t;
        p_tau += t; p_log += log(t);
      }
      bsync<NT>();
      if (pr.sng_theta_scale > 0.0) {     // MH_SNG_theta, R/20_mh.R:116-146
        const double st = bsum<NT>(p_tau, w.red), sl = bsum<NT>(p_log, w.red);
        const double prop = exp(scale_t * ph.normal(it, SLOT_SNG_THETA_PROP)) * theta;
        auto post = [&](double th) {
          const double r = th * lam_s / 2.0;
          return (double)m * (th * log(r) - lgamma(th)) + (th - 1.0) * sl - r * st + log(pr.sng_theta_a) - pr.sng_theta_a * th;
        };
        const double lp = post(prop) - post(theta) + log(prop) - log(theta);
        const bool acc = ph.uniform(it, SLOT_SNG_THETA_ACC) < exp(lp);
        if (acc) theta = prop;
        mh_count(cnt[1], acc);
      }
      if (tid == 0) { aux[2 * m] = lam_s; aux[2 * m + 1] = theta; }
    }
  }

  // ---- spatial parameter ----
  double ldet_cur = aux[2 * m + 2], rss_cur = aux[2 * m + 3];
  if (mode != MODE_STEP && (sar || sem)) {
    ldet_cur = logdet_nodes<NT>(dd, lam, w.red);
    rss_cur = sem ? sem_rss<NT>(g, w, lam, m) : 0.0;
  }
  if (mode == MODE_STEP && sar) {
    if (!have_b) {
      factor_system<NT>(g, w, 0.0, sigma2, m, 2, 1);
      for (int j = tid; j < m; j += NT) {
        const double *Y = w.T + (m + 2 + j) * ld;
        w.b0[j] = upper_dot(Y, w.T + m * ld, j, m);
        w.b1[j] = upper_dot(Y, w.T + (m + 1) * ld, j, m);
      }
      bsync<NT>();
    }
    double e0e0, e1e0, e1e1;
    sar_coeffs<NT>(g, w, m, e0e0, e1e0, e1e1);
    const double prop = ph.truncated_rw(it, SLOT_LAMBDA_PROP, lam, scale_l, pr.lam_lo, pr.lam_hi);
    const double ldet_prop = logdet_nodes<NT>(dd, prop, w.red);
    const double rss_p = e0e0 - 2.0 * prop * e1e0 + prop * prop * e1e1;
    const double rss_c = e0e0 - 2.0 * lam * e1e0 + lam * lam * e1e1;
    const double p = exp(logpost_lambda(prop, ldet_prop, rss_p, n, m, pr) - logpost_lambda(lam, ldet_cur, rss_c, n, m, pr));
    const bool acc = ph.uniform(it, SLOT_LAMBDA_ACC) < p;     // NaN compares false: R/20_mh.R:74
    if (acc) { lam = prop; ldet_cur = ldet_prop; }
    mh_count(cnt[0], acc);
  } else if (mode == MODE_STEP && sem) {
    const double prop = ph.truncated_rw(it, SLOT_RHO_PROP, lam, scale_l, pr.lam_lo, pr.lam_hi);
    const double ldet_prop = logdet_nodes<NT>(dd, prop, w.red);
    const double rss_p = sem_rss<NT>(g, w, prop, m);
    const double p = exp(logpost_lambda(prop, ldet_prop, rss_p, n, m, pr) - logpost_lambda(lam, ldet_cur, rss_cur, n, m, pr));
    const bool acc = ph.uniform(it, SLOT_RHO_ACC) < p;
    if (acc) { lam = prop; ldet_cur = ldet_prop; rss_cur = rss_p; }
    mh_count(cnt[0], acc);
  }

  // ---- finalize, tuning, draw store ----
  if (mode == MODE_STEP) {
    iters += 1;
    const RunCtl ctl = *ctlp;
    const int step = step0 + 1;
    if (step <= ctl.n_burn && step % 10 == 0 && step <= ctl.tune_limit) {
      if (sar || sem) mh_tune(scale_l, cnt[0], ctl);
      if (sng && pr.sng_theta_scale > 0.0) mh_tune(scale_t, cnt[1], ctl);
    }
    if (step > ctl.n_burn && ctl.store) {
      const int row = step - ctl.n_burn - 1 - ctl.save_base;
      if (row >= 0 && row < ctl.save_cap) {
        const int P = m + 1 + ((sar || sem) ? 1 : 0);
        double *dst = draws + (size_t)c * P * ctl.save_cap + row;
        for (int a = tid; a < m; a += NT) dst[(size_t)a * ctl.save_cap] = w.beta[a];
        if (tid == 0) {
          dst[(size_t)m * ctl.save_cap] = sqrt(sigma2);
          if (sar || sem) dst[(size_t)(m + 1) * ctl.save_cap] = lam;
        }
      }
    }
    if (tid == 0) cs.step[c] = step;
  }
  for (int a = tid; a < m; a += NT) {
    cs.beta[(size_t)c * m + a] = w.beta[a];
    cs.p0[(size_t)c * m + a] = w.p0[a];
    if (conj && mode != MODE_REFRESH) vecA[a] = w.mu1[a];
  }
  if (tid == 0) {
    cs.sigma2[c] = sigma2; cs.lam[c] = lam;
    cs.mh_scale[2 * c] = scale_l; cs.mh_scale[2 * c + 1] = scale_t;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) cs.mh_cnt[(size_t)c * 8 + k * 4 + q] = cnt[k][q];
    cs.iters[c] = iters;
    aux[2 * m + 2] = ldet_cur; aux[2 * m + 3] = rss_cur;
  }
}

// ---------------------------------------------------------------------------
// Fast chain step: Independent prior, M <= FAST_MAX_M, LM / SAR / SEM.  Same arithmetic as
// chain_kernel, re-scheduled so that only the sigma^2 -> beta -> lambda dependency chain is
// serial.  One CTA of 4 warps per chain:
//   warps 0-1  critical path: rss(beta) -> sigma^2 -> P1 -> bordered Cholesky (row = thread,
//              pivot recomputed by every thread, one barrier per column) -> solves
//   warp 2     random numbers that do not depend on this iteration's draws: the unit-rate gamma
//              behind sigma^2, the M normals of the beta draw, the acceptance uniform
//   warp 3     lambda proposal, log|I - v W| and log prior at the proposal (SEM: also the OLS
//              rss at the proposal, its own (M+1)-row Cholesky)
// ---------------------------------------------------------------------------
#ifdef BSREG_TRACE
__device__ long long g_trace[32];
#define TRACE(i) do { if (blockIdx.x == 0 && tid == 0) g_trace[i] = clock64(); } while (0)
#else
#define TRACE(i) do { } while (0)
#endif
constexpr int FAST_MAX_M = 30;          // 2M + 3 bordered rows fit two warps
constexpr int FAST_NT = 128;

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

// sum_{k=j}^{m-1} Y[k] v[k] with four independent accumulators (FP64 latency, not throughput, binds here)
__device__ __forceinline__ double dot_from(const double *Y, const double *v, int j, int m) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = j;
  for (; k + 3 < m; k += 4) {
    s0 = fma(Y[k], v[k], s0); s1 = fma(Y[k + 1], v[k + 1], s1); s2 = fma(Y[k + 2], v[k + 2], s2); s3 = fma(Y[k + 3], v[k + 3], s3);
  }
  for (; k < m; ++k) s0 = fma(Y[k], v[k], s0);
  return (s0 + s1) + (s2 + s3);
}

enum { X_UNIT = 0, X_SIGMA2, X_SD, X_UACC, X_PROP, X_LDETP, X_LPP, X_LPC, X_RSSP, X_S0, X_S1, X_S2, X_COUNT = 16 };

template <int MMAX>
__host__ __device__ inline size_t fast_doubles(int m, int model) {
  const size_t d = gram_dim(m, model), ld = work_ld(m);
  return d * d + 2 * 64 + (size_t)64 * (MMAX | 1) + 3 * MMAX + 2 * MMAX + (size_t)(m + 1) * ld + 10 * (size_t)m + X_COUNT;
}

// The K x K system is solved in the sigma^2-scaled form Q = X*'X* + sigma^2 P0 (so nothing on the
// critical path divides by sigma^2; only the diagonal and the right-hand sides wait for the draw):
//   mu1 = Q^-1 (X*'y* + sigma^2 P0 mu0),  beta = mu1 + sigma Q^-1/2 eps,  b0, b1 alike (R/15_sar.R:105-109).
// Q = L D L' is eliminated right-looking with ONE ROW PER THREAD IN REGISTERS, bordered by the
// right-hand sides and an identity block (so z = L^-1 rhs and D^-1 L^-1 fall out of the same sweep);
// per column: publish the column, one barrier, reciprocal of the pivot, rank-1 update.
template <int MMAX>
__global__ void __launch_bounds__(FAST_NT) chain_fast_kernel(DeviceData dd, Priors pr, ChainState cs, const RunCtl *ctlp,
                                                            double *draws) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, m = dd.m, n = dd.n, d = dd.d;
  const int ld = work_ld(m), A = aux_len(m);
  constexpr int LDY = MMAX | 1;
  const bool sar = dd.model == 1, sem = dd.model == 2;
  double *G = sm, *col = G + d * d, *Yd = col + 128, *Z = Yd + 64 * LDY, *dvec = Z + 3 * MMAX, *sq = dvec + MMAX;   // Yd doubles as the staging area of the bordered rows
  double *T2 = sq + MMAX, *p0 = T2 + (m + 1) * ld, *mu0 = p0 + m, *beta = mu0 + m, *mu1 = beta + m, *bd = mu1 + m;
  double *b0 = bd + m, *b1 = b0 + m, *eps = b1 + m, *t0 = eps + m, *t1 = t0 + m, *x = t1 + m;
  double *aux = cs.aux + (size_t)c * A;
  const Philox ph(cs.seed, (uint32_t)(cs.chain_offset + c));
  const GView g{G, nullptr, nullptr, d, m, dd.model};
  const int nrhs = sar ? 3 : 1, nrows = 2 * m + nrhs;

  TRACE(0);
  // ---- phase 0: one round trip to L2 for the state and the Gram matrix; warp 2 draws meanwhile ----
  const long long iters0 = cs.iters[c];
  const uint32_t it = (uint32_t)(iters0 + 1);
  const double lam = cs.lam[c];
  if (wid == 2) {
    // random numbers that do not depend on this iteration's state (R/10_lm.R:178, R/00_aux.R:65, R/20_mh.R:74)
    const double unit = ph.gamma(it, SLOT_SIGMA, pr.shape1, 1.0);
    if (lane == 0) x[X_UNIT] = unit;
  } else {
    const int t3 = wid == 3 ? 64 + lane : tid;                       // 96 staging threads
    const long long *gf = dd.Gfix + (size_t)dd.gcur * d * d;
    for (int i = t3 >> 5; i < d; i += 3)
      for (int j = lane; j < d; j += 32) {
        const int e = i <= j ? i * d + j : j * d + i;
        G[i * d + j] = (double)gf[e] * dd.Gscale[e];
      }
    for (int a = t3; a < m; a += 96) {
      p0[a] = cs.p0[(size_t)c * m + a];
      beta[a] = cs.beta[(size_t)c * m + a];
      mu0[a] = cs.mu0[a];
    }
  }
  const double scale_l0 = cs.mh_scale[2 * c];
  const int step0 = cs.step[c];
  double ldet_cur = aux[2 * m + 2], rss_cur = aux[2 * m + 3];
  __syncthreads();
  TRACE(1);

  double a[MMAX];                                                    // this thread's row of the bordered system
  if (wid != 3) {
    // sigma^2-free part of [Q; rhs'; I], built cooperatively (no divergence), one row per later owner
    const int t3 = tid;                                              // 96 threads
    for (int r = t3 >> 5; r < nrows; r += 3)
      for (int j = lane; j < m; j += 32) {
        double v;
        if (r < m) v = j <= r ? g.xx_eff(r, j, lam) : 0.0;
        else if (r == m) v = g.xy_eff(j, lam);
        else if (r < m + nrhs) v = r == m + 1 ? g.Xy(j) : g.XWy(j);
        else v = (r - m - nrhs == j) ? 1.0 : 0.0;
        Yd[r * LDY + j] = v;
      }
  }
  if (wid == 2) {
    __threadfence_block();
    named_arrive(3, 96);
    if (lane < m) eps[lane] = ph.normal(it, SLOT_BETA, (uint32_t)lane);
    if (lane == 31 && (sar || sem)) x[X_UACC] = ph.uniform(it, sar ? SLOT_LAMBDA_ACC : SLOT_RHO_ACC);
  } else if (wid == 3) {
    // ---- sigma^2 (R/10_lm.R:174-179), then the proposal side of the Metropolis-Hastings step ----
    double part = 0.0;
    if (lane < m) {
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
      int cc = 0;
      for (; cc + 3 < m; cc += 4) {
        u0 = fma(g.xx_eff(lane, cc, lam), beta[cc], u0); u1 = fma(g.xx_eff(lane, cc + 1, lam), beta[cc + 1], u1);
        u2 = fma(g.xx_eff(lane, cc + 2, lam), beta[cc + 2], u2); u3 = fma(g.xx_eff(lane, cc + 3, lam), beta[cc + 3], u3);
      }
      for (; cc < m; ++cc) u0 = fma(g.xx_eff(lane, cc, lam), beta[cc], u0);
      part = beta[lane] * (((u0 + u1) + (u2 + u3)) - 2.0 * g.xy_eff(lane, lam));
    }
    const double rss = g.yy_eff(lam) + warp_sum_c(part);
    const double sigma2 = (pr.rate0 + 0.5 * rss) / x[X_UNIT];        // 1 / Gamma(shape1, rate1)
    if (lane == 0) x[X_SIGMA2] = sigma2;
    __threadfence_block();
    named_arrive(1, 96);
    if (lane == 0) x[X_SD] = sqrt(sigma2);
    if (sar || sem) {
      const double prop = ph.truncated_rw(it, sar ? SLOT_LAMBDA_PROP : SLOT_RHO_PROP, lam, scale_l0, pr.lam_lo, pr.lam_hi);
      double acc = 0.0;
      for (int k = lane; k < dd.n_nodes; k += 32) {
        const double aa = 1.0 - prop * dd.node_re[k], bb = prop * dd.node_im[k];
        acc = fma(dd.node_w[k], 0.5 * log(aa * aa + bb * bb), acc);
      }
      const double ldet_prop = warp_sum_c(acc);
      const double lp = lane == 0 ? log_prior_lambda(prop, pr) : log_prior_lambda(lam, pr);
      if (lane == 0) { x[X_PROP] = prop; x[X_LDETP] = ldet_prop; x[X_LPP] = lp; }
      if (lane == 1) x[X_LPC] = lp;
      if (sem) {   // OLS rss of (y - v Wy) on (X - v WX) at the proposal, R/15_sem.R:116-120
        for (int e = lane; e < m * m; e += 32) {
          const int r = e / m, b = e - r * m;
          if (b <= r) T2[r * ld + b] = g.xx_eff(r, b, prop);
        }
        for (int r = lane; r < m; r += 32) T2[m * ld + r] = g.xy_eff(r, prop);
        __syncwarp();
        chol_bordered<32>(T2, t0, m, m + 1, ld, lane);               // t0 is free until phase 3 (SAR only)
        double q2 = 0.0;
        for (int r = lane; r < m; r += 32) { const double q = T2[m * ld + r]; q2 = fma(q, q, q2); }
        const double r = g.yy_eff(prop) - warp_sum_c(q2);
        if (lane == 0) x[X_RSSP] = r;
      }
    }
  } else {
    // ---- critical path: one bordered row per thread, in registers ----
    const int r = tid;
    named_sync(3, 96);                                               // staged rows complete
    TRACE(2);
#pragma unroll
    for (int j = 0; j < MMAX; ++j) a[j] = (j < m && r < nrows) ? Yd[r * LDY + j] : 0.0;
    named_sync(1, 96);                                               // sigma^2 from warp 3
    TRACE(3);
    const double s2 = x[X_SIGMA2];
#pragma unroll
    for (int j = 0; j < MMAX; ++j)
      if (j < m) {
        if (r == j) a[j] = fma(s2, p0[j], a[j]);
        else if (r >= m && r < m + nrhs) a[j] = fma(s2, p0[j] * mu0[j], a[j]);
      }
#pragma unroll
    for (int k = 0; k < MMAX; ++k) {
      if (k < m) {
        double *cb = col + 64 * (k & 1);
        cb[r] = a[k];
        if (r >= m && r < m + nrhs) Z[(r - m) * MMAX + k] = a[k];   // z = L^-1 rhs, before scaling
        if (r == k) dvec[k] = a[k];                                 // pivot d_k
        named_sync(2, 64);
        const double rk = __drcp_rn(cb[k]);
        if (r > k && r < nrows) {
          const double l = a[k] * rk;
          a[k] = l;
#pragma unroll
          for (int j = k + 1; j < MMAX; ++j)
            if (j < m) a[j] = fma(-l, cb[j], a[j]);
        }
      }
    }
    TRACE(4);
    if (r >= m + nrhs && r < nrows) {
      double *y = Yd + (r - m - nrhs) * LDY;
#pragma unroll
      for (int k = 0; k < MMAX; ++k) if (k < m) y[k] = a[k];
    }
  }
  __syncthreads();
  TRACE(5);

  // ---- phase 2: mu1, the beta draw, b0, b1 (one vector per warp, one entry per lane) ----
  if (wid == 1) {                                                    // D^1/2 eps for the beta draw
    if (lane < m) sq[lane] = sqrt(dvec[lane]) * eps[lane];
    __syncwarp();
  }
  if (lane < m && (wid < 2 || sar)) {
    const double *v = wid == 0 ? Z : wid == 1 ? sq : Z + (wid - 1) * MMAX;
    const double r = dot_from(Yd + lane * LDY, v, lane, m);
    (wid == 0 ? mu1 : wid == 1 ? bd : wid == 2 ? b0 : b1)[lane] = r;
  }
  __syncthreads();
  const double sigma2 = x[X_SIGMA2];
  if (tid < m) beta[tid] = fma(x[X_SD], bd[tid], mu1[tid]);
  TRACE(6);

  // ---- phase 3: the Metropolis-Hastings decision ----
  double lam_new = lam, scale_l = scale_l0;
  long long cnt[4];
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) cnt[q] = cs.mh_cnt[(size_t)c * 8 + q];
  }
  if (sar) {
    if (wid < 2 && lane < m) {      // t0 = X'X b0, t1 = X'X b1
      const double *b = wid == 0 ? b0 : b1;
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
      int cc = 0;
      for (; cc + 3 < m; cc += 4) {
        u0 = fma(g.XX(lane, cc), b[cc], u0); u1 = fma(g.XX(lane, cc + 1), b[cc + 1], u1);
        u2 = fma(g.XX(lane, cc + 2), b[cc + 2], u2); u3 = fma(g.XX(lane, cc + 3), b[cc + 3], u3);
      }
      for (; cc < m; ++cc) u0 = fma(g.XX(lane, cc), b[cc], u0);
      (wid == 0 ? t0 : t1)[lane] = (u0 + u1) + (u2 + u3);
    }
    __syncthreads();
    if (wid < 3) {                  // e0'e0, e1'e0, e1'e1 of R/15_sar.R:110-114, one per warp
      double part = 0.0;
      if (lane < m) {
        if (wid == 0) part = b0[lane] * (t0[lane] - 2.0 * g.Xy(lane));
        else if (wid == 1) part = b1[lane] * t0[lane] - b0[lane] * g.XWy(lane) - b1[lane] * g.Xy(lane);
        else part = b1[lane] * (t1[lane] - 2.0 * g.XWy(lane));
      }
      const double tot = (wid == 0 ? g.yy() : wid == 1 ? g.yWy() : g.WyWy()) + warp_sum_c(part);
      if (lane == 0) x[X_S0 + wid] = tot;
    }
    __syncthreads();
  }
  TRACE(7);
  if ((sar || sem) && wid == 0) {
    const double prop = x[X_PROP], ldet_prop = x[X_LDETP];
    double rss_p, rss_c;
    if (sar) {
      const double e0e0 = x[X_S0], e1e0 = x[X_S1], e1e1 = x[X_S2];
      rss_p = e0e0 - 2.0 * prop * e1e0 + prop * prop * e1e1;
      rss_c = e0e0 - 2.0 * lam * e1e0 + lam * lam * e1e1;
    } else {
      rss_p = x[X_RSSP];
      rss_c = rss_cur;
    }
    // lane 0: proposal, lane 1: current value (the two log rss side by side)
    const double rr = lane == 0 ? rss_p : rss_c;
    const double lpl = lane == 0 ? x[X_LPP] : x[X_LPC];
    const double ld_ = lane == 0 ? ldet_prop : ldet_cur;
    const double post = !(rr > 0.0) ? nan("") : ld_ - 0.5 * (double)(n - m) * log(rr) + lpl;
    const double post_c = __shfl_sync(0xffffffffu, post, 1);
    if (lane == 0) {
      const double pacc = exp(post - post_c);
      const bool acc = x[X_UACC] < pacc;                  // NaN compares false: R/20_mh.R:74
      if (acc) { lam_new = prop; ldet_cur = ldet_prop; if (sem) rss_cur = rss_p; }
      mh_count(cnt, acc);
    }
  }
  TRACE(8);

  // ---- finalize, tuning, draw store (thread 0 owns the scalars) ----
  const RunCtl ctl = *ctlp;
  const int step = step0 + 1;
  const bool storing = step > ctl.n_burn && ctl.store;
  const int row = step - ctl.n_burn - 1 - ctl.save_base;
  const int P = m + 1 + ((sar || sem) ? 1 : 0);
  double *dst = draws + (size_t)c * P * ctl.save_cap + row;
  const bool row_ok = storing && row >= 0 && row < ctl.save_cap;
  if (tid < m) {
    const double b = beta[tid];
    cs.beta[(size_t)c * m + tid] = b;
    if (row_ok) dst[(size_t)tid * ctl.save_cap] = b;
  }
  if (tid == 0) {
    if (step <= ctl.n_burn && step % 10 == 0 && step <= ctl.tune_limit && (sar || sem)) mh_tune(scale_l, cnt, ctl);
    if (row_ok) {
      dst[(size_t)m * ctl.save_cap] = x[X_SD];
      if (sar || sem) dst[(size_t)(m + 1) * ctl.save_cap] = lam_new;
    }
    cs.step[c] = step;
    cs.sigma2[c] = sigma2; cs.lam[c] = lam_new;
    cs.mh_scale[2 * c] = scale_l;
#pragma unroll
    for (int q = 0; q < 4; ++q) cs.mh_cnt[(size_t)c * 8 + q] = cnt[q];
    cs.iters[c] = iters0 + 1;
    aux[2 * m + 2] = ldet_cur; aux[2 * m + 3] = rss_cur;
  }
  TRACE(9);
}

#ifdef BSREG_TRACE
extern "C" int bsreg_debug_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 32); }
#endif

bool chain_fast_supported(const DeviceData &dd, const Priors &pr) {
  return pr.prior == 0 && dd.m <= FAST_MAX_M && g_staged(dd.m, dd.model);
}

size_t chain_smem_bytes(int m, int model, int) { return work_doubles(m, model) * sizeof(double); }
int chain_threads(int m) { return m <= 32 ? 32 : 128; }

template <int NT>
static void launch_chain_nt(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                            int mode, const InitArgs &ia, cudaStream_t s) {
  const size_t smem = chain_smem_bytes(dd.m, dd.model, pr.prior);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(chain_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  chain_kernel<NT><<<cs.n_chains, NT, smem, s>>>(dd, pr, cs, ctl, draws, mode, ia);
}

static void launch_chain(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                         int mode, const InitArgs &ia, cudaStream_t s) {
  if (chain_threads(dd.m) == 32) launch_chain_nt<32>(dd, pr, cs, ctl, draws, mode, ia, s);
  else launch_chain_nt<128>(dd, pr, cs, ctl, draws, mode, ia, s);
}

void launch_chain_init(const DeviceData &dd, const Priors &pr, const ChainState &cs, const ChainInit &ci, cudaStream_t s) {
  InitArgs ia{ci.beta0, ci.sigma0, ci.lam0, ci.scale0, ci.hs_lambda, ci.hs_tau, ci.hs_zeta, ci.hs_nu,
              ci.sng_lambda, ci.sng_tau, ci.sng_theta, ci.sng_theta_scale, ci.p0_init};
  launch_chain(dd, pr, cs, nullptr, nullptr, ci.refresh_only ? MODE_REFRESH : MODE_INIT, ia, s);
}

void launch_chain_step(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                       cudaStream_t s) {
  static const bool no_fast = getenv("BSREG_CHAIN_GENERIC") != nullptr;   // A/B switch for tests and profiling
  if (!no_fast && chain_fast_supported(dd, pr)) {
    if (dd.m <= 24) chain_fast_kernel<24><<<cs.n_chains, FAST_NT, fast_doubles<24>(dd.m, dd.model) * sizeof(double), s>>>(dd, pr, cs, ctl, draws);
    else chain_fast_kernel<32><<<cs.n_chains, FAST_NT, fast_doubles<32>(dd.m, dd.model) * sizeof(double), s>>>(dd, pr, cs, ctl, draws);
    return;
  }
  InitArgs ia{};
  launch_chain(dd, pr, cs, ctl, draws, MODE_STEP, ia, s);
}

// ---------------------------------------------------------------------------
// parity kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) eval_logdet_kernel(DeviceData dd, const double *v, double *out) {
  __shared__ double red[8];
  const double r = logdet_nodes<128>(dd, v[blockIdx.x], red);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

void launch_eval_logdet(const DeviceData &dd, int n_eval, const double *v, double *out, cudaStream_t s) {
  eval_logdet_kernel<<<n_eval, 128, 0, s>>>(dd, v, out);
}

template <int NT>
__global__ void __launch_bounds__(NT) eval_cond_kernel(DeviceData dd, Priors pr, const double *mu0, const double *beta,
                                                       const double *sigma2, const double *lam, const double *p0,
                                                       const double *v, double *mu1, double *rate1, double *rss,
                                                       double *logpost) {
  extern __shared__ double sm[];
  const int e = blockIdx.x, tid = threadIdx.x, m = dd.m, ld = work_ld(m);
  Work w = carve(sm, m, dd.model);
  stage_gram<NT>(dd, w.G);
  const GView g{w.G, dd.Gfix + (size_t)dd.gcur * dd.d * dd.d, dd.Gscale, dd.d, m, dd.model};
  const bool conj = pr.prior == 1;
  for (int a = tid; a < m; a += NT) {
    w.mu0[a] = mu0[a]; w.p0[a] = p0[(size_t)e * m + a]; w.beta[a] = beta[(size_t)e * m + a];
  }
  bsync<NT>();
  const double s2 = sigma2[e], l = lam[e], vv = v[e];
  // sigma^2 rate at (beta, lambda)
  const double r = resid_ss<NT>(g, w.beta, l, m, w.red);
  if (tid == 0) rate1[e] = pr.rate0 + 0.5 * r;
  // beta conditional mean at (sigma^2, lambda, P0)
  factor_system<NT>(g, w, l, conj ? 1.0 : s2, m, 1, 0);
  for (int j = tid; j < m; j += NT) mu1[(size_t)e * m + j] = upper_dot(w.T + (m + 1 + j) * ld, w.T + m * ld, j, m);
  bsync<NT>();
  double rs = nan(""), lp = nan("");
  if (dd.model == 1) {
    factor_system<NT>(g, w, 0.0, s2, m, 2, 1);
    for (int j = tid; j < m; j += NT) {
      const double *Y = w.T + (m + 2 + j) * ld;
      w.b0[j] = upper_dot(Y, w.T + m * ld, j, m);
      w.b1[j] = upper_dot(Y, w.T + (m + 1) * ld, j, m);
    }
    bsync<NT>();
    double e0e0, e1e0, e1e1;
    sar_coeffs<NT>(g, w, m, e0e0, e1e0, e1e1);
    rs = e0e0 - 2.0 * vv * e1e0 + vv * vv * e1e1;
  } else if (dd.model == 2) {
    rs = sem_rss<NT>(g, w, vv, m);
  }
  if (dd.model != 0) {
    const double ldv = logdet_nodes<NT>(dd, vv, w.red);
    lp = logpost_lambda(vv, ldv, rs, dd.n, m, pr);
  }
  if (tid == 0) { rss[e] = rs; logpost[e] = lp; }
}

void launch_eval_conditionals(const DeviceData &dd, const Priors &pr, const double *mu0, int n_eval, const double *beta,
                              const double *sigma2, const double *lam, const double *p0, const double *v, double *mu1,
                              double *rate1, double *rss, double *logpost, cudaStream_t s) {
  const size_t smem = chain_smem_bytes(dd.m, dd.model, pr.prior);
  if (chain_threads(dd.m) == 32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(eval_cond_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eval_cond_kernel<32><<<n_eval, 32, smem, s>>>(dd, pr, mu0, beta, sigma2, lam, p0, v, mu1, rate1, rss, logpost);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(eval_cond_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eval_cond_kernel<128><<<n_eval, 128, smem, s>>>(dd, pr, mu0, beta, sigma2, lam, p0, v, mu1, rate1, rss, logpost);
  }
}

__global__ void eval_rng_kernel(int kind, uint64_t seed, int chain, int slot, int it0, int n, const double *p, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Philox ph(seed, (uint32_t)chain);
  const uint32_t it = (uint32_t)(it0 + i);
  double r;
  switch (kind) {
    case 0: r = ph.uniform(it, slot); break;
    case 1: r = ph.normal(it, slot); break;
    case 2: r = ph.gamma(it, slot, p[0], p[1]); break;
    case 3: r = ph.exponential(it, slot, p[0]); break;
    case 4: r = ph.gig(it, slot, p[0], p[1], p[2]); break;
    default: r = ph.truncated_rw(it, slot, p[0], p[1], p[2], p[3]); break;
  }
  out[i] = r;
}

void launch_eval_rng(int kind, uint64_t seed, int chain, int slot, int it0, int n, const double *params, double *out,
                     cudaStream_t s) {
  eval_rng_kernel<<<(n + 127) / 128, 128, 0, s>>>(kind, seed, chain, slot, it0, n, params, out);
}

}  // namespace bsreg
