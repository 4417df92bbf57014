// Shared device pieces of the per-chain Gibbs step (chain.cu kernels and the persistent kernel in persist.cu).
#pragma once
#include "internal.cuh"
#include "rng.cuh"

namespace bsreg {

enum { MODE_STEP = 0, MODE_INIT = 1, MODE_REFRESH = 2 };

#ifdef BSREG_TRACE
static __device__ long long g_trace[64];
#define TRACE(i) do { if (blockIdx.x == 0 && tid == 0) g_trace[i] = clock64(); } while (0)
#define FTRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_trace[i] = clock64(); } while (0)
#else
#define TRACE(i) do { } while (0)
#define FTRACE(i) do { } while (0)
#endif

__device__ __forceinline__ double warp_sum_c(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NT>
__device__ __forceinline__ void bsync() {
  // a named barrier over the chain's NT threads (threads 0 .. NT - 1 of the CTA): the same code runs as a kernel of NT threads
  // and inside the chain CTAs of the persistent kernel, whose other warps have retired
  if (NT == 32) __syncwarp(); else asm volatile("bar.sync 15, %0;\n" ::"n"(NT) : "memory");
}

// fixed-order block sum, result in every thread
template <int NT>
__device__ __forceinline__ double bsum(double v, double *red) {
  v = warp_sum_c(v);
  if (NT == 32) return v;
  bsync<NT>();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  bsync<NT>();
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
// Left-looking: a dot product of length k per row and column, two block barriers per column; any m, any NT.
template <int NT>
__device__ void chol_bordered_ll(double *T, double *dg, int m, int nrows, int ld, int tid = threadIdx.x) {
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

// Right-looking form for one bordered row per thread (nrows <= NT, m <= MMAX), the row in registers: column k costs ONE
// block barrier and no dependent chain of length k.  A dependent FP64 operation takes ~22 cycles here, so what counts is
// the number of operations between two barriers: every thread forms the reciprocal square root of the pivot itself (from the
// published diagonal of the pivot row: rsqrt() is one MUFU seed + a short Newton tail, ~100 cycles, where sqrt followed by a
// division is ~180 and a hand-written Newton iteration from the single-precision seed ~470), scales its entry, the rows of A
// publish theirs, and the trailing entries are updated with independent FMAs.  20 600 -> ~6 000 cycles per factorisation at
// M = 20 (clock64 trace, profiles/r2_chain_general_trace.log).
// pub: [2][MMAX + 2] doubles of shared memory (double buffered by the parity of the column).
template <int NT, int MMAX>
__device__ void chol_bordered_rr(double *T, double *pub, int m, int nrows, int ld) {
  constexpr int LP = MMAX + 2;
  const int r = threadIdx.x;
  const bool on = r < nrows;
  double a[MMAX];
#pragma unroll
  for (int j = 0; j < MMAX; ++j) a[j] = (on && j < m && (r >= m || j <= r)) ? T[r * ld + j] : 0.0;
  double dk = T[0];
#pragma unroll
  for (int k = 0; k < MMAX; ++k) {
    if (k < m) {
      double *pb = pub + (k & 1) * LP;
      const double rinv = rsqrt(dk);
      const double l = r == k ? dk * rinv : a[k] * rinv;
      a[k] = l;
      if (r > k && r < m) pb[r] = l;
      if (r == k + 1) pb[MMAX] = a[k + 1 < MMAX ? k + 1 : k];          // the next pivot row's diagonal, before this column's update
      bsync<NT>();
      if (k + 1 < m) {
        const double l1 = pb[k + 1 < MMAX ? k + 1 : k];
        dk = fma(-l1, l1, pb[MMAX]);                                   // the same arithmetic as row k + 1's own update below
      }
      if (r > k) {
#pragma unroll
        for (int j = k + 1; j < MMAX; ++j) a[j] = fma(-l, pb[j], a[j]);   // columns >= m carry garbage nobody reads
      }
    }
  }
  if (on) {
#pragma unroll
    for (int j = 0; j < MMAX; ++j) if (j < m && (r >= m || j <= r)) T[r * ld + j] = a[j];
  }
  bsync<NT>();
}

constexpr int CHOL_RR_MMAX = 24;
template <int NT>
__device__ __forceinline__ void chol_bordered(double *T, double *dg, double *pub, int m, int nrows, int ld, int tid = threadIdx.x) {
  if constexpr (NT >= 128) {
    if (nrows <= NT && m <= CHOL_RR_MMAX) { chol_bordered_rr<NT, CHOL_RR_MMAX>(T, pub, m, nrows, ld); return; }
  }
  chol_bordered_ll<NT>(T, dg, m, nrows, ld, tid);
}

// x[j] = sum_{k>=j} Y[j][k] w[k]  (Y = L^-T stored by rows, upper triangular)
__device__ __forceinline__ double upper_dot(const double *Yrow, const double *w, int j, int m) {
  double s0 = 0.0, s1 = 0.0;
  int k = j;
  for (; k + 1 < m; k += 2) { s0 = fma(Yrow[k], w[k], s0); s1 = fma(Yrow[k + 1], w[k + 1], s1); }
  if (k < m) s0 = fma(Yrow[k], w[k], s0);
  return s0 + s1;
}

// FMM's SEVAL on the determinant grid (splinefun(pars, ldets)(lambda), R/15_sar.R:86-87): the piece of the largest knot
// <= v; left of the grid the first piece, right of it the last knot's cubic
__device__ __forceinline__ double spline_ldet(const DeviceData &dd, double v) {
  const int n = dd.spl_n;
  const double *x = dd.spl;
  int lo = 0, hi = n;                                   // number of knots <= v
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] <= v) lo = mid + 1; else hi = mid; }
  const int i = lo > 0 ? (lo - 1 < n - 1 ? lo - 1 : n - 1) : 0;
  const double dx = v - x[i];
  return x[n + i] + dx * (x[2 * n + i] + dx * (x[3 * n + i] + dx * x[4 * n + i]));
}

template <int NT>
__device__ double logdet_nodes(const DeviceData &dd, double v, double *red) {
  if (dd.spl_n > 0) return spline_ldet(dd, v);          // uniform across the block: no reduction needed
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
  double *T, *dg, *p0, *mu0, *beta, *mu1, *b0, *b1, *eps, *t0, *t1, *red, *pub, *G;
};

constexpr int G_STAGE_MAX_D = 48;     // stage G in shared memory up to this Gram dimension (18 KB)
__host__ __device__ inline int gram_dim(int m, int model) { return model == 2 ? 2 * m + 2 : m + 2; }
__host__ __device__ inline bool g_staged(int m, int model) { return gram_dim(m, model) <= G_STAGE_MAX_D; }
__host__ __device__ inline int work_ld(int m) { return m | 1; }
constexpr int CHOL_PUB = 2 * (24 + 2);   // published column + next pivot's diagonal of the right-looking factorisation, double buffered
__host__ __device__ inline size_t work_doubles(int m, int model) {
  const size_t d = gram_dim(m, model);
  return (size_t)(2 * m + 3) * work_ld(m) + 10 * (size_t)m + 16 + CHOL_PUB + (g_staged(m, model) ? d * d : 0);
}

__device__ __forceinline__ Work carve(double *sm, int m, int model) {
  Work w;
  w.T = sm;
  w.dg = w.T + (size_t)(2 * m + 3) * work_ld(m);
  w.p0 = w.dg + m; w.mu0 = w.p0 + m; w.beta = w.mu0 + m; w.mu1 = w.beta + m;
  w.b0 = w.mu1 + m; w.b1 = w.b0 + m; w.eps = w.b1 + m; w.t0 = w.eps + m; w.t1 = w.t0 + m;
  w.red = w.t1 + m;
  w.pub = w.red + 16;
  w.G = g_staged(m, model) ? w.pub + CHOL_PUB : nullptr;
  return w;
}

// copy the Gram matrix of this launch's buffer into shared memory (both triangles)
template <int NT>
__device__ __forceinline__ void stage_gram(const DeviceData &dd, double *Gs, int chain) {
  if (Gs == nullptr) return;
  const int d = dd.d;
  if (dd.Gchain != nullptr) {                             // sampled delta: the chain's own Gram matrix (chain = CTA)
    const double *gc = dd.Gchain + (size_t)chain * d * d;
    for (int e = threadIdx.x; e < d * d; e += NT) Gs[e] = gc[e];
    return;
  }
  for (int e = threadIdx.x; e < d * d; e += NT) Gs[e] = g_load(dd, e / d, e - (e / d) * d);
}

// || y* - X* b ||^2 = y*'y* - 2 b'X*'y* + b'X*'X* b   (R/10_lm.R:176 via R/10_lm.R:192)
template <int NT>
__device__ double resid_ss(const GView &g, const double *b, double lam, int m, double *red) {
  double part = 0.0;
  if (NT > 32) {                       // the m x m products over all threads (fixed assignment, fixed-order block sum)
    for (int e = threadIdx.x; e < m * m; e += NT) {
      const int a = e / m, c = e - a * m;
      part = fma(b[a] * g.xx_eff(a, c, lam), b[c], part);
    }
    for (int a = threadIdx.x; a < m; a += NT) part = fma(-2.0 * b[a], g.xy_eff(a, lam), part);
  } else {
    for (int a = threadIdx.x; a < m; a += NT) {
      double t = 0.0;
      for (int c = 0; c < m; ++c) t = fma(g.xx_eff(a, c, lam), b[c], t);
      part = fma(b[a], t - 2.0 * g.xy_eff(a, lam), part);
    }
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
  const double si = 1.0 / s;
  auto rhs_value = [&](int kind, int a) {
    if (kind == 0) return fma(g.xy_eff(a, lam_eff), si, w.p0[a] * w.mu0[a]);
    if (kind == 1) return fma(g.Xy(a), si, w.p0[a] * w.mu0[a]);
    return fma(g.XWy(a), si, w.p0[a] * w.mu0[a]);
  };
  for (int e = threadIdx.x; e < m * m; e += NT) {
    const int a = e / m, b = e - a * m;
    if (b <= a) w.T[a * ld + b] = fma(g.xx_eff(a, b, lam_eff), si, a == b ? w.p0[a] : 0.0);
  }
  for (int e = threadIdx.x; e < nrhs * m; e += NT) {
    const int t = e / m, a = e - t * m;
    w.T[(m + t) * ld + a] = rhs_value(first_rhs_kind + t, a);
  }
  if (first_rhs_kind == 0) FTRACE(26);
  fill_identity<NT>(w.T, m + nrhs, m, ld);
  bsync<NT>();
  if (first_rhs_kind == 0) FTRACE(27);
  chol_bordered<NT>(w.T, w.dg, w.pub, m, 2 * m + nrhs, ld);
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
  chol_bordered<NT>(w.T, w.dg, w.pub, m, m + 1, ld);
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


}  // namespace bsreg
