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
#include <algorithm>
#include <cstdlib>

#include "internal.cuh"
#include "chainlib.cuh"
#include "rng.cuh"
#include "chainstep.cuh"

namespace bsreg {

template <int NT>
__global__ void __launch_bounds__(NT) chain_kernel(DeviceData dd, Priors pr, ChainState cs, const RunCtl *ctlp,
                                                   double *draws, int mode, InitArgs ia, int nrep) {
  extern __shared__ double sm[];
  chain_body<NT>(dd, pr, cs, ctlp, draws, mode, ia, nrep, (int)blockIdx.x, sm, NoChainHook{});
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
// The K x K system is solved in the sigma^2-scaled form Q = X*'X* + sigma^2 P0 (so nothing on the
// critical path divides by sigma^2; only the diagonal and the right-hand sides wait for the draw):
//   mu1 = Q^-1 (X*'y* + sigma^2 P0 mu0),  beta = mu1 + sigma Q^-1/2 eps,  b0, b1 alike (R/15_sar.R:105-109).
// Q = L D L' is eliminated right-looking with ONE ROW PER THREAD IN REGISTERS, bordered by the
// right-hand sides and an identity block (so z = L^-1 rhs and D^-1 L^-1 fall out of the same sweep);
// per column: publish the column, one barrier, reciprocal of the pivot, rank-1 update.
template <int MMAX>
__global__ void __launch_bounds__(FAST_NT) chain_fast_kernel(DeviceData dd, Priors pr, ChainState cs, const RunCtl *ctlp,
                                                            double *draws, int nrep) {
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

  // nrep > 1 (cached statistics): that many iterations in one launch, the state going through global memory between them as
  // between launches; the Gram matrix is staged once
  for (int rep = 0; rep < nrep; ++rep) {
  if (rep > 0) { __threadfence_block(); __syncthreads(); }
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
    for (int i = t3 >> 5; i < (rep == 0 ? d : 0); i += 3)
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
      double ldet_prop;
      if (dd.spl_n > 0) {
        ldet_prop = spline_ldet(dd, prop);
      } else {
        double acc = 0.0;
        for (int k = lane; k < dd.n_nodes; k += 32) {
          const double aa = 1.0 - prop * dd.node_re[k], bb = prop * dd.node_im[k];
          acc = fma(dd.node_w[k], 0.5 * log(aa * aa + bb * bb), acc);
        }
        ldet_prop = warp_sum_c(acc);
      }
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
        chol_bordered_ll<32>(T2, t0, m, m + 1, ld, lane);               // t0 is free until phase 3 (SAR only)
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
}

#ifdef BSREG_TRACE
extern "C" int bsreg_debug_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 64); }
#endif

bool chain_fast_supported(const DeviceData &dd, const Priors &pr) {
  return pr.prior == 0 && dd.m <= FAST_MAX_M && g_staged(dd.m, dd.model) && dd.Gchain == nullptr;
}

size_t chain_smem_bytes(int m, int model, int) { return work_doubles(m, model) * sizeof(double); }
int chain_threads(int m) {
  static const int forced = getenv("BSREG_CHAIN_NT") ? atoi(getenv("BSREG_CHAIN_NT")) : 0;   // A/B switch: 32 or 128
  if (forced == 128 || (forced == 32 && m <= 32)) return forced;
  // four warps from M = 12 on: one bordered row per thread (2M + 3 rows); M = 20 SAR + Horseshoe: 54 -> 40 us per step
  return m < 12 ? 32 : 128;
}

template <int NT>
static void launch_chain_nt(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                            int mode, const InitArgs &ia, cudaStream_t s, int nrep) {
  const size_t smem = chain_smem_bytes(dd.m, dd.model, pr.prior);
  if (smem_optin((const void *)chain_kernel<NT>, smem) != cudaSuccess) return;   // the error stays pending for cudaGetLastError
  chain_kernel<NT><<<cs.n_chains, NT, smem, s>>>(dd, pr, cs, ctl, draws, mode, ia, nrep);
}

static void launch_chain(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                         int mode, const InitArgs &ia, cudaStream_t s, int nrep = 1) {
  if (chain_threads(dd.m) == 32) launch_chain_nt<32>(dd, pr, cs, ctl, draws, mode, ia, s, nrep);
  else launch_chain_nt<128>(dd, pr, cs, ctl, draws, mode, ia, s, nrep);
}

void launch_chain_init(const DeviceData &dd, const Priors &pr, const ChainState &cs, const ChainInit &ci, cudaStream_t s) {
  InitArgs ia{ci.beta0, ci.sigma0, ci.lam0, ci.scale0, ci.hs_lambda, ci.hs_tau, ci.hs_zeta, ci.hs_nu,
              ci.sng_lambda, ci.sng_tau, ci.sng_theta, ci.sng_theta_scale, ci.p0_init};
  launch_chain(dd, pr, cs, nullptr, nullptr, ci.refresh_only ? MODE_REFRESH : MODE_INIT, ia, s);
}

static bool chain_uses_fast(const DeviceData &dd, const Priors &pr) {
  static const bool no_fast = getenv("BSREG_CHAIN_GENERIC") != nullptr;   // A/B switch for tests and profiling
  return !no_fast && chain_fast_supported(dd, pr);
}

bool chain_step_loops(const DeviceData &dd, const Priors &pr) {
  const bool off = getenv("BSREG_NO_CHAIN_LOOP") != nullptr;              // A/B switch (read per call: tests flip it)
  return !off && dd.Gchain == nullptr;
}

void launch_chain_step(const DeviceData &dd, const Priors &pr, const ChainState &cs, const RunCtl *ctl, double *draws,
                       cudaStream_t s, int nrep) {
  if (chain_uses_fast(dd, pr)) {
    if (dd.m <= 24) chain_fast_kernel<24><<<cs.n_chains, FAST_NT, fast_doubles<24>(dd.m, dd.model) * sizeof(double), s>>>(dd, pr, cs, ctl, draws, nrep);
    else chain_fast_kernel<32><<<cs.n_chains, FAST_NT, fast_doubles<32>(dd.m, dd.model) * sizeof(double), s>>>(dd, pr, cs, ctl, draws, nrep);
    return;
  }
  InitArgs ia{};
  launch_chain(dd, pr, cs, ctl, draws, MODE_STEP, ia, s, nrep);
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
  stage_gram<NT>(dd, w.G, 0);
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
  DeviceData dv = dd;
  dv.Gchain = nullptr;                                    // parity entry point: the shared Gram matrix
  if (chain_threads(dd.m) == 32) {
    if (smem_optin((const void *)eval_cond_kernel<32>, smem) != cudaSuccess) return;
    eval_cond_kernel<32><<<n_eval, 32, smem, s>>>(dv, pr, mu0, beta, sigma2, lam, p0, v, mu1, rate1, rss, logpost);
  } else {
    if (smem_optin((const void *)eval_cond_kernel<128>, smem) != cudaSuccess) return;
    eval_cond_kernel<128><<<n_eval, 128, smem, s>>>(dv, pr, mu0, beta, sigma2, lam, p0, v, mu1, rate1, rss, logpost);
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
