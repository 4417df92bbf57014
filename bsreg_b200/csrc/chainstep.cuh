// The general per-chain Gibbs step (all models x priors x modes) as a device function: the kernel of csrc/chain.cu and the
// chain CTAs of the persistent kernel for the priors other than Independent (csrc/persist.cu) run the same code.
#pragma once
#include "internal.cuh"
#include "chainlib.cuh"
#include "rng.cuh"

namespace bsreg {

// The body of the general chain kernel: chain c on the NT threads 0 .. NT - 1 of the calling CTA, workspace sm.
// hook.before(rep, dd) runs ahead of every iteration (all threads; it may wait for a Gram matrix and set dd.gcur),
// hook.after(rep) behind it; hook.restage: the Gram matrix changes between the iterations and is staged every time.
struct NoChainHook {
  static constexpr bool restage = false;
  __device__ __forceinline__ void before(int, DeviceData &) const {}
  __device__ __forceinline__ void after(int) const {}
};

template <int NT, class Hook>
__device__ void chain_body(DeviceData dd, const Priors &pr, const ChainState &cs, const RunCtl *ctlp, double *draws, int mode,
                           const InitArgs &ia, int nrep, int c, double *sm, const Hook &hook) {
  const int tid = threadIdx.x, m = dd.m, n = dd.n;
  const int A = aux_len(m), ld = work_ld(m);
  Work w = carve(sm, m, dd.model);
  const bool conj = pr.prior == 1, hs = pr.prior == 3, sng = pr.prior == 2;
  const bool sar = dd.model == 1, sem = dd.model == 2;
  double *aux = cs.aux + (size_t)c * A;
  double *vecA = aux, *vecB = aux + m;
  const Philox ph(cs.seed, (uint32_t)(cs.chain_offset + c));

  // nrep > 1 (cached statistics: the Gram matrix does not change): that many Gibbs iterations in one launch.  The state goes
  // through global memory between them exactly as between launches; the loop saves the launches (38 -> 35 us per iteration
  // at M = 20, SAR + Conjugate: the step itself is bound by its dependent FP64 arithmetic, profiles/r2_chain_general_trace.log).
  for (int rep = 0; rep < nrep; ++rep) {
  if (rep > 0) { __threadfence_block(); bsync<NT>(); }
  hook.before(rep, dd);
  if (Hook::restage || rep == 0) stage_gram<NT>(dd, w.G, c);            // visible behind the barrier that follows the state load
  const GView g{w.G, dd.Gfix + (size_t)dd.gcur * dd.d * dd.d, dd.Gscale, dd.d, m, dd.model};
  double sigma2, lam, scale_l, scale_t;
  TRACE(16);
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
        chol_bordered<NT>(w.T, w.dg, w.pub, m, 2 * m + 1, ld);
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
  TRACE(17);

  const bool do_sweep = (mode == MODE_STEP) || (mode == MODE_INIT && conj);   // Conjugate starts with one beta, sigma draw
  const bool reuse = sar && pr.prior == 0;     // P0 and sigma^2 unchanged between the beta and lambda steps
  bool have_b = false;

  if (do_sweep) {
    // ---- sample_sigma (skipped at init; Conjugate draws beta first, R/10_lm.R:218-221) ----
    if (mode == MODE_STEP) {
      const double rss = resid_ss<NT>(g, conj ? w.mu1 : w.beta, lam, m, w.red);
      sigma2 = 1.0 / ph.gamma(it, SLOT_SIGMA, pr.shape1, pr.rate0 + 0.5 * rss);
    }
    TRACE(18);
    // ---- sample_beta ----
    const int nrhs = reuse ? 3 : 1;
    factor_system<NT>(g, w, lam, conj ? 1.0 : sigma2, m, nrhs, 0);
    TRACE(19);
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

  TRACE(20);
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

  TRACE(21);
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
    TRACE(22);
    double e0e0, e1e0, e1e1;
    sar_coeffs<NT>(g, w, m, e0e0, e1e0, e1e1);
    TRACE(23);
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

  TRACE(24);
  // ---- finalize, tuning, draw store ----
  if (mode == MODE_STEP) {
    iters += 1;
    const RunCtl ctl = *ctlp;
    const int step = step0 + 1;
    if (step <= ctl.n_burn && step % 10 == 0 && step <= ctl.tune_limit) {
      if (sar || sem) mh_tune(scale_l, cnt[0], ctl);
      if (sng && pr.sng_theta_scale > 0.0) mh_tune(scale_t, cnt[1], ctl);
    }
    if (step > ctl.n_burn && ctl.store && dd.Gchain == nullptr) {   // sampled delta: delta_decide_kernel stores the draw
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
  TRACE(25);
  hook.after(rep);
  }
}

}  // namespace bsreg
