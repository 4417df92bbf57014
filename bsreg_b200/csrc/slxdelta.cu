// Sampled connectivity parameter delta of the SLX model: `Psi_SLX` given as a FUNCTION of delta.
//
// Reference: set_delta R/12_slx.R:26-36 (W <- Psi(delta); WX <- W %*% X_SLX; X <- cbind(X, WX); XX, Xy refreshed),
// sample_extra3 R/12_slx.R:87-106 (rss(v) re-solves for beta with X(v) at the current sigma^2 and P0), MH_SLX_delta
// R/20_mh.R:229-278 (truncated random walk, posterior -(N - M)/2 log rss(v) + dgamma(1/v, a, rate = b, log), no Jacobian).
// The reference takes any R closure; the C ABI offers the family the paper uses (paper/0_data.R:31-35): inverse-distance
// decay on a fixed sparsity pattern, W(delta)_ij = dist_ij ^ -delta, optionally row-normalised.
//
// This is the one place of the hot path where N-scale work is chain-specific: every chain proposes its own delta, so
//   sweep_delta_kernel   streams the pattern, log-distances, X_SLX, X and y ONCE per Gibbs iteration for all chains and
//                        forms, per chain c, the lags W(delta'_c) X_SLX on the fly and the blocks of the Gram matrix they
//                        touch: lag' [y | X | lag]  (L x (1 + K + L) per chain); per-CTA partials, combined in fixed order
//   delta_decide_kernel  per chain: Gram matrix at the proposal = shared base blocks + the chain's lag blocks; beta(v) and
//                        rss(v) at the proposal and at the current value (two K x K solves), MH decision, tuning, draw store
#include "internal.cuh"
#include "chainlib.cuh"
#include "rng.cuh"

namespace bsreg {

constexpr int DL_MAXL = 8;            // lagged columns
constexpr int DL_T = 128;             // (chain, lag column) pairs per CTA of the sweep

__device__ __forceinline__ double psi_weight(double logd, double delta) { return exp(-delta * logd); }

// X[:, K + l] = (W(delta0) X_SLX)[:, l]: the design at the starting value (setup_2SLX -> set_delta(priors$delta), R/12_slx.R:74)
__global__ void slx_power_lag_kernel(DeltaData dl, double delta, double *__restrict__ X, int m, int k_base, int n) {
  const int L = dl.L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double lag[DL_MAXL], ws = 0.0;
#pragma unroll
    for (int l = 0; l < DL_MAXL; ++l) lag[l] = 0.0;
    for (int p = dl.rp[i], pe = dl.rp[i + 1]; p < pe; ++p) {
      const double w = psi_weight(dl.logd[p], delta);
      ws += w;
      const double *xl = dl.Xl + (size_t)dl.ci[p] * L;
#pragma unroll
      for (int l = 0; l < DL_MAXL; ++l) if (l < L) lag[l] = fma(w, xl[l], lag[l]);
    }
    const double sc = (dl.norm == 1 && ws > 0.0) ? 1.0 / ws : 1.0;
#pragma unroll
    for (int l = 0; l < DL_MAXL; ++l) if (l < L) X[(size_t)i * m + k_base + l] = lag[l] * sc;
  }
}

void launch_slx_power_lag(const DeltaData &dl, double delta, double *X, int m, int k_base, int n, cudaStream_t s) {
  slx_power_lag_kernel<<<592, 256, 0, s>>>(dl, delta, X, m, k_base, n);
}

// One pass over the pattern, log-distances, X_SLX, X (base columns) and y for ALL chains of the shard: thread = (chain c,
// lag column l); every thread of a CTA walks the same rows, so the loads of a row are broadcasts.  partial[cta][pair][J],
// J = 1 + K + L: lag_l' y, lag_l' X_k, lag_l' lag_l'.
template <int KB>     // base columns handled in register blocks of KB
__global__ void __launch_bounds__(DL_T) sweep_delta_kernel(DeltaData dl, const double *__restrict__ X, const double *__restrict__ y,
                                                           int n, int m, int k_base, int n_chains, const double *__restrict__ delta,
                                                           double *__restrict__ partial) {
  const int L = dl.L, J = 1 + k_base + L;
  const int pair = blockIdx.y * DL_T + threadIdx.x, npairs = n_chains * L;
  const bool live = pair < npairs;
  const int c = live ? pair / L : 0, l = live ? pair - c * L : 0;
  const double dc = delta[c];
  double acc[1 + KB + DL_MAXL];
#pragma unroll
  for (int j = 0; j < 1 + KB + DL_MAXL; ++j) acc[j] = 0.0;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    double lag[DL_MAXL], ws = 0.0;
#pragma unroll
    for (int q = 0; q < DL_MAXL; ++q) lag[q] = 0.0;
    for (int p = dl.rp[i], pe = dl.rp[i + 1]; p < pe; ++p) {
      const double w = psi_weight(dl.logd[p], dc);
      ws += w;
      const double *xl = dl.Xl + (size_t)dl.ci[p] * L;
#pragma unroll
      for (int q = 0; q < DL_MAXL; ++q) if (q < L) lag[q] = fma(w, xl[q], lag[q]);
    }
    const double sc = (dl.norm == 1 && ws > 0.0) ? 1.0 / ws : 1.0;
    double mine = 0.0;
#pragma unroll
    for (int q = 0; q < DL_MAXL; ++q) { lag[q] *= sc; if (q == l) mine = lag[q]; }
    acc[0] = fma(mine, y[i], acc[0]);
    const double *xr = X + (size_t)i * m;
#pragma unroll
    for (int k = 0; k < KB; ++k) if (k < k_base) acc[1 + k] = fma(mine, xr[k], acc[1 + k]);
#pragma unroll
    for (int q = 0; q < DL_MAXL; ++q) if (q < L) acc[1 + KB + q] = fma(mine, lag[q], acc[1 + KB + q]);
  }
  if (live) {
    double *out = partial + ((size_t)blockIdx.x * npairs + pair) * J;
    out[0] = acc[0];
#pragma unroll
    for (int k = 0; k < KB; ++k) if (k < k_base) out[1 + k] = acc[1 + k];
#pragma unroll
    for (int q = 0; q < DL_MAXL; ++q) if (q < L) out[1 + k_base + q] = acc[1 + KB + q];
  }
}

int delta_sweep_ctas(int sm_count) { return sm_count * 4; }

void launch_sweep_delta(const DeltaData &dl, const double *X, const double *y, int n, int m, int k_base, int n_chains,
                        const double *delta, double *partial, int sm_count, cudaStream_t s) {
  const dim3 grid(delta_sweep_ctas(sm_count), (n_chains * dl.L + DL_T - 1) / DL_T);
  if (k_base <= 8) sweep_delta_kernel<8><<<grid, DL_T, 0, s>>>(dl, X, y, n, m, k_base, n_chains, delta, partial);
  else if (k_base <= 24) sweep_delta_kernel<24><<<grid, DL_T, 0, s>>>(dl, X, y, n, m, k_base, n_chains, delta, partial);
  else sweep_delta_kernel<48><<<grid, DL_T, 0, s>>>(dl, X, y, n, m, k_base, n_chains, delta, partial);
}

// dgamma(1 / v, a, rate = b, log = TRUE): R/20_mh.R:269 (quirk Q7: no Jacobian)
__device__ __forceinline__ double log_prior_delta(double v, double a, double b) {
  const double x = 1.0 / v;
  return a * log(b) - lgamma(a) + (a - 1.0) * log(x) - b * x;
}

// Per chain: the MH step on delta and the end of the iteration (draw store).  mode 0 = step, 1 = adopt the proposal's
// Gram matrix unconditionally, 2 = delta stays at its starting value (delta_scale = 0): only the draw is stored.
template <int NT>
__global__ void __launch_bounds__(NT) delta_decide_kernel(DeviceData dd, DeltaData dl, Priors pr, ChainState cs,
                                                          const RunCtl *ctlp, double *draws, const double *partial,
                                                          int n_parts, int mode) {
  extern __shared__ double sm[];
  const int c = blockIdx.x, tid = threadIdx.x, m = dd.m, d = dd.d, n = dd.n, L = dl.L, kb = m - L, J = 1 + kb + L;
  const int ld = work_ld(m);
  Work w = carve(sm, m, 0);                              // w.G: the chain's Gram matrix at the current delta
  double *Gp = w.G + d * d;                              // ... and at the proposal
  double *Gc_glob = dl.Gchain + (size_t)c * d * d;
  const int npairs = cs.n_chains * L;
  for (int e = tid; e < d * d; e += NT) { w.G[e] = Gc_glob[e]; Gp[e] = Gc_glob[e]; }
  for (int a = tid; a < m; a += NT) { w.mu0[a] = cs.mu0[a]; w.p0[a] = cs.p0[(size_t)c * m + a]; }
  bsync<NT>();
  // the chain's lag blocks at the proposal: per-CTA partials added in fixed order (mode 2: delta is not sampled)
  for (int e = tid; e < (mode == 2 ? 0 : L * J); e += NT) {
    const int l = e / J, j = e - l * J;
    const double *src = partial + ((size_t)c * L + l) * J + j;
    double s0 = 0.0, s1 = 0.0;
    int b = 0;
    for (; b + 1 < n_parts; b += 2) { s0 += src[(size_t)b * npairs * J]; s1 += src[(size_t)(b + 1) * npairs * J]; }
    if (b < n_parts) s0 += src[(size_t)b * npairs * J];
    const double v = s0 + s1;
    const int gi = 2 + kb + l;                           // Gram index of the lag column (y = 0, Wy = 1 (zero), X_a = 2 + a)
    const int gj = j == 0 ? 0 : (j <= kb ? 2 + (j - 1) : 2 + kb + (j - 1 - kb));
    Gp[gi * d + gj] = v;
    if (gj < 2 + kb) Gp[gj * d + gi] = v;                // base x lag: mirror; lag x lag is produced in both orders
  }
  bsync<NT>();
  if (mode == 1) {
    for (int e = tid; e < d * d; e += NT) Gc_glob[e] = Gp[e];
    return;
  }
  const double s2 = cs.sigma2[c];
  const double delta_c = dl.delta[c], prop = dl.delta_prop[c];
  const RunCtl ctl = *ctlp;
  const int step = cs.step[c];
  double delta_new = delta_c;
  if (mode == 0) {
    // rss(v) = || y - X(v) beta(v) ||^2, beta(v) = (P0 + X'X / s2)^-1 (X'y / s2 + P0 mu0): R/12_slx.R:91-97
    double rss[2];
    for (int which = 0; which < 2; ++which) {
      const GView g{which == 0 ? w.G : Gp, nullptr, nullptr, d, m, 0};
      factor_system<NT>(g, w, 0.0, s2, m, 1, 0);
      for (int j = tid; j < m; j += NT) w.mu1[j] = upper_dot(w.T + (m + 1 + j) * ld, w.T + m * ld, j, m);
      bsync<NT>();
      rss[which] = resid_ss<NT>(g, w.mu1, 0.0, m, w.red);
      bsync<NT>();
    }
    const Philox ph(cs.seed, (uint32_t)(cs.chain_offset + c));
    const uint32_t it = (uint32_t)cs.iters[c];             // the chain kernel has already counted this iteration
    auto post = [&](double v, double r) {
      return !(r > 0.0) ? nan("") : -0.5 * (double)(n - m) * log(r) + log_prior_delta(v, dl.prior_a, dl.prior_b);
    };
    const double p = exp(post(prop, rss[1]) - post(delta_c, rss[0]));
    const bool acc = ph.uniform(it, SLOT_DELTA_ACC) < p;   // NaN compares false: R/20_mh.R:74
    if (acc) {
      for (int e = tid; e < d * d; e += NT) Gc_glob[e] = Gp[e];
      delta_new = prop;
    }
    if (tid == 0) {
      long long cnt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cnt[k] = dl.cnt[(size_t)c * 4 + k];
      double scale = dl.scale[c];
      mh_count(cnt, acc);
      if (step <= ctl.n_burn && step % 10 == 0 && step <= ctl.tune_limit) mh_tune(scale, cnt, ctl);
#pragma unroll
      for (int k = 0; k < 4; ++k) dl.cnt[(size_t)c * 4 + k] = cnt[k];
      dl.scale[c] = scale;
      dl.delta[c] = delta_new;
    }
  }
  // draw store: c(beta, sigma, delta_SLX), R/10_lm.R:193 + R/12_slx.R:117-121
  if (step > ctl.n_burn && ctl.store) {
    const int row = step - ctl.n_burn - 1 - ctl.save_base;
    if (row >= 0 && row < ctl.save_cap) {
      double *dst = draws + (size_t)c * (m + 2) * ctl.save_cap + row;
      for (int a = tid; a < m; a += NT) dst[(size_t)a * ctl.save_cap] = cs.beta[(size_t)c * m + a];
      if (tid == 0) {
        dst[(size_t)m * ctl.save_cap] = sqrt(s2);
        dst[(size_t)(m + 1) * ctl.save_cap] = delta_new;
      }
    }
  }
}

// proposals of the next MH step on delta, one thread per chain: truncated random walk, R/20_mh.R:245-250
__global__ void delta_propose_kernel(DeltaData dl, ChainState cs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cs.n_chains) return;
  const Philox ph(cs.seed, (uint32_t)(cs.chain_offset + c));
  const uint32_t it = (uint32_t)(cs.iters[c] + 1);       // the iteration about to run
  dl.delta_prop[c] = dl.sampled ? ph.truncated_rw(it, SLOT_DELTA_PROP, dl.delta[c], dl.scale[c], dl.lo, dl.hi) : dl.delta[c];
}

void launch_delta_propose(const DeltaData &dl, const ChainState &cs, cudaStream_t s) {
  delta_propose_kernel<<<(cs.n_chains + 63) / 64, 64, 0, s>>>(dl, cs);
}

size_t delta_decide_smem(int m) { return (work_doubles(m, 0) + (size_t)(m + 2) * (m + 2)) * sizeof(double); }

void launch_delta_decide(const DeviceData &dd, const DeltaData &dl, const Priors &pr, const ChainState &cs, const RunCtl *ctl,
                         double *draws, const double *partial, int n_parts, int mode, cudaStream_t s) {
  const size_t smem = delta_decide_smem(dd.m);
  if (dd.m <= 32) {
    if (smem_optin((const void *)delta_decide_kernel<32>, smem) != cudaSuccess) return;
    delta_decide_kernel<32><<<cs.n_chains, 32, smem, s>>>(dd, dl, pr, cs, ctl, draws, partial, n_parts, mode);
  } else {
    if (smem_optin((const void *)delta_decide_kernel<128>, smem) != cudaSuccess) return;
    delta_decide_kernel<128><<<cs.n_chains, 128, smem, s>>>(dd, dl, pr, cs, ctl, draws, partial, n_parts, mode);
  }
}

}  // namespace bsreg
