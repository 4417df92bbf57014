// Device side of the counter-based RNG contract (mirror of oracle/philox.py).
//
// Replaces the reference's rnorm / rgamma / runif / GIGrvg::rgig calls
// (R/10_lm.R:178, R/00_aux.R:65, R/20_mh.R:74,126,175, R/11_shrinkage.R:56-60,129-137)
// with Philox4x32-10 addressed by (sub, slot, iteration, global chain id), so a
// chain's draws do not depend on how chains are batched or on the GPU count.
#pragma once
#include <cstdint>

namespace bsreg {

enum Slot : uint32_t {
  SLOT_SIGMA = 0, SLOT_BETA = 1, SLOT_LAMBDA_PROP = 2, SLOT_LAMBDA_ACC = 3,
  SLOT_HS_LAMBDA = 4, SLOT_HS_TAU = 5, SLOT_HS_NU = 6, SLOT_HS_ZETA = 7,
  SLOT_SNG_LAMBDA = 8, SLOT_SNG_TAU = 9, SLOT_SNG_THETA_PROP = 10, SLOT_SNG_THETA_ACC = 11,
  SLOT_RHO_PROP = 12, SLOT_RHO_ACC = 13, SLOT_DELTA_PROP = 14, SLOT_DELTA_ACC = 15, SLOT_PROBE = 64
};

constexpr int RNG_MAX_CALLS = 1024;
constexpr int RNG_MAX_ATTEMPTS = 500;
constexpr int RNG_BOOST_CALL = 1023;

struct Philox {
  uint32_t k0, k1, chain;
  __host__ __device__ Philox(uint64_t seed, uint32_t chain_)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), chain(chain_) {}

  __device__ __forceinline__ uint4 raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }

  // two 52-bit uniforms in the open interval (0,1)
  __device__ __forceinline__ void uniforms(uint32_t it, uint32_t slot, uint32_t elem, uint32_t call,
                                           double &ua, double &ub) const {
    const uint4 r = raw((elem << 10) | call, slot, it, chain);
    const uint64_t a = (((uint64_t)r.y << 32) | r.x) >> 12;
    const uint64_t b = (((uint64_t)r.w << 32) | r.z) >> 12;
    ua = (double)a * 0x1p-52 + 0x1p-53;
    ub = (double)b * 0x1p-52 + 0x1p-53;
  }
  __device__ __forceinline__ double uniform(uint32_t it, uint32_t slot, uint32_t elem = 0, uint32_t call = 0) const {
    double a, b; uniforms(it, slot, elem, call, a, b); return a;
  }
  __device__ __forceinline__ double normal(uint32_t it, uint32_t slot, uint32_t elem = 0, uint32_t call = 0) const {
    double a, b; uniforms(it, slot, elem, call, a, b);
    return sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
  }
  __device__ __forceinline__ double exponential(uint32_t it, uint32_t slot, double rate, uint32_t elem = 0) const {
    return -log(uniform(it, slot, elem, 0)) / rate;
  }
  // Marsaglia-Tsang; attempt t = calls 2t (normal), 2t+1 (uniform); shape<1 boosts with call 1023
  __device__ double gamma(uint32_t it, uint32_t slot, double shape, double rate, uint32_t elem = 0) const {
    if (shape == 1.0) return exponential(it, slot, rate, elem);
    const double a = shape < 1.0 ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    double g = nan("");
    for (int t = 0; t < RNG_MAX_ATTEMPTS; ++t) {
      const double x = normal(it, slot, elem, 2 * t);
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = uniform(it, slot, elem, 2 * t + 1);
      if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
    }
    if (shape < 1.0) g *= exp(log(uniform(it, slot, elem, RNG_BOOST_CALL)) / shape);
    return g / rate;
  }
  // R/20_mh.R:173-177: redraw N(loc, scale) until lower < v < upper
  __device__ double truncated_rw(uint32_t it, uint32_t slot, double loc, double scale, double lower, double upper) const {
    for (int t = 0; t < RNG_MAX_CALLS; ++t) {
      const double v = loc + scale * normal(it, slot, 0, t);
      if (v > lower && v < upper) return v;
    }
    return loc;
  }
  __device__ double gig(uint32_t it, uint32_t slot, double lam, double chi, double psi, uint32_t elem = 0) const;
};

// ---- generalised inverse Gaussian (Hoermann & Leydold 2014, as in GIGrvg::rgig) ----
__device__ __forceinline__ double gig_mode(double lam, double omega) {
  if (lam >= 1.0) return (sqrt((lam - 1.0) * (lam - 1.0) + omega * omega) + (lam - 1.0)) / omega;
  return omega / (sqrt((1.0 - lam) * (1.0 - lam) + omega * omega) + (1.0 - lam));
}

__device__ inline double Philox::gig(uint32_t it, uint32_t slot, double lam, double chi, double psi, uint32_t elem) const {
  const double ZTOL = 2.220446049250313e-16 * 10.0;
  if (chi < ZTOL) {
    return lam > 0.0 ? gamma(it, slot, lam, 1.0, elem) * 2.0 / psi : 1.0 / (gamma(it, slot, -lam, 1.0, elem) * 2.0 / psi);
  }
  if (psi < ZTOL) {
    return lam > 0.0 ? 1.0 / (gamma(it, slot, lam, 1.0, elem) * 2.0 / chi) : gamma(it, slot, -lam, 1.0, elem) * 2.0 / chi;
  }
  const double lam_old = lam;
  lam = fabs(lam);
  const double alpha = sqrt(chi / psi), omega = sqrt(psi * chi);
  double x, u, v;
  if (lam > 2.0 || omega > 3.0) {            // ratio-of-uniforms with mode shift
    const double t_ = 0.5 * (lam - 1.0), s = 0.25 * omega, xm = gig_mode(lam, omega);
    const double nc = t_ * log(xm) - s * (xm + 1.0 / xm);
    const double a = -(2.0 * (lam + 1.0) / omega + xm), b = 2.0 * (lam - 1.0) * xm / omega - 1.0, c = xm;
    const double p = b - a * a / 3.0, q = (2.0 * a * a * a) / 27.0 - (a * b) / 3.0 + c;
    const double fi = acos(-q / (2.0 * sqrt(-(p * p * p) / 27.0))), fak = 2.0 * sqrt(-p / 3.0);
    const double y1 = fak * cos(fi / 3.0) - a / 3.0;
    const double y2 = fak * cos(fi / 3.0 + 4.0 / 3.0 * 3.141592653589793) - a / 3.0;
    const double uplus = (y1 - xm) * exp(t_ * log(y1) - s * (y1 + 1.0 / y1) - nc);
    const double uminus = (y2 - xm) * exp(t_ * log(y2) - s * (y2 + 1.0 / y2) - nc);
    x = xm;
    for (int t = 0; t < RNG_MAX_CALLS; ++t) {
      uniforms(it, slot, elem, t, u, v);
      x = (uminus + u * (uplus - uminus)) / v + xm;
      if (x > 0.0 && log(v) <= t_ * log(x) - s * (x + 1.0 / x) - nc) break;
      x = xm;
    }
  } else if (lam >= 1.0 - 2.25 * omega * omega || omega > 0.2) {   // ratio-of-uniforms without shift
    const double t_ = 0.5 * (lam - 1.0), s = 0.25 * omega, xm = gig_mode(lam, omega);
    const double nc = t_ * log(xm) - s * (xm + 1.0 / xm);
    const double ym = ((lam + 1.0) + sqrt((lam + 1.0) * (lam + 1.0) + omega * omega)) / omega;
    const double um = exp(0.5 * (lam + 1.0) * log(ym) - s * (ym + 1.0 / ym) - nc);
    x = xm;
    for (int t = 0; t < RNG_MAX_CALLS; ++t) {
      uniforms(it, slot, elem, t, u, v);
      x = um * u / v;
      if (log(v) <= t_ * log(x) - s * (x + 1.0 / x) - nc) break;
    }
  } else {                                    // "new approach 1": hat with three parts
    const double xm = gig_mode(lam, omega), x0 = omega / (1.0 - lam);
    const double k0 = exp((lam - 1.0) * log(xm) - 0.5 * omega * (xm + 1.0 / xm));
    const double a0 = k0 * x0;
    double k1, a1, k2, a2;
    if (x0 >= 2.0 / omega) {
      k1 = 0.0; a1 = 0.0; k2 = pow(x0, lam - 1.0); a2 = k2 * 2.0 * exp(-omega * x0 / 2.0) / omega;
    } else {
      k1 = exp(-omega);
      a1 = (lam == 0.0) ? k1 * log(2.0 / (omega * omega)) : k1 / lam * (pow(2.0 / omega, lam) - pow(x0, lam));
      k2 = pow(2.0 / omega, lam - 1.0); a2 = k2 * 2.0 * exp(-1.0) / omega;
    }
    const double atot = a0 + a1 + a2;
    x = xm;
    for (int t = 0; t < RNG_MAX_CALLS; ++t) {
      double hx;
      uniforms(it, slot, elem, t, u, v);
      double w = atot * u;
      if (w <= a0) { x = x0 * w / a0; hx = k0; }
      else {
        w -= a0;
        if (w <= a1) {
          if (lam == 0.0) { x = omega * exp(exp(omega) * w); hx = k1 / x; }
          else { x = pow(pow(x0, lam) + (lam / k1 * w), 1.0 / lam); hx = k1 * pow(x, lam - 1.0); }
        } else {
          w -= a1;
          const double a = x0 > 2.0 / omega ? x0 : 2.0 / omega;
          x = -2.0 / omega * log(exp(-omega / 2.0 * a) - omega / (2.0 * k2) * w);
          hx = k2 * exp(-omega / 2.0 * x);
        }
      }
      if (log(v * hx) <= (lam - 1.0) * log(x) - omega / 2.0 * (x + 1.0 / x)) break;
      x = xm;
    }
  }
  return lam_old < 0.0 ? alpha / x : alpha * x;
}

}  // namespace bsreg
