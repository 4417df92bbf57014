"""Counter-based RNG contract shared by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws with R's
Mersenne-Twister + inversion (rnorm/rgamma/runif: R/10_lm.R:178, R/00_aux.R:65,
R/20_mh.R:74,175; R/11_shrinkage.R:56-60,129-137), which cannot be reproduced
bit-for-bit; RNG parity with the reference is therefore distributional only
(SURVEY.md section 7).  What IS bit-reproducible is this contract, which the device
code (bsreg_b200/csrc/rng.cuh) implements identically so that oracle and CUDA
chains can be compared draw by draw:

* generator : Philox4x32-10 (Salmon et al. 2011), key = (seed_lo, seed_hi)
* counter   : (c0, c1, c2, c3) = (sub, slot, iteration, global_chain_id)
              sub = (element << 10) | call   (call < 1024 calls per element)
* uniforms  : one Philox call -> two 52-bit uniforms in the OPEN interval (0,1)
                u_a = ((r1:r0 >> 12) + 0.5) * 2^-52,  u_b likewise from r3:r2
* normal    : Box-Muller, z = sqrt(-2 ln u_a) * cos(2 pi u_b)   (one call)
* exp(rate) : -ln(u_a) / rate                                    (one call)
* gamma     : Marsaglia-Tsang (2000); attempt t uses call 2t (normal) and
              2t+1 (u_a); shape < 1 boosts with u_a of call 1023
* gig       : Hoermann-Leydold (2014) as implemented by GIGrvg::rgig;
              attempt t uses call t (u_a = first, u_b = second uniform)
* trunc. RW : attempt t uses call t; redraw until lower < v < upper

Slots (c1) per Gibbs iteration, see ``Slot``.
"""
from __future__ import annotations

import math

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF
MAX_CALLS = 1024          # calls addressable per element (10 bits)
MAX_ATTEMPTS = 500        # cap on rejection loops (2 calls per gamma attempt)
BOOST_CALL = 1023
TWO_M52 = 2.0 ** -52


class Slot:
    """Philox c1 values.  Order follows the reference's RNG consumption order
    per iteration (SURVEY.md 9.4 / R/10_lm.R:46-63)."""
    SIGMA = 0          # 1 gamma            R/10_lm.R:178
    BETA = 1           # M normals          R/00_aux.R:65
    LAMBDA_PROP = 2    # truncated RW       R/20_mh.R:173-177
    LAMBDA_ACC = 3     # accept uniform     R/20_mh.R:74
    HS_LAMBDA = 4      # M exponentials     R/11_shrinkage.R:129
    HS_TAU = 5         # 1 gamma            R/11_shrinkage.R:132
    HS_NU = 6          # M exponentials     R/11_shrinkage.R:135
    HS_ZETA = 7        # 1 exponential      R/11_shrinkage.R:137
    SNG_LAMBDA = 8     # 1 gamma            R/11_shrinkage.R:56
    SNG_TAU = 9        # M GIG              R/11_shrinkage.R:60
    SNG_THETA_PROP = 10  # 1 normal         R/20_mh.R:126
    SNG_THETA_ACC = 11   # 1 uniform        R/20_mh.R:74
    RHO_PROP = 12      # SEM truncated RW   R/20_mh.R:173-177 (MH_SEM_lambda alias :218)
    RHO_ACC = 13       # SEM accept uniform
    DELTA_PROP = 14    # SLX connectivity parameter, truncated RW   R/20_mh.R:245-250
    DELTA_ACC = 15     # accept uniform                             R/20_mh.R:74
    PROBE = 64         # Rademacher probes for the trace estimator (setup, iteration = 0)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Scalar Philox4x32-10 on Python ints.  Returns (r0, r1, r2, r3)."""
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def philox4x32_vec(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; counters broadcast as uint64 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    m32 = np.uint64(MASK32)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> s32, p0 & m32
        hi1, lo1 = p1 >> s32, p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(W0)) & m32
        k1 = (k1 + np.uint64(W1)) & m32
    return c0, c1, c2, c3


def _u52(lo, hi):
    return (((hi << 32) | lo) >> 12) * TWO_M52 + 0.5 * TWO_M52


class Stream:
    """Addressed draws for one (seed, global chain id).

    Every method takes the iteration and slot explicitly, so results do not
    depend on call order, on how chains are batched, or on the GPU count.
    """

    def __init__(self, seed: int, chain: int):
        self.k0 = seed & MASK32
        self.k1 = (seed >> 32) & MASK32
        self.chain = chain & MASK32

    # -- raw ---------------------------------------------------------------
    def uniforms(self, it: int, slot: int, elem: int = 0, call: int = 0):
        sub = ((elem << 10) | call) & MASK32
        r0, r1, r2, r3 = philox4x32(sub, slot, it & MASK32, self.chain, self.k0, self.k1)
        return _u52(r0, r1), _u52(r2, r3)

    def uniforms_vec(self, it: int, slot: int, n: int, call: int = 0):
        sub = (np.arange(n, dtype=np.uint64) << np.uint64(10)) | np.uint64(call)
        r0, r1, r2, r3 = philox4x32_vec(sub, slot, it & MASK32, self.chain, self.k0, self.k1)
        ua = (((r1 << np.uint64(32)) | r0) >> np.uint64(12)).astype(np.float64) * TWO_M52 + 0.5 * TWO_M52
        ub = (((r3 << np.uint64(32)) | r2) >> np.uint64(12)).astype(np.float64) * TWO_M52 + 0.5 * TWO_M52
        return ua, ub

    # -- distributions -----------------------------------------------------
    def uniform(self, it, slot, elem=0, call=0):
        return self.uniforms(it, slot, elem, call)[0]

    def normal(self, it, slot, elem=0, call=0):
        ua, ub = self.uniforms(it, slot, elem, call)
        return math.sqrt(-2.0 * math.log(ua)) * math.cos(2.0 * math.pi * ub)

    def normal_vec(self, it, slot, n):
        ua, ub = self.uniforms_vec(it, slot, n)
        return np.sqrt(-2.0 * np.log(ua)) * np.cos(2.0 * np.pi * ub)

    def exponential(self, it, slot, rate, elem=0):
        return -math.log(self.uniforms(it, slot, elem, 0)[0]) / rate

    def exponential_vec(self, it, slot, rate):
        rate = np.asarray(rate, dtype=np.float64)
        ua, _ = self.uniforms_vec(it, slot, rate.shape[0])
        return -np.log(ua) / rate

    def gamma(self, it, slot, shape, rate, elem=0):
        """Gamma(shape, rate) by Marsaglia-Tsang; shape == 1 is exponential."""
        if shape == 1.0:
            return self.exponential(it, slot, rate, elem)
        a = shape + 1.0 if shape < 1.0 else shape
        d = a - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        g = float("nan")
        for t in range(MAX_ATTEMPTS):
            x = self.normal(it, slot, elem, 2 * t)
            v = 1.0 + c * x
            if v <= 0.0:
                continue
            v = v * v * v
            u = self.uniforms(it, slot, elem, 2 * t + 1)[0]
            if math.log(u) < 0.5 * x * x + d - d * v + d * math.log(v):
                g = d * v
                break
        if shape < 1.0:
            u = self.uniforms(it, slot, elem, BOOST_CALL)[0]
            g *= math.exp(math.log(u) / shape)
        return g / rate

    def truncated_rw(self, it, slot, loc, scale, lower, upper):
        """R/20_mh.R:173-177: redraw N(loc, scale) until lower < v < upper."""
        for t in range(MAX_CALLS):
            v = loc + scale * self.normal(it, slot, 0, t)
            if lower < v < upper:
                return v
        return loc  # R would spin forever; bounded here (never reached in practice)

    def gig(self, it, slot, lam, chi, psi, elem=0):
        return gig_draw(lambda t: self.uniforms(it, slot, elem, t), lam, chi, psi,
                        lambda shp, t0: self._gamma_from(it, slot, elem, shp, t0))

    def _gamma_from(self, it, slot, elem, shape, _t0):
        # degenerate GIG branches (chi or psi ~ 0) fall back to a unit-rate gamma
        return self.gamma(it, slot, shape, 1.0, elem)


# ---------------------------------------------------------------------------
# Generalised inverse Gaussian: Hoermann & Leydold (2014), restating the
# algorithm shipped as GIGrvg::rgig (used by R/11_shrinkage.R:60; the package
# is NOT vendored in /root/reference and DESCRIPTION pins no version).
# Density: f(x) ~ x^(lam-1) exp(-(chi/x + psi x)/2).
# ---------------------------------------------------------------------------
GIG_ZTOL = 2.220446049250313e-16 * 10.0


def _gig_mode(lam, omega):
    if lam >= 1.0:
        return (math.sqrt((lam - 1.0) ** 2 + omega * omega) + (lam - 1.0)) / omega
    return omega / (math.sqrt((1.0 - lam) ** 2 + omega * omega) + (1.0 - lam))


def gig_draw(unif2, lam, chi, psi, gamma1):
    """One GIG(lam, chi, psi) variate.  ``unif2(t)`` returns the two uniforms of
    attempt t; ``gamma1(shape, t0)`` a Gamma(shape, 1) for the degenerate cases."""
    if chi < GIG_ZTOL:
        return gamma1(lam, 0) * 2.0 / psi if lam > 0.0 else 1.0 / (gamma1(-lam, 0) * 2.0 / psi)
    if psi < GIG_ZTOL:
        return 1.0 / (gamma1(lam, 0) * 2.0 / chi) if lam > 0.0 else gamma1(-lam, 0) * 2.0 / chi
    lam_old = lam
    lam = abs(lam)
    alpha = math.sqrt(chi / psi)
    omega = math.sqrt(psi * chi)
    if lam > 2.0 or omega > 3.0:
        x = _gig_rou_shift(unif2, lam, omega)
    elif lam >= 1.0 - 2.25 * omega * omega or omega > 0.2:
        x = _gig_rou_noshift(unif2, lam, omega)
    else:
        x = _gig_newapproach1(unif2, lam, omega)
    return alpha / x if lam_old < 0.0 else alpha * x


def _gig_rou_noshift(unif2, lam, omega):
    t_ = 0.5 * (lam - 1.0)
    s = 0.25 * omega
    xm = _gig_mode(lam, omega)
    nc = t_ * math.log(xm) - s * (xm + 1.0 / xm)
    ym = ((lam + 1.0) + math.sqrt((lam + 1.0) ** 2 + omega * omega)) / omega
    um = math.exp(0.5 * (lam + 1.0) * math.log(ym) - s * (ym + 1.0 / ym) - nc)
    x = xm
    for t in range(MAX_CALLS):
        u, v = unif2(t)
        x = um * u / v
        if math.log(v) <= t_ * math.log(x) - s * (x + 1.0 / x) - nc:
            break
    return x


def _gig_rou_shift(unif2, lam, omega):
    t_ = 0.5 * (lam - 1.0)
    s = 0.25 * omega
    xm = _gig_mode(lam, omega)
    nc = t_ * math.log(xm) - s * (xm + 1.0 / xm)
    a = -(2.0 * (lam + 1.0) / omega + xm)
    b = 2.0 * (lam - 1.0) * xm / omega - 1.0
    c = xm
    p = b - a * a / 3.0
    q = (2.0 * a * a * a) / 27.0 - (a * b) / 3.0 + c
    fi = math.acos(-q / (2.0 * math.sqrt(-(p * p * p) / 27.0)))
    fak = 2.0 * math.sqrt(-p / 3.0)
    y1 = fak * math.cos(fi / 3.0) - a / 3.0
    y2 = fak * math.cos(fi / 3.0 + 4.0 / 3.0 * math.pi) - a / 3.0
    uplus = (y1 - xm) * math.exp(t_ * math.log(y1) - s * (y1 + 1.0 / y1) - nc)
    uminus = (y2 - xm) * math.exp(t_ * math.log(y2) - s * (y2 + 1.0 / y2) - nc)
    x = xm
    for t in range(MAX_CALLS):
        u, v = unif2(t)
        x = (uminus + u * (uplus - uminus)) / v + xm
        if x > 0.0 and math.log(v) <= t_ * math.log(x) - s * (x + 1.0 / x) - nc:
            break
        x = xm
    return x


def _gig_newapproach1(unif2, lam, omega):
    xm = _gig_mode(lam, omega)
    x0 = omega / (1.0 - lam)
    k0 = math.exp((lam - 1.0) * math.log(xm) - 0.5 * omega * (xm + 1.0 / xm))
    a0 = k0 * x0
    if x0 >= 2.0 / omega:
        k1 = 0.0
        a1 = 0.0
        k2 = x0 ** (lam - 1.0)
        a2 = k2 * 2.0 * math.exp(-omega * x0 / 2.0) / omega
    else:
        k1 = math.exp(-omega)
        if lam == 0.0:
            a1 = k1 * math.log(2.0 / (omega * omega))
        else:
            a1 = k1 / lam * ((2.0 / omega) ** lam - x0 ** lam)
        k2 = (2.0 / omega) ** (lam - 1.0)
        a2 = k2 * 2.0 * math.exp(-1.0) / omega
    atot = a0 + a1 + a2
    x = xm
    for t in range(MAX_CALLS):
        u1, u2 = unif2(t)
        v = atot * u1
        if v <= a0:
            x = x0 * v / a0
            hx = k0
        else:
            v -= a0
            if v <= a1:
                if lam == 0.0:
                    x = omega * math.exp(math.exp(omega) * v)
                    hx = k1 / x
                else:
                    x = (x0 ** lam + (lam / k1 * v)) ** (1.0 / lam)
                    hx = k1 * x ** (lam - 1.0)
            else:
                v -= a1
                a = x0 if x0 > 2.0 / omega else 2.0 / omega
                x = -2.0 / omega * math.log(math.exp(-omega / 2.0 * a) - omega / (2.0 * k2) * v)
                hx = k2 * math.exp(-omega / 2.0 * x)
        if math.log(u2 * hx) <= (lam - 1.0) * math.log(x) - omega / 2.0 * (x + 1.0 / x):
            break
        x = xm
    return x
