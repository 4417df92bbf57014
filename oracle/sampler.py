"""One-chain CPU restatement of the bsreg Gibbs sampler (reference-algorithm port).

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned -- the reference
ships no tests and R is not available to run it).  The structure follows the
reference's R6 classes; citations are relative to /root/reference.

Differences from the literal R code, all result-preserving:
* W may be CSR (the reference requires a dense base matrix, R/12_slx.R:61).
* SEM filtering uses the cached lags, y - lam*Wy and X - lam*WX, instead of
  rebuilding the dense N x N `diag(N) - lambda * W` every iteration
  (R/15_sem.R:28-36); `dense_sem=True` does the literal dense rebuild.
* Random numbers come from an RNG adapter: `PhiloxRNG` (the addressed contract
  of oracle/philox.py, comparable draw-by-draw with the CUDA chains) or
  `NumpyRNG` (an independent stream, for distributional checks and timing).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.stats

from . import core
from .philox import Slot, Stream

MODELS = ("lm", "slx", "sar", "sem", "sdm", "sdem")
PRIORS = ("Independent", "Conjugate", "Shrinkage", "Horseshoe")


def default_priors() -> dict:
    """Defaults of R/41_set_options.R:51 (set_NG), :66-67 (set_SNG), :89 (set_HS),
    :116-118 (set_SAR = set_SLX = set_SEM)."""
    spatial = dict(lambda_a=1.01, lambda_b=1.01, lam=0.0, lambda_scale=0.1,
                   lambda_min=-1.0, lambda_max=1.0 - 1e-12)
    return dict(
        NG=dict(mu=0.0, precision=1e-8, shape=0.01, rate=0.01, beta=None, sigma=None),
        SNG=dict(lambda_a=0.01, lambda_b=0.01, theta_scale=0.0, theta_a=1.0, lam=1.0, tau=10.0, theta=0.1),
        HS=dict(lam=1.0, tau=1.0, zeta=1.0, nu=1.0),
        SAR=dict(spatial), SEM=dict(spatial),
        # set_SLX = set_SAR (R/41_set_options.R:149), delta part: delta_scale = 0 keeps the connectivity fixed
        SLX=dict(delta_a=1.01, delta_b=1.01, delta=1.0, delta_scale=0.0, delta_min=1e-12, delta_max=float("inf")),
    )


def default_mh() -> dict:
    """R/42_set_mh.R:13-26.  Quirk Q1: `acc_change` is read from `adjust_burn`
    (line 23), so the default adjustment factor is 0.8, not 0.01."""
    return dict(adjust_burn=0.8, acc_lo=0.20, acc_hi=0.45, acc_change=0.8)


@dataclass
class MHState:
    """Counters of R/20_mh.R:24-31, :73-98."""
    value: float
    scale: float
    accepted: int = 0
    proposed: int = 0
    accepted_tune: int = 0
    proposed_tune: int = 0

    def finalize(self, u: float, prob: float, proposal: float) -> bool:
        """R/20_mh.R:73-81: accept iff isTRUE(u < p); NaN rejects."""
        acc = bool(u < prob)  # NaN compares False
        if acc:
            self.value = proposal
            self.accepted += 1
            self.accepted_tune += 1
        self.proposed += 1
        self.proposed_tune += 1
        return acc

    def tune(self, mh: dict) -> None:
        """R/50_execute.R:62-71 with set_scale<- zeroing counters (R/20_mh.R:88-93);
        the rate is cumulative over *_tune (quirk Q3, R/20_mh.R:97)."""
        rate = self.accepted_tune / max(self.proposed_tune, 1)
        if rate < mh["acc_lo"]:
            self.scale = max(self.scale * (1.0 - mh["acc_change"]), 1e-12)
            self.accepted = self.proposed = 0
        elif rate > mh["acc_hi"]:
            self.scale = self.scale * (1.0 + mh["acc_change"])
            self.accepted = self.proposed = 0


class PhiloxRNG:
    def __init__(self, seed: int, chain: int):
        self.s = Stream(seed, chain)

    def gamma(self, it, slot, shape, rate):
        return self.s.gamma(it, slot, shape, rate)

    def normal_vec(self, it, slot, n):
        return self.s.normal_vec(it, slot, n)

    def normal(self, it, slot):
        return self.s.normal(it, slot)

    def uniform(self, it, slot):
        return self.s.uniform(it, slot)

    def exponential_vec(self, it, slot, rate):
        return self.s.exponential_vec(it, slot, rate)

    def exponential(self, it, slot, rate):
        return self.s.exponential(it, slot, rate)

    def truncated_rw(self, it, slot, loc, scale, lower, upper):
        return self.s.truncated_rw(it, slot, loc, scale, lower, upper)

    def gig_vec(self, it, slot, lam, chi, psi):
        return np.array([self.s.gig(it, slot, lam, float(c), psi, elem=j) for j, c in enumerate(chi)])


class NumpyRNG:
    """Independent stream with library samplers (R's own samplers are nmath
    inversion/rejection codes; any exact sampler is distributionally equal)."""

    def __init__(self, seed: int):
        self.g = np.random.default_rng(seed)

    def gamma(self, it, slot, shape, rate):
        return self.g.gamma(shape) / rate

    def normal_vec(self, it, slot, n):
        return self.g.standard_normal(n)

    def normal(self, it, slot):
        return float(self.g.standard_normal())

    def uniform(self, it, slot):
        return float(self.g.random())

    def exponential_vec(self, it, slot, rate):
        return self.g.exponential(size=len(rate)) / np.asarray(rate)

    def exponential(self, it, slot, rate):
        return float(self.g.exponential()) / rate

    def truncated_rw(self, it, slot, loc, scale, lower, upper):
        while True:
            v = loc + scale * float(self.g.standard_normal())
            if lower < v < upper:
                return v

    def gig_vec(self, it, slot, lam, chi, psi):
        chi = np.maximum(np.asarray(chi, dtype=np.float64), 1e-300)
        omega = np.sqrt(chi * psi)
        alpha = np.sqrt(chi / psi)
        return scipy.stats.geninvgauss.rvs(lam, omega, scale=alpha, random_state=self.g)


@dataclass
class RefChain:
    """Base -> NormalGamma -> {Conjugate, Shrinkage, Horseshoe} with the SLX / SAR /
    SEM mix-ins composed as in R/40_setup.R:11-123."""
    model: str
    y: np.ndarray
    X: np.ndarray
    W: object = None                 # Psi_SAR / Psi_SEM
    X_slx: np.ndarray = None
    W_slx: object = None             # Psi_SLX (defaults to W, R/40_setup.R:99,120)
    prior_type: str = "Independent"
    priors: dict = field(default_factory=default_priors)
    ldet: dict = field(default_factory=lambda: dict(kind="eigen", reps=1))
    rng: object = None
    dense_sem: bool = False
    psi_slx: dict = None             # connectivity FUNCTION Psi_SLX(delta) (R/12_slx.R:69-74): see psi_power()

    def __post_init__(self):
        assert self.model in MODELS and self.prior_type in PRIORS
        pri = default_priors()
        for k, v in (self.priors or {}).items():
            pri[k] = {**pri.get(k, {}), **v}
        self.priors = pri
        self.y = np.asarray(self.y, dtype=np.float64).ravel()
        self.X = np.asarray(self.X, dtype=np.float64)
        self.iterations = 0                       # R/10_lm.R:21
        self.MH = {}
        self.has_sar = self.model in ("sar", "sdm")
        self.has_sem = self.model in ("sem", "sdem")
        self.has_slx = self.model in ("slx", "sdm", "sdem")
        self.conjugate = self.prior_type == "Conjugate"
        self._setup()

    # -- setup: R/10_lm.R:30-44 and the setup_* chain in ls() order -----------
    def _setup(self):
        if self.W is not None:
            self.W = core.as_csr(self.W)
        # setup_2SLX: R/12_slx.R:51-84 (fixed matrix :60-68; connectivity function + set_delta :69-74, :26-36)
        self.delta = None
        if self.has_slx:
            Xl = np.asarray(self.X_slx, dtype=np.float64).reshape(self.X.shape[0], -1)
            if self.psi_slx is not None:
                assert self.model == "slx", "sampled delta is restated for the SLX model only"
                self.X_base, self.X_lag = self.X, Xl
                self.delta = float(self.priors["SLX"]["delta"])
                self.X = np.hstack([self.X_base, core.spmm(psi_power(self.psi_slx, self.delta), Xl)])
                if self.priors["SLX"]["delta_scale"] > 0:                  # R/12_slx.R:41-46
                    self.MH["delta"] = MHState(self.delta, float(self.priors["SLX"]["delta_scale"]))
            else:
                Ws = core.as_csr(self.W_slx) if self.W_slx is not None else self.W
                self.X = np.hstack([self.X, core.spmm(Ws, Xl)])
        self.N, self.M = self.X.shape
        N, M = self.N, self.M
        # setup_4NG: R/10_lm.R:142-153
        ng = self.priors["NG"]
        self.mu0 = np.broadcast_to(np.asarray(ng["mu"], dtype=np.float64), (M,)).copy()
        self.p0 = np.broadcast_to(np.asarray(ng["precision"], dtype=np.float64), (M,)).copy()
        self.shape0, self.rate0 = float(ng["shape"]), float(ng["rate"])
        self.shape1 = self.shape0 + N / 2.0
        self.lam = 0.0
        # setup_8SEM R/15_sem.R:53-110, setup_9SAR R/15_sar.R:47-98
        self.Wy = self.WX = None
        if self.has_sem or self.has_sar:
            key = "SEM" if self.has_sem else "SAR"
            sp_ = self.priors[key]
            self.Wy = core.spmm(self.W, self.y)
            if self.has_sem:
                self.WX = core.spmm(self.W, self.X)
            self.lam = float(sp_["lam"])
            self.MH["lambda"] = MHState(self.lam, float(sp_["lambda_scale"]))
            self.lam_a, self.lam_b = float(sp_["lambda_a"]), float(sp_["lambda_b"])
            self.lam_lo, self.lam_hi = float(sp_["lambda_min"]), float(sp_["lambda_max"])
            self.spline = None
            if self.ldet.get("kind", "eigen") == "eigen":
                self.nodes = core.eigen_nodes(self.W, int(self.ldet.get("reps", 1)))
            elif self.ldet["kind"] == "spline":                # grid = TRUE, R/15_sar.R:79-87, R/15_sem.R:91-99
                from .spline import SplineLogdet
                self.spline = SplineLogdet(self.W, self.ldet.get("i_lambda", (-1.0, 1.0 - 1e-12, 100)),
                                           int(self.ldet.get("reps", 1)), self.ldet.get("values"))
            else:
                self.nodes = (np.asarray(self.ldet["re"]), np.asarray(self.ldet["im"]), np.asarray(self.ldet["w"]))
        self._refresh()
        # setup_HS R/11_shrinkage.R:112-122, setup_SNG :35-49 (run after the numbered ones)
        if self.prior_type == "Horseshoe":
            hs = self.priors["HS"]
            self.hs_lambda = np.full(M, float(hs["lam"]))
            self.hs_nu = np.full(M, float(hs["nu"]))
            self.hs_tau, self.hs_zeta = float(hs["tau"]), float(hs["zeta"])
            self.p0 = 1.0 / (self.hs_lambda * self.hs_tau)
        if self.prior_type == "Shrinkage":
            sng = self.priors["SNG"]
            self.sng_tau = np.full(M, float(sng["tau"]))
            self.sng_lambda, self.sng_theta = float(sng["lam"]), float(sng["theta"])
            self.p0 = 1.0 / self.sng_tau
            if sng["theta_scale"] > 0:                       # R/11_shrinkage.R:27-30
                self.MH["theta"] = MHState(self.sng_theta, float(sng["theta_scale"]))
        # starting_NG: R/10_lm.R:155-164 (Conjugate: :218-221 draws beta then sigma)
        self.mu1 = np.zeros(M)
        if self.conjugate:
            self._sample_beta(0)
            self._sample_sigma(0)
        else:
            self.beta = (np.linalg.solve(self.XXs, self.Xys) if ng.get("beta") is None
                         else np.broadcast_to(np.asarray(ng["beta"], dtype=np.float64), (M,)).copy())
            self.sigma2 = (core.sq_sum(self.ys - self.Xs @ self.beta) / (N - M) if ng.get("sigma") is None
                           else float(ng["sigma"]))

    def _refresh(self):
        """Working (y*, X*, XX*, Xy*) at the current lambda: the active bindings of
        R/15_sar.R:132-133 (z, X'z) and R/15_sem.R:28-36, :136-139 (filtered data)."""
        if self.has_sar:
            self.ys = self.y - self.lam * self.Wy          # R/15_sar.R:29
            self.Xs = self.X
        elif self.has_sem:
            if self.dense_sem:                             # literal R/15_sem.R:29-32
                S = np.eye(self.N) - self.lam * np.asarray(self.W.todense())
                self.ys, self.Xs = S @ self.y, S @ self.X
            else:
                self.ys = self.y - self.lam * self.Wy
                self.Xs = self.X - self.lam * self.WX
        else:
            self.ys, self.Xs = self.y, self.X
        if not self.has_sar or not hasattr(self, "XXs"):
            self.XXs = self.Xs.T @ self.Xs                 # R/10_lm.R:36, R/15_sem.R:35
        self.Xys = self.Xs.T @ self.ys                     # R/15_sar.R:133, R/15_sem.R:34

    # -- Gibbs steps -----------------------------------------------------------
    def _sample_sigma(self, it):
        """R/10_lm.R:174-179; Conjugate R/10_lm.R:226-230 (residuals at mu1)."""
        b = self.mu1 if self.conjugate else self.beta
        self.rate1 = self.rate0 + core.sq_sum(self.ys - self.Xs @ b) / 2.0
        self.sigma2 = 1.0 / self.rng.gamma(it, Slot.SIGMA, self.shape1, self.rate1)

    def _sample_beta(self, it):
        """R/10_lm.R:166-172 + rmvn R/00_aux.R:54-70; Conjugate R/10_lm.R:222-225 (Q5)."""
        P1, self.mu1 = core.beta_moments(self.XXs, self.Xys, getattr(self, "sigma2", 1.0),
                                         self.p0, self.mu0, self.conjugate)
        R = np.linalg.cholesky(P1).T                       # chol(): upper, R'R = P1
        eps = self.rng.normal_vec(it, Slot.BETA, self.M)
        self.beta = self.mu1 + np.linalg.solve(R, eps)     # backsolve(R, eps)

    def _sample_shrinkage(self, it):
        if self.prior_type == "Horseshoe":                 # R/11_shrinkage.R:124-141
            b2 = self.beta ** 2
            self.hs_lambda = 1.0 / self.rng.exponential_vec(
                it, Slot.HS_LAMBDA, 1.0 / self.hs_nu + b2 / (2.0 * self.hs_tau * self.sigma2))
            self.hs_tau = 1.0 / self.rng.gamma(
                it, Slot.HS_TAU, (self.M + 1) / 2.0,
                1.0 / self.hs_zeta + float(np.sum(b2 / self.hs_lambda)) / (2.0 * self.sigma2))
            self.hs_nu = 1.0 / self.rng.exponential_vec(it, Slot.HS_NU, 1.0 + 1.0 / self.hs_lambda)
            self.hs_zeta = 1.0 / self.rng.exponential(it, Slot.HS_ZETA, 1.0 + 1.0 / self.hs_tau)
            self.p0 = 1.0 / np.maximum(self.hs_lambda * self.hs_tau, 1e-12)
        elif self.prior_type == "Shrinkage":               # R/11_shrinkage.R:51-72
            sng = self.priors["SNG"]
            self.sng_lambda = self.rng.gamma(
                it, Slot.SNG_LAMBDA, sng["lambda_a"] + self.M * self.sng_theta,
                sng["lambda_b"] + self.sng_theta * float(np.sum(self.sng_tau)) / 2.0)
            self.sng_tau = np.maximum(self.rng.gig_vec(
                it, Slot.SNG_TAU, self.sng_theta - 0.5, (self.beta - self.mu0) ** 2,
                self.sng_lambda * self.sng_theta), 1e-12)
            self.p0 = 1.0 / self.sng_tau
            if "theta" in self.MH:                         # R/20_mh.R:116-146
                mh = self.MH["theta"]
                prop = math.exp(mh.scale * self.rng.normal(it, Slot.SNG_THETA_PROP)) * mh.value
                rate = float(sng["theta_a"])
                lp = (core.logpost_theta(prop, self.sng_tau, self.sng_lambda, rate)
                      - core.logpost_theta(mh.value, self.sng_tau, self.sng_lambda, rate)
                      + math.log(prop) - math.log(mh.value))
                mh.finalize(self.rng.uniform(it, Slot.SNG_THETA_ACC), _exp(lp), prop)
                self.sng_theta = mh.value

    def _ldet(self, v):
        if self.spline is not None:
            return self.spline(v)
        return core.ldet_nodes(v, *self.nodes)

    def _sample_lambda(self, it):
        """SAR: sample_latent R/15_sar.R:101-125.  SEM: sample_volatility R/15_sem.R:113-129."""
        mh = self.MH["lambda"]
        if self.has_sar:
            e0e0, e1e0, e1e1 = core.sar_rss_coeffs(self.y, self.Wy, self.X, self.XXs,
                                                   self.sigma2, self.p0, self.mu0)
            rss = lambda v: e0e0 - 2.0 * v * e1e0 + v * v * e1e1     # noqa: E731
            ps, as_ = Slot.LAMBDA_PROP, Slot.LAMBDA_ACC
        else:
            rss = lambda v: core.sem_rss(v, self.y, self.Wy, self.X, self.WX)   # noqa: E731
            ps, as_ = Slot.RHO_PROP, Slot.RHO_ACC
        prop = self.rng.truncated_rw(it, ps, mh.value, mh.scale, self.lam_lo, self.lam_hi)
        post = lambda v: core.logpost_lambda(v, self._ldet(v), rss(v), self.N, self.M,   # noqa: E731
                                             self.lam_a, self.lam_b)
        prob = _exp(post(prop) - post(mh.value))
        mh.finalize(self.rng.uniform(it, as_), prob, prop)
        if self.has_sem or abs(mh.value - self.lam) > 1e-12:   # R/15_sar.R:124, R/15_sem.R:128
            self.lam = mh.value
            self._refresh()

    def _sample_delta(self, it):
        """sample_extra3, R/12_slx.R:87-106 + MH_SLX_delta R/20_mh.R:229-278: random-walk MH on the connectivity
        parameter; rss(v) re-solves for beta with X(v) = [X, Psi(v) X_SLX] at the current sigma^2 and P0 (:91-97);
        prior dgamma(1 / v, a, rate = b, log) without a Jacobian (quirk Q7, :269); truncated proposal, no Hastings
        correction (:245-263)."""
        mh = self.MH["delta"]
        sl = self.priors["SLX"]
        a, b = float(sl["delta_a"]), float(sl["delta_b"])

        def rss(v):
            X = np.hstack([self.X_base, core.spmm(psi_power(self.psi_slx, v), self.X_lag)])
            _, beta = core.beta_moments(X.T @ X, X.T @ self.y, self.sigma2, self.p0, self.mu0)
            return core.sq_sum(self.y - X @ beta)

        def post(v):
            r = rss(v)
            if not r > 0.0:
                return float("nan")
            x = 1.0 / v
            return -0.5 * (self.N - self.M) * math.log(r) + (a * math.log(b) - math.lgamma(a) + (a - 1.0) * math.log(x) - b * x)
        prop = self.rng.truncated_rw(it, Slot.DELTA_PROP, mh.value, mh.scale, float(sl["delta_min"]), float(sl["delta_max"]))
        prob = _exp(post(prop) - post(mh.value))
        mh.finalize(self.rng.uniform(it, Slot.DELTA_ACC), prob, prop)
        if abs(mh.value - self.delta) > 1e-12:                 # set_delta, R/12_slx.R:26-36, :104
            self.delta = mh.value
            self.X = np.hstack([self.X_base, core.spmm(psi_power(self.psi_slx, self.delta), self.X_lag)])
            self._refresh()

    def sample(self):
        """Base$sample, R/10_lm.R:46-63: sigma, beta, shrinkage, latent, volatility, extra1-3, finalize."""
        it = self.iterations + 1
        self._sample_sigma(it)
        self._sample_beta(it)
        self._sample_shrinkage(it)
        if self.has_sar or self.has_sem:
            self._sample_lambda(it)
        if "delta" in self.MH:
            self._sample_delta(it)
        self.iterations += 1                               # R/10_lm.R:73-77

    # -- R/10_lm.R:193, R/15_sar.R:138-142, R/15_sem.R:142-146 -----------------
    def get_parameters(self) -> np.ndarray:
        out = list(self.beta) + [math.sqrt(self.sigma2)]
        if self.delta is not None:                         # pars$delta_SLX, R/12_slx.R:117-121
            out.append(self.delta)
        if self.has_sar or self.has_sem:
            out.append(self.lam)
        return np.array(out)

    def parameter_names(self):
        names = ["beta"] if self.M == 1 else [f"beta{i + 1}" for i in range(self.M)]
        names.append("sigma")
        if self.delta is not None:
            names.append("delta_SLX")
        if self.has_sar:
            names.append("lambda_SAR")
        if self.has_sem:
            names.append("lambda_SEM")
        return names


def psi_power(psi: dict, delta: float) -> sp.csr_matrix:
    """The connectivity family the C ABI offers for `Psi_SLX(delta)` (the reference takes any R closure; the paper's is
    inverse-distance decay, paper/0_data.R:31-35): W(delta)_ij = dist_ij ^ -delta on a fixed sparsity pattern,
    normalise = "row" divides every row by its sum, "none" leaves it."""
    D = psi["dist"]
    W = sp.csr_matrix((np.power(D.data, -delta), D.indices, D.indptr), shape=D.shape)
    if psi.get("normalise", "row") == "row":
        rs = np.asarray(W.sum(axis=1)).ravel()
        rs[rs == 0.0] = 1.0
        W = sp.diags(1.0 / rs) @ W
    return sp.csr_matrix(W)


def _exp(x: float) -> float:
    try:
        return math.exp(x)
    except OverflowError:
        return float("inf")


def tune(chain: RefChain, n_burn: int, mh: dict | None = None) -> None:
    """R/50_execute.R:50-83."""
    mh = mh or default_mh()
    for i in range(1, n_burn + 1):
        chain.sample()
        if i % 10 == 0 and i <= mh["adjust_burn"] * n_burn:
            for obj in chain.MH.values():
                obj.tune(mh)


def sample(chain: RefChain, n_save: int = 1000, n_burn: int = 0, mh: dict | None = None) -> np.ndarray:
    """R/50_execute.R:12-42: burn-in/tune, then store unlist(get_parameters) per draw."""
    if n_burn > 0:
        tune(chain, n_burn, mh)
    storage = np.full((n_save, len(chain.get_parameters())), np.nan)
    for i in range(n_save):
        chain.sample()
        storage[i] = chain.get_parameters()
    return storage
