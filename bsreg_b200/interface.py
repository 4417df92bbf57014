"""Host-side mirror of bsreg's user interface, on top of the C ABI (include/bsreg_cuda.h).

The reference's host language is R, which this image does not have, so the layer an R
user touches is mirrored here in Python with the same names, argument meaning, defaults
and error messages (the R shell that binds the same C ABI is in rpkg/):

    bm / blm / bslx / bsar / bsem / bsdm / bsdem      R/60_interface.R:32-101
    set_options, set_NG, set_SNG, set_HS,
    set_SAR, set_SLX, set_SEM                          R/41_set_options.R:20-153
    set_mh                                             R/42_set_mh.R:13-26
    get_blm ... get_bsdem (model factories)            R/40_setup.R:11-123
    print / summary / as.mcmc on the `bm` object       R/80_methods.R, R/81_coda.R

What changes underneath: `sample()`/`tune()` (R/50_execute.R:12-83) do not loop in the
host language; the whole burn-in + sampling loop runs on the GPU behind `bsreg_run`.
There is no CPU fallback: without the CUDA library / a CUDA device every model call raises.

Extensions (reference-compatible defaults): `n_chains`, `seed`, `devices`, sparse `W`
(scipy.sparse), and `ldet_SAR/ldet_SEM = dict(method="chebyshev", order=, probes=)` for
large N next to the reference's eigenvalue method (`dict(grid=False, reps=)`).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import capi

TYPES = ("lm", "slx", "sar", "sem", "sdm", "sdem", "sv")
PRIOR_TYPES = ("Independent", "Conjugate", "Shrinkage", "Horseshoe")


# ---------------------------------------------------------------------------
# R/00_aux.R:15-23
# ---------------------------------------------------------------------------
def num_check(x, min=0.0, max=math.inf, msg="Please check the numeric parameters."):
    """`if(!is.numeric(x) || length(x) != 1 || x < min || x > max) stop(msg)`"""
    if isinstance(x, (bool, str)) or x is None:
        raise ValueError(msg)
    a = np.asarray(x)
    if a.size != 1 or not np.issubdtype(a.dtype, np.number):
        raise ValueError(msg)
    v = float(a.reshape(()))
    if math.isnan(v) or v < min or v > max:      # R: NA in `if` is an error as well
        raise ValueError(msg)
    return v


def _vcheck(x, **kw):
    """`vapply(x, num_check, numeric(1L), ...)`"""
    vals = [num_check(v, **kw) for v in np.atleast_1d(np.asarray(x, dtype=object))]
    return vals[0] if len(vals) == 1 else np.array(vals)


def match_arg(arg, choices, what="type"):
    """R's match.arg: NULL/default picks the first choice; partial matching allowed."""
    if arg is None or (isinstance(arg, (tuple, list)) and tuple(arg) == tuple(choices)):
        return choices[0]
    hits = [c for c in choices if c == arg] or [c for c in choices if c.startswith(str(arg))]
    if len(hits) != 1:
        raise ValueError(f"'{what}' should be one of " + ", ".join(f"'{c}'" for c in choices))
    return hits[0]


# ---------------------------------------------------------------------------
# R/42_set_mh.R, R/41_set_options.R
# ---------------------------------------------------------------------------
def set_mh(adjust_burn=0.8, acc_target=(0.20, 0.45), acc_change=0.01):
    """R/42_set_mh.R:13-26.  Reference quirk Q1 is replicated: `acc_change` is validated
    and TAKEN from `adjust_burn` (line 23), so the default adjustment factor is 0.8."""
    return dict(
        adjust_burn=num_check(adjust_burn, min=0, max=1, msg="Please provide a valid length for the MH tuning "
                              "period (in percent of burn-in) via 'adjust_burn'."),
        acc_target=tuple(num_check(a, min=0, max=1, msg="Please provide a valid target range for the MH tuning "
                                   "(in percent of acceptance) via 'acc_target'.") for a in acc_target),
        acc_change=num_check(adjust_burn, min=0, max=1e6, msg="Please provide a valid scale adjustment factor "
                             "for the MH tuning period (in percent) via 'acc_change'."),
        _class="mh_settings")


def set_NG(mu=0, precision=1e-8, shape=0.01, rate=0.01, beta=None, sigma=None):
    """R/41_set_options.R:50-61."""
    return dict(
        mu=_vcheck(mu, min=-math.inf, max=math.inf, msg="Please provide a valid value for the prior mean via 'mu'."),
        precision=_vcheck(precision, min=0, max=math.inf, msg="Please provide a valid prior variance via 'precision'."),
        shape=num_check(shape, min=1e-12, msg="Please provide a valid prior shape via 'shape'."),
        rate=num_check(rate, min=1e-12, msg="Please provide a valid prior rate via 'rate'."),
        beta=beta, sigma=sigma, _class="prior_NG")


def set_SNG(lambda_a=0.01, lambda_b=0.01, theta_scale=0, theta_a=1, lambda_=1, tau=10, theta=0.1, **kw):
    """R/41_set_options.R:65-84 (`lambda` is a Python keyword: pass `lambda_=` or `**{"lambda": v}`)."""
    lambda_ = kw.pop("lambda", lambda_)
    if kw:
        raise TypeError(f"unused argument(s): {sorted(kw)}")
    return dict(
        lambda_a=num_check(lambda_a, min=1e-12, msg="Please provide a valid shape for lambda (Normal-Gamma) via 'lambda_a'."),
        lambda_b=num_check(lambda_b, min=1e-12, msg="Please provide a valid rate for lambda (Normal-Gamma) via 'lambda_b'."),
        theta_scale=num_check(theta_scale, min=0, msg="Please provide a valid proposal scale for theta (Normal-Gamma) via 'theta_scale'."),
        theta_a=num_check(theta_a, min=0, msg="Please provide a valid rate for theta (Normal-Gamma) via 'theta_a'."),
        lam=num_check(lambda_, min=1e-12, msg="Please provide a valid starting value for 'lambda' (Normal-Gamma)."),
        tau=_vcheck(tau, min=1e-12, msg="Please provide a valid starting value for 'tau' (Normal-Gamma)."),
        theta=num_check(theta, min=1e-12, msg="Please provide a valid starting value for 'theta' (Normal-Gamma)."),
        _class="prior_SNG")


def set_HS(lambda_=1, tau=1, zeta=1, nu=1, **kw):
    """R/41_set_options.R:88-100."""
    lambda_ = kw.pop("lambda", lambda_)
    if kw:
        raise TypeError(f"unused argument(s): {sorted(kw)}")
    return dict(
        lam=_vcheck(lambda_, min=1e-12, msg="Please provide a valid starting value for 'lambda' (Horseshoe)."),
        tau=num_check(tau, min=1e-12, msg="Please provide a valid starting value for 'tau' (Horseshoe)."),
        zeta=num_check(zeta, min=1e-12, msg="Please provide a valid starting value for 'zeta' (Horseshoe)."),
        nu=_vcheck(nu, min=1e-12, msg="Please provide a valid starting value for 'nu' (Horseshoe)."),
        _class="prior_HS")


def set_SAR(lambda_a=1.01, lambda_b=1.01, lambda_=0, lambda_scale=0.1, lambda_min=-1, lambda_max=1 - 1e-12,
            delta_a=1.01, delta_b=1.01, delta=1, delta_scale=0, delta_min=1e-12, delta_max=math.inf, **kw):
    """R/41_set_options.R:115-145; `set_SLX` and `set_SEM` are the same function (:149-152)."""
    lambda_ = kw.pop("lambda", lambda_)
    if kw:
        raise TypeError(f"unused argument(s): {sorted(kw)}")
    lo = num_check(lambda_min, min=-math.inf, msg="Please provide a valid lower bound for lambda (spatial) via 'lambda_min'.")
    hi = num_check(lambda_max, min=-math.inf, msg="Please provide a valid upper bound for lambda (spatial) via 'lambda_max'.")
    dlo = num_check(delta_min, min=-math.inf, msg="Please provide a valid lower bound for delta (spatial) via 'delta_min'.")
    dhi = num_check(delta_max, min=-math.inf, msg="Please provide a valid upper bound for delta (spatial) via 'delta_max'.")
    return dict(
        lambda_a=num_check(lambda_a, min=1e-12, msg="Please provide a valid shape for lambda (spatial) via 'lambda_a'."),
        lambda_b=num_check(lambda_b, min=1e-12, msg="Please provide a valid rate for lambda (spatial) via 'lambda_b'."),
        lambda_min=lo, lambda_max=hi,
        lam=num_check(lambda_, min=lo, max=hi, msg="Please provide a valid starting value for lambda (spatial) via 'lambda'."),
        lambda_scale=num_check(lambda_scale, min=1e-12, msg="Please provide a valid proposal scale for lambda (spatial) via 'lambda_scale'."),
        delta_a=num_check(delta_a, min=1e-12, msg="Please provide a valid shape for delta (spatial) via 'delta_a'."),
        delta_b=num_check(delta_b, min=1e-12, msg="Please provide a valid rate for delta (spatial) via 'delta_b'."),
        delta_min=dlo, delta_max=dhi,
        delta=num_check(delta, min=dlo, max=dhi, msg="Please provide a valid starting value for delta (spatial) via 'delta'."),
        delta_scale=num_check(delta_scale, min=0, msg="Please provide a valid proposal scale for delta (spatial) via 'delta_scale'."),
        _class="prior_SAR")


set_SLX = set_SAR
set_SEM = set_SAR


def set_SV(*_a, **_k):
    """R/41_set_options.R:160-185 wraps the third-party `stochvol` sampler; the stochastic
    volatility model is outside the spatial hot path."""
    raise NotImplementedError("the stochastic volatility model (bsv / set_SV, R/13_stochvol.R) is out of scope "
                              "of the GPU hot path")


def set_options(type=PRIOR_TYPES, NG=None, SNG=None, HS=None, SAR=None, SLX=None, SEM=None, SV=None, **extra):
    """R/41_set_options.R:20-36.  Unlike the reference (quirk Q8) the default does not call
    set_SV(), so no model call depends on `stochvol`."""
    return dict(type=match_arg(type, PRIOR_TYPES),
                priors=dict(NG=NG or set_NG(), SNG=SNG or set_SNG(), HS=HS or set_HS(), SAR=SAR or set_SAR(),
                            SLX=SLX or set_SLX(), SEM=SEM or set_SEM(), SV=SV, **extra),
                _class="priors")


# ---------------------------------------------------------------------------
# formula -> (y, X): the subset of model.frame/model.matrix the reference's examples use
# ---------------------------------------------------------------------------
_FUNCS = dict(log=np.log, exp=np.exp, sqrt=np.sqrt, abs=np.abs, log10=np.log10, log2=np.log2, I=lambda v: v)


def _split_terms(rhs):
    terms, depth, cur = [], 0, ""
    for ch in rhs:
        depth += ch == "("
        depth -= ch == ")"
        if ch in "+-" and depth == 0 and cur.strip():
            terms.append(cur.strip())
            cur = ch if ch == "-" else ""
        else:
            cur += ch
    if cur.strip():
        terms.append(cur.strip())
    return terms


def model_matrix(formula, data):
    """`y ~ a + log(b) + factor(g)` on a mapping of columns (dict / DataFrame).  Intercept
    first unless `- 1` / `+ 0`; factors expand to treatment dummies (first level dropped when
    there is an intercept), levels in sorted order as R's factor() does."""
    lhs, rhs = (s.strip() for s in str(formula).split("~", 1))
    env = dict(_FUNCS)
    env.update({k: np.asarray(data[k]) for k in (data.keys() if hasattr(data, "keys") else data.dtype.names)})

    def ev(expr):
        return np.asarray(eval(expr.replace("^", "**"), {"__builtins__": {}}, env))   # noqa: S307 (formula DSL)
    y = ev(lhs).astype(np.float64)
    intercept, cols, names = True, [], []
    for t in _split_terms(rhs):
        if t in ("0", "-1", "- 1"):
            intercept = False
            continue
        if t == "1":
            continue
        mfac = re.fullmatch(r"factor\((.*)\)", t)
        if mfac:
            v = ev(mfac.group(1))
            levels = sorted(set(v.tolist()))
            for lv in (levels[1:] if (intercept or cols) else levels):
                cols.append((v == lv).astype(np.float64))
                names.append(f"{t}{lv}")
        else:
            cols.append(ev(t).astype(np.float64))
            names.append(t)
    if intercept:
        cols.insert(0, np.ones_like(y))
        names.insert(0, "(Intercept)")
    return y, np.column_stack(cols), names


# ---------------------------------------------------------------------------
# W handling and the log-determinant set-up (R/15_sar.R:47-98 = R/15_sem.R:53-110)
# ---------------------------------------------------------------------------
def _as_csr(W, n, what):
    if W is None:
        raise ValueError(f"Please provide a connectivity matrix '{what}'.")          # R/15_sar.R:52, R/15_sem.R:58
    W = sp.csr_matrix(W, dtype=np.float64)
    if W.shape != (n, n):
        raise ValueError(f"'{what}' must be an N x N matrix (N = {n}), got {W.shape}")
    W.sum_duplicates()
    W.eliminate_zeros()
    W.sort_indices()
    return W


def is_symmetric(W):
    """R/00_aux.R:92-95."""
    return abs(W - W.T).max() <= 100 * np.finfo(float).eps if W.nnz else True


DEFAULT_LDET = dict(grid=False, i_lambda=(-1, 1 - 1e-12, 100), reps=1)
EIGEN_MAX_N = 6000        # dense eigen() beyond this is the reference's O(N^3) wall; use Chebyshev


def ldet_nodes(W, ldet):
    """Nodes (re, im, weight) for `Re(sum(log(1 - lambda ev))) * reps`, R/15_sar.R:68-76,91-96.
    As in the reference a partial `ldet_*` list REPLACES the default list (`grid` missing = FALSE)."""
    reps = int(ldet.get("reps", 1) or 1)
    n = W.shape[0]
    size = n // reps
    if size * reps != n:
        raise ValueError("'reps' must divide the number of observations")
    Wsub = W[:size, :size].toarray()                                   # get_W(), R/15_sar.R:72-76
    ev = (np.linalg.eigvalsh(Wsub) if is_symmetric(W) else np.linalg.eigvals(Wsub)).astype(np.complex128)
    return ev.real.copy(), ev.imag.copy(), np.full(size, float(reps))


# ---------------------------------------------------------------------------
# model object: what `sample()` and the methods see of the R6 classes
# ---------------------------------------------------------------------------
class MHView:
    """get_* bindings of R/20_mh.R:85-98 for MH object `k` (0 = spatial lambda, 1 = SNG theta) of chain `c`."""

    def __init__(self, model, k, name):
        self._m, self._k, self.name = model, k, name

    def _st(self):
        return self._m.handle.get_state()

    def get_value(self, c=0):
        return float(self._st()["mh_value"][c, self._k])

    def get_scale(self, c=0):
        return float(self._st()["mh_scale"][c, self._k])

    def get_accepted(self, c=0):
        return int(self._st()["mh_counters"][c, self._k, 0])

    def get_proposed(self, c=0):
        return int(self._st()["mh_counters"][c, self._k, 1])

    def get_acceptance(self, c=0):
        cnt = self._st()["mh_counters"][c, self._k]
        return cnt[0] / max(cnt[1], 1)

    def get_tuning(self, c=0):
        cnt = self._st()["mh_counters"][c, self._k]
        return cnt[2] / max(cnt[3], 1)


class DeltaMHView(MHView):
    """MH object SLX_delta (R/12_slx.R:41-46): same get_* bindings, state kept per chain on the device."""

    def __init__(self, model):
        self._m, self.name = model, "SLX_delta"

    def _st(self):
        d = self._m.handle.get_delta_state()
        return dict(mh_value=d["delta"][:, None], mh_scale=d["scale"][:, None], mh_counters=d["counters"][:, None, :])

    _k = 0


class PsiPower:
    """Connectivity function `Psi(delta) = dist ^ -delta` on the sparsity pattern of `dist` (a sparse matrix of positive
    distances), rows divided by their sum if normalise = "row": what `Psi_SLX` may be besides a fixed matrix
    (R/12_slx.R:69-74 takes any closure; this is the paper's inverse-distance family, paper/0_data.R:31-35)."""

    def __init__(self, dist, normalise="row"):
        self.dist = sp.csr_matrix(dist, dtype=np.float64)
        self.dist.sort_indices()
        if normalise not in ("row", "none"):
            raise ValueError("normalise should be one of 'row', 'none'")
        if self.dist.nnz and not np.all(self.dist.data > 0):
            raise ValueError("distances must be positive")
        self.normalise = normalise

    def __call__(self, delta):
        W = sp.csr_matrix((np.power(self.dist.data, -float(delta)), self.dist.indices, self.dist.indptr), shape=self.dist.shape)
        if self.normalise == "row":
            rs = np.asarray(W.sum(axis=1)).ravel()
            rs[rs == 0] = 1.0
            W = sp.diags(1.0 / rs) @ W
        return sp.csr_matrix(W)


@dataclass
class Model:
    """The object stored as `bm$model`; continuing a fit (`bm.bm`, R/60_interface.R:71-78) reuses it."""
    handle: capi.Handle
    type: str
    prior: str
    colnames: list
    n_chains: int
    MH: dict = field(default_factory=dict)

    @property
    def get_parameters(self):
        """unlist(get_parameters) of chain 1: c(beta, sigma = sqrt(sigma^2), lambda_*), R/10_lm.R:193."""
        return self.parameters()[0]

    def parameters(self):
        p = self.handle.get_state()["params"].copy()
        p[:, self.handle.m] = np.sqrt(p[:, self.handle.m])
        return p

    @property
    def get_meta(self):
        """R/10_lm.R:90 + the print method's text (R/10_lm.R:116-127)."""
        st = self.handle.get_state()
        out = f"{int(st['iterations'][0])} iterations"
        for name, mh in self.MH.items():
            out += f", {name} acceptance {mh.get_acceptance():.3f}"
        return out

    def trace_inverse(self, lam):
        """tr((I - lam W)^-1) for an array of lam, from the SAME spectral nodes the sampler's log-determinant uses
        (eigenvalues with weight `reps`, or the Chebyshev nodes with weights folded from the Hutchinson trace moments,
        SURVEY App. B): sum_k w_k Re 1 / (1 - lam omega_k).  Replaces the dense
        `sum(diag(solve(diag(N) - lambda * W)))` of R/15_sar.R:145-146 (marked "To-do: ... more efficient" there)."""
        lam = np.atleast_1d(np.asarray(lam, dtype=np.float64))
        x, y, b, c, d = self.handle.spline()
        if len(x):
            # grid + spline log-determinant: tr((I - lam W)^-1) = N + lam tr(W (I - lam W)^-1) = N - lam d/dlam log|I - lam W|,
            # with the derivative of the same spline the sampler evaluates
            i = np.clip(np.searchsorted(x, lam, side="right") - 1, 0, len(x) - 1)
            dx = lam - x[i]
            return self.handle.n - lam * (b[i] + dx * (2.0 * c[i] + dx * 3.0 * d[i]))
        re, im, w = self.handle.nodes()
        z = 1.0 - lam[:, None] * (re[None, :] + 1j * im[None, :])
        return np.sum(w[None, :] * (1.0 / z).real, axis=1)

    @property
    def get_effects(self):
        """`mdl$get_effects` at the current state of chain 1, R/15_sar.R:143-149 (SAR / SDM only, as in the reference)."""
        if self.type not in ("sar", "sdm"):
            raise AttributeError("get_effects is defined for the SAR / SDM models only (R/15_sar.R:143)")
        p = self.get_parameters
        m = self.handle.m
        return effects_from_draws(p[None, :m], p[None, -1:].ravel(), self.trace_inverse(p[-1]) / self.handle.n)

    def sample(self, n_save=1000, n_burn=0, mh=None, verbose=True):
        """`sample(mdl, n_save, n_burn, mh, verbose)`, R/50_execute.R:12-42 -- on the device.
        Returns (n_chains, n_save, P)."""
        mh = mh or set_mh()
        prog = None
        if verbose:
            print(f"Starting sampler with {n_burn} burn-in and {n_save} saved draws"
                  + (f" for {self.n_chains} chains" if self.n_chains > 1 else "") + ".")
            prog = lambda done, total: False          # noqa: E731  (hook for a progress bar / interrupt)
        d = self.handle.run(int(n_burn), int(n_save),
                            (mh["adjust_burn"], mh["acc_target"][0], mh["acc_target"][1], mh["acc_change"]),
                            progress=prog)
        return np.ascontiguousarray(d.transpose(0, 2, 1))


def effects_from_draws(beta, lam, tau):
    """R/15_sar.R:144-148 per draw: total = beta / (1 - lambda), direct = tr((I - lambda W)^-1) / N * beta,
    indirect = total - direct.  beta (S, M), lam (S,), tau = tr(.)/N (S,)."""
    total = beta / (1.0 - lam)[:, None]
    direct = tau[:, None] * beta
    return {"total": total, "direct": direct, "indirect": total - direct}


class BM:
    """`structure(list(draws, model, call), class = "bm")`, R/60_interface.R:65.
    `draws` is the n_save x P matrix of the first chain (exactly the reference object for
    n_chains = 1); `chains` holds all chains as (n_chains, n_save, P)."""

    def __init__(self, chains, model, call):
        self.chains, self.model, self.call = chains, model, call

    @property
    def draws(self):
        return self.chains[0]

    @property
    def colnames(self):
        return self.model.colnames

    def __str__(self):   # print.bm, R/80_methods.R:3-11
        coefs = self.draws.mean(0)
        body = "  ".join(f"{n}={v:.3g}" for n, v in zip(self.colnames, coefs))
        return (f"\nCall:\n{self.call}\n\nDraws (total): {self.draws.shape[0]} ({self.model.get_meta})\n"
                f"Coefficients (mean):\n{body}\n")

    def summary(self):   # summary.bm -> summary(matrix), R/80_methods.R:15-17
        q = np.quantile(self.draws, [0, 0.25, 0.5, 0.75, 1], axis=0)
        return {n: dict(zip(("Min.", "1st Qu.", "Median", "Mean", "3rd Qu.", "Max."),
                            (q[0, j], q[1, j], q[2, j], self.draws[:, j].mean(), q[3, j], q[4, j])))
                for j, n in enumerate(self.colnames)}

    def effects(self, chain=0):
        """Direct / indirect / total effects of every saved draw of a SAR / SDM fit (what `get_effects`, R/15_sar.R:143-149,
        gives for one state): dict of (n_save, M) arrays."""
        if self.model.type not in ("sar", "sdm"):
            raise AttributeError("effects are defined for the SAR / SDM models only (R/15_sar.R:143)")
        d = self.chains[chain]
        m = self.model.handle.m
        lam = d[:, -1]
        return effects_from_draws(d[:, :m], lam, self.model.trace_inverse(lam) / self.model.handle.n)

    def as_mcmc(self):   # as.mcmc.bm, R/81_coda.R:17-30: the draws matrix with column names
        return self.draws, self.colnames

    def as_mcmc_list(self):
        return [c for c in self.chains], self.colnames


# ---------------------------------------------------------------------------
# factories + bm(): R/40_setup.R, R/60_interface.R
# ---------------------------------------------------------------------------
_MODEL = dict(lm=capi.MODEL_LM, slx=capi.MODEL_LM, sar=capi.MODEL_SAR, sdm=capi.MODEL_SAR,
              sem=capi.MODEL_SEM, sdem=capi.MODEL_SEM)
_PRIOR = dict(Independent=capi.PRIOR_INDEPENDENT, Conjugate=capi.PRIOR_CONJUGATE,
              Shrinkage=capi.PRIOR_SHRINKAGE, Horseshoe=capi.PRIOR_HORSESHOE)


def _scalar(v, what):
    a = np.atleast_1d(np.asarray(v, dtype=np.float64))
    if not np.all(a == a[0]):
        raise NotImplementedError(f"vector-valued starting values for {what} are not supported by the C ABI yet")
    return float(a[0])


def get_model(type, y, X, options=None, Psi=None, X_SLX=None, Psi_SLX=None, ldet_SAR=None, ldet_SEM=None,
              n_chains=1, chain_offset=0, seed=0, devices=None, stats="stream"):
    """get_blm / get_bslx / get_bsar / get_bsem / get_bsdm / get_bsdem, R/40_setup.R:11-123."""
    options = options or set_options()
    pri = options["priors"]
    y = np.asarray(y, dtype=np.float64).ravel()
    X = np.asarray(X, dtype=np.float64)
    X = X.reshape(len(y), -1)
    n = len(y)
    has_slx = type in ("slx", "sdm", "sdem")
    spatial = "SAR" if type in ("sar", "sdm") else ("SEM" if type in ("sem", "sdem") else None)
    W = W_slx = None
    kw = {}
    psi_fun = None                                                      # Psi_SLX given as a connectivity FUNCTION of delta
    if has_slx:
        cand = Psi_SLX if Psi_SLX is not None else Psi
        if isinstance(cand, PsiPower):
            if type != "slx":
                raise NotImplementedError("a connectivity function Psi(delta) is built for the SLX model; SDM / SDEM take a "
                                          "fixed connectivity matrix (R/12_slx.R:87-106 is model-agnostic in the reference)")
            psi_fun, Psi_SLX, Psi = cand, cand.dist, (cand.dist if Psi_SLX is None else Psi)
            sl = pri["SLX"]
            kw["slx_psi"] = dict(normalise=cand.normalise, delta=sl["delta"], delta_scale=sl["delta_scale"],
                                 delta_a=sl["delta_a"], delta_b=sl["delta_b"], delta_min=sl["delta_min"],
                                 delta_max=sl["delta_max"])
        elif callable(cand):
            raise NotImplementedError("arbitrary connectivity closures cannot cross the C ABI: use PsiPower(dist) "
                                      "(W(delta) = dist ^ -delta, the paper's family, paper/0_data.R:31-35)")
    if (psi_fun is None and pri["SLX"]["delta_scale"] > 0 and has_slx):
        raise ValueError("Connectivity function 'Psi' (SLX) required to sample 'delta'.")           # R/12_slx.R:79
    if spatial and pri[spatial]["delta_scale"] > 0:
        raise NotImplementedError("sampling a connectivity parameter of Psi_SAR / Psi_SEM: the reference itself keeps "
                                  "these fixed (Psi_fixed <- TRUE, R/15_sar.R:54, R/15_sem.R:60)")
    if spatial:
        W = _as_csr(Psi, n, f"Psi_{spatial}")
        ld = (ldet_SAR if spatial == "SAR" else ldet_SEM)
        ld = dict(DEFAULT_LDET) if ld is None else dict(ld)
        method = ld.get("method", "eigen" if n <= EIGEN_MAX_N * int(ld.get("reps", 1) or 1) else "chebyshev")
        if ld.get("grid") is True:                                        # isTRUE(ldet$grid), R/15_sar.R:69
            # spline over a determinant grid, R/15_sar.R:79-87: the dense LUs run on the device (BSREG_LDET_SPLINE)
            i_lambda = ld.get("i_lambda", DEFAULT_LDET["i_lambda"])
            kw.update(ldet=capi.LDET_SPLINE, ldet_grid=tuple(i_lambda), ldet_reps=int(ld.get("reps", 1) or 1))
        elif method == "eigen":
            kw["nodes"] = ldet_nodes(W, ld)
        elif method == "chebyshev":
            kw.update(ldet=capi.LDET_CHEBYSHEV, cheb_order=int(ld.get("order", 50)), cheb_probes=int(ld.get("probes", 64)))
        else:
            raise ValueError("ldet method should be one of 'eigen', 'chebyshev'")
    if has_slx:
        if X_SLX is None:
            raise ValueError("Please provide the regressors to lag via 'X_SLX'.")
        X_SLX = np.asarray(X_SLX, dtype=np.float64).reshape(n, -1)
        psi_slx = Psi_SLX if Psi_SLX is not None else Psi                # R/40_setup.R:99,120
        Wl = _as_csr(psi_slx, n, "Psi_SLX")
        if W is None:
            W = Wl
        elif Psi_SLX is not None:
            W_slx = Wl
    sng, hs = pri["SNG"], pri["HS"]
    h = capi.Handle(
        y, X, model=_MODEL[type], prior=_PRIOR[options["type"]], W=W, X_slx=X_SLX if has_slx else None, W_slx=W_slx,
        stats_mode=capi.STATS_STREAM if stats == "stream" else capi.STATS_CACHED,
        n_chains=int(n_chains), chain_offset=int(chain_offset), seed=int(seed), devices=devices,
        ng={k: pri["NG"][k] for k in ("mu", "precision", "shape", "rate", "beta", "sigma")},
        sng=dict(lambda_a=sng["lambda_a"], lambda_b=sng["lambda_b"], theta_scale=sng["theta_scale"], theta_a=sng["theta_a"],
                 lam=sng["lam"], tau=_scalar(sng["tau"], "tau (Normal-Gamma)"), theta=sng["theta"]),
        hs=dict(lam=_scalar(hs["lam"], "lambda (Horseshoe)"), tau=hs["tau"], zeta=hs["zeta"],
                nu=_scalar(hs["nu"], "nu (Horseshoe)")),
        spatial={k: pri[spatial][k] for k in ("lambda_a", "lambda_b", "lam", "lambda_scale", "lambda_min", "lambda_max")}
        if spatial else None, **kw)
    m = h.m
    names = (["beta"] if m == 1 else [f"beta{i + 1}" for i in range(m)]) + ["sigma"]     # unlist() naming
    if psi_fun is not None:
        names.append("delta_SLX")                                       # R/12_slx.R:117-121
    if spatial:
        names.append(f"lambda_{spatial}")
    mdl = Model(h, type, options["type"], names, int(n_chains))
    if spatial:
        mdl.MH[f"{spatial}_lambda"] = MHView(mdl, 0, f"{spatial}_lambda")
    if options["type"] == "Shrinkage" and sng["theta_scale"] > 0:
        mdl.MH["SNG_theta"] = MHView(mdl, 1, "SNG_theta")
    if psi_fun is not None and pri["SLX"]["delta_scale"] > 0:
        mdl.MH["SLX_delta"] = DeltaMHView(mdl)
    return mdl


def bm(x, data=None, n_save=1000, n_burn=None, options=None, mh=None, verbose=True, W=None, X_SLX=None,
       type=TYPES, **kw):
    """bm.formula / bm.bm, R/60_interface.R:38-78.  `x` is a formula string evaluated on `data`,
    a `(y, X)` pair of arrays, or a previous `BM` object (continue its chains)."""
    if isinstance(x, BM):                                             # bm.bm: n_burn defaults to 0
        n_burn = 0 if n_burn is None else n_burn
        more = x.model.sample(n_save=n_save, n_burn=n_burn, mh=mh, verbose=verbose)
        return BM(np.concatenate([x.chains, more], axis=1), x.model, x.call)
    n_burn = 500 if n_burn is None else n_burn                        # bm.formula default
    type = match_arg(type, TYPES)
    if type == "sv":
        set_SV()
    call = f"bm({x if isinstance(x, str) else '(y, X)'}, type = '{type}', n_save = {n_save}, n_burn = {n_burn})"
    if isinstance(x, str):
        if data is None:
            raise ValueError("a formula needs 'data'")
        y, X, _ = model_matrix(x, data)
    else:
        y, X = x
        X = np.asarray(X, dtype=np.float64).reshape(len(np.asarray(y).ravel()), -1)
    if X.shape[1] > 1 and np.all(X[:, 0] == X[:, 1]):
        X = X[:, 1:]                                                  # drop double intercept, R/60_interface.R:56
    if type in ("slx", "sdm", "sdem") and X_SLX is None:
        X_SLX = X[:, 1:]                                              # all regressors except the intercept, :57
    mdl = get_model(type, y, X, options=options, Psi=W, X_SLX=X_SLX, **kw)
    chains = mdl.sample(n_save=n_save, n_burn=n_burn, mh=mh, verbose=verbose)
    return BM(chains, mdl, call)


def blm(*a, **k):
    return bm(*a, **k, type="lm")


def bslx(*a, **k):
    return bm(*a, **k, type="slx")


def bsar(*a, **k):
    return bm(*a, **k, type="sar")


def bsem(*a, **k):
    return bm(*a, **k, type="sem")


def bsdm(*a, **k):
    return bm(*a, **k, type="sdm")


def bsdem(*a, **k):
    return bm(*a, **k, type="sdem")


def bsv(*a, **k):
    return bm(*a, **k, type="sv")
