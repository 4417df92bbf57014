"""CUDA path through the PUBLIC interface (bsreg_b200.bm / blm / bsar / bsdem) on the real-data
configs 1-3 of BASELINE.json, against the committed golden vectors and against the oracle."""
import numpy as np
import pytest

import bsreg_b200 as bs
from oracle import core
from oracle.sampler import PhiloxRNG, RefChain, sample
from tests.golden.make_golden import configs, design, load_cigarettes, panel_w

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cig():
    return load_cigarettes()


def _fit(cfg, seed, n_save, n_burn, **kw):
    name, model, prior, pri, y, X, W, X_slx, reps = cfg
    opts = bs.set_options(prior, SNG=bs.set_SNG(**pri["SNG"]) if "SNG" in pri else None)
    ld = dict(reps=reps)
    return bs.bm((y, X), type=model, W=W, options=opts, n_save=n_save, n_burn=n_burn, verbose=False, seed=seed,
                 ldet_SAR=ld, ldet_SEM=ld, **kw)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_interface_reproduces_golden(cig, idx):
    cfg = list(configs(cig))[idx]
    g = np.load(f"tests/golden/golden_{cfg[0]}.npz")
    fit = _fit(cfg, int(g["seed"]), int(g["n_save"]), int(g["n_burn"]))
    assert fit.colnames == list(g["names"])
    assert fit.draws.shape == g["draws"].shape
    assert np.max(np.abs(fit.draws - g["draws"])) < 1e-7 * np.max(np.abs(g["draws"]))
    G = fit.model.handle.eval_gram()
    scale = np.sqrt(np.outer(np.diag(g["gram"]), np.diag(g["gram"])))
    scale[scale == 0] = 1.0                      # LM carries a zero Wy column
    assert np.max(np.abs(G - g["gram"]) / scale) < 1e-12
    if "ldet" in g:
        got = fit.model.handle.eval_logdet(g["ldet_v"])
        assert np.max(np.abs(got - g["ldet"])) < 1e-8 * max(1.0, np.max(np.abs(g["ldet"])))
        st = fit.model.handle.get_state()
        assert np.allclose([st["mh_scale"][0, 0], *st["mh_counters"][0, 0]], g["mh"])
    assert "iterations" in fit.model.get_meta and str(fit).startswith("\nCall:")
    assert set(fit.summary()) == set(fit.colnames)


def test_continue_a_fit_like_bm_bm(cig):
    """bm(x, n_save) on a `bm` object continues the chain and row-binds (R/60_interface.R:71-78)."""
    cfg = list(configs(cig))[1]
    a = _fit(cfg, 5, 30, 20)
    b = bs.bm(a, n_save=15, verbose=False)
    assert b.draws.shape == (45, a.draws.shape[1]) and np.array_equal(b.draws[:30], a.draws)
    whole = _fit(cfg, 5, 45, 20)
    assert np.max(np.abs(whole.draws - b.draws)) < 1e-9                 # same trajectory as an uninterrupted run
    assert "65 iterations" in b.model.get_meta


def test_multi_chain_object_and_chain_offset(cig):
    cfg = list(configs(cig))[1]
    fit = _fit(cfg, 9, 20, 20, n_chains=4)
    assert fit.chains.shape == (4, 20, 5) and np.array_equal(fit.draws, fit.chains[0])
    one = _fit(cfg, 9, 20, 20, n_chains=1, chain_offset=2)
    assert np.array_equal(one.draws, fit.chains[2])                      # keyed by GLOBAL chain id
    chains, names = fit.as_mcmc_list()
    assert len(chains) == 4 and names[-1] == "lambda_SAR"


@pytest.mark.parametrize("model", ["sar", "sem"])
def test_paper_design_with_fixed_effects(cig, model):
    """paper/1a_estimate_cont.R: M = 78 (state + year dummies).  Checks the on-chip K x K path at
    large M and the Cholesky-based SEM rss against the oracle's Householder QR (cond^2 caveat)."""
    y, X = design(cig, fixed_effects=True)
    W = panel_w(cig)
    assert X.shape == (1380, 3 + 45 + 29)
    ch = RefChain(model, y, X, W=W, ldet=dict(kind="eigen", reps=30), rng=PhiloxRNG(3, 0))
    ref = sample(ch, 10, 10)
    fit = bs.bm((y, X), type=model, W=W, n_save=10, n_burn=10, verbose=False, seed=3,
                ldet_SAR=dict(reps=30), ldet_SEM=dict(reps=30))
    assert np.max(np.abs(fit.draws - ref)) < 1e-6 * np.max(np.abs(ref))
    h = fit.model.handle
    Wy, WX = core.spmm(ch.W, y), core.spmm(ch.W, X)
    v = np.array([-0.5, 0.3, 0.9])
    m = X.shape[1]
    got = h.eval_conditionals(np.tile(ch.beta, (3, 1)), np.full(3, ch.sigma2), np.zeros(3), np.full((3, m), 1e-8), v)
    for e in range(3):
        if model == "sar":
            c = core.sar_rss_coeffs(y, Wy, X, X.T @ X, ch.sigma2, np.full(m, 1e-8), np.zeros(m))
            rss = c[0] - 2 * v[e] * c[1] + v[e] ** 2 * c[2]
        else:
            rss = core.sem_rss(v[e], y, Wy, X, WX)
        assert abs(got["rss"][e] - rss) / rss < 1e-9


def test_spatial_effects_match_the_reference_formula(cig):
    """`get_effects`, R/15_sar.R:143-149: total = beta / (1 - lambda), direct = tr((I - lambda W)^-1) / N * beta with the
    reference's dense solve, indirect = total - direct -- here from the spectral nodes, for every saved draw."""
    cfg = list(configs(cig))[1]                       # bsar on the cigarettes panel (eigenvalue nodes, reps = 30)
    name, model, prior, pri, y, X, W, X_slx, reps = cfg
    fit = _fit(cfg, 3, 12, 10)
    eff = fit.effects()
    m = X.shape[1]
    Wd = W.toarray() if hasattr(W, "toarray") else np.asarray(W)
    n = Wd.shape[0]
    for s in (0, 5, 11):
        beta, lam = fit.draws[s, :m], fit.draws[s, -1]
        direct = np.trace(np.linalg.inv(np.eye(n) - lam * Wd)) / n * beta
        total = beta / (1.0 - lam)
        assert np.max(np.abs(eff["direct"][s] - direct)) < 1e-9 * max(1.0, np.max(np.abs(direct)))
        assert np.max(np.abs(eff["total"][s] - total)) < 1e-12 * max(1.0, np.max(np.abs(total)))
        assert np.allclose(eff["indirect"][s], total - direct, rtol=0, atol=1e-9)
    cur = fit.model.get_effects                      # the active binding: current state of chain 1 = the last draw
    assert np.allclose(cur["direct"][0], eff["direct"][-1]) and np.allclose(cur["total"][0], eff["total"][-1])
    with pytest.raises(AttributeError):
        _fit(list(configs(cig))[0], 1, 5, 5).effects()          # blm: no spatial effects in the reference either


def test_spatial_effects_from_chebyshev_nodes():
    """Large-N path: the same quadrature nodes that give log|I - vW| give tr((I - vW)^-1); against a dense inverse."""
    from bsreg_b200 import capi
    from bsreg_b200.synthetic import make_problem
    p = make_problem("sar", n=1500, k_reg=4, knn=6)
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=40, cheb_probes=256,
                     n_chains=1, seed=3) as h:
        mdl = bs.interface.Model(h, "sar", "Independent", [], 1)
        lam = np.array([-0.4, 0.2, 0.6])
        got = mdl.trace_inverse(lam) / h.n
    Wd = p.W.toarray()
    want = np.array([np.trace(np.linalg.inv(np.eye(1500) - v * Wd)) / 1500 for v in lam])
    assert np.max(np.abs(got - want)) < 5e-3         # Hutchinson error of 256 probes at N = 1500 (exact moments 0-2)


def test_interface_grid_true_runs_the_spline_logdet(cig):
    """`bsar(..., ldet_SAR = list(grid = TRUE, i_lambda = c(-1, 1 - 1e-12, 100), reps = 30))` on the cigarettes panel:
    the reference's spline-over-a-determinant-grid option (R/15_sar.R:79-87).  Posterior of lambda close to the
    eigenvalue-method fit; effects available through the spline's derivative."""
    name, model, prior, pri, y, X, W, X_slx, reps = list(configs(cig))[1]
    assert model == "sar"
    kw = dict(type=model, W=W, options=bs.set_options(prior), n_save=300, n_burn=200, verbose=False, seed=5)
    fit_e = bs.bm((y, X), ldet_SAR=dict(reps=reps), **kw)
    fit_g = bs.bm((y, X), ldet_SAR=dict(grid=True, i_lambda=(-1, 1 - 1e-12, 100), reps=reps), **kw)
    assert fit_g.model.handle.spline()[0].shape == (100,)
    assert abs(fit_g.draws[:, -1].mean() - fit_e.draws[:, -1].mean()) < 0.05
    eg, ee = fit_g.effects(), fit_e.effects()
    assert np.allclose(eg["total"].mean(0), ee["total"].mean(0), rtol=0.1, atol=0.02)
    assert np.allclose(eg["direct"].mean(0), ee["direct"].mean(0), rtol=0.1, atol=0.02)


# ---------------------------------------------------------------------------
# the paper's own cross-checks (paper/1a_estimate_cont.R:14-36: blm vs lm, bsar vs lagsarlm, bsem vs errorsarlm),
# with the frequentist side computed here independently of the oracle: OLS, and concentrated-likelihood ML for SAR / SEM
# ---------------------------------------------------------------------------
def _ml_spatial(model, y, X, W, reps):
    """Concentrated log-likelihood maximised over the spatial parameter (Ord 1975; what spatialreg::lagsarlm /
    errorsarlm do): l(v) = log|I - v W| - N/2 log(rss(v) / N)."""
    from scipy.optimize import minimize_scalar
    n = len(y)
    size = n // reps
    ev = np.linalg.eigvals(W[:size, :size].toarray())
    ldet = lambda v: reps * float(np.sum(np.log(1 - v * ev)).real)      # noqa: E731
    Wy, WX = W @ y, W @ X

    def ols(yy, XX):
        b = np.linalg.lstsq(XX, yy, rcond=None)[0]
        e = yy - XX @ b
        return b, float(e @ e)

    def negll(v):
        rss = ols(y - v * Wy, X)[1] if model == "sar" else ols(y - v * Wy, X - v * WX)[1]
        return -(ldet(v) - 0.5 * n * np.log(rss / n))
    v = minimize_scalar(negll, bounds=(-0.99, 0.99), method="bounded", options=dict(xatol=1e-10)).x
    b, rss = ols(y - v * Wy, X) if model == "sar" else ols(y - v * Wy, X - v * WX)
    return v, b, np.sqrt(rss / n)


def test_posterior_means_agree_with_ols_and_ml(cig):
    """parity unpinned (no R here): what CAN be checked against something that is neither the oracle nor the CUDA code.
    Flat-ish default priors, N = 1380: posterior means sit on the least-squares / maximum-likelihood estimates, well
    within one posterior standard deviation (plus Monte-Carlo error)."""
    from bsreg_b200.diagnostics import summarize
    cfgs = list(configs(cig))
    # cfg1: blm vs lm
    name, model, prior, pri, y, X, W, X_slx, reps = cfgs[0]
    fit = bs.bm((y, X), type="lm", n_save=2000, n_burn=500, n_chains=8, seed=11, verbose=False)
    s = summarize(fit.chains)
    b = np.linalg.lstsq(X, y, rcond=None)[0]
    e = y - X @ b
    ols = np.concatenate([b, [np.sqrt(e @ e / (len(y) - X.shape[1]))]])
    assert np.all(np.abs(s["mean"] - ols) < 0.1 * s["sd"] + 5 * s["mcse"]), (s["mean"], ols)
    # cfg2-like: bsar vs lagsarlm, bsem vs errorsarlm (same data)
    name, model, prior, pri, y, X, W, X_slx, reps = cfgs[1]
    for mdl in ("sar", "sem"):
        v_ml, b_ml, s_ml = _ml_spatial(mdl, y, X, W.tocsr(), reps)
        fit = bs.bm((y, X), type=mdl, W=W, n_save=3000, n_burn=1000, n_chains=8, seed=12, verbose=False,
                    ldet_SAR=dict(reps=reps), ldet_SEM=dict(reps=reps))
        s = summarize(fit.chains)
        ml = np.concatenate([b_ml, [s_ml], [v_ml]])
        assert np.all(s["rhat"] < 1.05)
        assert np.all(np.abs(s["mean"] - ml) < 0.5 * s["sd"] + 5 * s["mcse"]), (mdl, s["mean"], ml, s["sd"])


def test_dense_literal_sem_oracle_on_golden_config(cig):
    """The reference rebuilds the dense N x N filter `diag(N) - lambda W` every iteration (R/15_sem.R:28-36); the oracle
    normally uses the cached lags instead.  On BASELINE config 3 (bsdem + Shrinkage, N = 1380) the CUDA chain must also
    follow the LITERAL dense restatement."""
    cfg = list(configs(cig))[2]
    name, model, prior, pri, y, X, W, X_slx, reps = cfg
    n_burn, n_save, seed = 20, 15, 5
    ch = RefChain(model, y, X, W=W, X_slx=X_slx, prior_type=prior, priors=pri, ldet=dict(kind="eigen", reps=reps),
                  rng=PhiloxRNG(seed, 0), dense_sem=True)
    ref = sample(ch, n_save, n_burn)
    fit = _fit(cfg, seed, n_save, n_burn)
    assert np.max(np.abs(fit.draws - ref)) < 1e-7 * np.max(np.abs(ref))
