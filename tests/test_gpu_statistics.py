"""Second and third correctness legs of the north star: posterior means / standard deviations of the CUDA
chains agree with the CPU oracle's own chains (independent NumPy RNG stream) within Monte-Carlo standard
error, and the convergence diagnostics (split R-hat, ESS per draw, Geweke z) are comparable."""
import numpy as np
import pytest

import bsreg_b200 as bs
from bsreg_b200.diagnostics import geweke_z, summarize
from bsreg_b200.synthetic import make_problem
from oracle.sampler import NumpyRNG, RefChain, sample

pytestmark = pytest.mark.gpu

CASES = [("sar", "Independent", 5), ("sem", "Independent", 5), ("sar", "Horseshoe", 6), ("lm", "Shrinkage", 6),
         ("sar", "Independent", 14),          # runs the persistent kernel (d = 16)
         ("lm", "Conjugate", 14),             # the general chain step inside the persistent kernel
         ("sar", "Shrinkage", 13)]            # four-warp general chain kernel, two factorisations per iteration


@pytest.mark.parametrize("model,prior,k", CASES)
def test_posterior_moments_and_diagnostics_match_the_oracle(model, prior, k):
    p = make_problem(model, n=1500, k_reg=k, knn=6, symmetric=True)
    n_burn, n_save = 400, 800
    W = p.W if model != "lm" else None
    opts = bs.set_options(prior, SNG=bs.set_SNG(theta_scale=0.3) if prior == "Shrinkage" else None)
    fit = bs.bm((p.y, p.X), type=model, W=W, options=opts, n_save=n_save, n_burn=n_burn, verbose=False,
                n_chains=16, seed=2026)
    gpu = fit.chains                                                   # (16, n_save, P)
    pri = {"SNG": dict(theta_scale=0.3)} if prior == "Shrinkage" else {}
    cpu = np.stack([sample(RefChain(model, p.y, p.X, W=W, prior_type=prior, priors=pri, rng=NumpyRNG(100 + c)),
                           n_save, n_burn) for c in range(4)])         # (4, n_save, P), independent stream
    g, c = summarize(gpu), summarize(cpu)
    se = np.sqrt(g["mcse"] ** 2 + c["mcse"] ** 2)
    assert np.all(np.abs(g["mean"] - c["mean"]) < 4.5 * se + 1e-12), (g["mean"], c["mean"], se)
    assert np.all(np.abs(g["sd"] / c["sd"] - 1) < 0.2), (g["sd"], c["sd"])
    # diagnostics comparable: both converged, similar mixing
    assert np.all(g["rhat"] < 1.05) and np.all(c["rhat"] < 1.1)
    eff_g, eff_c = g["ess"] / gpu[:, :, 0].size, c["ess"] / cpu[:, :, 0].size
    assert np.all(eff_g / eff_c < 2.5) and np.all(eff_c / eff_g < 2.5), (eff_g, eff_c)
    z = np.array([geweke_z(ch) for ch in gpu])
    assert np.mean(np.abs(z) < 3) > 0.95
    # truth is inside the posterior mass
    truth = np.concatenate([p.beta, [p.sigma], [p.lam] if model != "lm" else []])
    assert np.all(np.abs(g["mean"] - truth) < 5 * g["sd"] + 0.02)
