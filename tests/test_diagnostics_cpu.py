"""Diagnostics helpers on synthetic draws with known properties (no GPU)."""
import numpy as np

from bsreg_b200.diagnostics import ess, geweke_z, split_rhat, summarize


def ar1(rng, C, n, P, rho):
    x = np.zeros((C, n, P))
    e = rng.standard_normal((C, n, P)) * np.sqrt(1 - rho ** 2)
    x[:, 0] = rng.standard_normal((C, P))
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    return x


def test_iid_draws_have_rhat_one_and_full_ess():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 1000, 3))
    assert np.all(np.abs(split_rhat(x) - 1) < 0.01)
    assert np.all(np.abs(ess(x) / 8000 - 1) < 0.15)
    assert np.all(np.abs(geweke_z(x[0])) < 4)
    s = summarize(x)
    assert np.all(np.abs(s["mean"]) < 5 * s["mcse"]) and np.all(np.abs(s["sd"] - 1) < 0.05)


def test_autocorrelated_draws_lose_ess_as_theory_says():
    rng = np.random.default_rng(1)
    rho = 0.8
    x = ar1(rng, 8, 4000, 2, rho)
    want = 8 * 4000 * (1 - rho) / (1 + rho)
    assert np.all(np.abs(ess(x) / want - 1) < 0.25)


def test_rhat_flags_chains_that_disagree():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 500, 1))
    x[0] += 3.0
    assert split_rhat(x)[0] > 1.5
    drift = rng.standard_normal((1, 2000, 1)) + np.linspace(0, 3, 2000)[None, :, None]
    assert abs(geweke_z(drift[0])[0]) > 5
