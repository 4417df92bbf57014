"""Parity at the size of the named synthetic configurations (BASELINE.json configs[3], configs[4]) -- what the small cases do
not exercise: the tile-range partition over the sweep CTAs with its remainder, the fixed-point head-room of the Gram
accumulators at N = 1e6, the breadth-first relabelling at scale, 22 chain CTAs x 3 chains beside 126 sweep CTAs.

cfg4: SAR, N = 1e5, kNN(10), K = 20, 64 chains, Chebyshev log-determinant (order 50, 64 probes), both BSREG_REORDER modes.
cfg5: SEM, N = 1e6, kNN(10), K = 20, one GPU's share of the chains.
Oracle cost: about 5 ms per SAR iteration at N = 1e5, about 3 s per SEM iteration at N = 1e6 (two thin QRs).
parity unpinned: the oracle is a restatement of the R code (no reference golden vectors exist, R cannot run here)."""
import numpy as np
import pytest

from oracle import core
from bsreg_b200.synthetic import make_problem
from tests.helpers import oracle_chain, oracle_draws

pytestmark = pytest.mark.gpu

CHEB_ORDER, CHEB_PROBES, SEED = 50, 64, 2024


def relnorm(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def capi_mod():
    from bsreg_b200 import capi
    return capi


@pytest.fixture(scope="module")
def cfg4():
    p = make_problem("sar", n=100_000, k_reg=20, knn=10)
    Wy = core.spmm(p.W, p.y)
    return dict(p=p, Wy=Wy, G=core.gram(p.y, p.X, Wy, None), refs={})


@pytest.mark.parametrize("reorder", ["", "0", "1"])
def test_cfg4_deterministic_pieces_and_trajectories(cfg4, reorder, monkeypatch):
    capi = capi_mod()
    p, Wy, G_ref = cfg4["p"], cfg4["Wy"], cfg4["G"]
    n, m = p.X.shape
    monkeypatch.setenv("BSREG_REORDER", reorder)
    n_burn, n_save, chains = 20, 10, (0, 31, 63)
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=CHEB_ORDER,
                     cheb_probes=CHEB_PROBES, n_chains=64, seed=SEED) as h:
        # K1 + K2: Gram matrix of the fused sweep (R/10_lm.R:36, R/15_sar.R:57,106-108,133-135)
        G = h.eval_gram()
        scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
        assert np.max(np.abs(G - G_ref) / scale) < 1e-12
        assert relnorm(G, G_ref) < 1e-10
        assert np.array_equal(G, h.eval_gram())
        # K4 + K3: trace moments and log-determinant on identical moments (1e-8)
        if "mom" not in cfg4:
            cfg4["mom"] = core.trace_moments(p.W, CHEB_ORDER, CHEB_PROBES, SEED)
        mom = h.eval_trace_moments(CHEB_ORDER, CHEB_PROBES, SEED)
        assert np.max(np.abs(mom - cfg4["mom"])) / n < 1e-12
        nodes = h.nodes()
        x, _, w = core.chebyshev_nodes(cfg4["mom"], n)
        assert np.allclose(nodes[0], x, atol=1e-14) and np.allclose(nodes[2], w, rtol=0, atol=1e-9 * n)
        v = np.array([-0.9, -0.3, 0.2, 0.5, 0.85, 0.97])
        want = np.array([core.ldet_nodes(t, *nodes) for t in v])
        assert np.max(np.abs(h.eval_logdet(v) - want) / np.maximum(np.abs(want), 1.0)) < 1e-8
        # K5: conditional pieces at fixed parameters (R/10_lm.R:168-169,176; R/15_sar.R:104-116; R/20_mh.R:192-199)
        rng = np.random.default_rng(3)
        ne = 4
        beta = p.beta + 0.01 * rng.standard_normal((ne, m))
        s2, lam, vv = rng.uniform(0.8, 1.3, ne), rng.uniform(0.3, 0.6, ne), rng.uniform(0.3, 0.6, ne)
        p0 = rng.uniform(1e-8, 2.0, (ne, m))
        got = h.eval_conditionals(beta, s2, lam, p0, vv)
        XX = p.X.T @ p.X
        for e in range(ne):
            ys = p.y - lam[e] * Wy
            _, mu1 = core.beta_moments(XX, p.X.T @ ys, s2[e], p0[e], np.zeros(m))
            assert relnorm(got["mu1"][e], mu1) < 1e-10
            rate1 = 0.01 + 0.5 * core.sq_sum(ys - p.X @ beta[e])
            assert abs(got["rate1"][e] - rate1) / rate1 < 1e-10
            c = core.sar_rss_coeffs(p.y, Wy, p.X, XX, s2[e], p0[e], np.zeros(m))
            rss = c[0] - 2 * vv[e] * c[1] + vv[e] ** 2 * c[2]
            assert abs(got["rss"][e] - rss) / rss < 1e-10
            lp = core.logpost_lambda(vv[e], core.ldet_nodes(vv[e], *nodes), rss, n, m, 1.01, 1.01)
            assert abs(got["logpost"][e] - lp) / abs(lp) < 1e-10
        # the loop itself: 64 chains through the persistent kernel (22 chain CTAs x 3 chains, 126 sweep CTAs)
        draws = h.run(n_burn, n_save)
        assert h.get_state()["iterations"].tolist() == [n_burn + n_save] * 64
    assert np.all(np.isfinite(draws))
    for c in chains:
        if c not in cfg4["refs"]:
            cfg4["refs"][c] = oracle_draws(oracle_chain("sar", p, seed=SEED, chain=c, nodes=nodes), n_save, n_burn)
        assert relnorm(draws[c].T, cfg4["refs"][c]) < 1e-7, (reorder, c)


def test_cfg5_share_sem_n1e6():
    """One GPU's share of cfg5: SEM, N = 1e6.  Gram matrix (d = 42: [y | Wy | X | WX]), conditional pieces, and 8
    iterations of two chains against the oracle (whose OLS rss at the proposal is a Householder QR of 1e6 x 20,
    R/15_sem.R:116-120)."""
    capi = capi_mod()
    p = make_problem("sem", n=1_000_000, k_reg=20, knn=10)
    n, m = p.X.shape
    Wy, WX = core.spmm(p.W, p.y), core.spmm(p.W, p.X)
    Z = np.hstack([p.y[:, None], Wy[:, None], p.X, WX])
    G_ref = Z.T @ Z                                     # float64 BLAS (the long-double form takes minutes at this size)
    n_burn, n_save, n_chains = 5, 3, 4
    with capi.Handle(p.y, p.X, model=capi.MODEL_SEM, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=20, cheb_probes=8,
                     n_chains=n_chains, chain_offset=60, seed=SEED) as h:
        G = h.eval_gram()
        scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
        assert np.max(np.abs(G - G_ref) / scale) < 1e-11   # the float64 reference itself carries ~1e-13 of rounding
        assert np.array_equal(G, G.T) and np.array_equal(G, h.eval_gram())
        nodes = h.nodes()
        rng = np.random.default_rng(5)
        ne = 2
        beta = p.beta + 0.003 * rng.standard_normal((ne, m))
        s2, lam, vv = rng.uniform(0.9, 1.1, ne), rng.uniform(0.4, 0.6, ne), rng.uniform(0.4, 0.6, ne)
        p0 = rng.uniform(1e-8, 2.0, (ne, m))
        got = h.eval_conditionals(beta, s2, lam, p0, vv)
        for e in range(ne):
            ys, Xs = p.y - lam[e] * Wy, p.X - lam[e] * WX
            _, mu1 = core.beta_moments(Xs.T @ Xs, Xs.T @ ys, s2[e], p0[e], np.zeros(m))
            assert relnorm(got["mu1"][e], mu1) < 1e-10
            rate1 = 0.01 + 0.5 * core.sq_sum(ys - Xs @ beta[e])
            assert abs(got["rate1"][e] - rate1) / rate1 < 1e-10
            rss = core.sem_rss(vv[e], p.y, Wy, p.X, WX)
            assert abs(got["rss"][e] - rss) / rss < 1e-10
            lp = core.logpost_lambda(vv[e], core.ldet_nodes(vv[e], *nodes), rss, n, m, 1.01, 1.01)
            assert abs(got["logpost"][e] - lp) / abs(lp) < 1e-10
        draws = h.run(n_burn, n_save)
    for c in (0, 3):
        ref = oracle_draws(oracle_chain("sem", p, seed=SEED, chain=60 + c, nodes=nodes), n_save, n_burn)
        assert relnorm(draws[c].T, ref) < 1e-7, c
