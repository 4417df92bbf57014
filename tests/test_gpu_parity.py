"""CUDA path vs oracle on the same seeded inputs, through the C ABI.  Tolerances:
1e-10 relative for the deterministic kernels, 1e-8 for log-determinants (north_star)."""
import numpy as np
import pytest

from oracle import core
from oracle.philox import Slot, Stream
from bsreg_b200.synthetic import make_problem
from tests.helpers import cuda_handle, oracle_chain, oracle_draws, small_problem

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relnorm(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("model", ["sar", "sem"])
def test_spmm_and_gram(model):
    p = small_problem(model, n=1037, k=7)
    with cuda_handle(model, p) as h:
        B = np.random.default_rng(0).standard_normal((p.X.shape[0], 5))
        assert relnorm(h.eval_spmm(B), core.spmm(p.W, B)) < 1e-13
        Wy = core.spmm(p.W, p.y)
        WX = core.spmm(p.W, p.X) if model == "sem" else None
        G_ref = core.gram(p.y, p.X, Wy, WX)
        G = h.eval_gram()
        assert G.shape == G_ref.shape
        scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
        assert np.max(np.abs(G - G_ref) / scale) < 1e-13
        assert np.array_equal(G, G.T)
        assert np.array_equal(G, h.eval_gram())          # bit-reproducible


def test_gram_lm():
    p = small_problem("lm", n=777, k=3)
    with cuda_handle("lm", p) as h:
        G = h.eval_gram()
        G_ref = core.gram(p.y, p.X, np.zeros_like(p.y), None)
        assert np.max(np.abs(G - G_ref)) / np.max(np.abs(G_ref)) < 1e-13


@pytest.mark.parametrize("model", ["lm", "sar", "sem"])
def test_residual_ss_direct(model):
    p = small_problem(model, n=901, k=6)
    rng = np.random.default_rng(1)
    with cuda_handle(model, p) as h:
        ne = 7
        beta = p.beta + 0.1 * rng.standard_normal((ne, len(p.beta)))
        lam = rng.uniform(-0.5, 0.9, ne) if model != "lm" else np.zeros(ne)
        rho = lam if model == "sem" else np.zeros(ne)
        got = h.eval_residual_ss(beta, lam, rho)
        Wy = core.spmm(p.W, p.y) if model != "lm" else np.zeros_like(p.y)
        WX = core.spmm(p.W, p.X) if model == "sem" else np.zeros_like(p.X)
        want = [core.residual_ss(p.y, p.X, Wy, WX, beta[e], lam[e], rho[e]) for e in range(ne)]
        assert rel(got, want) < 1e-10


def test_logdet_eigen_real_and_complex():
    for symmetric in (True, False):
        p = small_problem("sar", n=300, k=3, symmetric=symmetric)
        nodes = core.eigen_nodes(p.W, 1)
        with cuda_handle("sar", p, nodes=nodes) as h:
            v = np.linspace(-0.95, 0.98, 41)
            want = np.array([core.ldet_nodes(x, *nodes) for x in v])
            got = h.eval_logdet(v)
            assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < 1e-8
            dense = np.array([core.ldet_dense(x, p.W) for x in v[::8]])
            assert np.max(np.abs(got[::8] - dense) / np.maximum(np.abs(dense), 1.0)) < 1e-8


@pytest.mark.parametrize("model", ["sar", "sem"])
def test_conditionals(model):
    p = small_problem(model, n=650, k=5)
    rng = np.random.default_rng(3)
    nodes = core.eigen_nodes(p.W, 1)
    with cuda_handle(model, p, nodes=nodes) as h:
        ne, m = 6, p.X.shape[1]
        beta = p.beta + 0.05 * rng.standard_normal((ne, m))
        s2 = rng.uniform(0.5, 2.0, ne)
        lam = rng.uniform(-0.3, 0.8, ne)
        v = rng.uniform(-0.3, 0.8, ne)
        p0 = rng.uniform(1e-8, 2.0, (ne, m))
        got = h.eval_conditionals(beta, s2, lam, p0, v)
        Wy, WX = core.spmm(p.W, p.y), core.spmm(p.W, p.X)
        for e in range(ne):
            if model == "sar":
                ys, Xs = p.y - lam[e] * Wy, p.X
            else:
                ys, Xs = p.y - lam[e] * Wy, p.X - lam[e] * WX
            _, mu1 = core.beta_moments(Xs.T @ Xs, Xs.T @ ys, s2[e], p0[e], np.zeros(m))
            assert relnorm(got["mu1"][e], mu1) < 1e-10
            rate1 = 0.01 + 0.5 * core.sq_sum(ys - Xs @ beta[e])
            assert abs(got["rate1"][e] - rate1) / rate1 < 1e-10
            if model == "sar":
                c = core.sar_rss_coeffs(p.y, Wy, p.X, p.X.T @ p.X, s2[e], p0[e], np.zeros(m))
                rss = c[0] - 2 * v[e] * c[1] + v[e] ** 2 * c[2]
            else:
                rss = core.sem_rss(v[e], p.y, Wy, p.X, WX)
            assert abs(got["rss"][e] - rss) / rss < 1e-10
            lp = core.logpost_lambda(v[e], core.ldet_nodes(v[e], *nodes), rss, len(p.y), m, 1.01, 1.01)
            assert abs(got["logpost"][e] - lp) / abs(lp) < 1e-10


def test_device_rng_matches_contract(capi):
    s = Stream(seed=0x1234567890ABCDEF, chain=5)
    n = 64
    u = capi.eval_rng(0, 0x1234567890ABCDEF, 5, Slot.LAMBDA_ACC, 1, n)
    assert np.array_equal(u, [s.uniform(1 + i, Slot.LAMBDA_ACC) for i in range(n)])
    z = capi.eval_rng(1, 0x1234567890ABCDEF, 5, Slot.BETA, 1, n)
    assert rel(z, [s.normal(1 + i, Slot.BETA) for i in range(n)]) < 1e-12
    for shape, rate in [(50000.01, 3.0), (2.5, 0.7), (0.3, 2.0), (1.0, 4.0)]:
        g = capi.eval_rng(2, 0x1234567890ABCDEF, 5, Slot.SIGMA, 1, n, (shape, rate))
        assert rel(g, [s.gamma(1 + i, Slot.SIGMA, shape, rate) for i in range(n)]) < 1e-11
    for lam, chi, psi in [(-0.4, 0.3, 0.1), (-0.4, 1e-4, 0.05), (0.5, 2.0, 3.0), (3.0, 1.0, 1.0), (-0.4, 30.0, 2.0)]:
        g = capi.eval_rng(4, 0x1234567890ABCDEF, 5, Slot.SNG_TAU, 1, n, (lam, chi, psi))
        assert rel(g, [s.gig(1 + i, Slot.SNG_TAU, lam, chi, psi) for i in range(n)]) < 1e-10
    t = capi.eval_rng(5, 0x1234567890ABCDEF, 5, Slot.LAMBDA_PROP, 1, n, (0.95, 0.2, -1.0, 1 - 1e-12))
    assert rel(t, [s.truncated_rw(1 + i, Slot.LAMBDA_PROP, 0.95, 0.2, -1.0, 1 - 1e-12) for i in range(n)]) < 1e-12


CASES = [("lm", "Independent"), ("sar", "Independent"), ("sem", "Independent"),
         ("lm", "Conjugate"), ("sar", "Conjugate"), ("sem", "Conjugate"),
         ("lm", "Horseshoe"), ("sar", "Horseshoe"), ("sem", "Horseshoe"),
         ("lm", "Shrinkage"), ("sar", "Shrinkage"), ("sem", "Shrinkage")]


@pytest.mark.parametrize("model,prior", CASES)
def test_chain_trajectory_matches_oracle(model, prior):
    """Same Philox stream => the CUDA chain must reproduce the oracle chain draw by draw
    (burn-in with MH tuning included)."""
    p = small_problem(model, n=500, k=4)
    priors = {"SNG": dict(theta_scale=0.3)} if prior == "Shrinkage" else {}
    n_burn, n_save, chain_id = 60, 80, 3
    ref = oracle_draws(oracle_chain(model, p, prior, priors, seed=99, chain=chain_id), n_save, n_burn)
    with cuda_handle(model, p, prior, priors, seed=99, n_chains=1, chain_offset=chain_id) as h:
        got = h.run(n_burn, n_save)[0].T
    assert got.shape == ref.shape
    assert relnorm(got, ref) < 1e-7


@pytest.mark.parametrize("model,prior,k", [("sar", "Horseshoe", 14), ("sar", "Shrinkage", 20), ("lm", "Conjugate", 16),
                                           ("sem", "Horseshoe", 13), ("sar", "Conjugate", 26), ("sem", "Shrinkage", 22)])
def test_general_chain_step_wide_designs(model, prior, k):
    """From M = 12 on the general chain step runs with four warps (one bordered row of the factorisation per thread, block-wide
    reductions): same trajectories as the oracle; M = 26 still one row per thread, three right-hand sides."""
    p = small_problem(model, n=600, k=k)
    priors = {"SNG": dict(theta_scale=0.3)} if prior == "Shrinkage" else {}
    n_burn, n_save, chain_id = 30, 50, 1
    ref = oracle_draws(oracle_chain(model, p, prior, priors, seed=7, chain=chain_id), n_save, n_burn)
    with cuda_handle(model, p, prior, priors, seed=7, n_chains=2, chain_offset=chain_id - 1) as h:
        got = h.run(n_burn, n_save)[1].T
    assert got.shape == ref.shape
    assert relnorm(got, ref) < 1e-7


@pytest.mark.parametrize("model,prior,k", [("sar", "Horseshoe", 14), ("lm", "Shrinkage", 12), ("sem", "Conjugate", 5),
                                           ("sem", "Independent", 20), ("sar", "Independent", 5), ("lm", "Independent", 28)])
def test_looped_chain_kernel_equals_one_launch_per_step(model, prior, k, monkeypatch):
    """With cached statistics the chain kernels (general and fast) run a whole block of iterations in one launch (csrc/chain.cu, nrep):
    the draws must be bit-identical to one launch per step (BSREG_NO_CHAIN_LOOP), to the streamed run (the sweeps
    recompute the same Gram matrix), and must not depend on how the run is cut into blocks."""
    from bsreg_b200 import capi
    p = small_problem(model, n=500, k=k)
    priors = {"SNG": dict(theta_scale=0.3)} if prior == "Shrinkage" else {}
    kw = dict(seed=13, n_chains=3)
    with cuda_handle(model, p, prior, priors, stats_mode=capi.STATS_CACHED, **kw) as h:
        looped = h.run(25, 35)
    with cuda_handle(model, p, prior, priors, stats_mode=capi.STATS_CACHED, **kw) as h:
        a = h.run(25, 10)
        b = h.run(0, 1)
        c = h.run(0, 24)
    assert np.array_equal(np.concatenate([a, b, c], axis=2), looped)
    with cuda_handle(model, p, prior, priors, stats_mode=capi.STATS_STREAM, **kw) as h:
        # the streamed run may sweep with another kernel (persistent sweep CTAs): the same Gram matrix to rounding
        streamed = h.run(25, 35)
        assert max(relnorm(streamed[c].T, looped[c].T) for c in range(3)) < 1e-9
    monkeypatch.setenv("BSREG_NO_CHAIN_LOOP", "1")
    with cuda_handle(model, p, prior, priors, stats_mode=capi.STATS_CACHED, **kw) as h:
        assert np.array_equal(h.run(25, 35), looped)


@pytest.mark.parametrize("model,prior,k,n", [("sar", "Horseshoe", 20, 3000), ("sar", "Shrinkage", 14, 700), ("lm", "Conjugate", 16, 5000),
                                             ("sar", "Conjugate", 11, 900), ("lm", "Horseshoe", 22, 2500)])
def test_priors_in_the_persistent_kernel(model, prior, k, n, monkeypatch):
    """Streamed statistics, 13 <= d <= 24, a prior other than Independent: the run is ONE cooperative launch per block
    (csrc/persist.cu: gibbs_persistent_general_kernel, one chain per chain CTA running the general chain step beside the sweep
    CTAs).  The draws do not depend on the cut into blocks (bit for bit), agree with the CUDA-graph path (one sweep + one chain-step
    launch per iteration) and with the cached-statistics run to rounding, and follow the oracle."""
    from bsreg_b200 import capi
    p = small_problem(model, n=n, k=k)
    priors = {"SNG": dict(theta_scale=0.3)} if prior == "Shrinkage" else {}
    kw = dict(seed=17, n_chains=5)
    if model != "lm":
        monkeypatch.setenv("BSREG_PERSISTENT_GENERAL", "1")      # default: the models without a spatial lag only (slower for SAR)
    with cuda_handle(model, p, prior, priors, **kw) as h:
        l0 = h.kernel_launches()
        got = h.run(30, 40)
        assert h.kernel_launches() - l0 <= 4, "not the persistent path"
    with cuda_handle(model, p, prior, priors, **kw) as h:
        parts = [h.run(30, 7), h.run(0, 3), h.run(0, 30)]
    assert np.array_equal(np.concatenate(parts, axis=2), got)
    # the other paths sweep with the one-launch kernel on all SMs: the same Gram matrix to rounding (the fixed-point sums are
    # exact across CTAs, the partial sums inside a CTA depend on its tiles)
    with cuda_handle(model, p, prior, priors, stats_mode=capi.STATS_CACHED, **kw) as h:
        cached = h.run(30, 40)
    monkeypatch.setenv("BSREG_PERSISTENT_GENERAL", "0")
    with cuda_handle(model, p, prior, priors, **kw) as h:
        l0 = h.kernel_launches()
        graph = h.run(30, 40)
        assert h.kernel_launches() - l0 >= 2 * 70
    assert np.array_equal(graph, cached)
    assert max(relnorm(graph[c].T, got[c].T) for c in range(5)) < 1e-9
    ref = oracle_draws(oracle_chain(model, p, prior, priors, seed=17, chain=3), 40, 30)
    assert relnorm(got[3].T, ref) < 1e-7


@pytest.mark.parametrize("model,n,k", [("lm", 50, 1), ("sar", 50, 1), ("sem", 65, 2), ("sar", 129, 3), ("sem", 31, 1)])
def test_tiny_problems(model, n, k):
    """Edge of the tiling: fewer rows than one 64-row tile, one row beyond a tile boundary, a single regressor.  Gram matrix
    and a short trajectory against the oracle."""
    p = small_problem(model, n=n, k=k, knn=4)
    with cuda_handle(model, p, seed=3, n_chains=2) as h:
        G = h.eval_gram()
        got = h.run(20, 30)[1].T
    Wy = core.spmm(p.W, p.y) if model != "lm" else None
    WX = core.spmm(p.W, p.X) if model == "sem" else None
    G_ref = core.gram(p.y, p.X, Wy, WX)
    if model == "lm":                                   # the device matrix keeps an (empty) Wy slot
        keep = [0] + list(range(2, 2 + k))
        G = G[np.ix_(keep, keep)]
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    assert np.max(np.abs(G - G_ref) / scale) < 1e-13
    ref = oracle_draws(oracle_chain(model, p, seed=3, chain=1), 30, 20)
    assert got.shape == ref.shape and relnorm(got, ref) < 1e-7


@pytest.mark.parametrize("model,own_psi", [("slx", False), ("sdm", False), ("sdm", True), ("sdem", False), ("sdem", True)])
def test_slx_family_with_fixed_connectivity(model, own_psi):
    """SLX / SDM / SDEM (R/40_setup.R:99-120): X <- [X, Psi_SLX %*% X_SLX] at set-up (R/12_slx.R:60-68), Psi_SLX = the model's W by
    default or a matrix of its own (`nnz_slx` of the C ABI).  The CUDA chain follows the oracle draw by draw."""
    from bsreg_b200 import capi
    from bsreg_b200.synthetic import knn_weights
    from oracle.sampler import PhiloxRNG, RefChain
    from tests.helpers import MODEL_CODE
    p = small_problem(model, n=480, k=3)
    Xl = np.random.default_rng(5).standard_normal((len(p.y), 2))
    W_slx = knn_weights(len(p.y), 4, seed=77)[0] if own_psi else None
    n_burn, n_save, chain_id = 40, 60, 2
    ch = RefChain(model, p.y, p.X, W=p.W, X_slx=Xl, W_slx=W_slx, rng=PhiloxRNG(21, chain_id))
    ref = oracle_draws(ch, n_save, n_burn)
    nodes = core.eigen_nodes(p.W, 1) if model != "slx" else None
    with capi.Handle(p.y, p.X, model=MODEL_CODE[model], W=p.W, X_slx=Xl, W_slx=W_slx, nodes=nodes, n_chains=1,
                     chain_offset=chain_id, seed=21) as h:
        got = h.run(n_burn, n_save)[0].T
    assert got.shape == ref.shape == (n_save, 3 + 2 + 1 + (model != "slx"))
    assert relnorm(got, ref) < 1e-7


def test_chain_batching_and_modes_are_invariant():
    """Draws of a chain id do not depend on batching, chain_offset or stats mode."""
    from bsreg_b200 import capi
    p = small_problem("sar", n=450, k=4)
    with cuda_handle("sar", p, seed=5, n_chains=6) as h:
        all6 = h.run(40, 30)
    with cuda_handle("sar", p, seed=5, n_chains=2, chain_offset=3) as h:
        part = h.run(40, 30)
    assert np.array_equal(all6[3:5], part)
    with cuda_handle("sar", p, seed=5, n_chains=6, stats_mode=capi.STATS_CACHED) as h:
        cached = h.run(40, 30)
    assert np.array_equal(all6, cached)


def test_continue_and_state_roundtrip():
    p = small_problem("sem", n=420, k=3)
    with cuda_handle("sem", p, seed=8, n_chains=3) as h:
        full = h.run(20, 40)
    with cuda_handle("sem", p, seed=8, n_chains=3) as h:
        a = h.run(20, 25)
        st = h.get_state()
        b = h.run(0, 15)                      # bm.bm continuation, R/60_interface.R:71-78
        assert st["iterations"].tolist() == [45, 45, 45]
    assert np.array_equal(np.concatenate([a, b], axis=2), full)
    with cuda_handle("sem", p, seed=8, n_chains=3) as h2:
        h2.set_state(st)
        assert np.array_equal(h2.run(0, 15), b)


def test_chebyshev_moments_and_nodes():
    from bsreg_b200 import capi
    p = small_problem("sar", n=800, k=3, symmetric=True)
    order, probes, seed = 30, 32, 77
    with cuda_handle("sar", p, ldet=capi.LDET_CHEBYSHEV, cheb_order=order, cheb_probes=probes, seed=seed) as h:
        mom = h.eval_trace_moments(order, probes, seed)
        want = core.trace_moments(p.W, order, probes, seed)
        assert np.max(np.abs(mom - want)) / len(p.y) < 1e-12
        re, im, w = h.nodes()
        x, _, ww = core.chebyshev_nodes(want, len(p.y))
        assert np.allclose(re, x, atol=1e-14) and np.allclose(w, ww, rtol=0, atol=1e-9 * len(p.y))
        v = np.array([-0.5, 0.1, 0.6, 0.9])
        got = h.eval_logdet(v)
        same_input = np.array([core.ldet_nodes(t, re, im, w) for t in v])
        assert np.max(np.abs(got - same_input) / np.maximum(np.abs(same_input), 1.0)) < 1e-8
        exact = np.array([core.ldet_dense(t, p.W) for t in v])     # approximation error, reported separately
        assert np.max(np.abs(got - exact)) < 0.05 * len(p.y) ** 0.5


# ---------------------------------------------------------------------------
# persistent kernel (csrc/persist.cu): used for Independent-prior LM / SAR with 13 <= d <= 24
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("model,k", [("sar", 12), ("sar", 20), ("lm", 14), ("sar", 22)])
def test_persistent_kernel_trajectories(model, k):
    """Several chains per chain CTA, sweeps running ahead on the other SMs: every chain must still
    reproduce the oracle chain of its global id, burn-in tuning and continuation included."""
    p = small_problem(model, n=3000, k=k, knn=7)
    n_burn, n_save, n_chains, off = 40, 50, 6, 2
    with cuda_handle(model, p, seed=123, n_chains=n_chains, chain_offset=off) as h:
        a = h.run(n_burn, 30)
        b = h.run(0, n_save - 30)                                   # second launch continues on-chip state
        got = np.concatenate([a, b], axis=2)
        st = h.get_state()
    assert st["iterations"].tolist() == [n_burn + n_save] * n_chains
    for c in (0, 3, 5):
        ref = oracle_draws(oracle_chain(model, p, seed=123, chain=off + c), n_save, n_burn)
        assert relnorm(got[c].T, ref) < 1e-7, (model, k, c)


def test_persistent_equals_cached_statistics_path():
    """Stream mode (persistent kernel, Gram matrix rebuilt every iteration) and cached mode (one sweep,
    one launch per chain step) see the same Gram matrix bit for bit and run the same schedule; the
    two kernels are compiled separately, so draws agree to rounding, not bit for bit."""
    from bsreg_b200 import capi
    p = small_problem("sar", n=2500, k=16, knn=6)
    with cuda_handle("sar", p, seed=9, n_chains=5) as h:
        s = h.run(30, 25)
        G = h.eval_gram()
    with cuda_handle("sar", p, seed=9, n_chains=5, stats_mode=capi.STATS_CACHED) as h:
        c = h.run(30, 25)
        assert np.array_equal(G, h.eval_gram())
    assert relnorm(s, c) < 1e-9


# ---------------------------------------------------------------------------
# internal ordering of the observations (breadth-first relabelling at create) and per-tile distinct-column lists
# ---------------------------------------------------------------------------
def _scrambled(model, n, k, knn, seed=3):
    """A problem whose observations are in random order (no locality in W): triggers the relabelling."""
    from bsreg_b200.synthetic import Problem
    p = make_problem(model, n=n, k_reg=k, knn=knn, sort=True)
    q = np.random.default_rng(seed).permutation(n)
    W = p.W[q][:, q].tocsr()
    W.sort_indices()
    W.indices = W.indices.astype(np.int32)
    W.indptr = W.indptr.astype(np.int32)
    return Problem(p.y[q].copy(), np.ascontiguousarray(p.X[q]), W, p.beta, p.sigma, p.lam, p.model, p.coords[q])


@pytest.mark.parametrize("n", [4200, 9001])
def test_gram_is_invariant_to_internal_ordering(n, monkeypatch):
    """Z'Z does not depend on the order of the observations: the sweep over relabelled, de-duplicated tiles must give the
    oracle's Gram matrix, and the same one (to rounding) with the relabelling switched off or forced on."""
    p = _scrambled("sar", n, 9, 8)
    Wy = core.spmm(p.W, p.y)
    G_ref = core.gram(p.y, p.X, Wy, None)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    got = {}
    for mode in ("", "0", "1"):
        monkeypatch.setenv("BSREG_REORDER", mode)
        with cuda_handle("sar", p, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4) as h:
            got[mode] = h.eval_gram()
            assert relnorm(h.eval_spmm(p.y[:, None])[:, 0], Wy) < 1e-13     # K1 stays in the caller's ordering
        assert np.max(np.abs(got[mode] - G_ref) / scale) < 1e-13, mode
    assert np.max(np.abs(got["0"] - got["1"]) / scale) < 1e-14


def test_chains_are_invariant_to_internal_ordering(monkeypatch):
    """Same draws (to rounding) with and without the relabelling, persistent kernel included (d = 16)."""
    p = _scrambled("sar", 5000, 14, 7)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("BSREG_REORDER", mode)
        with cuda_handle("sar", p, seed=21, n_chains=5, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=20, cheb_probes=8) as h:
            out[mode] = h.run(30, 30)
    assert relnorm(out["0"], out["1"]) < 1e-9


def test_tiles_with_long_and_empty_rows(monkeypatch):
    """Ragged W: empty rows and one very long row (a tile with 1800 nonzeros: three ring stages of 56 KB)."""
    import scipy.sparse as sp
    n = 4500
    rng = np.random.default_rng(5)
    p = make_problem("sar", n=n, k_reg=6, knn=5)
    W = p.W.tolil()
    for i in (0, 63, 64, 1000, n - 1):
        W.rows[i] = []; W.data[i] = []
    cols = np.sort(rng.choice(n, 1500, replace=False))
    W.rows[2000] = cols.tolist(); W.data[2000] = (np.full(1500, 1.0 / 1500)).tolist()
    W = sp.csr_matrix(W)
    W.indices = W.indices.astype(np.int32); W.indptr = W.indptr.astype(np.int32)
    Wy = core.spmm(W, p.y)
    G_ref = core.gram(p.y, p.X, Wy, None)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    for mode in ("0", "1"):
        monkeypatch.setenv("BSREG_REORDER", mode)
        with capi_mod().Handle(p.y, p.X, model=capi_mod().MODEL_SAR, W=W, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=6,
                               cheb_probes=2, n_chains=1, seed=1) as h:
            assert np.max(np.abs(h.eval_gram() - G_ref) / scale) < 1e-13, mode


def capi_mod():
    from bsreg_b200 import capi
    return capi


def test_tensor_core_and_vector_gram_updates_agree(monkeypatch):
    """The persistent sweep forms Z'Z on the FP64 tensor cores (mma.m8n8k4) for 19 <= d <= 24; the vector-FMA form of the
    same kernel (BSREG_NO_MMA=1) must give the same chains to rounding, and both follow the oracle."""
    p = small_problem("sar", n=6000, k=20, knn=8)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("BSREG_NO_MMA", mode)
        with cuda_handle("sar", p, seed=77, n_chains=7, chain_offset=1) as h:
            out[mode] = h.run(25, 35)
    assert relnorm(out["0"], out["1"]) < 1e-9
    ref = oracle_draws(oracle_chain("sar", p, seed=77, chain=1 + 4), 35, 25)
    assert relnorm(out["0"][4].T, ref) < 1e-7


@pytest.mark.parametrize("n_chains", [1, 75])
def test_persistent_kernel_chain_counts(n_chains):
    """One resident chain (the reference's default) and more than 72 (four chains per chain CTA instead of three)."""
    p = small_problem("sar", n=3000, k=13, knn=7)
    with cuda_handle("sar", p, seed=31, n_chains=n_chains) as h:
        got = h.run(20, 25)
    for c in sorted({0, n_chains - 1, n_chains // 2}):
        ref = oracle_draws(oracle_chain("sar", p, seed=31, chain=c), 25, 20)
        assert relnorm(got[c].T, ref) < 1e-7, c


def test_sem_sweep_in_internal_ordering(monkeypatch):
    """SEM (and wide designs) sweep W, X, y directly; without locality they get the internal ordering too.  Same Gram matrix
    (to rounding) and same chains with the relabelling off / on, and K1 still answers in the caller's ordering."""
    p = _scrambled("sem", 6100, 5, 7)
    Wy, WX = core.spmm(p.W, p.y), core.spmm(p.W, p.X)
    G_ref = core.gram(p.y, p.X, Wy, WX)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    G, draws = {}, {}
    for mode in ("0", "1"):
        monkeypatch.setenv("BSREG_REORDER", mode)
        with cuda_handle("sem", p, seed=4, n_chains=3, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=12, cheb_probes=8) as h:
            G[mode] = h.eval_gram()
            assert relnorm(h.eval_spmm(p.X), WX) < 1e-13
            draws[mode] = h.run(15, 20)
        assert np.max(np.abs(G[mode] - G_ref) / scale) < 1e-13, mode
    assert relnorm(draws["0"], draws["1"]) < 1e-9


def test_persistent_kernel_with_cached_statistics():
    """BSREG_STATS_CACHED (the reference's schedule: statistics formed once) also runs the persistent kernel -- the sweeps
    fill the three Gram buffers once per launch -- and must follow the oracle, continuation included, and equal the
    streamed run bit for bit (same Gram matrix in every buffer, same chain code)."""
    from bsreg_b200 import capi
    p = small_problem("sar", n=3000, k=20, knn=7)
    with cuda_handle("sar", p, seed=19, n_chains=5, stats_mode=capi.STATS_CACHED) as h:
        a = h.run(30, 20)
        b = h.run(0, 15)
        got = np.concatenate([a, b], axis=2)
        assert h.get_state()["iterations"].tolist() == [65] * 5
    ref = oracle_draws(oracle_chain("sar", p, seed=19, chain=3), 35, 30)
    assert relnorm(got[3].T, ref) < 1e-7
    with cuda_handle("sar", p, seed=19, n_chains=5) as h:
        streamed = h.run(30, 35)
    assert np.array_equal(streamed, got)


def test_persistent_kernel_short_and_uneven_blocks():
    """Continuing a fit (bm.bm, R/60_interface.R:71-78) in launches of 3, 4, 5 and 7 iterations (every phase of the
    triple-buffered Gram rotation) reproduces the uninterrupted chain bit for bit; with one- and two-iteration blocks in
    between (they run the one-launch-per-stage kernels, compiled separately) it agrees to rounding.  (The burn-in stays in
    one call: the tuning schedule of R/50_execute.R:62-71 counts iterations per call.)"""
    p = small_problem("sar", n=2500, k=15, knn=6)
    with cuda_handle("sar", p, seed=2, n_chains=4) as h:
        whole = h.run(21, 19)
    with cuda_handle("sar", p, seed=2, n_chains=4) as h:
        parts = [h.run(21, 3)] + [h.run(0, k) for k in (4, 5, 7)]
    assert np.array_equal(np.concatenate(parts, axis=2), whole)
    with cuda_handle("sar", p, seed=2, n_chains=4) as h:
        parts = [h.run(21, 1)] + [h.run(0, k) for k in (3, 2, 6, 7)]
    assert relnorm(np.concatenate(parts, axis=2), whole) < 1e-9


def test_refused_cooperative_launch_falls_back_to_single_launches(monkeypatch):
    """If the device cannot hold every CTA of the persistent kernel (MPS share, MIG slice) the cooperative launch is
    refused before anything runs; the handle then keeps to one launch per stage -- still on the GPU, same chains."""
    p = small_problem("sar", n=2600, k=16, knn=6)
    with cuda_handle("sar", p, seed=6, n_chains=4) as h:
        want = h.run(20, 20)
    monkeypatch.setenv("BSREG_FORCE_COOP_FAIL", "1")
    with cuda_handle("sar", p, seed=6, n_chains=4) as h:
        got = np.concatenate([h.run(20, 8), h.run(0, 12)], axis=2)
        G = h.eval_gram()
    monkeypatch.delenv("BSREG_FORCE_COOP_FAIL")
    assert relnorm(got, want) < 1e-9
    G_ref = core.gram(p.y, p.X, core.spmm(p.W, p.y), None)
    assert np.max(np.abs(G - G_ref)) / np.max(np.abs(G_ref)) < 1e-12


def test_wide_design_sweep_in_internal_ordering(monkeypatch):
    """24 < d <= 88 (here K = 30) also sweeps W, X, y directly: same Gram matrix with the relabelling off / on."""
    p = _scrambled("sar", 5200, 30, 6)
    G_ref = core.gram(p.y, p.X, core.spmm(p.W, p.y), None)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    for mode in ("0", "1"):
        monkeypatch.setenv("BSREG_REORDER", mode)
        with cuda_handle("sar", p, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4) as h:
            assert np.max(np.abs(h.eval_gram() - G_ref) / scale) < 1e-13, mode


def test_slot_major_tiles_wider_than_the_unrolled_gather():
    """Rows of 13 nonzeros: the slot-major (ELL) tile layout with more slots than the gather loop unrolls (10)."""
    p = make_problem("sar", n=4500, k_reg=9, knn=13)
    G_ref = core.gram(p.y, p.X, core.spmm(p.W, p.y), None)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    with cuda_handle("sar", p, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4, n_chains=2, seed=3) as h:
        assert np.max(np.abs(h.eval_gram() - G_ref) / scale) < 1e-13
    p20 = make_problem("sar", n=4500, k_reg=20, knn=13)          # the same through the persistent kernel (d = 22, mma)
    nodes = None
    with cuda_handle("sar", p20, ldet=capi_mod().LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4, n_chains=2, seed=3) as h:
        nodes = h.nodes()
        got = h.run(10, 12)
    ref = oracle_draws(oracle_chain("sar", p20, seed=3, chain=1, nodes=nodes), 12, 10)
    assert relnorm(got[1].T, ref) < 1e-7


# ---------------------------------------------------------------------------
# one process driving several devices (bsreg_options.devices): shards are set up by one host thread each
# ---------------------------------------------------------------------------
def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0, 0], [0, 1], [0, 1, 2, 3]])
def test_one_handle_over_several_devices(devices):
    """Chains are split in contiguous blocks over the listed devices (the same device twice exercises the multi-shard
    host code on a single-GPU box); a chain id must give the same draws wherever it runs.  Each shard here holds as many
    chains as the single-device comparison run, so the two execute the same kernels on the same grid: bit for bit."""
    if max(devices) >= _n_devices():
        pytest.skip(f"needs {max(devices) + 1} CUDA devices")
    capi = capi_mod()
    p = small_problem("sar", n=5000, k=20, knn=8)               # d = 22: persistent kernel, > 48 KB of dynamic shared memory
    per, nd = 5, len(devices)
    kw = dict(model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=16, cheb_probes=8, seed=41)
    with capi.Handle(p.y, p.X, n_chains=per * nd, chain_offset=7, devices=devices, **kw) as h:
        G = h.eval_gram()
        a = h.run(30, 20)
        b = h.run(0, 10)                                       # continuation on every shard
        st = h.get_state()
    assert st["iterations"].tolist() == [60] * (per * nd)
    multi = np.concatenate([a, b], axis=2)
    for g in range(nd):
        with capi.Handle(p.y, p.X, n_chains=per, chain_offset=7 + per * g, devices=[devices[g]], **kw) as h1:
            single = h1.run(30, 30)
            assert np.array_equal(G, h1.eval_gram())
        assert np.array_equal(multi[per * g:per * (g + 1)], single), g
    # ... and a SEM design (one launch per stage, CUDA graphs per shard)
    q = small_problem("sem", n=1500, k=4)
    kw = dict(model=capi.MODEL_SEM, W=q.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=12, cheb_probes=8, seed=5)
    with capi.Handle(q.y, q.X, n_chains=2 * nd, devices=devices, **kw) as h:
        multi = h.run(30, 30)
    with capi.Handle(q.y, q.X, n_chains=2, chain_offset=2 * (nd - 1), devices=[devices[-1]], **kw) as h1:
        assert np.array_equal(multi[2 * (nd - 1):], h1.run(30, 30))


def test_chebyshev_rejects_unnormalised_w():
    """T_j(W) explodes unless the spectrum of W lies in [-1, 1]: a binary contiguity matrix must be refused (the
    reference's eigenvalue method has no such condition; eigenvalue nodes still work)."""
    capi = capi_mod()
    p = small_problem("sar", n=600, k=3)
    Wb = p.W.copy()
    Wb.data[:] = 1.0
    with pytest.raises(capi.BsregError, match="spectrum of W"):
        capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=Wb, ldet=capi.LDET_CHEBYSHEV)
    # column-stochastic W (||W||_1 = 1, rows do not sum to one) is fine
    Wc = sp_csr(p.W.T)
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=Wc, ldet=capi.LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4) as h:
        assert np.all(np.isfinite(h.eval_logdet(np.array([0.3]))))


def sp_csr(W):
    import scipy.sparse as sp
    W = sp.csr_matrix(W)
    W.sort_indices()
    W.indices = W.indices.astype(np.int32)
    W.indptr = W.indptr.astype(np.int32)
    return W


def test_chebyshev_against_exact_logdet_asymmetric():
    """Approximation error of the Chebyshev / Hutchinson log-determinant against the exact one on a small ASYMMETRIC W
    (kNN, complex spectrum inside the unit disc) across the lambda range: reported separately from same-input parity."""
    capi = capi_mod()
    p = small_problem("sar", n=1500, k=3, knn=8, symmetric=False)
    with capi.Handle(p.y, p.X, model=capi.MODEL_SAR, W=p.W, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=64,
                     seed=3) as h:
        v = np.array([-0.9, -0.5, 0.0, 0.3, 0.6, 0.8, 0.9])
        got = h.eval_logdet(v)
    exact = np.array([core.ldet_dense(t, p.W) for t in v])
    err = np.abs(got - exact)
    assert err[2] < 1e-9                                        # log|I| = 0
    assert np.all(err[np.abs(v) <= 0.6] < 0.15)                 # probe noise ~ sqrt(2 / (N R)) per moment
    assert np.all(err < 1.0)                                    # measured ~0.25 at |v| = 0.9: small beside
    #                                                             the (N - M)/2 log rss term the MH ratio adds it to


# ---------------------------------------------------------------------------
# SEM ring sweep (csrc/semsweep.cu): de-duplicated row gathers + FP64 mma Gram update of [y | Wy | X | WX]
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("k,knn,n", [(4, 6, 3100), (9, 8, 4500), (13, 13, 4200), (20, 10, 9001), (22, 7, 2500)])
def test_sem_ring_sweep_gram(k, knn, n, monkeypatch):
    """2, 4 and 6 column groups of 8; slot-major tiles wider than the unrolled loop (kNN 13); a last tile that is not full.
    The Gram matrix must match the oracle and (to rounding) the direct sweep of round 1, in the caller's ordering and in
    the breadth-first one."""
    capi = capi_mod()
    p = _scrambled("sem", n, k, knn)
    Wy, WX = core.spmm(p.W, p.y), core.spmm(p.W, p.X)
    G_ref = core.gram(p.y, p.X, Wy, WX)
    scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
    got = {}
    for ring, reorder in (("0", "1"), ("0", "0"), ("1", "1")):
        monkeypatch.setenv("BSREG_NO_SEM_RING", ring)
        monkeypatch.setenv("BSREG_REORDER", reorder)
        with cuda_handle("sem", p, ldet=capi.LDET_CHEBYSHEV, cheb_order=8, cheb_probes=4) as h:
            got[ring, reorder] = h.eval_gram()
            assert np.array_equal(got[ring, reorder], h.eval_gram())
        assert np.max(np.abs(got[ring, reorder] - G_ref) / scale) < 1e-13, (ring, reorder)
    assert np.max(np.abs(got["0", "1"] - got["1", "1"]) / scale) < 1e-14


def test_sem_ring_sweep_ragged_rows(monkeypatch):
    """Empty rows, rows of very different length (CSR tile layout) and a row too long for a ring stage (the handle then
    keeps the direct sweep): same Gram matrix, same chains."""
    import scipy.sparse as sp
    capi = capi_mod()
    n = 4500
    rng = np.random.default_rng(5)
    p = make_problem("sem", n=n, k_reg=6, knn=5)
    for long_row in (40, 1500):
        W = p.W.tolil()
        for i in (0, 63, 64, 1000, n - 1):
            W.rows[i] = []; W.data[i] = []
        cols = np.sort(rng.choice(n, long_row, replace=False))
        W.rows[2000] = cols.tolist(); W.data[2000] = (np.full(long_row, 1.0 / long_row)).tolist()
        W = sp_csr(sp.csr_matrix(W))
        Wy, WX = core.spmm(W, p.y), core.spmm(W, p.X)
        G_ref = core.gram(p.y, p.X, Wy, WX)
        scale = np.sqrt(np.outer(np.diag(G_ref), np.diag(G_ref)))
        for mode in ("0", "1"):
            monkeypatch.setenv("BSREG_REORDER", mode)
            with capi.Handle(p.y, p.X, model=capi.MODEL_SEM, W=W, ldet=capi.LDET_CHEBYSHEV, cheb_order=6, cheb_probes=2,
                             n_chains=2, seed=1) as h:
                assert np.max(np.abs(h.eval_gram() - G_ref) / scale) < 1e-13, (long_row, mode)
                assert np.all(np.isfinite(h.run(5, 5)))


def test_sem_ring_sweep_chains_follow_the_oracle():
    """SEM at K = 20 (d = 42) in stream mode: every iteration's Gram matrix comes from the ring sweep."""
    p = small_problem("sem", n=6000, k=20, knn=8)
    capi = capi_mod()
    with cuda_handle("sem", p, seed=17, n_chains=5, chain_offset=2, ldet=capi.LDET_CHEBYSHEV, cheb_order=16, cheb_probes=8) as h:
        nodes = h.nodes()
        got = h.run(25, 30)
    ref = oracle_draws(oracle_chain("sem", p, seed=17, chain=2 + 3, nodes=nodes), 30, 25)
    assert relnorm(got[3].T, ref) < 1e-7
    with cuda_handle("sem", p, seed=17, n_chains=5, chain_offset=2, ldet=capi.LDET_CHEBYSHEV, cheb_order=16, cheb_probes=8,
                     stats_mode=capi.STATS_CACHED) as h:
        assert np.array_equal(h.run(25, 30), got)


# ---------------------------------------------------------------------------
# grid + spline log-determinant, ldet_* = list(grid = TRUE) (R/15_sar.R:79-87, R/15_sem.R:91-99): csrc/logdet.cu
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("symmetric,reps", [(True, 1), (False, 1), (True, 3)])
def test_grid_spline_logdet(symmetric, reps):
    """The determinant grid (dense LU with partial pivoting on the device, one CTA per grid point) against LAPACK's, the
    FMM coefficients against the oracle's restatement, and the per-proposal evaluation -- 1e-8 on log-determinants."""
    import scipy.sparse as sp
    from oracle.spline import SplineLogdet, spline_eval
    capi = capi_mod()
    p = small_problem("sar", n=231, k=3, symmetric=symmetric)
    W = sp_csr(sp.kron(sp.eye(reps), p.W)) if reps > 1 else p.W      # panel: kronecker(diag(reps), W), paper/0_data.R:27
    y, X = np.tile(p.y, reps), np.tile(p.X, (reps, 1))
    grid = (-1.0, 1.0 - 1e-12, 100)
    with capi.Handle(y, X, model=capi.MODEL_SAR, W=W, ldet=capi.LDET_SPLINE, ldet_grid=grid, ldet_reps=reps, seed=2) as h:
        x, ld, b, c, d = h.spline()
        ref = SplineLogdet(W, grid, reps)
        assert np.array_equal(x, ref.x)
        fin = np.isfinite(ref.y)
        assert np.array_equal(np.isfinite(ld), fin)
        # 1e-8 wherever I - x W is well conditioned; at x = 1 - 1e-12 (W row-stochastic: a pivot of ~1e-12 left after
        # cancellation) the determinant itself is only defined to ~1e-4 relative, whichever LU computes it
        well = fin & (np.abs(x) < 1 - 1e-6)
        assert np.max(np.abs(ld[well] - ref.y[well]) / np.maximum(np.abs(ref.y[well]), 1.0)) < 1e-8
        assert np.max(np.abs(ld[fin] - ref.y[fin])) < 1e-2
        if fin.all():
            from oracle.spline import fmm_coefficients
            for got, want in zip((b, c, d), fmm_coefficients(x, ld)):    # same grid values in: same spline out
                assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
            v = np.array([-0.93, -0.4, 0.0, 0.31, 0.77, 0.985])
            want = np.array([spline_eval(t, x, ld, b, c, d) for t in v])               # same coefficients: 1e-8
            assert np.max(np.abs(h.eval_logdet(v) - want) / np.maximum(np.abs(want), 1.0)) < 1e-8
            dense = np.array([core.ldet_dense(t, W) for t in v[:5]])
            assert np.max(np.abs(h.eval_logdet(v[:5]) - dense)) < 5e-3 * reps * np.max(np.abs(dense))


@pytest.mark.parametrize("model", ["sar", "sem"])
def test_chain_with_grid_spline_logdet_matches_oracle(model):
    capi = capi_mod()
    p = small_problem(model, n=300, k=4)
    grid = (-1.0, 1.0 - 1e-12, 60)
    with capi.Handle(p.y, p.X, model=MODEL_CODE_[model], W=p.W, ldet=capi.LDET_SPLINE, ldet_grid=grid, seed=77,
                     n_chains=3, chain_offset=4) as h:
        got = h.run(40, 50)
        x, ld = h.spline()[:2]
    from oracle.sampler import PhiloxRNG, RefChain, sample
    ref = sample(RefChain(model, p.y, p.X, W=p.W, ldet=dict(kind="spline", i_lambda=grid, reps=1, values=(x, ld)),
                          rng=PhiloxRNG(77, 5)), 50, 40)
    assert relnorm(got[1].T, ref) < 1e-7


MODEL_CODE_ = {"sar": 1, "sem": 2}


# ---------------------------------------------------------------------------
# sampled connectivity parameter delta of the SLX model (R/12_slx.R:87-106, R/20_mh.R:229-278): csrc/slxdelta.cu
# ---------------------------------------------------------------------------
def _delta_problem(n=900, knn=6, seed=0, k=3):
    import scipy.sparse as sp
    from bsreg_b200.synthetic import knn_weights
    from oracle.sampler import psi_power
    W, pts = knn_weights(n, knn)
    r = np.repeat(np.arange(n), np.diff(W.indptr))
    D = sp_csr(sp.csr_matrix((np.linalg.norm(pts[r] - pts[W.indices], axis=1), W.indices, W.indptr), shape=W.shape))
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(n), rng.standard_normal((n, k - 1))])
    Xl = X[:, 1:3].copy()
    y = X @ rng.standard_normal(k) + psi_power(dict(dist=D, normalise="row"), 3.0) @ Xl @ np.array([0.8, -0.5]) \
        + 0.3 * rng.standard_normal(n)
    return y, X, Xl, D


@pytest.mark.parametrize("prior,normalise", [("Independent", "row"), ("Horseshoe", "row"), ("Independent", "none")])
def test_sampled_delta_chain_matches_oracle(prior, normalise):
    """Every chain proposes its own delta: the per-chain sweep forms W(delta') X_SLX and the Gram blocks it touches, the
    decision kernel re-solves for beta at the proposal and at the current value.  Trajectories (beta, sigma, delta_SLX)
    against the oracle, burn-in tuning of the delta proposal scale and continuation included."""
    from oracle.sampler import PhiloxRNG, RefChain, sample
    capi = capi_mod()
    y, X, Xl, D = _delta_problem()
    slx = dict(delta=2.0, delta_scale=0.3, delta_a=2.0, delta_b=2.0, delta_min=1e-12, delta_max=float("inf"))
    n_burn, n_save, off = 50, 40, 3
    with capi.Handle(y, X, model=capi.MODEL_LM, prior=PRIOR_CODE_[prior], W=D, X_slx=Xl, n_chains=5, chain_offset=off, seed=13,
                     slx_psi=dict(normalise=normalise, **slx)) as h:
        assert h.P == X.shape[1] + 2 + 2
        a = h.run(n_burn, 25)
        b = h.run(0, n_save - 25)
        got = np.concatenate([a, b], axis=2)
        st = h.get_delta_state()
        assert h.get_state()["iterations"].tolist() == [n_burn + n_save] * 5
    for c in (0, 4):
        ch = RefChain("slx", y, X, X_slx=Xl, psi_slx=dict(dist=D, normalise=normalise), prior_type=prior,
                      priors={"SLX": slx}, rng=PhiloxRNG(13, off + c))
        ref = sample(ch, n_save, n_burn)
        assert ch.parameter_names()[-1] == "delta_SLX"
        assert relnorm(got[c].T, ref) < 1e-7, (prior, c)
        mh = ch.MH["delta"]
        assert abs(st["scale"][c] - mh.scale) < 1e-12 * mh.scale
        assert st["counters"][c].tolist() == [mh.accepted, mh.proposed, mh.accepted_tune, mh.proposed_tune]
        assert abs(st["delta"][c] - mh.value) < 1e-9


PRIOR_CODE_ = {"Independent": 0, "Conjugate": 1, "Shrinkage": 2, "Horseshoe": 3}


def test_connectivity_function_with_fixed_delta_equals_the_fixed_matrix():
    """delta_scale = 0 (the default): Psi(delta0) is used as it is -- same beta and sigma draws as the fixed-matrix SLX
    model with W = Psi(delta0) (to rounding: the lags are formed by another kernel), plus a constant delta_SLX column."""
    from oracle.sampler import psi_power
    capi = capi_mod()
    y, X, Xl, D = _delta_problem(n=700)
    with capi.Handle(y, X, model=capi.MODEL_LM, W=D, X_slx=Xl, n_chains=2, seed=5,
                     slx_psi=dict(normalise="row", delta=2.5, delta_scale=0.0)) as h:
        fun = h.run(20, 30)
    Wfix = sp_csr(psi_power(dict(dist=D, normalise="row"), 2.5))
    with capi.Handle(y, X, model=capi.MODEL_LM, W=Wfix, X_slx=Xl, n_chains=2, seed=5) as h:
        fix = h.run(20, 30)
    assert np.all(fun[:, -1, :] == 2.5)
    assert relnorm(fun[:, :-1, :], fix) < 1e-9


def test_delta_state_round_trip_and_interface():
    import bsreg_b200 as bs
    capi = capi_mod()
    y, X, Xl, D = _delta_problem(n=600)
    kw = dict(model=capi.MODEL_LM, W=D, X_slx=Xl, n_chains=3, seed=21,
              slx_psi=dict(normalise="row", delta=2.0, delta_scale=0.2, delta_a=2.0, delta_b=2.0))
    with capi.Handle(y, X, **kw) as h:
        whole = h.run(30, 30)
    with capi.Handle(y, X, **kw) as h:
        first = h.run(30, 12)
        st, dst = h.get_state(), h.get_delta_state()
    with capi.Handle(y, X, **kw) as h2:                   # a fresh handle continues from the saved state
        h2.set_state(st)
        h2.set_delta_state(dst)
        rest = h2.run(0, 18)
    assert relnorm(np.concatenate([first, rest], axis=2), whole) < 1e-9
    # public interface: bslx with a connectivity function, the paper's call (paper/1b_estimate_dist.R:19-22)
    fit = bs.bslx((y, X), W=bs.PsiPower(D), X_SLX=Xl, n_save=200, n_burn=200, verbose=False, seed=4, n_chains=4,
                  options=bs.set_options(SLX=bs.set_SLX(delta=3, delta_scale=0.05, delta_a=2, delta_b=2)))
    assert fit.colnames[-2:] == ["sigma", "delta_SLX"] and fit.chains.shape == (4, 200, X.shape[1] + 2 + 2)
    assert abs(fit.chains[:, :, -1].mean() - 3.0) < 0.5                  # simulated with delta = 3
    assert 0.0 < fit.model.MH["SLX_delta"].get_acceptance() < 1.0 and "SLX_delta" in fit.model.get_meta
