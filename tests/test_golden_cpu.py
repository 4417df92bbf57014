"""Oracle vs the committed golden vectors of BASELINE.json configs 1-3 (real data: the
reference's cigarettes panel with the queen-contiguity W derived from its us_states polygons).
Regenerate with `python tests/golden/make_golden.py` only when the oracle changes on purpose."""
import numpy as np
import pytest

from oracle import core
from oracle.sampler import PhiloxRNG, RefChain, sample
from tests.golden.make_golden import configs, load_cigarettes, panel_w


@pytest.fixture(scope="module")
def cig():
    return load_cigarettes()


def test_fixture_shape_matches_the_reference_dataset(cig):
    assert cig["sales"].shape == (1380,) and cig["contiguity"].shape == (46, 46)
    assert np.array_equal(cig["year"][:46], np.full(46, 1963)) and cig["year"][-1] == 1992    # year-major
    A = cig["contiguity"]
    assert np.array_equal(A, A.T) and A.sum() == 188 and A.diagonal().sum() == 0              # 94 links
    deg = A.sum(1)
    names = list(cig["state_names"])
    assert deg.min() == 1 and deg.max() == 8 and deg[names.index("Missouri")] == 8 and deg[names.index("Maine")] == 1
    W = panel_w(cig)
    assert W.shape == (1380, 1380) and np.allclose(W.sum(1), 1.0)
    ev = np.linalg.eigvals((A / A.sum(1, keepdims=True)))
    assert np.max(np.abs(ev.imag)) < 1e-12 and abs(ev.real.max() - 1) < 1e-12 and abs(ev.real.min() + 0.7182) < 1e-3


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_reproduces_golden(cig, idx):
    name, model, prior, pri, y, X, W, X_slx, reps = list(configs(cig))[idx]
    g = np.load(f"tests/golden/golden_{name}.npz")
    ch = RefChain(model, y, X, W=W, X_slx=X_slx, prior_type=prior, priors=pri,
                  ldet=dict(kind="eigen", reps=reps), rng=PhiloxRNG(int(g["seed"]), 0))
    assert np.allclose(ch.beta, g["beta0"], rtol=1e-9) and abs(ch.sigma2 - g["sigma2_0"]) < 1e-12 * g["sigma2_0"] + 1e-15
    if W is not None:
        got = np.array([core.ldet_nodes(v, *ch.nodes) for v in g["ldet_v"]])
        assert np.max(np.abs(got - g["ldet"])) < 1e-8 * max(1.0, np.max(np.abs(g["ldet"])))
        # the Kronecker shortcut (reps = 30) equals the full 1380 x 1380 determinant
        full = core.ldet_dense(0.5, ch.W)
        assert abs(core.ldet_nodes(0.5, *ch.nodes) - full) < 1e-8 * abs(full)
    d = sample(ch, int(g["n_save"]), int(g["n_burn"]))
    assert list(ch.parameter_names()) == list(g["names"])
    assert np.max(np.abs(d - g["draws"])) < 1e-9 * np.max(np.abs(g["draws"]))
    if "mh" in g:
        mh = ch.MH["lambda"]
        assert np.allclose([mh.scale, mh.accepted, mh.proposed, mh.accepted_tune, mh.proposed_tune], g["mh"])


def test_cfg1_posterior_agrees_with_ols(cig):
    """README demo blm(log(sales) ~ log(price)): flat prior => posterior mean = OLS."""
    _, model, prior, pri, y, X, *_ = list(configs(cig))[0]
    from oracle.sampler import NumpyRNG
    d = sample(RefChain(model, y, X, rng=NumpyRNG(0)), 2000, 200)
    ols = np.linalg.lstsq(X, y, rcond=None)[0]
    se = np.sqrt(np.diag(np.linalg.inv(X.T @ X)) * np.sum((y - X @ ols) ** 2) / (len(y) - 2))
    assert np.all(np.abs(d[:, :2].mean(0) - ols) < 4 * se / np.sqrt(2000) * 2)
