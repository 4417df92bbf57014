"""Host-side mirror of the R interface: argument handling, defaults, error messages and the
reference quirks it replicates (no GPU needed; model calls are covered by the gpu tests)."""
import math

import numpy as np
import pytest

import bsreg_b200 as bs
from bsreg_b200 import interface as bi


def test_set_mh_defaults_and_quirk_q1():
    mh = bs.set_mh()
    assert mh["adjust_burn"] == 0.8 and mh["acc_target"] == (0.20, 0.45)
    assert mh["acc_change"] == 0.8                       # R/42_set_mh.R:23 reads adjust_burn
    assert bs.set_mh(0.5, (0.1, 0.5), .05)["acc_change"] == 0.5      # the roxygen example, R/42_set_mh.R:11-12
    with pytest.raises(ValueError, match="MH tuning period"):
        bs.set_mh(adjust_burn=1.5)
    with pytest.raises(ValueError, match="target range"):
        bs.set_mh(acc_target=(0.2, 1.2))


def test_set_options_defaults_match_reference():
    o = bs.set_options()
    assert o["type"] == "Independent" and set(o["priors"]) >= {"NG", "SNG", "HS", "SAR", "SLX", "SEM", "SV"}
    ng = o["priors"]["NG"]
    assert (ng["mu"], ng["precision"], ng["shape"], ng["rate"], ng["beta"], ng["sigma"]) == (0, 1e-8, 0.01, 0.01, None, None)
    sar = o["priors"]["SAR"]
    assert (sar["lambda_a"], sar["lambda_b"], sar["lam"], sar["lambda_scale"]) == (1.01, 1.01, 0, 0.1)
    assert (sar["lambda_min"], sar["lambda_max"]) == (-1, 1 - 1e-12)
    assert (sar["delta"], sar["delta_scale"], sar["delta_min"], sar["delta_max"]) == (1, 0, 1e-12, math.inf)
    sng = o["priors"]["SNG"]
    assert (sng["lambda_a"], sng["lambda_b"], sng["theta_scale"], sng["theta_a"], sng["lam"], sng["tau"], sng["theta"]) == \
        (0.01, 0.01, 0, 1, 1, 10, 0.1)
    hs = o["priors"]["HS"]
    assert (hs["lam"], hs["tau"], hs["zeta"], hs["nu"]) == (1, 1, 1, 1)
    assert bs.set_SLX is bs.set_SAR and bs.set_SEM is bs.set_SAR             # R/41_set_options.R:149-152
    assert bs.set_options("Shrinkage", SNG=bs.set_SNG(lambda_a=1, lambda_b=1))["type"] == "Shrinkage"   # :18-19
    assert bs.set_options("Hors")["type"] == "Horseshoe"                     # match.arg partial matching
    with pytest.raises(ValueError, match="should be one of"):
        bs.set_options("Ridge")
    assert o["priors"]["SV"] is None                                         # quirk Q8 not inherited


@pytest.mark.parametrize("call,msg", [
    (lambda: bs.set_NG(precision=-1), "valid prior variance via 'precision'"),
    (lambda: bs.set_NG(shape=0), "valid prior shape"),
    (lambda: bs.set_NG(mu="a"), "prior mean via 'mu'"),
    (lambda: bs.set_SAR(lambda_=2), "valid starting value for lambda \\(spatial\\)"),
    (lambda: bs.set_SAR(lambda_scale=0), "proposal scale for lambda"),
    (lambda: bs.set_SAR(**{"lambda": -3}), "starting value for lambda"),
    (lambda: bs.set_SNG(theta=0), "starting value for 'theta'"),
    (lambda: bs.set_HS(tau=-1), "starting value for 'tau' \\(Horseshoe\\)"),
    (lambda: bs.set_SNG(lambda_a=[1, 2]), "shape for lambda \\(Normal-Gamma\\)"),
])
def test_validation_messages_are_the_reference_ones(call, msg):
    with pytest.raises(ValueError, match=msg):
        call()


def test_vector_valued_prior_arguments():
    ng = bs.set_NG(mu=[0, 1, 2], precision=[1, 2, 3])
    assert list(ng["mu"]) == [0, 1, 2] and list(ng["precision"]) == [1, 2, 3]


def test_model_matrix_subset():
    rng = np.random.default_rng(0)
    data = dict(sales=rng.uniform(50, 200, 12), price=rng.uniform(20, 60, 12), ndi=rng.uniform(1, 2, 12),
                cpi=rng.uniform(1, 2, 12), name=np.repeat(["b", "a", "c"], 4), year=np.tile([1963, 1964, 1965, 1966], 3))
    y, X, names = bi.model_matrix("log(sales) ~ log(price)", data)
    assert names == ["(Intercept)", "log(price)"] and np.allclose(y, np.log(data["sales"]))
    assert np.all(X[:, 0] == 1) and np.allclose(X[:, 1], np.log(data["price"]))
    y, X, names = bi.model_matrix("log(sales) ~ log(price) + log(ndi / cpi) + factor(name) + factor(year)", data)
    assert names == ["(Intercept)", "log(price)", "log(ndi / cpi)", "factor(name)b", "factor(name)c",
                     "factor(year)1964", "factor(year)1965", "factor(year)1966"]
    assert X.shape == (12, 8) and np.linalg.matrix_rank(X) == 8
    _, X0, names0 = bi.model_matrix("sales ~ price - 1", data)
    assert names0 == ["price"] and X0.shape == (12, 1)
    _, X2, _ = bi.model_matrix("sales ~ I(price^2)", data)
    assert np.allclose(X2[:, 1], data["price"] ** 2)


def test_errors_that_precede_device_work():
    y, X = np.zeros(10), np.ones((10, 2))
    with pytest.raises(ValueError, match="Please provide a connectivity matrix 'Psi_SAR'"):     # R/15_sar.R:52
        bs.bsar((y, X), verbose=False)
    with pytest.raises(ValueError, match="Please provide a connectivity matrix 'Psi_SEM'"):     # R/15_sem.R:58
        bs.bsem((y, X), verbose=False)
    with pytest.raises(ValueError, match="N x N"):
        bs.bsar((y, X), W=np.eye(4), verbose=False)
    with pytest.raises(NotImplementedError, match="stochastic volatility"):
        bs.bsv((y, X))
    with pytest.raises(ValueError, match=r"Connectivity function 'Psi' \(SLX\) required to sample 'delta'"):   # R/12_slx.R:79
        bs.bslx((y, np.column_stack([np.ones(10), np.arange(10.)])), W=np.eye(10),
                options=bs.set_options(SLX=bs.set_SLX(delta_scale=0.1)), verbose=False)
    with pytest.raises(NotImplementedError, match="SLX model"):
        bs.bsdm((y, np.column_stack([np.ones(10), np.arange(10.)])), W=bs.PsiPower(np.ones((10, 10)) - np.eye(10)), verbose=False)
    assert np.allclose(bs.PsiPower(np.array([[0, 2.0], [4.0, 0]]), "none")(2).toarray(), [[0, 0.25], [1 / 16, 0]])
    with pytest.raises(ValueError, match="should be one of"):
        bs.bm((y, X), type="probit")
    with pytest.raises(ValueError, match="'reps' must divide"):
        bi.ldet_nodes(bi._as_csr(np.eye(4) * 0.0 + np.diag(np.ones(3), 1), 4, "Psi_SAR"), dict(reps=3))


def test_ldet_nodes_reps_and_symmetry():
    A = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=float)
    W0 = A / A.sum(1, keepdims=True)                                   # asymmetric, real spectrum
    W = bi._as_csr(np.kron(np.eye(4), W0), 12, "Psi_SAR")
    re, im, w = bi.ldet_nodes(W, dict(reps=4))                         # partial list replaces the default list
    assert len(re) == 3 and np.all(w == 4.0)
    v = 0.6
    dense = np.linalg.slogdet(np.eye(12) - v * W.toarray())[1]
    assert abs(float(np.sum(w * 0.5 * np.log((1 - v * re) ** 2 + (v * im) ** 2))) - dense) < 1e-12
    assert not bi.is_symmetric(W) and bi.is_symmetric(bi._as_csr(A, 3, "W"))
    with pytest.raises(ValueError, match="divide"):
        bi.ldet_nodes(W, dict(reps=5))
