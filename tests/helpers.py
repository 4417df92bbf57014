"""Shared builders for the parity tests (oracle side and CUDA side on the same inputs)."""
from __future__ import annotations

import numpy as np

from bsreg_b200 import capi
from bsreg_b200.synthetic import make_problem
from oracle import core
from oracle.sampler import PhiloxRNG, RefChain, default_mh, sample

MODEL_CODE = {"lm": capi.MODEL_LM, "slx": capi.MODEL_LM, "sar": capi.MODEL_SAR, "sdm": capi.MODEL_SAR,
              "sem": capi.MODEL_SEM, "sdem": capi.MODEL_SEM}
PRIOR_CODE = {"Independent": capi.PRIOR_INDEPENDENT, "Conjugate": capi.PRIOR_CONJUGATE,
              "Shrinkage": capi.PRIOR_SHRINKAGE, "Horseshoe": capi.PRIOR_HORSESHOE}


def small_problem(model, n=400, k=5, knn=6, symmetric=True):
    return make_problem(model, n=n, k_reg=k, knn=knn, symmetric=symmetric)


def oracle_chain(model, prob, prior="Independent", priors=None, seed=11, chain=0, nodes=None, reps=1, x_slx=None):
    ldet = dict(kind="eigen", reps=reps) if nodes is None else dict(kind="nodes", re=nodes[0], im=nodes[1], w=nodes[2])
    return RefChain(model, prob.y, prob.X, W=prob.W if model != "lm" else None, X_slx=x_slx,
                    prior_type=prior, priors=priors or {}, ldet=ldet, rng=PhiloxRNG(seed, chain))


def cuda_handle(model, prob, prior="Independent", priors=None, seed=11, n_chains=1, chain_offset=0, nodes=None,
                x_slx=None, stats_mode=capi.STATS_STREAM, **kw):
    priors = priors or {}
    spatial = priors.get("SAR") or priors.get("SEM")
    if nodes is None and model not in ("lm", "slx") and "ldet" not in kw:
        nodes = core.eigen_nodes(prob.W, 1)
    return capi.Handle(prob.y, prob.X, model=MODEL_CODE[model], prior=PRIOR_CODE[prior],
                       W=prob.W if model != "lm" else None, X_slx=x_slx, nodes=nodes,
                       stats_mode=stats_mode, n_chains=n_chains, chain_offset=chain_offset, seed=seed,
                       ng=priors.get("NG"), sng=priors.get("SNG"), hs=priors.get("HS"), spatial=spatial, **kw)


def oracle_draws(chain, n_save, n_burn, mh=None):
    return sample(chain, n_save, n_burn, mh or default_mh())
