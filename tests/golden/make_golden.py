"""Generate the committed fixtures under tests/golden/ (run in the build container only):

    python tests/golden/make_golden.py

1. cigarettes.npz -- the reference's bundled panel (data/cigarettes.rda: 46 states x 30 years,
   year-major) plus the 46 x 46 queen-contiguity matrix derived from data/us_states.rda (states
   sharing at least one polygon vertex).  The reference bundles no W.
2. golden_cfg{1,2,3}.npz -- frozen oracle outputs on BASELINE.json configs 1-3 (deterministic
   pieces + a short chain with the addressed Philox stream), so that a change in the oracle or in
   the CUDA path shows up as a diff against committed numbers.

The reference itself cannot be run (no R in the image), so these vectors freeze the ORACLE, not R:
parity with R stays unpinned (oracle/__init__.py).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

REF_DATA = Path("/root/reference/data")
NUMERIC = ("id", "year", "price", "pop", "pop16", "cpi", "ndi", "sales", "pimin", "longitude", "latitude")


def build_cigarettes():
    from tests.golden import rda
    cig = rda.data_frame(rda.read_rda(REF_DATA / "cigarettes.rda")["cigarettes"])
    us = rda.data_frame(rda.read_rda(REF_DATA / "us_states.rda")["us_states"])
    states = cig["name"][:46]
    assert len(set(states)) == 46 and cig["name"][46:92] == states          # year-major order
    verts = {}
    for name, geom in zip(us["name"], us["geometry"]):
        pts = set()
        for poly in geom:
            for ring in poly:
                pts.update(map(tuple, np.round(np.asarray(ring).reshape(-1, 2), 12)))
        verts[name] = pts
    A = np.zeros((46, 46), dtype=np.uint8)
    for i, a in enumerate(states):
        for j, b in enumerate(states):
            if i < j and verts[a] & verts[b]:
                A[i, j] = A[j, i] = 1
    assert A.sum() == 2 * 94 and A.sum(1).min() >= 1, (A.sum(), A.sum(1).min())
    out = {k: np.asarray(cig[k], dtype=np.float64) for k in NUMERIC}
    out["state"] = np.tile(np.arange(46), 30)
    out["state_names"] = np.array(states)
    out["contiguity"] = A
    np.savez_compressed(HERE / "cigarettes.npz", **out)
    return out


def load_cigarettes():
    return dict(np.load(HERE / "cigarettes.npz", allow_pickle=False))


def panel_w(d, n_time=30):
    """W_cont = kronecker(diag(n_time), contig / rowSums(contig)), paper/0_data.R:24-27."""
    A = d["contiguity"].astype(np.float64)
    return sp.kron(sp.identity(n_time), sp.csr_matrix(A / A.sum(1, keepdims=True)), format="csr")


def design(d, fixed_effects=False):
    """log(sales) ~ log(price) [+ log(ndi / cpi) + factor(name) + factor(year)], R/60_interface.R:26-31."""
    y = np.log(d["sales"])
    cols = [np.ones_like(y), np.log(d["price"])]
    if fixed_effects:
        cols.append(np.log(d["ndi"] / d["cpi"]))
        st, yr = d["state"].astype(int), d["year"].astype(int)
        names = d["state_names"]
        order = np.argsort(names)                           # factor() levels are sorted
        cols += [(st == s).astype(float) for s in order[1:]]
        cols += [(yr == t).astype(float) for t in sorted(set(yr))[1:]]
    return y, np.column_stack(cols)


def configs(d):
    """(name, model, prior, priors, y, X, W, X_slx, reps) for BASELINE.json configs 1-3."""
    y, X = design(d)
    W = panel_w(d)
    yield "cfg1", "lm", "Independent", {}, y, X, None, None, 1
    y2, X2 = design(d)
    X2 = np.column_stack([X2, np.log(d["ndi"] / d["cpi"])])
    yield "cfg2", "sar", "Independent", {}, y2, X2, W, None, 30
    yield "cfg3", "sdem", "Shrinkage", {"SNG": dict(theta_scale=0.5)}, y2, X2, W, X2[:, 1:], 30


def make_golden(d, n_burn=40, n_save=25, seed=20260101):
    from oracle import core
    from oracle.sampler import PhiloxRNG, RefChain, sample
    for name, model, prior, pri, y, X, W, X_slx, reps in configs(d):
        ch = RefChain(model, y, X, W=W, X_slx=X_slx, prior_type=prior, priors=pri,
                      ldet=dict(kind="eigen", reps=reps), rng=PhiloxRNG(seed, 0))
        out = dict(beta0=ch.beta.copy(), sigma2_0=ch.sigma2)
        if W is not None:
            Wy = core.spmm(ch.W, ch.y)
            WX = core.spmm(ch.W, ch.X) if ch.has_sem else None
            out["gram"] = core.gram(ch.y, ch.X, Wy, WX)
            v = np.array([-0.9, -0.3, 0.0, 0.2, 0.5, 0.8, 0.95])
            out["ldet_v"] = v
            out["ldet"] = np.array([core.ldet_nodes(x, *ch.nodes) for x in v])
            out["nodes_re"], out["nodes_im"], out["nodes_w"] = ch.nodes
        else:
            out["gram"] = core.gram(ch.y, ch.X, np.zeros_like(ch.y))
        out["draws"] = sample(ch, n_save, n_burn)
        out["names"] = np.array(ch.parameter_names())
        out["n_burn"], out["n_save"], out["seed"] = n_burn, n_save, seed
        if "lambda" in ch.MH:
            mh = ch.MH["lambda"]
            out["mh"] = np.array([mh.scale, mh.accepted, mh.proposed, mh.accepted_tune, mh.proposed_tune])
        np.savez_compressed(HERE / f"golden_{name}.npz", **out)
        print(name, out["draws"].shape, out["draws"].mean(0)[[0, 1, -2, -1]])


if __name__ == "__main__":
    data = build_cigarettes() if REF_DATA.exists() else load_cigarettes()
    make_golden(load_cigarettes())
