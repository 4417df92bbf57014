"""Synthetic SAR / SEM problems of the shape BASELINE.json names.

Host-side harness code (NumPy/SciPy): coordinates U(0,1)^2 (seed 1001), W = kNN
graph (k = 10, self excluded, row-stochastic weights 1/k, CSR int32 + float64),
X = [1, N(0,1)^(K-1)] (seed 1002), beta ~ N(0,1) (seed 1003), sigma = 1,
lambda = rho = 0.5, eps seed 1004.  The spatial filter inverse is applied by the
Neumann iteration v <- b + lam*W v, SpMV only, so it works at N = 1e6.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree


@dataclass
class Problem:
    y: np.ndarray          # (N,)
    X: np.ndarray          # (N, K) float64
    W: sp.csr_matrix       # (N, N) CSR, int32 indices
    beta: np.ndarray
    sigma: float
    lam: float
    model: str
    coords: np.ndarray


def knn_weights(n: int, k: int = 10, seed: int = 1001, symmetric: bool = False, sort: bool = False):
    """Row-stochastic k-nearest-neighbour W.  `symmetric=True` gives the union-kNN
    D^-1 A variant (real spectrum) used as the exactness fixture at small N.
    `sort=True` orders the points along a Morton (Z-order) curve first; this is a
    relabelling of the observations only (the data are i.i.d. draws) and makes
    neighbours close in memory."""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    if sort:
        pts = pts[np.argsort(_morton(pts), kind="stable")]
    tree = cKDTree(pts)
    _, idx = tree.query(pts, k=k + 1, workers=-1)
    idx = idx[:, 1:].astype(np.int64)
    rows = np.repeat(np.arange(n), k)
    A = sp.csr_matrix((np.ones(n * k), (rows, idx.ravel())), shape=(n, n))
    if symmetric:
        A = ((A + A.T) > 0).astype(np.float64).tocsr()
    deg = np.asarray(A.sum(axis=1)).ravel()
    W = sp.diags(1.0 / deg) @ A
    W = sp.csr_matrix(W)
    W.sort_indices()
    W.indices = W.indices.astype(np.int32)
    W.indptr = W.indptr.astype(np.int32)
    return W, pts


def _morton(pts: np.ndarray) -> np.ndarray:
    q = np.minimum((pts * 65536.0).astype(np.uint64), np.uint64(65535))

    def spread(v):
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))


def neumann_solve(W: sp.csr_matrix, b: np.ndarray, lam: float, iters: int = 80) -> np.ndarray:
    """(I - lam W)^-1 b by v <- b + lam W v  (||lam W||_inf = |lam| < 1)."""
    v = b.copy()
    for _ in range(iters):
        v = b + lam * (W @ v)
    return v


def make_problem(model: str = "sar", n: int = 100_000, k_reg: int = 20, knn: int = 10,
                 lam: float = 0.5, sigma: float = 1.0, symmetric: bool = False, sort: bool = False) -> Problem:
    W, pts = knn_weights(n, knn, 1001, symmetric, sort)
    X = np.random.default_rng(1002).standard_normal((n, k_reg))
    X[:, 0] = 1.0
    beta = np.random.default_rng(1003).standard_normal(k_reg)
    eps = sigma * np.random.default_rng(1004).standard_normal(n)
    if model in ("sar", "sdm"):
        y = neumann_solve(W, X @ beta + eps, lam)
    elif model in ("sem", "sdem"):
        y = X @ beta + neumann_solve(W, eps, lam)
    else:
        y = X @ beta + eps
    return Problem(y, X, W, beta, sigma, lam, model, pts)


def with_model(p: Problem, model: str) -> Problem:
    """The same W, X, beta and innovations as `p`, with y regenerated for another model (the kNN graph of a large
    problem is the expensive part of make_problem)."""
    n = len(p.y)
    eps = p.sigma * np.random.default_rng(1004).standard_normal(n)
    if model in ("sar", "sdm"):
        y = neumann_solve(p.W, p.X @ p.beta + eps, p.lam)
    elif model in ("sem", "sdem"):
        y = p.X @ p.beta + neumann_solve(p.W, eps, p.lam)
    else:
        y = p.X @ p.beta + eps
    return Problem(y, p.X, p.W, p.beta, p.sigma, p.lam, model, p.coords)
