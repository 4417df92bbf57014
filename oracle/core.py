"""Deterministic pieces of the bsreg hot path, restated in NumPy/SciPy float64.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned -- no reference
tests exist).  Every function cites the reference lines it follows; paths are
relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
from scipy.special import betaln, gammaln

from .philox import MASK32, Slot, philox4x32_vec


# ---------------------------------------------------------------------------
# spatial lags (K1)
# ---------------------------------------------------------------------------
def as_csr(W) -> sp.csr_matrix:
    """The reference requires a dense base matrix (R/12_slx.R:61, R/15_sem.R:29);
    CSR is the sparse-W extension.  Explicit zeros are dropped and
    column indices sorted so that the CSR handed to the device is canonical."""
    W = sp.csr_matrix(W, dtype=np.float64)
    W.sum_duplicates()
    W.eliminate_zeros()
    W.sort_indices()
    return W


def spmm(W: sp.csr_matrix, B: np.ndarray) -> np.ndarray:
    """`W %*% y`, `W %*% X`: R/15_sar.R:57, R/15_sem.R:63-64, R/12_slx.R:65."""
    return np.asarray(W @ B)


# ---------------------------------------------------------------------------
# quadratic forms (K2)
# ---------------------------------------------------------------------------
def _sum(x) -> float:
    """R's sum() accumulates in long double (R/00_aux.R:118); so does this."""
    return float(np.sum(x, dtype=np.longdouble))


def sq_sum(x) -> float:
    """R/00_aux.R:117-119."""
    x = np.asarray(x, dtype=np.float64).ravel()
    return _sum(x * x)


def residual_ss(y, X, Wy, WX, beta, lam, rho) -> float:
    """|| (y - lam Wy) - (X - rho WX) beta ||^2.

    LM:  lam = rho = 0            R/10_lm.R:192, :176
    SAR: rho = 0, z = y - lam Wy  R/15_sar.R:29, :132
    SEM: lam = rho                R/15_sem.R:29-32, :136-137"""
    ys = y - lam * Wy if lam != 0.0 else y
    Xs = X - rho * WX if rho != 0.0 else X
    return sq_sum(ys - Xs @ beta)


def gram(y, X, Wy=None, WX=None) -> np.ndarray:
    """Z'Z for Z = [y | Wy | X | WX] (blocks present only if given): the
    chain-independent sufficient statistics behind R/10_lm.R:36 (XX, Xy),
    R/15_sar.R:133-135 (X'z, X'Wy) and R/15_sem.R:34-35 (filtered XX, Xy)."""
    cols = [y.reshape(-1, 1)]
    if Wy is not None:
        cols.append(Wy.reshape(-1, 1))
    cols.append(X)
    if WX is not None:
        cols.append(WX)
    Z = np.hstack(cols).astype(np.longdouble)
    return np.asarray(Z.T @ Z, dtype=np.float64)


# ---------------------------------------------------------------------------
# log-determinant (K3, K4)
# ---------------------------------------------------------------------------
def eigen_nodes(W, reps: int = 1):
    """R/15_sar.R:68-76, :91-92 (= R/15_sem.R:80-88, :103-104): eigenvalues of the
    leading N/reps block, symmetric solver iff W is symmetric.  Returns the
    node form (re, im, weight) with weight = reps."""
    W = np.asarray(W.todense()) if sp.issparse(W) else np.asarray(W, dtype=np.float64)
    size = W.shape[0] // reps
    Wsub = W[:size, :size]
    if np.allclose(Wsub, Wsub.T, rtol=0, atol=100 * np.finfo(float).eps):  # isSymmetric tolerance
        ev = np.linalg.eigvalsh(Wsub).astype(np.complex128)
    else:
        ev = np.linalg.eigvals(Wsub).astype(np.complex128)
    return ev.real.copy(), ev.imag.copy(), np.full(size, float(reps))


def ldet_nodes(lam: float, re, im, w) -> float:
    """Re(sum(log(1 - lam*ev))) * reps, R/15_sar.R:94-96, R/15_sem.R:106-108,
    written on nodes: sum_k w_k * 0.5*log((1 - lam re_k)^2 + (lam im_k)^2)."""
    a = 1.0 - lam * np.asarray(re)
    b = lam * np.asarray(im)
    return _sum(np.asarray(w) * 0.5 * np.log(a * a + b * b))


def ldet_dense(lam: float, W) -> float:
    """`determinant(diag(n) - x * W, logarithm = TRUE)$modulus`, R/15_sar.R:83."""
    W = np.asarray(W.todense()) if sp.issparse(W) else np.asarray(W)
    return float(np.linalg.slogdet(np.eye(W.shape[0]) - lam * W)[1])


def rademacher_probes(n: int, n_probes: int, seed: int) -> np.ndarray:
    """n x n_probes matrix of +-1 from Philox: call (c0=row, c1=Slot.PROBE,
    c2=probe block of 128, c3=0xffffffff) -> bit j of word w is probe 128b+32w+j."""
    U = np.empty((n, n_probes), dtype=np.float64)
    rows = np.arange(n, dtype=np.uint64)
    for b in range((n_probes + 127) // 128):
        words = philox4x32_vec(rows, Slot.PROBE, b, MASK32, seed & MASK32, (seed >> 32) & MASK32)
        for w in range(4):
            for j in range(32):
                p = 128 * b + 32 * w + j
                if p < n_probes:
                    bit = (words[w] >> np.uint64(j)) & np.uint64(1)
                    U[:, p] = 1.0 - 2.0 * bit.astype(np.float64)
    return U


def trace_moments(W: sp.csr_matrix, order: int, n_probes: int, seed: int) -> np.ndarray:
    """m_j ~ tr T_j(W), j = 0..order (not in the reference).
    m_0 = N, m_1 = tr W, m_2 = 2 tr(W^2) - N are exact; higher orders use
    Hutchinson probes through the three-term recurrence T_{j+1}U = 2 W T_j U - T_{j-1} U."""
    n = W.shape[0]
    m = np.zeros(order + 1)
    m[0] = n
    if order >= 1:
        m[1] = W.diagonal().sum()
    if order >= 2:
        m[2] = 2.0 * float(W.multiply(W.T).sum()) - n
    if order >= 3:
        U = rademacher_probes(n, n_probes, seed)
        Tm, T = U, np.asarray(W @ U)
        for j in range(2, order + 1):
            Tm, T = T, 2.0 * np.asarray(W @ T) - Tm
            if j >= 3:
                m[j] = float(np.sum(U * T)) / n_probes
    return m


def chebyshev_nodes(moments: np.ndarray, n: int):
    """Fold the moments into node form:  log|I - lam W| ~ sum_k w_k log(1 - lam x_k)
    with x_k = cos(pi (k-1/2)/(q+1)) and
    w_k = 2/(q+1) * ( N/2 + sum_{j>=1} m_j cos(pi j (k-1/2)/(q+1)) ),
    which is algebraically the Pace-LeSage Chebyshev form of Appendix B."""
    q = len(moments) - 1
    k = np.arange(1, q + 2)
    theta = math.pi * (k - 0.5) / (q + 1)
    x = np.cos(theta)
    j = np.arange(1, q + 1)
    w = 2.0 / (q + 1) * (0.5 * n + np.cos(np.outer(theta, j)) @ moments[1:])
    return x, np.zeros_like(x), w


# ---------------------------------------------------------------------------
# conditional pieces of the Gibbs sweep
# ---------------------------------------------------------------------------
def log_prior_lambda(v: float, a: float, b: float) -> float:
    """dbeta((v+1)/2, a, b, log=TRUE) - log(2), R/20_mh.R:197-199."""
    x = 0.5 * (v + 1.0)
    return (a - 1.0) * math.log(x) + (b - 1.0) * math.log1p(-x) - betaln(a, b) - math.log(2.0)


def chol_solve(A, b):
    L = np.linalg.cholesky(A)
    return np.linalg.solve(L.T, np.linalg.solve(L, b))


def beta_moments(XX, Xy, sigma2, p0, mu0, conjugate=False):
    """P1, mu1 of R/10_lm.R:168-169 (independent) or :223-224 (conjugate, no sigma^2)."""
    s = 1.0 if conjugate else sigma2
    P1 = XX / s + np.diag(p0)
    mu1 = chol_solve(P1, Xy / s + p0 * mu0)
    return P1, mu1


def sar_rss_coeffs(y, Wy, X, XX, sigma2, p0, mu0):
    """get_rss closure of R/15_sar.R:104-116: returns (e0e0, e1e0, e1e1) so that
    rss(v) = e0e0 - 2 v e1e0 + v^2 e1e1.  Uses RAW y, current sigma^2 and P0."""
    A = np.diag(p0) + XX / sigma2
    b0 = chol_solve(A, p0 * mu0 + (X.T @ y) / sigma2)
    b1 = chol_solve(A, p0 * mu0 + (X.T @ Wy) / sigma2)
    e0 = y - X @ b0
    e1 = Wy - X @ b1
    return _sum(e0 * e0), _sum(e1 * e0), _sum(e1 * e1)


def sem_rss(v, y, Wy, X, WX):
    """get_rss of R/15_sem.R:116-120: ||y_s||^2 - ||Q(X_s)' y_s||^2 (Householder QR)."""
    ys = y - v * Wy
    Xs = X - v * WX
    Q, _ = np.linalg.qr(Xs)
    return sq_sum(ys) - sq_sum(Q.T @ ys)


def logpost_lambda(v, ldet, rss, n, m, a, b):
    """posterior of R/20_mh.R:192-195 (SAR and, via the alias at :218, SEM)."""
    if not rss > 0.0:
        return float("nan")
    return ldet - 0.5 * (n - m) * math.log(rss) + log_prior_lambda(v, a, b)


def logpost_theta(theta, tau, lam, rate):
    """posterior of MH_SNG_theta, R/20_mh.R:133-140:
    sum dgamma(tau; theta, rate = theta*lam/2, log) + dexp(theta; rate, log)."""
    r = theta * lam / 2.0
    tau = np.asarray(tau)
    return float(np.sum(theta * math.log(r) - gammaln(theta) + (theta - 1.0) * np.log(tau) - r * tau)
                 + math.log(rate) - rate * theta)
