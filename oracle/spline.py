"""Grid + spline log-determinant of the reference (`ldet_SAR / ldet_SEM = list(grid = TRUE)`).

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).

Reference: R/15_sar.R:79-87 (= R/15_sem.R:91-99):
    pars  <- i_seq(i_lambda)                                   # R/00_aux.R:105-107: seq.int(from, to, length.out)
    ldets <- determinant(diag(size) - x * get_W(), logarithm = TRUE)$modulus * reps
    splinefun(pars, y = ldets)                                 # stats::splinefun, default method "fmm"

`stats::splinefun(method = "fmm")` is third-party arithmetic outside /root/reference (base R, unversioned by
DESCRIPTION beyond R >= 3.5.0).  It is the cubic spline of Forsythe, Malcolm and Moler, "Computer Methods for
Mathematical Computations" (1977), routines SPLINE / SEVAL: interior knots C2, and at each end the third derivative of
the spline equals that of the cubic through the four nearest data points.  Restated here from the published
algorithm; pinned by its defining properties in tests/test_oracle_cpu.py (interpolation, C2 continuity, the end
condition, exact reproduction of cubics).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def i_seq(i_lambda) -> np.ndarray:
    """R/00_aux.R:105-107: seq.int(x[1], x[2], length.out = x[3])."""
    lo, hi, n = float(i_lambda[0]), float(i_lambda[1]), int(i_lambda[2])
    if n == 1:
        return np.array([lo])
    return lo + (hi - lo) * np.arange(n) / (n - 1)          # seq.int's "from + (0:n1) * by" with by = (to - from) / n1


def fmm_coefficients(x, y):
    """(b, c, d) with s(u) = y_i + dx (b_i + dx (c_i + dx d_i)), dx = u - x_i on [x_i, x_{i+1}) -- FMM's SPLINE."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    b, c, d = np.zeros(n), np.zeros(n), np.zeros(n)
    if n < 2:
        raise ValueError("need at least two knots")
    if n < 3:
        b[:] = (y[1] - y[0]) / (x[1] - x[0])
        return b, c, d
    # tridiagonal system: b = diagonal, d = off-diagonal, c = right-hand side
    d[0] = x[1] - x[0]
    c[1] = (y[1] - y[0]) / d[0]
    for i in range(1, n - 1):
        d[i] = x[i + 1] - x[i]
        b[i] = 2.0 * (d[i - 1] + d[i])
        c[i + 1] = (y[i + 1] - y[i]) / d[i]
        c[i] = c[i + 1] - c[i]
    # end conditions: third derivatives at the ends from divided differences
    b[0] = -d[0]
    b[n - 1] = -d[n - 2]
    c[0] = c[n - 1] = 0.0
    if n > 3:
        c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
        c[n - 1] = c[n - 2] / (x[n - 1] - x[n - 3]) - c[n - 3] / (x[n - 2] - x[n - 4])
        c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
        c[n - 1] = -c[n - 1] * d[n - 2] * d[n - 2] / (x[n - 1] - x[n - 4])
    # forward elimination, back substitution
    for i in range(1, n):
        t = d[i - 1] / b[i - 1]
        b[i] = b[i] - t * d[i - 1]
        c[i] = c[i] - t * c[i - 1]
    c[n - 1] = c[n - 1] / b[n - 1]
    for i in range(n - 2, -1, -1):
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
    # polynomial coefficients
    b[n - 1] = (y[n - 1] - y[n - 2]) / d[n - 2] + d[n - 2] * (c[n - 2] + 2.0 * c[n - 1])
    for i in range(n - 1):
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
        d[i] = (c[i + 1] - c[i]) / d[i]
        c[i] = 3.0 * c[i]
    c[n - 1] = 3.0 * c[n - 1]
    d[n - 1] = d[n - 2]
    return b, c, d


def spline_eval(u: float, x, y, b, c, d, deriv: int = 0) -> float:
    """FMM's SEVAL: the piece of the largest knot <= u (first piece left of the grid, last knot's cubic right of it)."""
    n = len(x)
    i = int(np.searchsorted(x, u, side="right")) - 1
    i = min(max(i, 0), n - 1)
    dx = u - x[i]
    if deriv == 1:
        return b[i] + dx * (2.0 * c[i] + dx * 3.0 * d[i])
    return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


def grid_logdets(W, i_lambda=(-1.0, 1.0 - 1e-12, 100), reps: int = 1):
    """The determinant grid of R/15_sar.R:81-85: log|det(I - x W_sub)| * reps at the grid points (dense LU)."""
    Wd = np.asarray(W.todense()) if sp.issparse(W) else np.asarray(W, dtype=np.float64)
    size = Wd.shape[0] // reps
    Wsub = Wd[:size, :size]
    pars = i_seq(i_lambda)
    with np.errstate(divide="ignore"):
        ld = np.array([np.linalg.slogdet(np.eye(size) - t * Wsub)[1] * reps for t in pars])
    return pars, ld


class SplineLogdet:
    """get_ldet of the grid branch: splinefun(pars, ldets)(lambda)."""

    def __init__(self, W, i_lambda=(-1.0, 1.0 - 1e-12, 100), reps: int = 1, values=None):
        # values = (x, y): a determinant grid computed elsewhere (tests: the device's, so that both sides interpolate the
        # same numbers -- at a nearly singular grid point two LUs agree to ~1e-4 only)
        self.x, self.y = grid_logdets(W, i_lambda, reps) if values is None else (np.asarray(values[0]), np.asarray(values[1]))
        self.b, self.c, self.d = fmm_coefficients(self.x, self.y)

    def __call__(self, lam: float) -> float:
        return spline_eval(lam, self.x, self.y, self.b, self.c, self.d)
