"""Convergence diagnostics for the multi-chain draws (host side, NumPy only).

The reference leaves diagnostics to coda (`as.mcmc.bm`, R/81_coda.R:17-30; `geweke.diag`, `HPDinterval` in
paper/9_paper.R:34-41).  With many chains per fit the cross-chain versions are the natural ones; these are the
textbook definitions (Gelman et al. BDA3 for split R-hat and the autocorrelation ESS; Geweke 1992 with a
batch-means spectral estimate), used by the tests to compare the GPU chains with the CPU oracle's chains.
"""
from __future__ import annotations

import numpy as np


def split_rhat(chains: np.ndarray) -> np.ndarray:
    """chains: (C, n, P) -> split-R-hat per parameter (each chain cut in halves)."""
    C, n, P = chains.shape
    h = n // 2
    x = np.concatenate([chains[:, :h], chains[:, h:2 * h]], axis=0)           # (2C, h, P)
    w = x.var(axis=1, ddof=1).mean(axis=0)
    b = h * x.mean(axis=1).var(axis=0, ddof=1)
    return np.sqrt(((h - 1) / h * w + b / h) / w)


def ess(chains: np.ndarray) -> np.ndarray:
    """Effective sample size per parameter from the chain-averaged autocorrelations (Geyer's initial
    positive sequence truncation)."""
    C, n, P = chains.shape
    out = np.empty(P)
    xc = chains - chains.mean(axis=1, keepdims=True)
    var = xc.var(axis=1).mean(axis=0)
    for p in range(P):
        f = np.fft.rfft(xc[:, :, p], 2 * n, axis=1)
        acov = np.fft.irfft(f * np.conj(f), axis=1)[:, :n].mean(axis=0) / n
        rho = acov / max(var[p], 1e-300)
        s, t = 0.0, 1
        while t + 1 < n:
            pair = rho[t] + rho[t + 1]
            if pair < 0:
                break
            s += pair
            t += 2
        out[p] = C * n / (1.0 + 2.0 * s)
    return out


def geweke_z(chain: np.ndarray, first=0.1, last=0.5, batches=20) -> np.ndarray:
    """Geweke z-score per parameter of one chain (n, P): mean of the first 10 % vs the last 50 %, variances
    from batch means."""
    n = chain.shape[0]
    a, b = chain[: int(first * n)], chain[n - int(last * n):]

    def mean_var(x):
        k = max(2, min(batches, x.shape[0] // 5))
        m = x[: x.shape[0] // k * k].reshape(k, -1, x.shape[1]).mean(axis=1)
        return x.mean(axis=0), m.var(axis=0, ddof=1) / k
    ma, va = mean_var(a)
    mb, vb = mean_var(b)
    return (ma - mb) / np.sqrt(va + vb)


def summarize(chains: np.ndarray) -> dict:
    """Posterior mean, sd, Monte-Carlo standard error, split R-hat and ESS per parameter."""
    e = ess(chains)
    sd = chains.reshape(-1, chains.shape[2]).std(axis=0, ddof=1)
    return dict(mean=chains.mean(axis=(0, 1)), sd=sd, mcse=sd / np.sqrt(e), rhat=split_rhat(chains), ess=e)
