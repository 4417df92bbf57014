"""ctypes binding of include/bsreg_cuda.h (the C ABI an R `.Call` shim binds the same way).

Loading is strict: if libbsreg_cuda.so is missing the import raises -- there is
no CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np

import os

# BSREG_CUDA_LIB selects another build of the same library (profiling variants); the default is the in-tree build
LIB_PATH = Path(os.environ.get("BSREG_CUDA_LIB") or Path(__file__).resolve().parent / "lib" / "libbsreg_cuda.so")

MODEL_LM, MODEL_SAR, MODEL_SEM = 0, 1, 2
PRIOR_INDEPENDENT, PRIOR_CONJUGATE, PRIOR_SHRINKAGE, PRIOR_HORSESHOE = 0, 1, 2, 3
LDET_NONE, LDET_NODES, LDET_CHEBYSHEV, LDET_SPLINE = 0, 1, 2, 3
STATS_STREAM, STATS_CACHED = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("y", c_dp), ("X", c_dp),
                ("k_slx", C.c_int32), ("X_slx", c_dp),
                ("nnz", C.c_int64), ("row_ptr", c_ip), ("col_idx", c_ip), ("val", c_dp),
                ("nnz_slx", C.c_int64), ("row_ptr_slx", c_ip), ("col_idx_slx", c_ip), ("val_slx", c_dp),
                ("n_nodes", C.c_int32), ("node_re", c_dp), ("node_im", c_dp), ("node_w", c_dp)]


class Options(C.Structure):
    _fields_ = [("model", C.c_int32), ("prior", C.c_int32), ("ldet", C.c_int32), ("stats_mode", C.c_int32),
                ("ng_mu", c_dp), ("ng_mu_len", C.c_int32), ("ng_precision", c_dp), ("ng_precision_len", C.c_int32),
                ("ng_shape", C.c_double), ("ng_rate", C.c_double), ("ng_beta0", c_dp), ("ng_sigma0", C.c_double),
                ("sng_lambda_a", C.c_double), ("sng_lambda_b", C.c_double), ("sng_theta_scale", C.c_double),
                ("sng_theta_a", C.c_double), ("sng_lambda", C.c_double), ("sng_tau", C.c_double),
                ("sng_theta", C.c_double),
                ("hs_lambda", C.c_double), ("hs_tau", C.c_double), ("hs_zeta", C.c_double), ("hs_nu", C.c_double),
                ("sp_lambda_a", C.c_double), ("sp_lambda_b", C.c_double), ("sp_lambda", C.c_double),
                ("sp_lambda_scale", C.c_double), ("sp_lambda_min", C.c_double), ("sp_lambda_max", C.c_double),
                ("cheb_order", C.c_int32), ("cheb_probes", C.c_int32),
                ("ldet_grid_from", C.c_double), ("ldet_grid_to", C.c_double),
                ("ldet_grid_len", C.c_int32), ("ldet_reps", C.c_int32),
                ("slx_psi", C.c_int32), ("slx_psi_norm", C.c_int32),
                ("slx_delta", C.c_double), ("slx_delta_scale", C.c_double), ("slx_delta_a", C.c_double),
                ("slx_delta_b", C.c_double), ("slx_delta_min", C.c_double), ("slx_delta_max", C.c_double),
                ("n_chains", C.c_int32), ("chain_offset", C.c_int32), ("seed", C.c_uint64),
                ("n_devices", C.c_int32), ("devices", c_ip), ("stream", C.c_void_p)]


class MH(C.Structure):
    _fields_ = [("adjust_burn", C.c_double), ("acc_lo", C.c_double), ("acc_hi", C.c_double),
                ("acc_change", C.c_double)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_int64, C.c_int64, C.c_void_p)

# every symbol include/bsreg_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bsreg_abi_version": (C.c_int, []),
    "bsreg_last_error": (C.c_char_p, []),
    "bsreg_create": (C.c_int, [C.POINTER(Problem), C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "bsreg_destroy": (None, [C.c_void_p]),
    "bsreg_dims": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip, c_ip, c_ip]),
    "bsreg_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(MH), C.c_void_p, C.c_void_p, C.c_void_p]),
    "bsreg_fetch_draws": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "bsreg_aux_len": (C.c_int, [C.c_void_p]),
    "bsreg_get_state": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_lp, c_lp]),
    "bsreg_set_state": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_lp, c_lp]),
    "bsreg_get_delta_state": (C.c_int, [C.c_void_p, c_dp, c_dp, c_lp]),
    "bsreg_set_delta_state": (C.c_int, [C.c_void_p, c_dp, c_dp, c_lp]),
    "bsreg_eval_spmm": (C.c_int, [C.c_void_p, c_dp, C.c_int32, c_dp]),
    "bsreg_eval_gram": (C.c_int, [C.c_void_p, c_dp]),
    "bsreg_eval_residual_ss": (C.c_int, [C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp]),
    "bsreg_eval_logdet": (C.c_int, [C.c_void_p, C.c_int32, c_dp, c_dp]),
    "bsreg_n_nodes": (C.c_int, [C.c_void_p]),
    "bsreg_get_nodes": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "bsreg_n_spline": (C.c_int, [C.c_void_p]),
    "bsreg_get_spline": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "bsreg_eval_trace_moments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, c_dp]),
    "bsreg_eval_conditionals": (C.c_int, [C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp,
                                          c_dp, c_dp, c_dp, c_dp]),
    "bsreg_eval_rng": (C.c_int, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp]),
    "bsreg_stream": (C.c_void_p, [C.c_void_p]),
    "bsreg_launch_sweep": (C.c_int, [C.c_void_p, C.c_int32]),
    "bsreg_launch_chain_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "bsreg_sync": (C.c_int, [C.c_void_p]),
    "bsreg_sweep_bytes": (C.c_int64, [C.c_void_p]),
    "bsreg_kernel_launches": (C.c_int64, [C.c_void_p]),
    "bsreg_release_cache": (None, []),
    "bsreg_debug_internal_order": (C.c_int, [C.c_int32, c_ip, c_ip, C.c_int32, c_ip]),
}

_lib = None


class BsregError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"bsreg_cuda error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Load the CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m bsreg_b200.build` "
                              "(the bsreg_b200 hot path has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise BsregError(rc, load().bsreg_last_error().decode())


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _f64(a, order="C"):
    return None if a is None else np.require(np.asarray(a, dtype=np.float64), requirements=[order, "A"])


class Handle:
    """Owns a bsreg_handle.  Arrays passed in are copied to the device by create()."""

    def __init__(self, y, X, *, model=MODEL_LM, prior=PRIOR_INDEPENDENT, W=None, X_slx=None, W_slx=None,
                 nodes=None, ldet=None, stats_mode=STATS_STREAM, n_chains=1, chain_offset=0, seed=0,
                 ng=None, sng=None, hs=None, spatial=None, cheb_order=50, cheb_probes=64,
                 ldet_grid=(-1.0, 1.0 - 1e-12, 100), ldet_reps=1, slx_psi=None, devices=None, stream=None):
        """slx_psi: dict(normalise="row"|"none", delta=, delta_scale=, delta_a=, delta_b=, delta_min=, delta_max=) makes
        W_slx (or W) a pattern of DISTANCES and Psi_SLX(delta) = dist ^ -delta a connectivity function."""
        lib = load()
        y = _f64(y).ravel()
        X = _f64(X, "F")
        if X.ndim == 1:
            X = X.reshape(-1, 1, order="F")
        n, k = X.shape
        prob = Problem()
        prob.n, prob.k, prob.y, prob.X = n, k, _dp(y), _dp(X)
        keep = [y, X]
        if X_slx is not None:
            X_slx = _f64(X_slx, "F").reshape(n, -1, order="F")
            X_slx = np.asfortranarray(X_slx)
            prob.k_slx, prob.X_slx = X_slx.shape[1], _dp(X_slx)
            keep.append(X_slx)

        def csr(Wm):
            rp = np.ascontiguousarray(Wm.indptr, dtype=np.int32)
            ci = np.ascontiguousarray(Wm.indices, dtype=np.int32)
            va = np.ascontiguousarray(Wm.data, dtype=np.float64)
            keep.extend([rp, ci, va])
            return len(va), rp.ctypes.data_as(c_ip), ci.ctypes.data_as(c_ip), _dp(va)
        if W is not None:
            prob.nnz, prob.row_ptr, prob.col_idx, prob.val = csr(W)
        if W_slx is not None:
            prob.nnz_slx, prob.row_ptr_slx, prob.col_idx_slx, prob.val_slx = csr(W_slx)
        if nodes is not None:
            re, im, w = (_f64(a).ravel() for a in nodes)
            keep.extend([re, im, w])
            prob.n_nodes, prob.node_re, prob.node_im, prob.node_w = len(re), _dp(re), _dp(im), _dp(w)

        ng = dict(mu=0.0, precision=1e-8, shape=0.01, rate=0.01, beta=None, sigma=None) | (ng or {})
        sng = dict(lambda_a=0.01, lambda_b=0.01, theta_scale=0.0, theta_a=1.0, lam=1.0, tau=10.0, theta=0.1) | (sng or {})
        hs = dict(lam=1.0, tau=1.0, zeta=1.0, nu=1.0) | (hs or {})
        sp_ = dict(lambda_a=1.01, lambda_b=1.01, lam=0.0, lambda_scale=0.1, lambda_min=-1.0,
                   lambda_max=1.0 - 1e-12) | (spatial or {})
        opt = Options()
        opt.model, opt.prior, opt.stats_mode = model, prior, stats_mode
        if ldet is None:
            ldet = LDET_NONE if model == MODEL_LM else (LDET_NODES if nodes is not None else LDET_CHEBYSHEV)
        opt.ldet = ldet
        mu = np.atleast_1d(_f64(ng["mu"]))
        prec = np.atleast_1d(_f64(ng["precision"]))
        keep.extend([mu, prec])
        opt.ng_mu, opt.ng_mu_len, opt.ng_precision, opt.ng_precision_len = _dp(mu), len(mu), _dp(prec), len(prec)
        opt.ng_shape, opt.ng_rate = ng["shape"], ng["rate"]
        if ng.get("beta") is not None:
            b0 = np.ascontiguousarray(np.broadcast_to(_f64(ng["beta"]), (k + prob.k_slx,)))
            keep.append(b0)
            opt.ng_beta0 = _dp(b0)
        opt.ng_sigma0 = math.nan if ng.get("sigma") is None else float(ng["sigma"])
        opt.sng_lambda_a, opt.sng_lambda_b = sng["lambda_a"], sng["lambda_b"]
        opt.sng_theta_scale, opt.sng_theta_a = sng["theta_scale"], sng["theta_a"]
        opt.sng_lambda, opt.sng_tau, opt.sng_theta = sng["lam"], sng["tau"], sng["theta"]
        opt.hs_lambda, opt.hs_tau, opt.hs_zeta, opt.hs_nu = hs["lam"], hs["tau"], hs["zeta"], hs["nu"]
        opt.sp_lambda_a, opt.sp_lambda_b, opt.sp_lambda = sp_["lambda_a"], sp_["lambda_b"], sp_["lam"]
        opt.sp_lambda_scale, opt.sp_lambda_min, opt.sp_lambda_max = sp_["lambda_scale"], sp_["lambda_min"], sp_["lambda_max"]
        opt.cheb_order, opt.cheb_probes = cheb_order, cheb_probes
        opt.ldet_grid_from, opt.ldet_grid_to, opt.ldet_grid_len = float(ldet_grid[0]), float(ldet_grid[1]), int(ldet_grid[2])
        opt.ldet_reps = int(ldet_reps)
        if slx_psi is not None:
            sp2 = dict(normalise="row", delta=1.0, delta_scale=0.0, delta_a=1.01, delta_b=1.01, delta_min=1e-12,
                       delta_max=math.inf) | slx_psi
            opt.slx_psi, opt.slx_psi_norm = 1, 1 if sp2["normalise"] == "row" else 0
            opt.slx_delta, opt.slx_delta_scale = sp2["delta"], sp2["delta_scale"]
            opt.slx_delta_a, opt.slx_delta_b = sp2["delta_a"], sp2["delta_b"]
            opt.slx_delta_min, opt.slx_delta_max = sp2["delta_min"], sp2["delta_max"]
        opt.n_chains, opt.chain_offset, opt.seed = n_chains, chain_offset, seed
        if devices is not None:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            keep.append(dv)
            opt.n_devices, opt.devices = len(dv), dv.ctypes.data_as(c_ip)
        if stream is not None:
            opt.stream = C.c_void_p(stream)
        self._h = C.c_void_p()
        check(lib.bsreg_create(C.byref(prob), C.byref(opt), C.byref(self._h)))
        self._lib = lib
        dims = [C.c_int32() for _ in range(5)]
        check(lib.bsreg_dims(self._h, *[C.byref(d) for d in dims]))
        self.n, self.m, self.P, self.n_chains, self.d = (d.value for d in dims)
        self.model, self.prior = model, prior
        self.spatial = model != MODEL_LM

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsreg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- sampling ---------------------------------------------------------------
    def run(self, n_burn, n_save, mh=None, out=None, keep_on_device=False, progress=None):
        """Returns draws of shape (n_chains, P, n_save): draws[c].T is the R matrix of chain c."""
        mhs = MH(*(mh or (0.8, 0.20, 0.45, 0.8)))
        cb = PROGRESS_FN(lambda done, total, _u: int(bool(progress(done, total)))) if progress else None
        ptr = None
        if not keep_on_device:
            if out is None:
                out = np.empty((self.n_chains, self.P, max(n_save, 0)), dtype=np.float64)
            assert out.shape == (self.n_chains, self.P, n_save) and out.flags.c_contiguous
            ptr = C.c_void_p(out.ctypes.data)
        check(self._lib.bsreg_run(self._h, n_burn, n_save, C.byref(mhs), ptr,
                                  C.cast(cb, C.c_void_p) if cb else None, None))
        return out

    def run_into(self, n_burn, n_save, mh, ptr):
        """bsreg_run into a raw host pointer (e.g. pinned torch memory); ptr=None keeps draws on device."""
        mhs = MH(*(mh or (0.8, 0.20, 0.45, 0.8)))
        check(self._lib.bsreg_run(self._h, n_burn, n_save, C.byref(mhs), C.c_void_p(ptr) if ptr else None, None, None))

    def fetch_draws(self, first, count):
        out = np.empty((self.n_chains, self.P, count))
        check(self._lib.bsreg_fetch_draws(self._h, first, count, C.c_void_p(out.ctypes.data)))
        return out

    def get_state(self):
        Cn, m, P, A = self.n_chains, self.m, self.P, self._lib.bsreg_aux_len(self._h)
        st = dict(params=np.empty((Cn, P)), p0=np.empty((Cn, m)), aux=np.empty((Cn, A)),
                  mh_value=np.empty((Cn, 2)), mh_scale=np.empty((Cn, 2)),
                  mh_counters=np.empty((Cn, 2, 4), dtype=np.int64), iterations=np.empty(Cn, dtype=np.int64))
        check(self._lib.bsreg_get_state(self._h, _dp(st["params"]), _dp(st["p0"]), _dp(st["aux"]), _dp(st["mh_value"]),
                                        _dp(st["mh_scale"]), st["mh_counters"].ctypes.data_as(c_lp),
                                        st["iterations"].ctypes.data_as(c_lp)))
        return st

    def get_delta_state(self):
        Cn = self.n_chains
        st = dict(delta=np.empty(Cn), scale=np.empty(Cn), counters=np.empty((Cn, 4), dtype=np.int64))
        check(self._lib.bsreg_get_delta_state(self._h, _dp(st["delta"]), _dp(st["scale"]), st["counters"].ctypes.data_as(c_lp)))
        return st

    def set_delta_state(self, st):
        a = {k: np.ascontiguousarray(v) for k, v in st.items()}
        check(self._lib.bsreg_set_delta_state(self._h, _dp(a["delta"]), _dp(a["scale"]), a["counters"].ctypes.data_as(c_lp)))

    def set_state(self, st):
        a = {k: np.ascontiguousarray(v) for k, v in st.items()}
        check(self._lib.bsreg_set_state(self._h, _dp(a["params"]), _dp(a["p0"]), _dp(a["aux"]), _dp(a["mh_value"]),
                                        _dp(a["mh_scale"]), a["mh_counters"].ctypes.data_as(c_lp),
                                        a["iterations"].ctypes.data_as(c_lp)))

    # -- parity entry points -------------------------------------------------------
    def eval_spmm(self, B):
        B = np.asfortranarray(_f64(B).reshape(self.n, -1))
        out = np.empty_like(B, order="F")
        check(self._lib.bsreg_eval_spmm(self._h, _dp(B), B.shape[1], _dp(out)))
        return out

    def eval_gram(self):
        G = np.empty((self.d, self.d))
        check(self._lib.bsreg_eval_gram(self._h, _dp(G)))
        return G

    def eval_residual_ss(self, beta, lam, rho):
        beta = _f64(beta).reshape(-1, self.m)
        ne = beta.shape[0]
        lam = np.ascontiguousarray(np.broadcast_to(_f64(lam), (ne,)))
        rho = np.ascontiguousarray(np.broadcast_to(_f64(rho), (ne,)))
        out = np.empty(ne)
        check(self._lib.bsreg_eval_residual_ss(self._h, ne, _dp(beta), _dp(lam), _dp(rho), _dp(out)))
        return out

    def eval_logdet(self, v):
        v = np.ascontiguousarray(np.atleast_1d(_f64(v)))
        out = np.empty(len(v))
        check(self._lib.bsreg_eval_logdet(self._h, len(v), _dp(v), _dp(out)))
        return out

    def nodes(self):
        nn = self._lib.bsreg_n_nodes(self._h)
        re, im, w = np.empty(nn), np.empty(nn), np.empty(nn)
        check(self._lib.bsreg_get_nodes(self._h, _dp(re), _dp(im), _dp(w)))
        return re, im, w

    def spline(self):
        """(x, y, b, c, d) of the grid + spline log-determinant (LDET_SPLINE)."""
        nn = self._lib.bsreg_n_spline(self._h)
        a = [np.empty(nn) for _ in range(5)]
        check(self._lib.bsreg_get_spline(self._h, *[_dp(v) for v in a]))
        return tuple(a)

    def eval_trace_moments(self, order, n_probes, seed):
        out = np.empty(order + 1)
        check(self._lib.bsreg_eval_trace_moments(self._h, order, n_probes, seed, _dp(out)))
        return out

    def eval_conditionals(self, beta, sigma2, lam, p0, v):
        beta = _f64(beta).reshape(-1, self.m)
        ne = beta.shape[0]
        p0 = np.ascontiguousarray(np.broadcast_to(_f64(p0), (ne, self.m)))
        s2, lam, v = (np.ascontiguousarray(np.broadcast_to(_f64(a), (ne,))) for a in (sigma2, lam, v))
        mu1, rate1, rss, lp = np.empty((ne, self.m)), np.empty(ne), np.empty(ne), np.empty(ne)
        check(self._lib.bsreg_eval_conditionals(self._h, ne, _dp(beta), _dp(s2), _dp(lam), _dp(p0), _dp(v),
                                                _dp(mu1), _dp(rate1), _dp(rss), _dp(lp)))
        return dict(mu1=mu1, rate1=rate1, rss=rss, logpost=lp)

    # -- measurement ---------------------------------------------------------------
    def stream(self):
        return self._lib.bsreg_stream(self._h)

    def launch_sweep(self, reps=1):
        check(self._lib.bsreg_launch_sweep(self._h, reps))

    def launch_chain_step(self, reps=1):
        check(self._lib.bsreg_launch_chain_step(self._h, reps))

    def sync(self):
        check(self._lib.bsreg_sync(self._h))

    def sweep_bytes(self):
        return self._lib.bsreg_sweep_bytes(self._h)

    def kernel_launches(self):
        return self._lib.bsreg_kernel_launches(self._h)


def eval_rng(kind, seed, chain, slot, it0, n, params=()):
    lib = load()
    p = np.zeros(4)
    p[:len(params)] = params
    out = np.empty(n)
    check(lib.bsreg_eval_rng(kind, seed, chain, slot, it0, n, _dp(p), _dp(out)))
    return out
