/* shim.c -- thin .Call layer between R and the C ABI of include/bsreg_cuda.h.
 *
 * WRITTEN TO SPEC, NOT COMPILED HERE: this image has no R (no R.h / libR).  It is what a maintainer of
 * nk027/bsreg adds as src/shim.c (plus `useDynLib(bsreg, .registration = TRUE)` in NAMESPACE and
 * PKG_LIBS = -L<dir> -lbsreg_cuda in src/Makevars) to route sample()/tune() (R/50_execute.R:12-83) through
 * the GPU.  No C++ exceptions and no longjmp cross the ABI: every bsreg_* call returns a code, temporaries
 * are released first, then Rf_error() is raised.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Utils.h>
#include <string.h>
#include "bsreg_cuda.h"

static double num(SEXP list, const char *name, double dflt) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) {
      SEXP v = VECTOR_ELT(list, i);
      return Rf_isNull(v) ? dflt : Rf_asReal(v);
    }
  return dflt;
}

static void handle_finalizer(SEXP ptr) {
  bsreg_handle *h = (bsreg_handle *)R_ExternalPtrAddr(ptr);
  if (h) { bsreg_destroy(h); R_ClearExternalPtr(ptr); }
}

static SEXP elt(SEXP list, const char *name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

/* .Call("C_bsreg_create", y, X, X_slx, W, W_slx (lists p, j, x: dgRMatrix slots, 0-based; or NULL),
 *       nodes (list re, im, w or NULL), model, prior, priors (the `priors` list of set_options(), by name),
 *       ldet (kind, reps, cached), grid (from, to, length), n_chains, seed, devices (integer or NULL))                 */
SEXP C_bsreg_create(SEXP y, SEXP X, SEXP X_slx, SEXP W, SEXP W_slx, SEXP nodes, SEXP model, SEXP prior, SEXP priors,
                    SEXP ldet, SEXP grid, SEXP n_chains, SEXP seed, SEXP devices) {
  bsreg_problem p; bsreg_options o;
  memset(&p, 0, sizeof p); memset(&o, 0, sizeof o);
  p.n = Rf_nrows(X); p.k = Rf_ncols(X); p.y = REAL(y); p.X = REAL(X);              /* column-major, as R stores it */
  if (!Rf_isNull(X_slx)) { p.k_slx = Rf_ncols(X_slx); p.X_slx = REAL(X_slx); }
  if (!Rf_isNull(W)) {
    p.nnz = XLENGTH(elt(W, "x")); p.row_ptr = INTEGER(elt(W, "p")); p.col_idx = INTEGER(elt(W, "j")); p.val = REAL(elt(W, "x"));
  }
  if (!Rf_isNull(W_slx)) {                                                         /* separate Psi_SLX, R/40_setup.R:99 */
    p.nnz_slx = XLENGTH(elt(W_slx, "x")); p.row_ptr_slx = INTEGER(elt(W_slx, "p"));
    p.col_idx_slx = INTEGER(elt(W_slx, "j")); p.val_slx = REAL(elt(W_slx, "x"));
  }
  if (!Rf_isNull(nodes)) {
    p.n_nodes = (int32_t)XLENGTH(VECTOR_ELT(nodes, 0));
    p.node_re = REAL(VECTOR_ELT(nodes, 0)); p.node_im = REAL(VECTOR_ELT(nodes, 1)); p.node_w = REAL(VECTOR_ELT(nodes, 2));
  }
  o.model = Rf_asInteger(model); o.prior = Rf_asInteger(prior);
  o.ldet = INTEGER(ldet)[0]; o.ldet_reps = INTEGER(ldet)[1];
  o.stats_mode = INTEGER(ldet)[2] ? BSREG_STATS_CACHED : BSREG_STATS_STREAM;
  o.ldet_grid_from = REAL(grid)[0]; o.ldet_grid_to = REAL(grid)[1]; o.ldet_grid_len = (int32_t)REAL(grid)[2];
  SEXP NG = elt(priors, "NG"), SNG = elt(priors, "SNG"), HS = elt(priors, "HS");
  SEXP SP = elt(priors, o.model == BSREG_MODEL_SEM ? "SEM" : "SAR");               /* R/41_set_options.R:29-31 */
  SEXP mu = elt(NG, "mu"), prec = elt(NG, "precision"), beta0 = elt(NG, "beta");
  o.ng_mu = REAL(mu); o.ng_mu_len = (int32_t)XLENGTH(mu);
  o.ng_precision = REAL(prec); o.ng_precision_len = (int32_t)XLENGTH(prec);       /* a scalar or the diagonal */
  o.ng_shape = num(NG, "shape", 0.01); o.ng_rate = num(NG, "rate", 0.01); o.ng_sigma0 = num(NG, "sigma", R_NaN);
  if (!Rf_isNull(beta0)) {                                                         /* starting beta, R/10_lm.R:158-160 */
    if (XLENGTH(beta0) != p.k + p.k_slx) Rf_error("Please provide a valid starting value for 'beta' (one per coefficient).");
    o.ng_beta0 = REAL(beta0);
  }
  o.sng_lambda_a = num(SNG, "lambda_a", 0.01); o.sng_lambda_b = num(SNG, "lambda_b", 0.01);
  o.sng_theta_scale = num(SNG, "theta_scale", 0); o.sng_theta_a = num(SNG, "theta_a", 1);
  o.sng_lambda = num(SNG, "lambda", 1); o.sng_tau = num(SNG, "tau", 10); o.sng_theta = num(SNG, "theta", 0.1);
  o.hs_lambda = num(HS, "lambda", 1); o.hs_tau = num(HS, "tau", 1); o.hs_zeta = num(HS, "zeta", 1); o.hs_nu = num(HS, "nu", 1);
  o.sp_lambda_a = num(SP, "lambda_a", 1.01); o.sp_lambda_b = num(SP, "lambda_b", 1.01); o.sp_lambda = num(SP, "lambda", 0);
  o.sp_lambda_scale = num(SP, "lambda_scale", 0.1); o.sp_lambda_min = num(SP, "lambda_min", -1);
  o.sp_lambda_max = num(SP, "lambda_max", 1 - 1e-12);
  o.cheb_order = 50; o.cheb_probes = 64;
  o.n_chains = Rf_asInteger(n_chains); o.seed = (uint64_t)Rf_asReal(seed);
  if (!Rf_isNull(devices)) { o.n_devices = (int32_t)XLENGTH(devices); o.devices = INTEGER(devices); }   /* chains split over GPUs */
  bsreg_handle *h = NULL;
  if (bsreg_create(&p, &o, &h) != BSREG_OK) Rf_error("bsreg_cuda: %s", bsreg_last_error());
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);                             /* GC owns the handle */
  UNPROTECT(1);
  return ptr;
}

static int progress_cb(int64_t done, int64_t total, void *user) {
  /* R_CheckUserInterrupt longjmps; run it under R_ToplevelExec so that the handle survives an interrupt */
  (void)done; (void)total; (void)user;
  return R_ToplevelExec((void (*)(void *))R_CheckUserInterrupt, NULL) ? 0 : 1;
}

/* .Call("C_bsreg_run", ptr, n_burn, n_save, mh) -> numeric array (n_save, P, n_chains); [, , 1] is the reference's `draws` */
SEXP C_bsreg_run(SEXP ptr, SEXP n_burn, SEXP n_save, SEXP mh) {
  bsreg_handle *h = (bsreg_handle *)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("bsreg_cuda: model handle was released");
  int32_t n = 0, m = 0, P = 0, C = 0, d = 0;
  bsreg_dims(h, &n, &m, &P, &C, &d);
  SEXP target = VECTOR_ELT(mh, 1);
  bsreg_mh s = { num(mh, "adjust_burn", 0.8), REAL(target)[0], REAL(target)[1], num(mh, "acc_change", 0.8) };
  const int ns = Rf_asInteger(n_save);
  SEXP out = PROTECT(Rf_alloc3DArray(REALSXP, ns, P, C));                          /* caller-allocated, GC-owned */
  const int rc = bsreg_run(h, Rf_asInteger(n_burn), ns, &s, REAL(out), progress_cb, NULL);
  UNPROTECT(1);
  if (rc != BSREG_OK) Rf_error("bsreg_cuda: %s", bsreg_last_error());
  return out;
}

/* .Call("C_bsreg_state", ptr) -> list(params (P x C), mh_scale (2 x C), mh_counters (4 x 2 x C), iterations (C)) */
SEXP C_bsreg_state(SEXP ptr) {
  bsreg_handle *h = (bsreg_handle *)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("bsreg_cuda: model handle was released");
  int32_t n = 0, m = 0, P = 0, C = 0, d = 0;
  bsreg_dims(h, &n, &m, &P, &C, &d);
  SEXP params = PROTECT(Rf_allocMatrix(REALSXP, P, C)), scale = PROTECT(Rf_allocMatrix(REALSXP, 2, C));
  SEXP cnt = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)8 * C)), iters = PROTECT(Rf_allocVector(REALSXP, C));
  int64_t *c64 = (int64_t *)R_alloc((size_t)8 * C, sizeof(int64_t)), *i64 = (int64_t *)R_alloc(C, sizeof(int64_t));
  if (bsreg_get_state(h, REAL(params), NULL, NULL, NULL, REAL(scale), c64, i64) != BSREG_OK) {
    UNPROTECT(4);
    Rf_error("bsreg_cuda: %s", bsreg_last_error());
  }
  for (int i = 0; i < 8 * C; ++i) REAL(cnt)[i] = (double)c64[i];
  for (int i = 0; i < C; ++i) REAL(iters)[i] = (double)i64[i];
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(res, 0, params); SET_VECTOR_ELT(res, 1, scale); SET_VECTOR_ELT(res, 2, cnt); SET_VECTOR_ELT(res, 3, iters);
  UNPROTECT(5);
  return res;
}

static const R_CallMethodDef call_methods[] = {
  {"C_bsreg_create", (DL_FUNC)&C_bsreg_create, 14},
  {"C_bsreg_run", (DL_FUNC)&C_bsreg_run, 4},
  {"C_bsreg_state", (DL_FUNC)&C_bsreg_state, 1},
  {NULL, NULL, 0}};

void R_init_bsreg(DllInfo *dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
