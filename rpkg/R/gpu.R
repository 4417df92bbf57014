# gpu.R -- the R side of the drop-in (WRITTEN TO SPEC, NOT RUN HERE: no R in the build image).
# Replaces get_b*()$setup + sample()/tune() (R/40_setup.R, R/50_execute.R) behind bm.formula (R/60_interface.R:38-66);
# set_options/set_mh/set_* and the print/summary/plot/as.mcmc methods of the reference stay as they are.
# Apply rpkg/patches/*.patch to the reference tree, copy this file to R/ and rpkg/src/* to src/ (see rpkg/README.md).

gpu_model <- function(y, X, type, options, Psi = NULL, X_SLX = NULL, Psi_SLX = NULL,
                      ldet_SAR = list(grid = FALSE, i_lambda = c(-1, 1 - 1e-12, 100L), reps = 1L),
                      ldet_SEM = ldet_SAR,
                      n_chains = 1L, seed = NULL, devices = NULL, stats = c("stream", "cached"), ...) {
  if(is.null(seed)) {seed <- sample.int(.Machine$integer.max, 1L)}
  stats <- match.arg(stats)
  model <- c(lm = 0L, slx = 0L, sar = 1L, sdm = 1L, sem = 2L, sdem = 2L)[[type]]
  prior <- c(Independent = 0L, Conjugate = 1L, Shrinkage = 2L, Horseshoe = 3L)[[options$type]]
  spatial <- if(model == 1L) "SAR" else if(model == 2L) "SEM" else NULL
  for(nm in c("SLX", spatial)) {
    if(options$priors[[nm]]$delta_scale > 0) {
      stop("Sampling the connectivity parameter 'delta' needs the CPU backend (options(bsreg.backend = 'r')).")
    }
  }
  csr <- function(W) { # dgRMatrix slots: p (row pointers), j (columns), x -- 0-based, what the C ABI takes
    if(is.null(W)) {return(NULL)}
    if(is.function(W)) {stop("Connectivity functions 'Psi' need the CPU backend (options(bsreg.backend = 'r')).")}
    W <- methods::as(methods::as(W, "generalMatrix"), "RsparseMatrix")
    list(p = W@p, j = W@j, x = W@x)
  }
  if(model > 0L && is.null(Psi)) {stop("Please provide a connectivity matrix 'Psi_", spatial, "'.")} # R/15_sar.R:52
  if(!is.null(X_SLX) && is.null(Psi_SLX)) {Psi_SLX <- Psi}                    # R/40_setup.R:99,120
  if(!is.null(X_SLX) && is.null(Psi_SLX)) {stop("Please provide a connecitivity function or matrix 'Psi_SLX'.")}
  W <- csr(if(model > 0L) Psi else Psi_SLX)
  W_slx <- if(model > 0L && !is.null(X_SLX) && !identical(Psi_SLX, Psi)) csr(Psi_SLX) else NULL
  ldet <- if(model == 2L) ldet_SEM else ldet_SAR
  reps <- if(is.null(ldet$reps)) 1L else as.integer(ldet$reps)
  nodes <- NULL; ldet_kind <- 0L; grid <- c(-1, 1 - 1e-12, 100)
  if(model > 0L) {
    if(isTRUE(ldet$grid)) {                                        # spline over a determinant grid, R/15_sar.R:79-87
      ldet_kind <- 3L; grid <- as.numeric(ldet$i_lambda)           # the dense LUs run on the device
    } else if(NROW(X) <= 6000L * reps) {                           # reference method, R/15_sar.R:68-76,91-96
      size <- NROW(X) / reps
      Wd <- as.matrix(Psi)
      ev <- eigen(Wd[seq(size), seq(size)], symmetric = isSymmetric(Wd), only.values = TRUE)$values
      nodes <- list(Re(ev), Im(ev), rep(as.numeric(reps), size)); ldet_kind <- 1L
    } else {ldet_kind <- 2L}                                       # Chebyshev / Hutchinson moments on the device
  }
  ptr <- .Call("C_bsreg_create", as.numeric(y), X, X_SLX, W, W_slx, nodes, model, prior, options$priors,
               c(ldet_kind, reps, as.integer(stats == "cached")), grid,
               as.integer(n_chains), as.numeric(seed), if(is.null(devices)) NULL else as.integer(devices))
  M <- NCOL(X) + if(is.null(X_SLX)) 0L else NCOL(X_SLX)
  names <- c(if(M == 1L) "beta" else paste0("beta", seq(M)), "sigma",
             if(model == 1L) "lambda_SAR", if(model == 2L) "lambda_SEM")  # R/10_lm.R:193, R/15_sar.R:138-142
  structure(list(ptr = ptr, names = names, n_chains = n_chains, type = type), class = "bsreg_gpu_model")
}

# sample(mdl, n_save, n_burn, mh, verbose): same signature and return value as R/50_execute.R:12-42
sample_gpu <- function(x, n_save = 1000L, n_burn = 0L, mh = set_mh(), verbose = TRUE) {
  if(verbose) {cat("Starting sampler with", n_burn, "burn-in and", n_save, "saved draws on the GPU.\n")}
  draws <- .Call("C_bsreg_run", x$ptr, as.integer(n_burn), as.integer(n_save), mh)   # n_save x P x n_chains
  out <- matrix(draws[, , 1L], nrow = dim(draws)[1L], dimnames = list(NULL, x$names))
  if(x$n_chains > 1L) {attr(out, "chains") <- draws}
  out
}

# print.bm shows `x$model$get_meta` (R/80_methods.R:5)
`$.bsreg_gpu_model` <- function(x, name) {
  if(name == "get_meta") {
    st <- .Call("C_bsreg_state", unclass(x)$ptr)
    return(list(iterations = st[[4]][1L]))
  }
  unclass(x)[[name]]
}
