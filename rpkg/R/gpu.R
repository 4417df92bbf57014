# gpu.R -- the R side of the drop-in (WRITTEN TO SPEC, NOT RUN HERE: no R in the build image).
# Replaces get_b*()$setup + sample()/tune() (R/40_setup.R, R/50_execute.R) behind bm.formula (R/60_interface.R:38-66);
# set_options/set_mh/set_* and the print/summary/plot/as.mcmc methods of the reference stay as they are.

gpu_model <- function(y, X, type, options, Psi = NULL, X_SLX = NULL, ldet = list(reps = 1L),
                      n_chains = 1L, seed = sample.int(.Machine$integer.max, 1L)) {
  model <- c(lm = 0L, slx = 0L, sar = 1L, sdm = 1L, sem = 2L, sdem = 2L)[[type]]
  prior <- c(Independent = 0L, Conjugate = 1L, Shrinkage = 2L, Horseshoe = 3L)[[options$type]]
  Wp <- Wj <- Wx <- nodes <- NULL
  if(!is.null(Psi)) {
    W <- methods::as(Psi, "RsparseMatrix")                       # dgRMatrix: p (row pointers), j, x -- 0-based
    Wp <- W@p; Wj <- W@j; Wx <- W@x
    if(model > 0L && NROW(X) <= 6000L * ldet$reps) {             # reference method, R/15_sar.R:68-76,91-96
      size <- NROW(X) / ldet$reps
      ev <- eigen(as.matrix(W[seq(size), seq(size)]), symmetric = isSymmetric(as.matrix(W)), only.values = TRUE)$values
      nodes <- list(Re(ev), Im(ev), rep(as.numeric(ldet$reps), size))
    }                                                            # otherwise: Chebyshev moments on the device
  }
  ptr <- .Call("C_bsreg_create", as.numeric(y), X, X_SLX, Wp, Wj, Wx, nodes, model, prior, options$priors,
               as.integer(n_chains), as.numeric(seed))
  M <- NCOL(X) + if(is.null(X_SLX)) 0L else NCOL(X_SLX)
  names <- c(if(M == 1L) "beta" else paste0("beta", seq(M)), "sigma",
             if(model == 1L) "lambda_SAR", if(model == 2L) "lambda_SEM")
  structure(list(ptr = ptr, names = names, n_chains = n_chains), class = "bsreg_gpu_model")
}

# sample(mdl, n_save, n_burn, mh, verbose): same signature and return value as R/50_execute.R:12-42
sample.bsreg_gpu_model <- function(x, n_save = 1000L, n_burn = 0L, mh = set_mh(), verbose = TRUE) {
  draws <- .Call("C_bsreg_run", x$ptr, as.integer(n_burn), as.integer(n_save), mh)
  out <- draws[, , 1L, drop = TRUE]; dim(out) <- dim(draws)[1:2]
  colnames(out) <- x$names
  if(x$n_chains > 1L) attr(out, "chains") <- draws
  out
}
