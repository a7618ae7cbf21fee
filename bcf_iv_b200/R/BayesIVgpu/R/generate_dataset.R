#' Synthetic instrumental-variable data with heterogeneous complier effects
#'
#' Same arguments, distribution and return value as BayesIV::generate_dataset (reference R/generate_dataset.R:22):
#' binary covariates 1{N(0, Sigma) > 0} with equicorrelation rho, one-sided non-compliance, effect +effect_size in the
#' cell x1 = x2 = 0 and -effect_size in x1 = x2 = 1.  Like the reference it needs p >= 10 and returns the first ten
#' covariates.  `keep_all = TRUE` (an addition) returns all p columns, which the p = 50 / 100 / 200 benchmarks need.
#' @export
generate_dataset <- function(n = 1000, p = 10, rho = 0, null = 0, effect_size = 2, compliance = 0.75,
                             keep_all = FALSE) {
  if (p < 10) stop("generate_dataset needs p >= 10")
  Sigma <- matrix(rho, p, p); diag(Sigma) <- 1
  X <- (MASS::mvrnorm(n, rep(0, p), Sigma) > 0) * 1
  if (!keep_all) X <- X[, 1:10]
  colnames(X) <- paste0("x", seq_len(ncol(X)))
  complier <- rbinom(n, 1, compliance)
  y0 <- rnorm(n)
  effect <- ifelse(X[, 1] == 0 & X[, 2] == 0, effect_size, ifelse(X[, 1] == 1 & X[, 2] == 1, -effect_size, null))
  z <- rbinom(n, 1, 0.5)
  w <- z * complier
  list(y = y0 + w * effect, z = z, w = w, X = X)
}
