#' Probit BART S-learner on the GPU: the stand-in for bartCause::bartc + extract(type = "ite")
#'
#' The reference (fbargaglistoffi/BCF-IV) calls bartCause::bartc(response, treatment, confounders, n.samples, n.burn,
#' n.chains = 2L) for the complier proportion (R/bcf_iv.R:78-80) and, with binary = TRUE, for the intention-to-treat
#' effect itself (R/bcf_iv.R:198-202, R/bcf_itt.R:166-170), and keeps only apply(extract(fit, type = "ite"), 2, mean).
#' This function returns that vector, from libbcfgpu's probit-BART model (BCFGPU_MODEL_PROBIT_BART):
#'   1. treatment model P(z = 1 | x) = pnorm(f_z(x)) -> p.score (its posterior mean),
#'   2. response model P(response = 1 | x, p.score, z) = pnorm(f(x, p.score, z)),
#'   3. per draw pnorm(f(x, ps, 1)) - pnorm(f(x, ps, 0)), the chains pooled,
#' with dbarts' defaults (75 trees, k = 2, base .95, power 2, binaryOffset 0).  The response must be 0/1.
#' @export
bartc_ite <- function(response, treatment, confounders, n.samples = 500L, n.burn = 500L, n.chains = 2L,
                      seed = sample.int(.Machine$integer.max, 1), devices = NULL) {
  if (!all(response %in% c(0, 1))) stop("bartc_ite needs a 0/1 response")
  if (!all(treatment %in% c(0, 1))) stop("treatment must be a vector of 0's and 1's")
  confounders <- as.matrix(confounders); storage.mode(confounders) <- "double"
  ps <- .probit_bart(treatment, confounders, NULL, n.samples, n.burn, n.chains, seed, 100L, devices)$prob_mean
  .probit_bart(response, cbind(confounders, ps), treatment, n.samples, n.burn, n.chains, seed, 200L, devices)$ite_mean
}

# dbarts-style cutpoints: fewer than 8 distinct values -> midpoints; else 100 equally spaced points inside the range
.bart_cutpoints <- function(x, numcut = 100, cat_levels = 8) {
  ux <- sort(unique(x))
  if (length(ux) == 1) return(ux)
  if (length(ux) < cat_levels) return(ux[-length(ux)] + diff(ux) / 2)
  ux[1] + (ux[length(ux)] - ux[1]) * seq_len(numcut) / (numcut + 1)
}

.probit_bart <- function(y01, x, treatment, n.samples, n.burn, n.chains, seed, first_chain, devices,
                         n.trees = 75L, k = 2, base = 0.95, power = 2) {
  n <- length(y01)
  xc <- if (is.null(treatment)) x else cbind(x, as.double(treatment))
  storage.mode(xc) <- "double"
  z <- if (is.null(treatment)) integer(n) else as.integer(treatment)
  cuts <- lapply(seq_len(ncol(xc)), function(j) .bart_cutpoints(xc[, j]))
  cfg <- list(model = 1L, ntree_control = n.trees, ntree_moderate = 0L, base_control = base, power_control = power,
              con_sd = 3 / k, use_muscale = 0L, use_tauscale = 0L, sigma_init = 1, lambda = 1, seed = seed,
              chain = as.integer(first_chain), device = -1L,
              treatment_col = if (is.null(treatment)) -1L else ncol(xc) - 1L, probit_offset = 0)
  if (is.null(devices)) devices <- seq_len(min(.Call("bcfgpu_R_device_count"), n.chains)) - 1L
  empty <- matrix(0, n, 0)
  chains <- .Call("bcfgpu_R_fit_chains", as.double(y01), z, xc, empty,
                  as.double(unlist(cuts)), as.integer(c(0L, cumsum(lengths(cuts)))), double(0), 0L,
                  cfg, as.integer(n.burn), as.integer(n.samples), 1L, FALSE, as.integer(n.chains), as.integer(devices),
                  matrix(0, 0, 0), matrix(0, 0, 0))
  avg <- function(f) Reduce(`+`, lapply(chains, `[[`, f)) / length(chains)
  list(prob_mean = avg("mu_mean"), ite_mean = avg("tau_mean"))
}
