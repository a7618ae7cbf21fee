#' Bayesian Causal Forest for the intention-to-treat effect (GPU sampler)
#'
#' Drop-in for BayesIV::bcf_itt (reference R/bcf_itt.R:33): same arguments; prints the node effects and returns
#' NULL invisibly, as the reference does.
#' @export
bcf_itt <- function(y, w, z, x, binary = FALSE, max_depth = 2, n_burn = 500, n_sim = 500, inference_ratio = 0.5,
                    seed = 42) {
  # binary = TRUE: the reference fits the ITT with bartCause::bartc (probit BART, R/bcf_itt.R:166-170): bartc_ite()
  if (binary && !all(y %in% c(0, 1))) stop("binary = TRUE needs a 0/1 outcome")
  x <- as.matrix(x)
  index <- .honest_split(nrow(x), inference_ratio, seed)
  inference <- data.frame(y = y, w = w, z = z, x)[index, ]
  xd <- x[-index, , drop = FALSE]
  pihat <- .logit_pihat(z[-index], xd)
  tauhat <- if (binary) bartc_ite(y[-index], z[-index], xd, n.samples = n_sim, n.burn = n_burn, n.chains = 2L, seed = seed)
            else bcf(y[-index], z[-index], xd, xd, pihat, nburn = n_burn, nsim = n_sim, random_seed = seed,
                     keep_draws = FALSE)$tau_mean
  tree <- rpart::rpart(tauhat ~ ., data = data.frame(tauhat = tauhat, xd), maxdepth = max_depth)
  for (i in as.numeric(row.names(tree$frame))) {
    rule <- if (i == 1) NULL else .node_rule(tree, i)
    sub <- if (i == 1) inference else inference[which(eval(parse(text = rule), inference)), ]
    if (i != 1 && length(unique(sub$w)) == 1 && length(unique(sub$z)) == 1) next
    r <- .node_iv(sub)
    if (i == 1) cat("The effect on the overall sample is", round(r[["effect"]], 4), "\n")
    else cat("The effect on the subpopulation", rule, "is", round(r[["effect"]], 4), "\n")
    cat("P-value", r[["p"]], "\n")
    cat("P-value Weak-Instrument Test", r[["weak"]], "\n")
    cat("Proportion of observations in the node: ", format(round(nrow(sub) / nrow(inference), 2), nsmall = 2), "\n")
  }
  invisible(NULL)
}
