#' Bayesian Causal Forest with Instrumental Variables (GPU sampler)
#'
#' Drop-in for BayesIV::bcf_iv (reference R/bcf_iv.R:47): same arguments, same returned data.frame
#' (node, CCACE, pvalue, Weak_IV_test, Pi_obs, ITT, Pi_compliers, Adj_pvalue; one row per rpart node number).
#' The ITT forest is fitted by bcf() of this package and the complier proportion (and, with binary = TRUE, the ITT) by
#' bartc_ite() -- both on the GPU; everything else is unchanged CPU code.
#' @export
bcf_iv <- function(y, w, z, x, binary = FALSE, n_burn = 500, n_sim = 500, inference_ratio = 0.5, max_depth = 2,
                   cp = 0.01, minsplit = 10, adj_method = "holm", seed = 42) {
  # the reference's second sampler, bartCause::bartc (probit BART S-learner), is bartc_ite() of this package: the complier
  # proportion of every call (R/bcf_iv.R:78-80) and, with binary = TRUE, the ITT itself (R/bcf_iv.R:198-202)
  if (binary && !all(y %in% c(0, 1))) stop("binary = TRUE needs a 0/1 outcome")
  x <- as.matrix(x)
  index <- .honest_split(nrow(x), inference_ratio, seed)
  all_rows <- data.frame(y = y, w = w, z = z, x)
  inference <- all_rows[index, ]
  xd <- x[-index, , drop = FALSE]
  pihat <- .logit_pihat(z[-index], xd)

  # the fits are independent samplers, as in the reference (bartc for pic, bcf for the ITT): each has its own Philox
  # streams, so the Monte Carlo noise of itt and pic is not shared
  fit <- function(outcome, chain) bcf(outcome, z[-index], xd, xd, pihat, nburn = n_burn, nsim = n_sim,
                                      random_seed = seed, keep_draws = FALSE, first_chain = chain)$tau_mean
  pic <- bartc_ite(w[-index], z[-index], xd, n.samples = n_sim, n.burn = n_burn, n.chains = 2L, seed = seed)
  itt <- if (binary) bartc_ite(y[-index], z[-index], xd, n.samples = n_sim, n.burn = n_burn, n.chains = 2L, seed = seed + 1L)
         else fit(y[-index], 0L)              # posterior mean intention-to-treat effect given x  (the hot path, R/bcf_iv.R:89)
  tauhat <- itt / pic

  tree <- rpart::rpart(tauhat ~ ., data = data.frame(tauhat = tauhat, xd), maxdepth = max_depth, cp = cp,
                       minsplit = minsplit)
  nodes <- as.numeric(row.names(tree$frame))
  # Rows are indexed by rpart NODE NUMBER, as the reference does (bcfivMat[i, ] with i the node number, R/bcf_iv.R:174;
  # root in row 1, R/bcf_iv.R:138): the table starts with length(nodes) rows and grows (NA rows in between) when a
  # node number exceeds that, exactly like the reference's data.frame assignment.
  out <- data.frame(node = rep(NA_character_, length(nodes)), CCACE = NA, pvalue = NA, Weak_IV_test = NA,
                    Pi_obs = NA, ITT = NA, Pi_compliers = NA)
  leaves <- numeric(length(nodes))
  leaves[nodes[tree$frame$var == "<leaf>"]] <- 1                        # R/bcf_iv.R:118-120
  for (i in nodes) {
    rule <- if (i == 1) NA_character_ else .node_rule(tree, i)
    sub <- if (i == 1) inference else inference[which(eval(parse(text = rule), inference)), ]
    if (i != 1 && length(unique(sub$w)) == 1 && length(unique(sub$z)) == 1) next      # R/bcf_iv.R:157
    r <- .node_iv(sub)
    out[i, ] <- list(rule, round(r[["effect"]], 4), round(r[["p"]], 4), round(r[["weak"]], 4),
                     round(nrow(sub) / nrow(inference), 4), round(r[["effect"]] * r[["compliers"]], 4),
                     round(r[["compliers"]], 4))
  }
  nr <- max(nrow(out), length(leaves))
  if (nrow(out) < nr) out[nr, ] <- NA
  length(leaves) <- nr
  is_leaf <- which(leaves == 1)
  out$Adj_pvalue <- NA
  out$Adj_pvalue[is_leaf] <- round(p.adjust(as.numeric(out$pvalue[is_leaf]), adj_method), 5)     # R/bcf_iv.R:182-185
  out
}
