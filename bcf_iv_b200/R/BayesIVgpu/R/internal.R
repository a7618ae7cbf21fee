# Shared pieces of bcf_iv() and bcf_itt(): the reference repeats them in both files (R/bcf_iv.R, R/bcf_itt.R of
# fbargaglistoffi/BCF-IV); here they are written once.

# honest split: `index` are the INFERENCE rows (reference R/bcf_iv.R:56-57), everything else is discovery
.honest_split <- function(n, inference_ratio, seed) {
  set.seed(seed)
  sample(n, n * inference_ratio, replace = FALSE)
}

# propensity of the instrument on the covariates, on the LINK scale (predict.glm's default; reference R/bcf_iv.R:72-75):
# glm.fit's IRLS on the device (bcfgpu_glm_logit; at n = 5e5, p = 100 stats::glm takes seconds, this 30 ms)
.logit_pihat <- function(z, x, device = 0L) {
  x <- as.matrix(x); storage.mode(x) <- "double"
  .Call("bcfgpu_R_glm_logit", as.integer(z), x, as.integer(device), PACKAGE = "BayesIVgpu")
}

# 2SLS of y on w with instrument z inside one subgroup: effect, p-value, weak-instrument p-value, compliers share
.node_iv <- function(d) {
  s <- summary(AER::ivreg(y ~ w | z, data = d), diagnostics = TRUE)
  comp <- mean(d$z == d$w)
  c(effect = s$coef[2, 1], p = s$coef[2, 4], weak = s$diagnostics[1, 4], compliers = comp)
}

# conditions on the path to rpart node `i`, joined as one expression ("x1< 0.5 & x2>=0.5")
.node_rule <- function(tree, i) {
  p <- rpart::path.rpart(tree, nodes = i, print.it = FALSE)[[1]]
  paste(p[-1], collapse = " & ")
}

quiet <- function(expr) {
  sink(tempfile()); on.exit(sink())
  invisible(force(expr))
}
