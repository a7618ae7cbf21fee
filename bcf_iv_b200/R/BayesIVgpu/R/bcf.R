#' Bayesian Causal Forest on the GPU: a bcf()-compatible front end
#'
#' Same formal arguments and return fields as bcf::bcf, which BayesIV calls at R/bcf_iv.R:89 and R/bcf_itt.R:71 of
#' fbargaglistoffi/BCF-IV; the Gibbs sampler runs in libbcfgpu.so (CUDA, sm_100a) through .Call (src/r_shim.c).
#' Arguments of bcf 2.x that have no meaning here (n_cores, n_threads, save_tree_directory, log_file, no_output,
#' verbose, update_interval) are accepted and ignored.  There is no CPU fallback.
#' @export
bcf <- function(y, z, x_control, x_moderate = x_control, pihat, nburn, nsim, nthin = 1, update_interval = 100,
                ntree_control = 200, sd_control = 2 * sd(y), base_control = 0.95, power_control = 2,
                ntree_moderate = 50, sd_moderate = sd(y), base_moderate = 0.25, power_moderate = 3,
                nu = 3, lambda = NULL, sigq = 0.9, sighat = NULL, include_pi = "control",
                use_muscale = TRUE, use_tauscale = TRUE,
                w = NULL, random_seed = sample.int(.Machine$integer.max, 1), n_chains = 1, n_cores = NULL,
                n_threads = NULL, no_output = TRUE, save_tree_directory = NULL, log_file = NULL, verbose = FALSE,
                device = -1L, devices = NULL, first_chain = 0L,
                x_predict_control = NULL, x_predict_moderate = x_predict_control, pihat_predict = NULL,
                keep_draws = 3 * 8 * as.double(length(y)) * nsim * n_chains <= 8 * 2^30) {
  # keep_draws: the three nsim x n matrices bcf returns (tau, mu, yhat) are kept only while they fit 8 GiB (the Python
  # twin's DRAW_BYTES_CAP); tau_mean / mu_mean / yhat_mean / tau_var -- what BayesIV consumes -- are always returned.
  # devices: CUDA ordinals the n_chains chains are dealt over, all running at the same time (one host thread per chain
  # inside the library: bcfgpu_run_chains); NULL = every visible device.  first_chain: index of the first chain's
  # Philox stream (bcf_iv gives the complier-proportion fit its own stream).
  if (!is.null(w)) stop("observation weights are not implemented")
  x_control <- as.matrix(x_control); x_moderate <- as.matrix(x_moderate); pihat <- as.matrix(pihat)
  n <- length(y)
  if (any(!is.finite(y)) || any(!is.finite(x_control)) || any(!is.finite(x_moderate)) || any(!is.finite(pihat)))
    stop("Missing or non-finite values in y, x_control, x_moderate or pihat")
  if (!all(sort(unique(z)) %in% c(0, 1))) stop("z must be a vector of 0's and 1's")
  if (length(z) != n || nrow(x_control) != n || nrow(x_moderate) != n || nrow(pihat) != n)
    stop("y, z, x_control, x_moderate and pihat must have the same number of rows")
  if (!(include_pi %in% c("control", "moderate", "both", "none"))) stop("bad include_pi")
  if (nburn < 100) warning("A low (<100) value for nburn was supplied")
  if (length(unique(y)) < 5) warning("y appears to be discrete")
  if (.Call("bcfgpu_R_device_count") < 1L) stop("no CUDA device available (libbcfgpu has no CPU fallback)")

  x_c <- if (include_pi %in% c("control", "both")) cbind(x_control, pihat) else x_control
  x_m <- if (include_pi %in% c("moderate", "both")) cbind(x_moderate, pihat) else x_moderate
  storage.mode(x_c) <- "double"; storage.mode(x_m) <- "double"
  cut_c <- .cutpoints_matrix(x_c)
  cut_m <- .cutpoints_matrix(x_m)
  muy <- mean(y); sdy <- sd(y); yscale <- (y - muy) / sdy
  if (is.null(sighat)) {   # summary(lm(yscale ~ z + x_c))$sigma, on the device (bcfgpu_lm_sigma)
    xs <- as.matrix(x_c); storage.mode(xs) <- "double"
    sighat <- .Call("bcfgpu_R_lm_sigma", as.double(yscale), as.integer(z), xs, 0L, PACKAGE = "BayesIVgpu")
  }
  if (is.null(lambda)) lambda <- sighat^2 * qchisq(1 - sigq, nu) / nu
  con_sd <- if (isTRUE(all.equal(sd_control, 2 * sdy))) 2 else sd_control / sdy
  mod_sd <- (if (isTRUE(all.equal(sd_moderate, sdy))) 1 else sd_moderate / sdy) / (if (use_tauscale) 0.674 else 1)
  perm <- order(z, decreasing = TRUE)           # reported only: the library permutes and un-permutes internally

  cfg <- list(ntree_control = ntree_control, ntree_moderate = ntree_moderate, base_control = base_control,
              power_control = power_control, base_moderate = base_moderate, power_moderate = power_moderate,
              con_sd = con_sd, mod_sd = mod_sd, nu = nu, lambda = lambda, sigma_init = sd(yscale),
              use_muscale = as.integer(use_muscale), use_tauscale = as.integer(use_tauscale),
              seed = random_seed, chain = as.integer(first_chain), device = device)
  if (is.null(devices)) devices <- if (n_chains > 1) seq_len(.Call("bcfgpu_R_device_count")) - 1L else as.integer(device)
  # bcf 2.x's predict() in the same call: new rows at which every saved sweep's forests are evaluated on the device
  xp_c <- xp_m <- matrix(0, 0, 0)
  if (!is.null(x_predict_control)) {
    php <- if (is.null(pihat_predict)) NULL else as.matrix(pihat_predict)
    if (include_pi != "none" && is.null(php)) stop("predicting needs pihat_predict")
    xp_c <- if (include_pi %in% c("control", "both")) cbind(as.matrix(x_predict_control), php) else as.matrix(x_predict_control)
    xp_m <- if (include_pi %in% c("moderate", "both")) cbind(as.matrix(x_predict_moderate), php) else as.matrix(x_predict_moderate)
    storage.mode(xp_c) <- "double"; storage.mode(xp_m) <- "double"
  }
  # all chains in ONE native call: the library runs them concurrently, chain c on devices[c mod length(devices)]
  chains <- .Call("bcfgpu_R_fit_chains", as.double(yscale), as.integer(z), x_c, x_m,
                  as.double(unlist(cut_c)), as.integer(c(0L, cumsum(lengths(cut_c)))),
                  as.double(unlist(cut_m)), as.integer(c(0L, cumsum(lengths(cut_m)))),
                  cfg, as.integer(nburn), as.integer(nsim), as.integer(nthin), isTRUE(keep_draws),
                  as.integer(n_chains), as.integer(devices), xp_c, xp_m)
  bind <- function(f) do.call(rbind, lapply(chains, function(r) t(r[[f]])))      # n x nsim -> nsim x n, chains stacked
  avg <- function(f) Reduce(`+`, lapply(chains, `[[`, f)) / length(chains)
  # variance of the STACKED draws (what fit$tau holds): within-chain sums of squares plus the spread of the chain means
  pooled_var <- function() {
    ns <- length(chains[[1]]$sigma); N <- ns * length(chains); m <- avg("tau_mean")
    if (N < 2) return(0 * m)
    Reduce(`+`, lapply(chains, function(r) (ns - 1) * r$tau_var + ns * (r$tau_mean - m)^2)) / (N - 1)
  }
  fit <- list(sigma = sdy * unlist(lapply(chains, `[[`, "sigma")),
              mu_scale = sdy * unlist(lapply(chains, `[[`, "mu_scale")),
              tau_scale = sdy * unlist(lapply(chains, `[[`, "tau_scale")),
              perm = perm,
              tau_mean = sdy * avg("tau_mean"), mu_mean = muy + sdy * avg("mu_mean"),
              yhat_mean = muy + sdy * avg("yhat_mean"), tau_var = sdy^2 * pooled_var())
  if (nrow(xp_c) > 0) {
    fit$tau_predict_mean <- sdy * avg("tau_predict_mean"); fit$mu_predict_mean <- muy + sdy * avg("mu_predict_mean")
  }
  if (isTRUE(keep_draws)) {
    fit$tau <- sdy * bind("tau"); fit$mu <- muy + sdy * bind("mu"); fit$yhat <- muy + sdy * bind("yhat")
  }
  fit
}

# every column's grid after ONE threaded pass of the library over the matrix (bcfgpu_scan_columns: min, max, two-valued,
# non-finite): a two-valued column -- the reference's covariates -- needs no sort, the others go through .cutpoints
.cutpoints_matrix <- function(x) {
  s <- .Call("bcfgpu_R_scan_columns", x, PACKAGE = "BayesIVgpu")
  if (any(s[3, ] >= 2)) stop("Missing or non-finite values in x_control, x_moderate or pihat")
  lapply(seq_len(ncol(x)), function(j)
    if (s[3, j] == 1 && s[1, j] != s[2, j]) s[1, j] + (s[2, j] - s[1, j]) / 2 else .cutpoints(x[, j]))
}

# bcf's .cp_quantile: one value -> itself; fewer than 8 distinct -> midpoints; else 10000 points of the empirical
# quantile function over an equispaced grid
.cutpoints <- function(x, num = 10000, cat_levels = 8) {
  ux <- sort(unique(x))
  if (length(ux) == 1) { warning("A supplied covariate contains a single distinct value."); return(x[1]) }
  if (length(ux) < cat_levels) return(ux[-length(ux)] + diff(ux) / 2)
  n <- length(x)
  q <- approxfun(sort(x), quantile(x, 0:(n - 1) / n), ties = mean)
  q(seq(min(x), max(x), length.out = num))
}
