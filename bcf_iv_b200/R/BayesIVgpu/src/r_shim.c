/*
 * r_shim.c -- the .Call boundary between R and libbcfgpu.so (include/bcfgpu.h).
 *
 * Replaces, for BayesIV, the native entry that bcf::bcf reaches through Rcpp:
 *     .Call("_bcf_bcfoverparRcppClean", y, z, t(x_c), t(x_m), cutpoint lists, nburn, nsim, nthin, ...)
 * (third-party package `bcf`; BayesIV call sites R/bcf_iv.R:89 and R/bcf_itt.R:71 of fbargaglistoffi/BCF-IV).
 * Differences by design: matrices stay in R's native column-major layout (no transpose, no copy), the caller's rows
 * stay in their order (the library applies and undoes bcf's treated-first permutation), outputs are allocated here
 * with allocVector/allocMatrix and filled by the library, errors become R errors carrying bcfgpu_last_error().
 *
 * Interrupts: the chains run on the library's own host threads (bcfgpu_fit_chains); this thread polls
 * R_CheckUserInterrupt() under R_ToplevelExec, so a Ctrl-C never longjmps past live handles: the library stops the
 * chains, destroys their handles and only then the shim raises the R error.  The whole run is ONE logical bcfgpu_run
 * per chain (nburn, nsim, nthin together), so thinning follows bcf's `iIter % thin` rule whatever nburn is.
 *
 * Not compiled in this repository's CI (no R in the image): tests/test_r_package.py syntax-checks it against a stub
 * of Rinternals.h; INTEGRATION.md shows the build line.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Utils.h>
#include <stdlib.h>
#include <string.h>

#include "bcfgpu.h"

static void destroy_all(bcfgpu_handle **hs, int nh)
{
    for (int c = 0; hs && c < nh; ++c) { if (hs[c]) bcfgpu_destroy(hs[c]); hs[c] = NULL; }
}

/* non-zero status -> R error, after every live handle is destroyed (Rf_error does not return) */
static void check(int rc, bcfgpu_handle **hs, int nh)
{
    if (rc != BCFGPU_OK) {
        char msg[512];
        strncpy(msg, bcfgpu_last_error(), sizeof(msg) - 1);
        msg[sizeof(msg) - 1] = 0;
        destroy_all(hs, nh);
        free(hs);
        Rf_error("libbcfgpu: %s", msg);
    }
}

static double num(SEXP list, const char *name, double dflt)
{
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    for (R_xlen_t i = 0; i < Rf_xlength(list); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(list, i));
    return dflt;
}

/* R_CheckUserInterrupt() longjmps on Ctrl-C: under R_ToplevelExec the jump is caught and reported as FALSE */
static void check_interrupt(void *unused) { (void)unused; R_CheckUserInterrupt(); }
static int interrupt_pending(void *unused) { (void)unused; return R_ToplevelExec(check_interrupt, NULL) == FALSE; }

/* .Call("bcfgpu_R_device_count") */
SEXP bcfgpu_R_device_count(void) { return Rf_ScalarInteger(bcfgpu_device_count()); }

/* results of one finished chain -> list(tau_mean, tau_var, mu_mean, yhat_mean, sigma, mu_scale, tau_scale, tau, mu, yhat,
 * device_ms) on the standardised scale; tau / mu / yhat are n x nsaved matrices when keep, else NULL */
static SEXP chain_result(bcfgpu_handle **hs, int nh, int c, R_xlen_t n, int keep, SEXP xpc, SEXP xpm)
{
    bcfgpu_handle *h = hs[c];
    double ms = 0;
    bcfgpu_last_run_ms(h, &ms);
    int32_t ns = 0;
    check(bcfgpu_num_saved(h, &ns), hs, nh);
    const char *nm[] = {"tau_mean", "tau_var", "mu_mean", "yhat_mean", "sigma", "mu_scale", "tau_scale", "tau", "mu",
                        "yhat", "device_ms", "tau_predict_mean", "tau_predict_var", "mu_predict_mean", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    SEXP v;
    v = PROTECT(Rf_allocVector(REALSXP, n)); check(bcfgpu_get_tau_mean(h, REAL(v)), hs, nh); SET_VECTOR_ELT(out, 0, v); UNPROTECT(1);
    v = PROTECT(Rf_allocVector(REALSXP, n)); check(bcfgpu_get_tau_var(h, REAL(v)), hs, nh); SET_VECTOR_ELT(out, 1, v); UNPROTECT(1);
    v = PROTECT(Rf_allocVector(REALSXP, n)); check(bcfgpu_get_mu_mean(h, REAL(v)), hs, nh); SET_VECTOR_ELT(out, 2, v); UNPROTECT(1);
    v = PROTECT(Rf_allocVector(REALSXP, n)); check(bcfgpu_get_yhat_mean(h, REAL(v)), hs, nh); SET_VECTOR_ELT(out, 3, v); UNPROTECT(1);
    v = PROTECT(Rf_allocVector(REALSXP, ns)); check(bcfgpu_get_sigma(h, REAL(v)), hs, nh); SET_VECTOR_ELT(out, 4, v); UNPROTECT(1);
    {
        SEXP a = PROTECT(Rf_allocVector(REALSXP, ns)), b = PROTECT(Rf_allocVector(REALSXP, ns));
        check(bcfgpu_get_scales(h, REAL(a), REAL(b)), hs, nh);
        SET_VECTOR_ELT(out, 5, a); SET_VECTOR_ELT(out, 6, b);
        UNPROTECT(2);
    }
    if (keep) {
        /* the library returns draw-major rows (nsaved x n, row-major) = an n x nsaved column-major R matrix: the R
         * wrapper transposes to bcf's nsim x n */
        for (int which = 0; which < 3; ++which) {
            v = PROTECT(Rf_allocMatrix(REALSXP, (int)n, ns));
            check(bcfgpu_get_draws(h, which, REAL(v)), hs, nh);
            SET_VECTOR_ELT(out, 7 + which, v);
            UNPROTECT(1);
        }
    }
    SET_VECTOR_ELT(out, 10, Rf_ScalarReal(ms));
    if (Rf_isMatrix(xpc) && Rf_nrows(xpc) > 0) {   /* bcf 2.x's predict(): the saved forests at new rows (keep_draws bit 3) */
        const int m = Rf_nrows(xpc);
        SEXP a = PROTECT(Rf_allocVector(REALSXP, m)), b = PROTECT(Rf_allocVector(REALSXP, m)), d = PROTECT(Rf_allocVector(REALSXP, m));
        check(bcfgpu_predict(h, (int64_t)m, REAL(xpc), Rf_isMatrix(xpm) && Rf_ncols(xpm) > 0 ? REAL(xpm) : NULL, REAL(a), REAL(b),
                             REAL(d), NULL, NULL), hs, nh);
        SET_VECTOR_ELT(out, 11, a); SET_VECTOR_ELT(out, 12, b); SET_VECTOR_ELT(out, 13, d);
        UNPROTECT(3);
    }
    UNPROTECT(1);
    return out;
}

/*
 * .Call("bcfgpu_R_fit_chains", yscale, z, x_c, x_m, cuts_c, off_c, cuts_m, off_m, cfg, nburn, nsim, nthin, keep_draws,
 *       n_chains, devices, x_predict_control, x_predict_moderate)
 *   yscale  numeric n      standardised outcome            z       integer n (0/1)
 *   x_c/x_m numeric matrices n x p_con / n x p_mod (column-major, as R stores them)
 *   cuts_*  numeric        concatenated cutpoints          off_*   integer p+1 offsets
 *   cfg     named list     the bcfgpu_config fields that differ from bcf's defaults (chain = first Philox stream)
 *   devices integer        CUDA ordinals the chains are dealt over (chain c on devices[c %% length]); they run at once
 *   x_predict_*  numeric matrices (0 rows: none): new rows at which the saved forests are evaluated (bcf 2.x's predict())
 * Returns a list of n_chains per-chain results (see chain_result).
 */
SEXP bcfgpu_R_fit_chains(SEXP y, SEXP z, SEXP xc, SEXP xm, SEXP cuts_c, SEXP off_c, SEXP cuts_m, SEXP off_m, SEXP cfg_list,
                         SEXP nburn_, SEXP nsim_, SEXP nthin_, SEXP keep_, SEXP nchains_, SEXP devices_, SEXP xpc, SEXP xpm)
{
    const R_xlen_t n = Rf_xlength(y);
    if (!Rf_isReal(y) || !Rf_isInteger(z) || !Rf_isMatrix(xc) || !Rf_isReal(xc) || !Rf_isMatrix(xm) || !Rf_isReal(xm))
        Rf_error("bcfgpu_R_fit_chains: y, x_control, x_moderate must be double, z integer");
    if (Rf_xlength(z) != n || Rf_nrows(xc) != n || Rf_nrows(xm) != n) Rf_error("bcfgpu_R_fit_chains: row counts differ");
    const int p_con = Rf_ncols(xc), p_mod = Rf_ncols(xm);
    if (Rf_xlength(off_c) != p_con + 1 || Rf_xlength(off_m) != p_mod + 1) Rf_error("bcfgpu_R_fit_chains: bad cutpoint offsets");
    const int nburn = Rf_asInteger(nburn_), nsim = Rf_asInteger(nsim_), nthin = Rf_asInteger(nthin_);
    const int keep = Rf_asLogical(keep_);
    const int nch = Rf_asInteger(nchains_);
    if (nch < 1 || !Rf_isInteger(devices_)) Rf_error("bcfgpu_R_fit_chains: need n_chains >= 1 and integer devices");
    int ndev = (int)Rf_xlength(devices_);
    if (ndev == 1 && INTEGER(devices_)[0] < 0) ndev = 0;      /* -1: the current device */

    bcfgpu_config cfg;
    bcfgpu_config_init(&cfg);
    cfg.ntree_con = (int32_t)num(cfg_list, "ntree_control", cfg.ntree_con);
    cfg.ntree_mod = (int32_t)num(cfg_list, "ntree_moderate", cfg.ntree_mod);
    cfg.alpha_con = num(cfg_list, "base_control", cfg.alpha_con);
    cfg.beta_con = num(cfg_list, "power_control", cfg.beta_con);
    cfg.alpha_mod = num(cfg_list, "base_moderate", cfg.alpha_mod);
    cfg.beta_mod = num(cfg_list, "power_moderate", cfg.beta_mod);
    cfg.con_sd = num(cfg_list, "con_sd", cfg.con_sd);
    cfg.mod_sd = num(cfg_list, "mod_sd", cfg.mod_sd);
    cfg.nu = num(cfg_list, "nu", cfg.nu);
    cfg.lambda = num(cfg_list, "lambda", cfg.lambda);
    cfg.sigma_init = num(cfg_list, "sigma_init", cfg.sigma_init);
    cfg.use_muscale = (int32_t)num(cfg_list, "use_muscale", 1);
    cfg.use_tauscale = (int32_t)num(cfg_list, "use_tauscale", 1);
    cfg.seed = (uint64_t)num(cfg_list, "seed", 42);
    cfg.chain = (uint32_t)num(cfg_list, "chain", 0);
    cfg.device = (int32_t)num(cfg_list, "device", -1);
    cfg.max_ctas = (int32_t)num(cfg_list, "max_ctas", 0);
    cfg.model = (int32_t)num(cfg_list, "model", BCFGPU_MODEL_BCF);            /* 1: probit BART (R/bartc.R) */
    cfg.treatment_col = (int32_t)num(cfg_list, "treatment_col", -1);
    cfg.probit_offset = num(cfg_list, "probit_offset", 0.0);
    cfg.keep_draws = keep ? 7 : 0;
    const int predict = Rf_isMatrix(xpc) && Rf_nrows(xpc) > 0;
    if (predict) {
        if (Rf_ncols(xpc) != p_con || (p_mod > 0 && (!Rf_isMatrix(xpm) || Rf_ncols(xpm) != p_mod || Rf_nrows(xpm) != Rf_nrows(xpc))))
            Rf_error("bcfgpu_R_fit_chains: the rows to predict at must have the training columns");
        cfg.keep_draws |= 8;
    }

    bcfgpu_handle **hs = (bcfgpu_handle **)calloc((size_t)nch, sizeof(bcfgpu_handle *));
    if (!hs) Rf_error("bcfgpu_R_fit_chains: out of memory");
    /* on failure (or interrupt) the library has already destroyed every handle */
    check(bcfgpu_fit_chains(&cfg, nch, ndev ? INTEGER(devices_) : NULL, ndev, (int64_t)n, p_con, p_mod, REAL(y), INTEGER(z),
                            REAL(xc), REAL(xm), REAL(cuts_c), INTEGER(off_c), REAL(cuts_m), INTEGER(off_m), nburn, nsim,
                            nthin, interrupt_pending, NULL, hs), hs, nch);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, nch));
    for (int c = 0; c < nch; ++c) SET_VECTOR_ELT(out, c, chain_result(hs, nch, c, n, keep, xpc, xpm));
    destroy_all(hs, nch);
    free(hs);
    UNPROTECT(1);
    return out;
}

/* .Call("bcfgpu_R_glm_logit", z, x, device): the linear predictor of glm(z ~ x, family = binomial) -- what the reference
 * computes with stats::glm + predict (R/bcf_iv.R:72-75, R/bcf_itt.R:58-61) -- by IRLS on the device.  x: numeric matrix
 * (column-major, handed over as it is). */
SEXP bcfgpu_R_glm_logit(SEXP z, SEXP x, SEXP device)
{
    if (!Rf_isReal(x) || !Rf_isMatrix(x) || !Rf_isInteger(z)) Rf_error("bcfgpu_R_glm_logit: x must be a double matrix, z an integer vector");
    const R_xlen_t n = Rf_nrows(x);
    const int p = Rf_ncols(x);
    if (Rf_xlength(z) != n) Rf_error("bcfgpu_R_glm_logit: z and x differ in length");
    SEXP eta = PROTECT(Rf_allocVector(REALSXP, n));
    int info[4] = {0, 0, 0, 0};
    const int rc = bcfgpu_glm_logit(Rf_asInteger(device), (int64_t)n, p, REAL(x), 0, INTEGER(z), 25, 1e-8, REAL(eta), NULL, NULL, info);
    if (rc != BCFGPU_OK) { UNPROTECT(1); Rf_error("libbcfgpu: %s", bcfgpu_last_error()); }
    if (!info[1]) Rf_warning("glm.fit: algorithm did not converge");
    UNPROTECT(1);
    return eta;
}

/* .Call("bcfgpu_R_lm_sigma", y, z, x, device): summary(lm(y ~ z + x))$sigma, bcf's default sighat, on the device */
SEXP bcfgpu_R_lm_sigma(SEXP y, SEXP z, SEXP x, SEXP device)
{
    if (!Rf_isReal(x) || !Rf_isMatrix(x) || !Rf_isReal(y) || !Rf_isInteger(z)) Rf_error("bcfgpu_R_lm_sigma: bad argument types");
    const R_xlen_t n = Rf_nrows(x);
    if (Rf_xlength(y) != n || Rf_xlength(z) != n) Rf_error("bcfgpu_R_lm_sigma: y, z and x differ in length");
    double sigma = 0.0;
    const int rc = bcfgpu_lm_sigma(Rf_asInteger(device), (int64_t)n, Rf_ncols(x), REAL(x), 0, INTEGER(z), REAL(y), &sigma, NULL, NULL);
    if (rc != BCFGPU_OK) Rf_error("libbcfgpu: %s", bcfgpu_last_error());
    return Rf_ScalarReal(sigma);
}

/* .Call("bcfgpu_R_scan_columns", x): 3 x p matrix (min, max, flags) of the columns of x in one threaded pass of the library:
 * flags 1 = two-valued or constant column (its cutpoint grid needs no sort), 2 = a NaN / Inf is present -- what bcf's R
 * preamble asks of every column before the native call (the NA check and .cp_quantile's first look) */
SEXP bcfgpu_R_scan_columns(SEXP x)
{
    if (!Rf_isReal(x) || !Rf_isMatrix(x)) Rf_error("bcfgpu_R_scan_columns: x must be a double matrix");
    const R_xlen_t n = Rf_nrows(x);
    const int p = Rf_ncols(x);
    if (n < 1 || p < 1) Rf_error("bcfgpu_R_scan_columns: empty matrix");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 3, p));
    double *lo = (double *)R_alloc((size_t)p, sizeof(double)), *hi = (double *)R_alloc((size_t)p, sizeof(double));
    uint8_t *fl = (uint8_t *)R_alloc((size_t)p, 1);
    const int rc = bcfgpu_scan_columns(REAL(x), (int64_t)n, p, (int64_t)n, lo, hi, fl);
    if (rc != BCFGPU_OK) { UNPROTECT(1); Rf_error("libbcfgpu: %s", bcfgpu_last_error()); }
    for (int j = 0; j < p; ++j) { REAL(out)[3 * j] = lo[j]; REAL(out)[3 * j + 1] = hi[j]; REAL(out)[3 * j + 2] = (double)fl[j]; }
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"bcfgpu_R_device_count", (DL_FUNC)&bcfgpu_R_device_count, 0},
    {"bcfgpu_R_fit_chains", (DL_FUNC)&bcfgpu_R_fit_chains, 17},
    {"bcfgpu_R_glm_logit", (DL_FUNC)&bcfgpu_R_glm_logit, 3},
    {"bcfgpu_R_lm_sigma", (DL_FUNC)&bcfgpu_R_lm_sigma, 4},
    {"bcfgpu_R_scan_columns", (DL_FUNC)&bcfgpu_R_scan_columns, 1},
    {NULL, NULL, 0}};

void R_init_BayesIVgpu(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
