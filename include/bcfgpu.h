/*
 * bcfgpu.h -- C-ABI of libbcfgpu.so, the B200 (sm_100a) implementation of the Bayesian Causal
 * Forest Gibbs sampler behind BayesIV's bcf_iv() / bcf_itt().
 *
 * What this boundary replaces in the reference (fbargaglistoffi/BCF-IV):
 *   R/bcf_iv.R:89   itt_bcf <- quiet(bcf(y[-index], z[-index], x[-index,], x[-index,], pihat, nburn=, nsim=))
 *   R/bcf_itt.R:71  bcf_fit <- quiet(bcf(...same...))
 * i.e. the native entry that bcf::bcf reaches through Rcpp's
 *   .Call("_bcf_bcfoverparRcppClean", y, z, t(x_c), t(x_m), cutpoint lists, nburn, nsim, nthin, ...)
 * (third-party package `bcf`, imported un-pinned at DESCRIPTION:23 / NAMESPACE:9; source not vendored,
 * see SURVEY.md 8b/8c).  The R shim that binds these entry points with .Call is
 * bcf_iv_b200/R/BayesIVgpu/src/r_shim.c; INTEGRATION.md shows the wiring.
 *
 * Conventions
 *   - plain C, no R / torch / CUDA types in any signature; every pointer is a HOST pointer unless the
 *     name says `_dev`.
 *   - every function returns a bcfgpu_status (0 = ok); bcfgpu_last_error() gives the message of the last
 *     failure on the calling thread.
 *   - the caller allocates every output buffer; the library owns all device memory behind the handle.
 *   - there is NO CPU fallback: with no CUDA device every compute call fails with BCFGPU_ERR_CUDA.
 *   - matrices are COLUMN-MAJOR doubles (R's native layout, no transpose needed): X[i + n*v].
 *   - rows may be in any order; the library applies bcf's `perm = order(z, decreasing=TRUE)` (treated rows
 *     first, stable) internally and un-permutes every per-observation output.
 */
#ifndef BCFGPU_H
#define BCFGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCFGPU_API __attribute__((visibility("default")))

#define BCFGPU_ABI_VERSION 1
#define BCFGPU_MAX_LEAVES 128   /* structural cap per tree (leaf slots are uint8, 255 = padding) */
#define BCFGPU_MAX_DEPTH 30     /* heap node ids are 32-bit */

typedef enum bcfgpu_status {
    BCFGPU_OK = 0,
    BCFGPU_ERR_BAD_ARG = 1,
    BCFGPU_ERR_CUDA = 2,        /* includes "no CUDA device": there is no CPU path */
    BCFGPU_ERR_NCCL = 3,
    BCFGPU_ERR_NUMERIC = 4,     /* NaN / fixed-point range guard tripped on device (bcf stop()s on NaN too) */
    BCFGPU_ERR_OOM = 5,
    BCFGPU_ERR_STATE = 6,       /* call order (e.g. run before set_data) */
    BCFGPU_ERR_INTERRUPTED = 7  /* bcfgpu_request_stop / the poll hook of bcfgpu_fit_chains ended the run */
} bcfgpu_status;

typedef enum bcfgpu_engine {
    BCFGPU_ENGINE_AUTO = 0,
    BCFGPU_ENGINE_GRAPH = 1,       /* one kernel per tree, a sweep captured in a CUDA graph */
    BCFGPU_ENGINE_PERSISTENT = 2,  /* one cooperative kernel per batch of sweeps, grid barrier per tree; the
                                      observation state lives in shared memory when the CTA's share fits */
    BCFGPU_ENGINE_PERSISTENT_HBM = 3, /* same schedule, observation state kept in HBM/L2 (any n) */
    BCFGPU_ENGINE_BATCHED = 4      /* cooperative kernel, up to 8 trees per observation pass and per grid barrier (exact
                                      integer cross-count correction), state in shared memory */
    , BCFGPU_ENGINE_PIPELINED = 5  /* BATCHED with the decisions off the pass's critical path: the CTA is split in 27 tile
                                      warps that stream the observations, 4 chain warps that decide the previous batch of
                                      4 trees (cross-batch counts keep it exact) and 1 utility warp that plans and stages
                                      two batches ahead; sweeps that hold a tree with more than 8 keys fall back to BATCHED.
                                      AUTO = PIPELINED when the observations of an SM fit in shared memory, else BATCHED
                                      (state in shared memory, else in HBM), else PERSISTENT */
} bcfgpu_engine;

/* Hyper-parameters of the sampler: bcf()'s arguments after the R wrapper's rescaling. */
typedef struct bcfgpu_config {
    int32_t struct_size;          /* = sizeof(bcfgpu_config), ABI guard */
    int32_t ntree_con;            /* ntree_control  (200) */
    int32_t ntree_mod;            /* ntree_moderate (50)  */
    double alpha_con, beta_con;   /* base_control .95, power_control 2   */
    double alpha_mod, beta_mod;   /* base_moderate .25, power_moderate 3 */
    double con_sd;                /* 2 when sd_control == 2*sd(y), else sd_control/sd(y) */
    double mod_sd;                /* (1 or sd_moderate/sd(y)) / 0.674 when use_tauscale */
    double nu, lambda;            /* sigma^2 prior: nu*lambda / chisq_nu */
    double sigma_init;            /* sd(yscale) */
    int32_t use_muscale, use_tauscale;
    int32_t min_leaf;             /* 5: birth needs >= min_leaf observations in each child */
    int32_t max_leaves;           /* 0 -> BCFGPU_MAX_LEAVES */
    uint64_t seed;                /* Philox key; reproducible for fixed (seed, chain) on any GPU count */
    uint32_t chain;               /* chain index: mixes into the key (one chain per GPU in chains mode) */
    int32_t device;               /* CUDA ordinal; -1 = current device */
    int32_t engine;               /* bcfgpu_engine */
    int32_t keep_draws;           /* bit0 tau, bit1 mu, bit2 yhat: keep nsim x n draws on device; bit3 (8): keep the FORESTS of
                                   * every saved sweep on the device (bcfgpu_predict) */
    int32_t max_ctas;             /* 0: one CTA per SM.  > 0: at most that many CTAs, so that several fits can share one GPU
                                   * side by side (bcf_iv fits the complier proportion and the ITT at the same time) */
    int32_t model;                /* bcfgpu_model: 0 = Bayesian Causal Forest (bcf), 1 = probit BART (bartCause::bartc's sampler) */
    int32_t treatment_col;        /* probit BART: column of x_control that holds the 0/1 treatment (S-learner: tau = the
                                   * individual effect Phi(f(x,1)) - Phi(f(x,0))), or -1: no treatment, only P(y = 1 | x) */
    int32_t reserved0;
    double probit_offset;         /* probit BART: P(y = 1 | x) = Phi(probit_offset + f(x))  (dbarts' binaryOffset, 0) */
    int32_t reserved[2];
} bcfgpu_config;

/* BCFGPU_MODEL_PROBIT_BART: the sampler behind bartCause::bartc(response, treatment, confounders) for a 0/1 response --
 * the reference's complier-proportion fit (R/bcf_iv.R:78-80) and its whole model when binary = TRUE (R/bcf_iv.R:198-202,
 * R/bcf_itt.R:166-170).  One forest (ntree_con trees; set ntree_mod = 0, use_muscale = 0, con_sd = 3 / k), sigma = 1,
 * Albert-Chib latent redrawn before every sweep.  set_data: y = the 0/1 response, z = the treatment (any 0/1 vector when
 * there is none), xcon = (confounders, ..., treatment) with treatment_col naming the treatment's column.  Results:
 * get_tau_mean / get_tau_var = posterior mean / variance of Phi(off + f(x, 1)) - Phi(off + f(x, 0)) per observation
 * (bartc's individual effects on the probability scale), get_mu_mean = posterior mean of P(y = 1 | x, z observed),
 * get_yhat_mean = posterior mean of f(x). */
typedef enum bcfgpu_model { BCFGPU_MODEL_BCF = 0, BCFGPU_MODEL_PROBIT_BART = 1 } bcfgpu_model;

typedef struct bcfgpu_handle bcfgpu_handle;

/* ---- lifecycle ------------------------------------------------------------------------------ */
BCFGPU_API int bcfgpu_abi_version(void);
BCFGPU_API int bcfgpu_device_count(void);                   /* 0 when no usable CUDA device */
BCFGPU_API const char *bcfgpu_last_error(void);
BCFGPU_API void bcfgpu_config_init(bcfgpu_config *cfg);     /* bcf()'s defaults */
BCFGPU_API int bcfgpu_create(const bcfgpu_config *cfg, bcfgpu_handle **out);
BCFGPU_API int bcfgpu_destroy(bcfgpu_handle *h);
/* The library parks the device (and pinned staging) blocks of destroyed handles, per device, and hands them to the next
 * fit of the same shape: cudaMalloc / cudaFree of a fit's buffers cost 5 - 35 ms per call at n = 1e6 and bcf_iv() fits
 * twice per call.  Bounded by BCFGPU_CACHE_MB per device (default 4096; 0 disables).  This empties the cache;
 * cache_stats: out2 = {device bytes parked over all devices, pinned host bytes parked}. */
BCFGPU_API int bcfgpu_release_cached_memory(void);
BCFGPU_API int bcfgpu_cache_stats(int64_t out2[2]);

/* ---- observation-sharded mode: call before set_data on every rank ------------------
 * nccl_unique_id: the 128 bytes of an ncclUniqueId made on rank 0 (bcfgpu_nccl_unique_id) and broadcast by
 * the caller (torch.distributed, MPI, ...).  Each rank then passes only ITS rows to set_data. */
BCFGPU_API int bcfgpu_nccl_unique_id(char id_out[128]);
BCFGPU_API int bcfgpu_set_shard(bcfgpu_handle *h, int rank, int nranks, const char nccl_unique_id[128]);
/* The same mode with every shard in ONE process: handles[r] becomes shard r of nranks (1 .. 8).  No NCCL, no IPC: the
 * handles reach each other's mailboxes directly (peer access is enabled between their devices); the few host-side
 * reductions meet in a barrier, so set_data / run / set_tree must then be called for all shards CONCURRENTLY, one host
 * thread per handle.  The shards may share a GPU: each then gets an equal share of its SMs (a few are left free for the
 * small kernels around the sweeps), so the whole sharded path, exchange included, runs on a single device. */
BCFGPU_API int bcfgpu_set_shard_local(bcfgpu_handle **handles, int nranks);

/* ---- data ------------------------------------------------------------------------------------
 * y: standardised outcome (n); z: 0/1 treatment = BayesIV's instrument (n);
 * xcon/xmod: column-major n x p_con / n x p_mod (xcon already has pihat appended when include_pi="control");
 * cuts_*: concatenated sorted cutpoints, cut_off_*[v]..cut_off_*[v+1] are variable v's (bcf's xinfo).
 * Bins X into column-major uint8/uint16 on device (kernel K0) and resets the chain to its initial state. */
BCFGPU_API int bcfgpu_set_data(bcfgpu_handle *h, int64_t n, int32_t p_con, int32_t p_mod, const double *y,
                               const int32_t *z, const double *xcon, const double *xmod,
                               const double *cuts_con, const int32_t *cut_off_con,
                               const double *cuts_mod, const int32_t *cut_off_mod);

/* ---- sampling --------------------------------------------------------------------------------
 * Runs nburn + nsim*nthin Gibbs sweeps (continuing the chain if called again); sweep it is saved when
 * it >= nburn and it % nthin == 0, as bcf does.  Blocking. */
BCFGPU_API int bcfgpu_run(bcfgpu_handle *h, int32_t nburn, int32_t nsim, int32_t nthin);
/* Asks the run in progress on `h` (another thread is inside bcfgpu_run) to end at its next launch boundary (at most 256
 * sweeps away); that bcfgpu_run returns BCFGPU_ERR_INTERRUPTED and the chain stays valid.  Callable from any thread. */
BCFGPU_API int bcfgpu_request_stop(bcfgpu_handle *h);
/* bcf 2.x's n_chains in ONE call: n_chains independent chains (Philox stream cfg->chain + c), chain c on CUDA device
 * devices[c % n_devices] (n_devices = 0: cfg->device for all), one host thread per device inside the library, chains
 * that share a device one after the other -- or side by side when cfg->max_ctas > 0 gives each a share of the SMs.
 * The same host buffers serve every chain.  `poll` (may be NULL) is called on the CALLING thread about every 20 ms
 * while the chains run; a non-zero return stops them (R shim: R_CheckUserInterrupt under R_ToplevelExec).
 * On success handles_out[0..n_chains) hold the finished chains: read them with the bcfgpu_get_* calls, then
 * bcfgpu_destroy each.  On failure every handle is destroyed and handles_out is all NULL. */
BCFGPU_API int bcfgpu_fit_chains(const bcfgpu_config *cfg, int32_t n_chains, const int32_t *devices, int32_t n_devices,
                                 int64_t n, int32_t p_con, int32_t p_mod, const double *y, const int32_t *z,
                                 const double *xcon, const double *xmod, const double *cuts_con,
                                 const int32_t *cut_off_con, const double *cuts_mod, const int32_t *cut_off_mod,
                                 int32_t nburn, int32_t nsim, int32_t nthin, int (*poll)(void *), void *poll_ctx,
                                 bcfgpu_handle **handles_out);
BCFGPU_API int bcfgpu_last_run_ms(bcfgpu_handle *h, double *device_ms);  /* CUDA-event time of the last run */
BCFGPU_API int bcfgpu_kernel_launches(bcfgpu_handle *h, int64_t *count); /* kernels launched since create */
/* bytes set_data copied host -> device.  When x_moderate IS the leading columns of x_control (same pointer, same
 * cutpoints -- what bcf(y, z, x, x, pihat) of R/bcf_iv.R:89 amounts to) the tau forest shares the mu forest's binned
 * columns and the matrix crosses the bus once. */
BCFGPU_API int bcfgpu_h2d_bytes(bcfgpu_handle *h, int64_t *bytes);
/* what the last run used: out[0] sweep kernel (0 graph of per-tree kernels, 1 k_persistent, 2 k_resident, 3 k_batched,
 * 4 k_pipe, 5 k_batched with the state in HBM), out[1] shards exchange (0 none, 1 ncclAllReduce per tree, 2 in-kernel peer-memory mailboxes) */
BCFGPU_API int bcfgpu_engine_in_use(bcfgpu_handle *h, int32_t out2[2]);
/* since create: out2 = {sweeps run by the pipelined kernel, sweeps it handed to the batched kernel because a tree had more
 * than 8 keys} (counted for multi-sweep launches; sweep-by-sweep callers queue the batched kernel without asking) */
BCFGPU_API int bcfgpu_engine_stats(bcfgpu_handle *h, int64_t out2[2]);

/* ---- results (standardised scale, original row order; the R wrapper rescales by sd(y)/mean(y)) ---- */
BCFGPU_API int bcfgpu_num_saved(bcfgpu_handle *h, int32_t *nsaved);
BCFGPU_API int bcfgpu_get_tau_mean(bcfgpu_handle *h, double *out_n);    /* colMeans(b_post): what BayesIV uses */
BCFGPU_API int bcfgpu_get_tau_var(bcfgpu_handle *h, double *out_n);     /* per-observation posterior variance */
BCFGPU_API int bcfgpu_get_mu_mean(bcfgpu_handle *h, double *out_n);
BCFGPU_API int bcfgpu_get_yhat_mean(bcfgpu_handle *h, double *out_n);
BCFGPU_API int bcfgpu_get_sigma(bcfgpu_handle *h, double *out_nsaved);
BCFGPU_API int bcfgpu_get_scales(bcfgpu_handle *h, double *mu_scale_nsaved, double *tau_scale_nsaved);
/* which: 0 tau, 1 mu, 2 yhat; out is nsaved x n ROW-major per draw (draw-major), needs keep_draws */
BCFGPU_API int bcfgpu_get_draws(bcfgpu_handle *h, int32_t which, double *out_nsaved_by_n);

/* ---- saved forests and out-of-sample evaluation (bcf 2.x's predict(); needs keep_draws bit 3 in the run) ----------
 * For every saved sweep s and every new row: mu(x) = mscale_s * sum of the mu trees' leaves, tau(x) = (bscale1_s -
 * bscale0_s) * sum of the tau trees' leaves, on the standardised scale (the R wrapper rescales as for the fit).
 * xcon_new / xmod_new: column-major n_new x p_con / n_new x p_mod, binned with the training cutpoints (K0).
 * Any output may be NULL; *_draws are nsaved x n_new, draw-major.  With keep_draws bit 3 the sampling part of the run is
 * launched sweep by sweep (the forest of each saved sweep is copied into a node pool between launches). */
BCFGPU_API int bcfgpu_num_saved_forests(bcfgpu_handle *h, int32_t *ndraws, int64_t *nnodes);
BCFGPU_API int bcfgpu_predict(bcfgpu_handle *h, int64_t n_new, const double *xcon_new, const double *xmod_new,
                              double *tau_mean, double *tau_var, double *mu_mean, double *tau_draws, double *mu_draws);

/* ---- chain state (parity tests, checkpoints, predict) ------------------------------------------ */
/* scalars: {sigma, mscale, bscale1, bscale0, delta_con, tau_con, tau_mod, sweeps_done} */
BCFGPU_API int bcfgpu_get_scalars(bcfgpu_handle *h, double out8[8]);
/* allfit_con = mscale * sum of mu trees, allfit_mod = bscale_z * sum of tau trees (original row order) */
BCFGPU_API int bcfgpu_get_fits(bcfgpu_handle *h, double *allfit_con_n, double *allfit_mod_n);
/* counters: {birth proposed, birth accepted, death proposed, death accepted} */
BCFGPU_API int bcfgpu_get_counters(bcfgpu_handle *h, int64_t out4[4]);
/* last sweep's move per tree: trace5[t] = {move (0 none,1 birth,2 death), accepted, node heap id, var, cut};
 * logr[t] = log MH ratio (NaN when not evaluated) */
BCFGPU_API int bcfgpu_get_trace(bcfgpu_handle *h, int32_t *trace5_T, double *logr_T);
BCFGPU_API int bcfgpu_tree_size(bcfgpu_handle *h, int32_t tree, int32_t *nnodes);
/* preorder serialisation: heap id (root 1, children 2i / 2i+1), split var (-1 = leaf), cut index, mu */
BCFGPU_API int bcfgpu_get_tree(bcfgpu_handle *h, int32_t tree, uint32_t *heap, int32_t *var, int32_t *cut, double *mu);
/* replace a tree's structure and leaf values; recomputes its leaf assignment (kernel K1) and the fits */
BCFGPU_API int bcfgpu_set_tree(bcfgpu_handle *h, int32_t tree, int32_t nnodes, const uint32_t *heap,
                               const int32_t *var, const int32_t *cut, const double *mu);
/* cached leaf of every observation for one tree, as the leaf's heap id (original row order) */
BCFGPU_API int bcfgpu_get_leaf_ids(bcfgpu_handle *h, int32_t tree, uint32_t *heap_n);
/* re-derive every cached leaf slot from the tree structures (K1) and count mismatches with the cache */
BCFGPU_API int bcfgpu_verify_leaf_cache(bcfgpu_handle *h, int64_t *mismatches);

/* ---- stand-alone operators on the handle's binned data (kernel-level parity) ----------- */
/* K0: bin one column: out[i] = #{c : cuts[c] <= x[i]}  (so x < cuts[c]  <=>  out <= c) */
BCFGPU_API int bcfgpu_op_bin_column(const double *x, int64_t n, const double *cuts, int32_t ncut, uint16_t *out);
/* K2: per-leaf sufficient statistics of tree `tree` for a caller-supplied residual r (original row order):
 * for each leaf in DFS order: heap id, count treated, count control, sum r treated, sum r control.
 * If birth_leaf_heap != 0 additionally splits that leaf by (var, cut): out_birth = {nl_trt, nl_ctl, sl_trt,
 * sl_ctl, nr_trt, nr_ctl, sr_trt, sr_ctl}. */
BCFGPU_API int bcfgpu_op_suffstats(bcfgpu_handle *h, int32_t tree, const double *r_n, uint32_t birth_leaf_heap,
                                   int32_t var, int32_t cut, int32_t *nleaves, uint32_t *heap, int64_t *cnt_trt,
                                   int64_t *cnt_ctl, double *sum_trt, double *sum_ctl, double out_birth[8]);
/* K2': all-candidate histogram: for leaf labels slot[i] in [0,nleaf) (or <0 = skip) and forest (0 con, 1 mod):
 * cnt/sum[leaf][cut_off[v] + v + b] over bins b = 0..ncut_v.  (count, residual sum) for every (var, cutpoint).
 * Two-valued columns with nleaf <= 12 go through k_hist_mma (tcgen05 int8 contraction over TMA-fed bins, int32
 * accumulators in tensor memory: sums exact in the 2^-44 fixed point, |r| < 2^15); BCFGPU_NO_HIST_MMA=1 selects the
 * CUDA-core kernels of round 1 instead. */
BCFGPU_API int bcfgpu_op_hist_all(bcfgpu_handle *h, int32_t forest, const int32_t *slot_n, int32_t nleaf,
                                  const double *r_n, int64_t *cnt, double *sum, double *device_ms);
/* K3: integrated log-likelihood of a leaf, weighted form (bcf's lilhet without the cancelling term) */
BCFGPU_API int bcfgpu_op_lil(const double *sum_phi, const double *sum_phi_r, const double *tau, int32_t m, double *out);
/* Philox stream probe: {ua, ub, normal} for (seed, chain, sweep, tree, purpose, idx), computed on device */
BCFGPU_API int bcfgpu_op_rng_probe(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t tree, uint32_t purpose,
                                   uint32_t idx, double out3[3]);
/* probit BART's latent draw: out[i] = ystar ~ N(m[i], 1) truncated by y[i] (1: > 0, 0: <= 0), from the uniform u[i] */
BCFGPU_API int bcfgpu_op_probit_latent(const double *m, const int32_t *y, const double *u, int32_t n, double *out);


/* ---- the dense pre-steps of bcf_iv() / bcf(): no handle, one call each ----------------------------
 * X is n x p doubles, column-major (row_major = 0: R's layout) or row-major (row_major = 1: NumPy's default; it is
 * transposed on the device, not on the host).  The n-sized passes and the weighted Gram matrix A' W A run on the
 * device in fp64 with block-ordered (deterministic) sums; the (p + 2)-square solve runs on the host.  Columns are
 * taken left to right as R's dqrdc2 does: a column that is a linear combination of the ones before it is ALIASED --
 * its coefficient is NaN (R's NA) and it does not enter the fit.
 *
 * bcfgpu_glm_logit replaces   p.score <- glm(z ~ x[-index,], family = binomial); pihat <- predict(p.score, ...)
 *   (R/bcf_iv.R:72-75, R/bcf_itt.R:58-61): stats::glm.fit's IRLS for binomial(logit) -- start mu = (z + 0.5) / 2,
 *   working weights max(mu (1 - mu), DBL_EPSILON), convergence |dev - dev_old| / (|dev| + 0.1) < epsilon (1e-8), at
 *   most max_iter (25) iterations.  eta_out[n] = the linear predictor (predict.glm's default type = "link");
 *   beta_out[p + 1] (intercept first; may be NULL); info[4] = {iterations, converged, rank, kernel launches}. */
BCFGPU_API int bcfgpu_glm_logit(int32_t device, int64_t n, int32_t p, const double *X, int32_t row_major,
                                const int32_t *z, int32_t max_iter, double epsilon, double *eta_out, double *beta_out,
                                double *deviance_out, int32_t *info);
/* bcfgpu_lm_sigma replaces    sighat <- summary(lm(yscale ~ z + x_c))$sigma   (bcf's R wrapper):
 *   sqrt(RSS / (n - rank)) of the least-squares fit of y on (1, z, X); z may be NULL (no such column).
 *   beta_out[1 + (z != NULL) + p] may be NULL; info[2] = {rank, kernel launches}. */
BCFGPU_API int bcfgpu_lm_sigma(int32_t device, int64_t n, int32_t p, const double *X, int32_t row_major,
                               const int32_t *z, const double *y, double *sigma_out, double *beta_out, int32_t *info);

/* The pass over the covariates that bcf's R wrapper makes before the native call (the NA check, and
 * .cp_quantile's look at every column), on the library's host threads; needs no device.  Per column v of the
 * column-major n x p matrix X (leading dimension ld >= n): lo[v] = min, hi[v] = max, flags[v] bit 0 = every value equals
 * lo or hi (a two-valued or constant column: its cutpoint grid is lo + (hi - lo) / 2, no sort), bit 1 = NaN / Inf present. */
BCFGPU_API int bcfgpu_scan_columns(const double *X, int64_t n, int32_t p, int64_t ld, double *lo, double *hi,
                                   uint8_t *flags);

#ifdef __cplusplus
}
#endif
#endif /* BCFGPU_H */
