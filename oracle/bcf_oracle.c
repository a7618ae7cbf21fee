/*
 * bcf_oracle.c -- CPU ORACLE for the Bayesian Causal Forest Gibbs sweep.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (bcf_iv_b200/, include/bcfgpu.h) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED: the algorithm restated here lives in the third-party R package
 * `bcf` (github.com/jaredsmurray/bcf; BayesIV imports it un-pinned, DESCRIPTION:23,
 * NAMESPACE:9, README.md:22).  Its source is not vendored in /root/reference, there
 * is no R in this image and the reference has no tests or golden vectors
 * (SURVEY.md section 4, section 8c).  What is restated is bcf's published algorithm
 * (Hahn, Murray & Carvalho 2020; Chipman, George & McCulloch 2010) as specified in
 * SURVEY.md Appendix A, anchored on the reference's own call sites
 * R/bcf_iv.R:89-91 and R/bcf_itt.R:71-75.
 *
 * Shape: deliberately the reference's own -- row-major double X (p doubles per
 * observation), pointer trees, no leaf cache, four O(n*depth) passes per tree
 * (fit, getsuff, allsuff, fit), sequential fp64 sums.  The only thing that is not
 * bcf's is the random stream: bcf draws from R's global RNG (not reproducible
 * outside R); the oracle draws from a counter-based Philox4x32-10 stream whose
 * addressing (sweep, tree, purpose, index) is the contract shared with the CUDA
 * path, so the two chains can be compared move by move.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11).  Known-answer vectors are checked in  */
/* tests/test_oracle_rng.py.                                                   */
/* ------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

/* Stream addressing shared with the CUDA path (DESIGN.md "Random stream"). */
enum {
    PUR_MOVE = 0,   /* ua: birth/death choice, ub: MH accept uniform            */
    PUR_NODE = 1,   /* ua: node choice, ub: variable choice                      */
    PUR_CUT = 2,    /* ua: cutpoint choice                                       */
    PUR_LEAF = 3,   /* idx = heap id of the leaf: normal for the leaf mean draw */
    PUR_BSCALE1 = 10, PUR_BSCALE0 = 11, PUR_MSCALE = 12,
    PUR_DELTA_CON = 13, PUR_DELTA_MOD = 14, PUR_SIGMA = 15,
    PUR_LATENT = 20 /* idx = the caller's row number: the uniform behind that row's truncated-normal latent */
};
#define TREE_SWEEP_LEVEL 0xFFFFFFFFu

typedef struct { uint32_t key[2]; } rng_t;

static void rng_init(rng_t *g, uint64_t seed, uint32_t chain)
{
    const uint64_t s = seed + (uint64_t)chain * 0x9E3779B97F4A7C15ull;
    g->key[0] = (uint32_t)s;
    g->key[1] = (uint32_t)(s >> 32);
}

/* two uniforms in (0,1), 53 bits each */
static void rng_u2(const rng_t *g, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx,
                   double *ua, double *ub)
{
    uint32_t ctr[4] = {sweep, tree, purpose, idx}, o[4];
    orc_philox4x32_10(ctr, g->key, o);
    const double inv53 = 1.0 / 9007199254740992.0;
    *ua = ((double)(((uint64_t)(o[0] >> 5) << 26) | (uint64_t)(o[1] >> 6)) + 0.5) * inv53;
    *ub = ((double)(((uint64_t)(o[2] >> 5) << 26) | (uint64_t)(o[3] >> 6)) + 0.5) * inv53;
}

static double rng_normal(const rng_t *g, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx)
{
    double ua, ub;
    rng_u2(g, sweep, tree, purpose, idx, &ua, &ub);
    return sqrt(-2.0 * log(ua)) * cos(6.283185307179586476925286766559 * ub);
}

/* Gamma(shape a >= 1/3, scale 1) by Marsaglia-Tsang; attempt t uses idx 2t (normal) and 2t+1 (uniform). */
static double rng_gamma(const rng_t *g, uint32_t sweep, uint32_t purpose, double a)
{
    double boost = 1.0;
    if (a < 1.0) { /* Gamma(a) = Gamma(a+1) * U^(1/a); idx 0xFFFFFFFF reserved for U */
        double ua, ub;
        rng_u2(g, sweep, TREE_SWEEP_LEVEL, purpose, 0xFFFFFFFFu, &ua, &ub);
        boost = pow(ua, 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0;; ++t) {
        const double z = rng_normal(g, sweep, TREE_SWEEP_LEVEL, purpose, 2 * t);
        double ua, ub;
        rng_u2(g, sweep, TREE_SWEEP_LEVEL, purpose, 2 * t + 1, &ua, &ub);
        const double w = 1.0 + c * z;
        if (w <= 0.0) continue;
        const double v = w * w * w;
        if (log(ua) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
        if (t > 1000000u) return boost * d; /* unreachable in practice */
    }
}

ORC_API void orc_rng_probe(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t tree, uint32_t purpose,
                           uint32_t idx, double out[3])
{
    rng_t g; rng_init(&g, seed, chain);
    rng_u2(&g, sweep, tree, purpose, idx, &out[0], &out[1]);
    out[2] = rng_normal(&g, sweep, tree, purpose, idx);
}

ORC_API double orc_rng_gamma_probe(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t purpose, double a)
{
    rng_t g; rng_init(&g, seed, chain);
    return rng_gamma(&g, sweep, purpose, a);
}

/* ------------------------------------------------------------------------- */
/* Pointer tree (bcf's `tree`: mu, v, c, p, l, r) [bcf-recall, SURVEY A.2]     */
/* ------------------------------------------------------------------------- */
typedef struct node {
    double mu;
    int v, c;              /* split variable and cutpoint index (internal nodes) */
    struct node *p, *l, *r;
    int tmp;               /* scratch: index in the current bottom-node list */
} node;

static node *node_new(node *parent, double mu)
{
    node *n = (node *)calloc(1, sizeof(node));
    n->mu = mu; n->p = parent; n->v = -1; n->c = -1;
    return n;
}
static void tree_free(node *n) { if (!n) return; tree_free(n->l); tree_free(n->r); free(n); }
static int node_depth(const node *n) { int d = 0; while (n->p) { n = n->p; ++d; } return d; }
static uint32_t node_heap_id(const node *n)
{   /* root 1, children 2i and 2i+1 */
    if (!n->p) return 1u;
    return 2u * node_heap_id(n->p) + (n->p->r == n ? 1u : 0u);
}
static int node_is_nog(const node *n) { return n->l && !n->l->l && !n->r->l; }
static int tree_size(const node *n) { return n->l ? 1 + tree_size(n->l) + tree_size(n->r) : 1; }

typedef struct { node **a; int n, cap; } nvec;
static void nvec_push(nvec *v, node *x)
{
    if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 16; v->a = (node **)realloc(v->a, sizeof(node *) * v->cap); }
    v->a[v->n++] = x;
}
static void tree_bots(node *n, nvec *out) { if (!n->l) nvec_push(out, n); else { tree_bots(n->l, out); tree_bots(n->r, out); } }
static void tree_nogs(node *n, nvec *out)
{
    if (!n->l) return;
    if (node_is_nog(n)) { nvec_push(out, n); return; }
    tree_nogs(n->l, out); tree_nogs(n->r, out);
}

/* cutpoint grid (bcf's xinfo): per variable a sorted vector of doubles */
typedef struct { int p; const double *cuts; const int32_t *off; } xinfo; /* cuts of var v: cuts[off[v] .. off[v+1]) */
static inline int xi_ncut(const xinfo *xi, int v) { return xi->off[v + 1] - xi->off[v]; }
static inline double xi_cut(const xinfo *xi, int v, int c) { return xi->cuts[xi->off[v] + c]; }

/* range of usable cutpoint indices for variable v at node n (tree::rg) */
static void node_rg(const node *n, int v, const xinfo *xi, int *lo, int *hi)
{
    *lo = 0; *hi = xi_ncut(xi, v) - 1;
    const node *child = n;
    for (const node *a = n->p; a; child = a, a = a->p) {
        if (a->v != v) continue;
        if (a->l == child) { if (a->c - 1 < *hi) *hi = a->c - 1; }
        else               { if (a->c + 1 > *lo) *lo = a->c + 1; }
    }
}
#define ORC_MAX_DEPTH 30 /* heap ids are 32-bit; shared structural cap, see DESIGN.md */
static int goodvars(const node *n, const xinfo *xi, int *list /* may be NULL */)
{
    if (node_depth(n) >= ORC_MAX_DEPTH) return 0;
    int cnt = 0;
    for (int v = 0; v < xi->p; ++v) {
        int lo, hi; node_rg(n, v, xi, &lo, &hi);
        if (hi >= lo) { if (list) list[cnt] = v; ++cnt; }
    }
    return cnt;
}
static int cansplit(const node *n, const xinfo *xi) { return goodvars(n, xi, NULL) > 0; }

/* bottom node an observation falls in (tree::bn): left iff x[v] < cut[v][c] */
static node *tree_bn(node *n, const double *x, const xinfo *xi)
{
    while (n->l) n = (x[n->v] < xi_cut(xi, n->v, n->c)) ? n->l : n->r;
    return n;
}

/* ------------------------------------------------------------------------- */
/* Sampler state                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    int32_t ntree_con, ntree_mod;
    double alpha_con, beta_con, alpha_mod, beta_mod;   /* base / power */
    double con_sd, mod_sd;                             /* 2 and 1/0.674 by default */
    double nu, lambda, sigma_init;
    int32_t use_muscale, use_tauscale;
    int32_t min_leaf;                                  /* 5 */
    int32_t max_leaves;                                /* 0 = unlimited (reference); else structural cap mirrored from the GPU */
    uint64_t seed; uint32_t chain;
    int32_t nthreads;                                  /* OpenMP threads over observations (1 = bcf 1.x) */
} orc_config;

typedef struct {
    double n0;   /* count */
    double n;    /* sum of weights phi */
    double sy;   /* sum of phi * R */
} sinfo;

typedef struct orc_sampler {
    orc_config cfg;
    int n, ntrt, p_con, p_mod;
    double *y;
    double *xc, *xm;              /* row-major: n x p */
    double *cuts_c, *cuts_m; int32_t *off_c, *off_m;
    xinfo xi_c, xi_m;
    node **t_con, **t_mod;
    double *allfit, *allfit_con, *allfit_mod, *ftemp, *rtemp, *weight;
    double sigma, mscale, bscale0, bscale1, delta_con, delta_mod, tau_con, tau_mod;
    rng_t rng;
    uint32_t sweep;
    /* counters */
    int64_t n_birth_prop, n_birth_acc, n_death_prop, n_death_acc;
    /* last-move trace (for parity tests) */
    int32_t *trace;  /* per tree: [move(0 none,1 birth,2 death), accepted, nodeheap, var, cut] */
    double *trace_logr; /* per tree: log MH ratio (log alpha1 + lil difference), NaN if not evaluated */
    /* probit BART (the bartc replacement, see "Probit BART" below): 0/1 response behind a latent Gaussian */
    int probit;          /* 1: y is the latent, redrawn before every sweep; sigma stays 1 */
    uint8_t *ybin;       /* the observed 0/1 response, internal order */
    int32_t *rowid;      /* caller's row number of every internal row: the Philox index of its latent draw */
    double offset;       /* binaryOffset: P(y = 1 | x) = Phi(offset + f(x)) */
    int trt_col;         /* column of the treatment in x_con (S-learner), or -1 */
} orc_sampler;

static void fit_tree(orc_sampler *s, node *t, const double *x, int p, const xinfo *xi, double *out)
{
    const int n = s->n;
#pragma omp parallel for schedule(static) num_threads(s->cfg.nthreads)
    for (int i = 0; i < n; ++i) out[i] = tree_bn(t, x + (size_t)i * p, xi)->mu;
}

static double lil(double sum_phi, double sum_phi_r, double tau)
{   /* integrated log-likelihood, weighted form; the -0.5*sum(phi R^2) term cancels in every ratio */
    const double d = 1.0 / (tau * tau) + sum_phi;
    return -log(tau) - 0.5 * log(d) + 0.5 * sum_phi_r * sum_phi_r / d;
}

/* getsuff, birth form: stats of the would-be children of leaf nx under rule (v,c) */
static void getsuff_birth(orc_sampler *s, node *t, node *nx, int v, int c, const double *x, int p,
                          const xinfo *xi, const double *phi, const double *R, sinfo *sl, sinfo *sr)
{
    double l0 = 0, ln = 0, ly = 0, r0 = 0, rn = 0, ry = 0;
    const double cut = xi_cut(xi, v, c);
    const int n = s->n;
#pragma omp parallel for schedule(static) num_threads(s->cfg.nthreads) reduction(+:l0,ln,ly,r0,rn,ry)
    for (int i = 0; i < n; ++i) {
        const double *xx = x + (size_t)i * p;
        if (tree_bn(t, xx, xi) != nx) continue;
        if (xx[v] < cut) { l0 += 1; ln += phi[i]; ly += phi[i] * R[i]; }
        else             { r0 += 1; rn += phi[i]; ry += phi[i] * R[i]; }
    }
    sl->n0 = l0; sl->n = ln; sl->sy = ly; sr->n0 = r0; sr->n = rn; sr->sy = ry;
}
/* getsuff, death form: stats of the two leaf children l and r */
static void getsuff_death(orc_sampler *s, node *t, node *nl, node *nr, const double *x, int p, const xinfo *xi,
                          const double *phi, const double *R, sinfo *sl, sinfo *sr)
{
    double l0 = 0, ln = 0, ly = 0, r0 = 0, rn = 0, ry = 0;
    const int n = s->n;
#pragma omp parallel for schedule(static) num_threads(s->cfg.nthreads) reduction(+:l0,ln,ly,r0,rn,ry)
    for (int i = 0; i < n; ++i) {
        node *b = tree_bn(t, x + (size_t)i * p, xi);
        if (b == nl) { l0 += 1; ln += phi[i]; ly += phi[i] * R[i]; }
        else if (b == nr) { r0 += 1; rn += phi[i]; ry += phi[i] * R[i]; }
    }
    sl->n0 = l0; sl->n = ln; sl->sy = ly; sr->n0 = r0; sr->n = rn; sr->sy = ry;
}
/* allsuff: stats for every bottom node */
static void allsuff(orc_sampler *s, node *t, const nvec *bots, const double *x, int p, const xinfo *xi,
                    const double *phi, const double *R, sinfo *out)
{
    const int nb = bots->n, n = s->n;
    for (int b = 0; b < nb; ++b) { bots->a[b]->tmp = b; out[b].n0 = out[b].n = out[b].sy = 0; }
    if (s->cfg.nthreads <= 1) {
        for (int i = 0; i < n; ++i) {
            const int b = tree_bn(t, x + (size_t)i * p, xi)->tmp;
            out[b].n0 += 1; out[b].n += phi[i]; out[b].sy += phi[i] * R[i];
        }
        return;
    }
#pragma omp parallel num_threads(s->cfg.nthreads)
    {
        sinfo *loc = (sinfo *)calloc((size_t)nb, sizeof(sinfo));
#pragma omp for schedule(static)
        for (int i = 0; i < n; ++i) {
            const int b = tree_bn(t, x + (size_t)i * p, xi)->tmp;
            loc[b].n0 += 1; loc[b].n += phi[i]; loc[b].sy += phi[i] * R[i];
        }
#pragma omp critical
        for (int b = 0; b < nb; ++b) { out[b].n0 += loc[b].n0; out[b].n += loc[b].n; out[b].sy += loc[b].sy; }
        free(loc);
    }
}

static double pgrow_at_depth(double alpha, double beta, int depth) { return alpha / pow(1.0 + (double)depth, beta); }

/* Structural part of the MH ratio (alpha1) of a BIRTH at leaf nx with rule (v, c) [bcf-recall, SURVEY A.3 step 2]:
 * PGnx (1-PGly)(1-PGry) PDy Pnogy / ((1-PGnx) PBx Pbotx). */
static double birth_alpha1(const orc_sampler *s, node *t, node *nx, int v, int c, const xinfo *xi, double alpha, double beta)
{
    nvec bots = {0}, good = {0}, nogs = {0};
    tree_bots(t, &bots);
    for (int b = 0; b < bots.n; ++b) if (cansplit(bots.a[b], xi)) nvec_push(&good, bots.a[b]);
    tree_nogs(t, &nogs);
    const int single = (t->l == NULL);
    const double PBx = single ? 1.0 : 0.5;          /* a birth is only proposed when some leaf can split */
    const int ngv = goodvars(nx, xi, NULL);
    int lo, hi; node_rg(nx, v, xi, &lo, &hi);
    const int dnx = node_depth(nx);
    const double PGnx = pgrow_at_depth(alpha, beta, dnx);
    double PGly, PGry;
    const int y_capped = (s->cfg.max_leaves > 0 && bots.n + 1 >= s->cfg.max_leaves);
    const int child_depth_ok = (dnx + 1 < ORC_MAX_DEPTH) && !y_capped;
    if (!child_depth_ok) { PGly = PGry = 0.0; }
    else if (ngv > 1) { PGly = PGry = pgrow_at_depth(alpha, beta, dnx + 1); }
    else {
        PGly = (c - 1 < lo) ? 0.0 : pgrow_at_depth(alpha, beta, dnx + 1);
        PGry = (hi < c + 1) ? 0.0 : pgrow_at_depth(alpha, beta, dnx + 1);
    }
    double PDy;
    if (y_capped) PDy = 1.0;
    else if (good.n > 1) PDy = 0.5;
    else PDy = (PGly == 0.0 && PGry == 0.0) ? 1.0 : 0.5;
    double Pnogy;
    if (!nx->p) Pnogy = 1.0;
    else Pnogy = node_is_nog(nx->p) ? 1.0 / nogs.n : 1.0 / (nogs.n + 1.0);
    const double Pbotx = 1.0 / good.n;
    free(bots.a); free(good.a); free(nogs.a);
    return (PGnx * (1.0 - PGly) * (1.0 - PGry) * PDy * Pnogy) / ((1.0 - PGnx) * PBx * Pbotx);
}
/* ... and of the DEATH of nog node nx: (1-PGny) PBy Pboty / (PGny (1-PGlx)(1-PGrx) PDx Pnogx). */
static double death_alpha1(const orc_sampler *s, node *t, node *nx, const xinfo *xi, double alpha, double beta)
{
    nvec bots = {0}, good = {0}, nogs = {0};
    tree_bots(t, &bots);
    for (int b = 0; b < bots.n; ++b) if (cansplit(bots.a[b], xi)) nvec_push(&good, bots.a[b]);
    tree_nogs(t, &nogs);
    const int capped = (s->cfg.max_leaves > 0 && bots.n >= s->cfg.max_leaves);
    const int ngood_x = capped ? 0 : good.n;
    const double PBx = (ngood_x == 0) ? 0.0 : 0.5;  /* the tree has a nog node, so it is not a single leaf */
    const int dny = node_depth(nx);
    const double PGny = pgrow_at_depth(alpha, beta, dny);
    /* pgrow of the two children as they stand in the current tree */
    const int rl = cansplit(nx->l, xi), rr = cansplit(nx->r, xi);
    const double PGlx = (rl && !capped) ? pgrow_at_depth(alpha, beta, dny + 1) : 0.0;
    const double PGrx = (rr && !capped) ? pgrow_at_depth(alpha, beta, dny + 1) : 0.0;
    const double PBy = nx->p ? 0.5 : 1.0;
    /* goodbots of the proposed tree: lose the children, gain nx (it splits on v, so v is usable there) */
    const int ngood = good.n - rl - rr + 1;
    const double Pboty = 1.0 / ngood;
    const double PDx = 1.0 - PBx;
    const double Pnogx = 1.0 / nogs.n;
    free(bots.a); free(good.a); free(nogs.a);
    return ((1.0 - PGny) * PBy * Pboty) / (PGny * (1.0 - PGlx) * (1.0 - PGrx) * PDx * Pnogx);
}

/* One birth/death Metropolis-Hastings step on tree *tp (bcf's bd/bdhet) [bcf-recall, SURVEY A.3 step 2] */
static void bd(orc_sampler *s, node **tp, uint32_t tree_id, const double *x, int p, const xinfo *xi,
               double alpha, double beta, double tau, const double *phi, const double *R)
{
    node *t = *tp;
    int32_t *tr = s->trace + 5 * tree_id;
    tr[0] = tr[1] = tr[2] = 0; tr[3] = tr[4] = -1;
    s->trace_logr[tree_id] = NAN;

    nvec bots = {0}, good = {0}, nogs = {0};
    tree_bots(t, &bots);
    /* good = leaves with a usable variable (bcf's goodbots).  A tree at the structural leaf cap
     * (cfg.max_leaves, 0 = none = the reference) has no splittable leaf at all. */
    for (int b = 0; b < bots.n; ++b) if (cansplit(bots.a[b], xi)) nvec_push(&good, bots.a[b]);
    const int capped = (s->cfg.max_leaves > 0 && bots.n >= s->cfg.max_leaves);
    const int ngood_x = capped ? 0 : good.n;
    const int single = (t->l == NULL);
    const double PBx = (ngood_x == 0) ? 0.0 : (single ? 1.0 : 0.5);
    double u_move, u_acc;
    rng_u2(&s->rng, s->sweep, tree_id, PUR_MOVE, 0, &u_move, &u_acc);

    if (u_move < PBx) {
        /* ---- birth proposal ---- */
        double ua, ub;
        rng_u2(&s->rng, s->sweep, tree_id, PUR_NODE, 0, &ua, &ub);
        int ni = (int)floor(ua * good.n); if (ni >= good.n) ni = good.n - 1;
        node *nx = good.a[ni];
        int *gv = (int *)malloc(sizeof(int) * (size_t)xi->p);
        const int ngv = goodvars(nx, xi, gv);
        int vi = (int)floor(ub * ngv); if (vi >= ngv) vi = ngv - 1;
        const int v = gv[vi];
        free(gv);
        int lo, hi; node_rg(nx, v, xi, &lo, &hi);
        rng_u2(&s->rng, s->sweep, tree_id, PUR_CUT, 0, &ua, &ub);
        int c = lo + (int)floor(ua * (hi - lo + 1)); if (c > hi) c = hi;

        const double alpha1 = birth_alpha1(s, t, nx, v, c, xi, alpha, beta);

        sinfo sl, sr;
        getsuff_birth(s, t, nx, v, c, x, p, xi, phi, R, &sl, &sr);
        s->n_birth_prop++;
        tr[0] = 1; tr[2] = (int32_t)node_heap_id(nx); tr[3] = v; tr[4] = c;
        if (sl.n0 >= s->cfg.min_leaf && sr.n0 >= s->cfg.min_leaf) {
            const double lill = lil(sl.n, sl.sy, tau), lilr = lil(sr.n, sr.sy, tau);
            const double lilt = lil(sl.n + sr.n, sl.sy + sr.sy, tau);
            const double logr = log(alpha1) + (lill + lilr - lilt);
            s->trace_logr[tree_id] = logr;
            const double a = logr >= 0 ? 1.0 : exp(logr);
            if (u_acc < a) {
                nx->v = v; nx->c = c;
                nx->l = node_new(nx, nx->mu); nx->r = node_new(nx, nx->mu);
                s->n_birth_acc++; tr[1] = 1;
            }
        }
    } else if (!single) {
        /* ---- death proposal ---- */
        tree_nogs(t, &nogs);
        double ua, ub;
        rng_u2(&s->rng, s->sweep, tree_id, PUR_NODE, 0, &ua, &ub);
        int ni = (int)floor(ua * nogs.n); if (ni >= nogs.n) ni = nogs.n - 1;
        node *nx = nogs.a[ni];
        const double alpha1 = death_alpha1(s, t, nx, xi, alpha, beta);
        sinfo sl, sr;
        getsuff_death(s, t, nx->l, nx->r, x, p, xi, phi, R, &sl, &sr);
        s->n_death_prop++;
        tr[0] = 2; tr[2] = (int32_t)node_heap_id(nx); tr[3] = nx->v; tr[4] = nx->c;
        const double lill = lil(sl.n, sl.sy, tau), lilr = lil(sr.n, sr.sy, tau);
        const double lilt = lil(sl.n + sr.n, sl.sy + sr.sy, tau);
        const double logr = log(alpha1) + (lilt - lill - lilr);
        s->trace_logr[tree_id] = logr;
        const double a = logr >= 0 ? 1.0 : exp(logr);
        if (u_acc < a) {
            /* merged leaf keeps a value; it is redrawn by drmu straight after */
            nx->mu = (sl.n0 * nx->l->mu + sr.n0 * nx->r->mu) / (sl.n0 + sr.n0);
            free(nx->l); free(nx->r); nx->l = nx->r = NULL; nx->v = nx->c = -1;
            s->n_death_acc++; tr[1] = 1;
        }
    }
    free(bots.a); free(good.a); free(nogs.a);
}

/* drmu / drmuhet: redraw every leaf mean from its conditional [SURVEY A.3 step 3] */
static void drmu(orc_sampler *s, node *t, uint32_t tree_id, const double *x, int p, const xinfo *xi,
                 double tau, const double *phi, const double *R)
{
    nvec bots = {0};
    tree_bots(t, &bots);
    sinfo *sv = (sinfo *)malloc(sizeof(sinfo) * (size_t)bots.n);
    allsuff(s, t, &bots, x, p, xi, phi, R, sv);
    const double a = 1.0 / (tau * tau);
    for (int b = 0; b < bots.n; ++b) {
        const double prec = a + sv[b].n;
        const double z = rng_normal(&s->rng, s->sweep, tree_id, PUR_LEAF, node_heap_id(bots.a[b]));
        bots.a[b]->mu = sv[b].sy / prec + z / sqrt(prec);
    }
    free(sv); free(bots.a);
}

/* ------------------------------------------------------------------------- */
/* Public sampler API                                                         */
/* ------------------------------------------------------------------------- */
ORC_API orc_sampler *orc_create(const orc_config *cfg, int n, int ntrt, int p_con, int p_mod, const double *y,
                                const double *xcon_rowmajor, const double *xmod_rowmajor,
                                const double *cuts_con, const int32_t *off_con,
                                const double *cuts_mod, const int32_t *off_mod)
{
    orc_sampler *s = (orc_sampler *)calloc(1, sizeof(orc_sampler));
    s->cfg = *cfg;
    if (s->cfg.nthreads < 1) s->cfg.nthreads = 1;
    s->n = n; s->ntrt = ntrt; s->p_con = p_con; s->p_mod = p_mod;
#define DUP(dst, src, cnt, T) do { dst = (T *)malloc(sizeof(T) * (size_t)(cnt)); memcpy(dst, src, sizeof(T) * (size_t)(cnt)); } while (0)
    DUP(s->y, y, n, double);
    DUP(s->xc, xcon_rowmajor, (size_t)n * p_con, double);
    DUP(s->xm, xmod_rowmajor, (size_t)n * p_mod, double);
    DUP(s->off_c, off_con, p_con + 1, int32_t);
    DUP(s->off_m, off_mod, p_mod + 1, int32_t);
    DUP(s->cuts_c, cuts_con, off_con[p_con], double);
    DUP(s->cuts_m, cuts_mod, off_mod[p_mod], double);
#undef DUP
    s->xi_c.p = p_con; s->xi_c.cuts = s->cuts_c; s->xi_c.off = s->off_c;
    s->xi_m.p = p_mod; s->xi_m.cuts = s->cuts_m; s->xi_m.off = s->off_m;
    const int T = cfg->ntree_con + cfg->ntree_mod;
    s->trace = (int32_t *)calloc((size_t)5 * T, sizeof(int32_t));
    s->trace_logr = (double *)calloc((size_t)T, sizeof(double));
    s->allfit = (double *)malloc(sizeof(double) * n); s->allfit_con = (double *)malloc(sizeof(double) * n);
    s->allfit_mod = (double *)malloc(sizeof(double) * n); s->ftemp = (double *)malloc(sizeof(double) * n);
    s->rtemp = (double *)malloc(sizeof(double) * n); s->weight = (double *)malloc(sizeof(double) * n);
    /* initial state [SURVEY A.2] */
    double ybar = 0; for (int i = 0; i < n; ++i) ybar += y[i]; ybar /= n;
    s->sigma = cfg->sigma_init; s->mscale = 1.0; s->bscale1 = 0.5; s->bscale0 = -0.5;
    s->delta_con = s->delta_mod = 1.0;
    s->tau_con = cfg->con_sd / (sqrt(s->delta_con) * sqrt((double)cfg->ntree_con));
    s->tau_mod = cfg->ntree_mod > 0 ? cfg->mod_sd / (sqrt(s->delta_mod) * sqrt((double)cfg->ntree_mod)) : 1.0;
    s->t_con = (node **)malloc(sizeof(node *) * (size_t)cfg->ntree_con);
    s->t_mod = (node **)malloc(sizeof(node *) * (size_t)(cfg->ntree_mod > 0 ? cfg->ntree_mod : 1));
    for (int j = 0; j < cfg->ntree_con; ++j) s->t_con[j] = node_new(NULL, ybar / cfg->ntree_con);
    const double trt_init = 1.0;
    for (int j = 0; j < cfg->ntree_mod; ++j) s->t_mod[j] = node_new(NULL, trt_init / cfg->ntree_mod);
    for (int i = 0; i < n; ++i) {
        s->allfit_con[i] = ybar;
        s->allfit_mod[i] = cfg->ntree_mod > 0 ? (i < ntrt ? s->bscale1 : s->bscale0) * trt_init : 0.0;
        s->allfit[i] = s->allfit_con[i] + s->allfit_mod[i];
    }
    rng_init(&s->rng, cfg->seed, cfg->chain);
    s->sweep = 0;
    return s;
}

ORC_API void orc_destroy(orc_sampler *s)
{
    if (!s) return;
    for (int j = 0; j < s->cfg.ntree_con; ++j) tree_free(s->t_con[j]);
    for (int j = 0; j < s->cfg.ntree_mod; ++j) tree_free(s->t_mod[j]);
    free(s->t_con); free(s->t_mod); free(s->y); free(s->xc); free(s->xm); free(s->cuts_c); free(s->cuts_m);
    free(s->off_c); free(s->off_m); free(s->allfit); free(s->allfit_con); free(s->allfit_mod); free(s->ftemp);
    free(s->rtemp); free(s->weight); free(s->trace); free(s->trace_logr); free(s->ybin); free(s->rowid); free(s);
}

/* One full Gibbs sweep [bcf-recall: the body of the outer loop of bcfoverparRcppClean; SURVEY 3.3, A.3, A.4] */
static void sweep_once(orc_sampler *s)
{
    const int n = s->n, ntrt = s->ntrt;
    const orc_config *c = &s->cfg;
    const int nth = c->nthreads;
    /* ---- prognostic (mu) trees ---- */
    for (int j = 0; j < c->ntree_con; ++j) {
        fit_tree(s, s->t_con[j], s->xc, s->p_con, &s->xi_c, s->ftemp);
        const double ms = s->mscale, ph = ms * ms / (s->sigma * s->sigma);
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int k = 0; k < n; ++k) {
            s->allfit[k] -= ms * s->ftemp[k];
            s->allfit_con[k] -= ms * s->ftemp[k];
            s->rtemp[k] = (s->y[k] - s->allfit[k]) / ms;
            s->weight[k] = ph;
        }
        bd(s, &s->t_con[j], (uint32_t)j, s->xc, s->p_con, &s->xi_c, c->alpha_con, c->beta_con, s->tau_con, s->weight, s->rtemp);
        drmu(s, s->t_con[j], (uint32_t)j, s->xc, s->p_con, &s->xi_c, s->tau_con, s->weight, s->rtemp);
        fit_tree(s, s->t_con[j], s->xc, s->p_con, &s->xi_c, s->ftemp);
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int k = 0; k < n; ++k) { s->allfit[k] += ms * s->ftemp[k]; s->allfit_con[k] += ms * s->ftemp[k]; }
    }
    /* ---- treatment-effect (tau) trees ---- */
    for (int j = 0; j < c->ntree_mod; ++j) {
        const uint32_t tid = (uint32_t)(c->ntree_con + j);
        fit_tree(s, s->t_mod[j], s->xm, s->p_mod, &s->xi_m, s->ftemp);
        const double b1 = s->bscale1, b0 = s->bscale0, s2 = s->sigma * s->sigma;
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int k = 0; k < n; ++k) {
            const double bs = k < ntrt ? b1 : b0;
            s->allfit[k] -= bs * s->ftemp[k];
            s->allfit_mod[k] -= bs * s->ftemp[k];
            s->rtemp[k] = (s->y[k] - s->allfit[k]) / bs;
            s->weight[k] = bs * bs / s2;
        }
        bd(s, &s->t_mod[j], tid, s->xm, s->p_mod, &s->xi_m, c->alpha_mod, c->beta_mod, s->tau_mod, s->weight, s->rtemp);
        drmu(s, s->t_mod[j], tid, s->xm, s->p_mod, &s->xi_m, s->tau_mod, s->weight, s->rtemp);
        fit_tree(s, s->t_mod[j], s->xm, s->p_mod, &s->xi_m, s->ftemp);
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int k = 0; k < n; ++k) {
            const double bs = k < ntrt ? b1 : b0;
            s->allfit[k] += bs * s->ftemp[k]; s->allfit_mod[k] += bs * s->ftemp[k];
        }
    }
    /* ---- tau scale (bscale1, bscale0): half-normal parameter expansion [SURVEY A.4] ---- */
    const double s2 = s->sigma * s->sigma;
    if (c->use_tauscale && c->ntree_mod > 0) {
        double ww1 = 0, rw1 = 0, ww0 = 0, rw0 = 0;
        for (int k = 0; k < n; ++k) {
            const double bs = k < ntrt ? s->bscale1 : s->bscale0;
            const double am = s->allfit_mod[k];
            const double iw = am * am / (s2 * bs * bs);                     /* 1/w */
            const double rw = (s->y[k] - s->allfit_con[k]) * am / (s2 * bs); /* r/w */
            if (k < ntrt) { ww1 += iw; rw1 += rw; } else { ww0 += iw; rw0 += rw; }
        }
        const double prec = 2.0; /* bscale_prec */
        const double b1o = s->bscale1, b0o = s->bscale0;
        double v = 1.0 / (ww1 + prec);
        s->bscale1 = v * rw1 + rng_normal(&s->rng, s->sweep, TREE_SWEEP_LEVEL, PUR_BSCALE1, 0) * sqrt(v);
        v = 1.0 / (ww0 + prec);
        s->bscale0 = v * rw0 + rng_normal(&s->rng, s->sweep, TREE_SWEEP_LEVEL, PUR_BSCALE0, 0) * sqrt(v);
        for (int k = 0; k < ntrt; ++k) s->allfit_mod[k] = s->allfit_mod[k] * s->bscale1 / b1o;
        for (int k = ntrt; k < n; ++k) s->allfit_mod[k] = s->allfit_mod[k] * s->bscale0 / b0o;
        /* half-normal tau scale: delta_mod stays 1, tau_mod unchanged */
    }
    /* ---- mu scale (mscale) + half-Cauchy mixing (delta_con) ---- */
    if (c->use_muscale) {
        double ww = 0, rw = 0;
        for (int k = 0; k < n; ++k) {
            const double ac = s->allfit_con[k];
            ww += ac * ac / (s2 * s->mscale * s->mscale);
            rw += (s->y[k] - s->allfit_mod[k]) * ac / (s2 * s->mscale);
        }
        const double mo = s->mscale, v = 1.0 / (ww + 1.0 /* mscale_prec */);
        s->mscale = v * rw + rng_normal(&s->rng, s->sweep, TREE_SWEEP_LEVEL, PUR_MSCALE, 0) * sqrt(v);
        for (int k = 0; k < n; ++k) s->allfit_con[k] = s->allfit_con[k] * s->mscale / mo;
        double ssq = 0, cnt = 0;
        for (int j = 0; j < c->ntree_con; ++j) {
            nvec b = {0}; tree_bots(s->t_con[j], &b);
            for (int q = 0; q < b.n; ++q) { const double m = b.a[q]->mu; ssq += m * m / (s->tau_con * s->tau_con); cnt += 1; }
            free(b.a);
        }
        s->delta_con = rng_gamma(&s->rng, s->sweep, PUR_DELTA_CON, 0.5 * (1.0 + cnt)) / (0.5 * (1.0 + ssq));
        s->tau_con = c->con_sd / (sqrt(s->delta_con) * sqrt((double)c->ntree_con));
    }
    /* ---- sigma ---- */
    double rss = 0;
    for (int k = 0; k < n; ++k) {
        s->allfit[k] = s->allfit_con[k] + s->allfit_mod[k];
        const double e = s->y[k] - s->allfit[k];
        rss += e * e;
    }
    if (!s->probit) {   /* probit BART: the latent has unit variance by construction */
        const double chi2 = 2.0 * rng_gamma(&s->rng, s->sweep, PUR_SIGMA, 0.5 * (c->nu + n));
        s->sigma = sqrt((c->nu * c->lambda + rss) / chi2);
    }
    s->sweep++;
}

ORC_API void orc_sweeps(orc_sampler *s, int nsweeps) { for (int i = 0; i < nsweeps; ++i) sweep_once(s); }

/* Run burn-in + sampling the way bcf does and accumulate what BayesIV consumes.
 * tau_draws / mu_draws / yhat_draws: nsim x n row-major or NULL; *_mean: length n or NULL;
 * sigma/mscale_out/bscale_out: length nsim or NULL.  Values are on the standardised scale. */
ORC_API int orc_run(orc_sampler *s, int nburn, int nsim, int nthin, double *tau_draws, double *mu_draws,
                    double *yhat_draws, double *tau_mean, double *mu_mean, double *yhat_mean, double *sigma,
                    double *msd, double *bsd)
{
    const int n = s->n, total = nsim * nthin + nburn;
    int saved = 0;
    if (tau_mean) memset(tau_mean, 0, sizeof(double) * n);
    if (mu_mean) memset(mu_mean, 0, sizeof(double) * n);
    if (yhat_mean) memset(yhat_mean, 0, sizeof(double) * n);
    for (int it = 0; it < total; ++it) {
        sweep_once(s);
        if (it >= nburn && it % nthin == 0 && saved < nsim) {
            for (int k = 0; k < n; ++k) {
                const double bs = k < s->ntrt ? s->bscale1 : s->bscale0;
                const double b = s->cfg.ntree_mod > 0 ? (s->bscale1 - s->bscale0) * s->allfit_mod[k] / bs : 0.0;
                if (tau_draws) tau_draws[(size_t)saved * n + k] = b;
                if (mu_draws) mu_draws[(size_t)saved * n + k] = s->allfit_con[k];
                if (yhat_draws) yhat_draws[(size_t)saved * n + k] = s->allfit[k];
                if (tau_mean) tau_mean[k] += b;
                if (mu_mean) mu_mean[k] += s->allfit_con[k];
                if (yhat_mean) yhat_mean[k] += s->allfit[k];
            }
            if (sigma) sigma[saved] = s->sigma;
            if (msd) msd[saved] = fabs(s->mscale) * s->cfg.con_sd;
            if (bsd) bsd[saved] = fabs(s->bscale1 - s->bscale0) * s->cfg.mod_sd;
            ++saved;
        }
    }
    if (saved > 0) for (int k = 0; k < n; ++k) {
        if (tau_mean) tau_mean[k] /= saved;
        if (mu_mean) mu_mean[k] /= saved;
        if (yhat_mean) yhat_mean[k] /= saved;
    }
    return saved;
}

/* The same run, evaluated at NEW covariate rows (bcf 2.x's predict()): per saved sweep mu(x) = mscale * sum of the mu
 * trees' leaves, tau(x) = (bscale1 - bscale0) * sum of the tau trees' leaves, by walking the pointer trees over the raw
 * doubles.  xcon_new / xmod_new row-major n_new x p; outputs: means over the saved sweeps, optional draws (nsim x n_new). */
ORC_API int orc_run_predict(orc_sampler *s, int nburn, int nsim, int nthin, int n_new, const double *xcon_new,
                            const double *xmod_new, double *tau_mean, double *mu_mean, double *tau_draws)
{
    const int total = nsim * nthin + nburn;
    int saved = 0;
    memset(tau_mean, 0, sizeof(double) * (size_t)n_new); memset(mu_mean, 0, sizeof(double) * (size_t)n_new);
    for (int it = 0; it < total; ++it) {
        sweep_once(s);
        if (it >= nburn && it % nthin == 0 && saved < nsim) {
            for (int i = 0; i < n_new; ++i) {
                double gc = 0, gm = 0;
                for (int j = 0; j < s->cfg.ntree_con; ++j) gc += tree_bn(s->t_con[j], xcon_new + (size_t)i * s->p_con, &s->xi_c)->mu;
                for (int j = 0; j < s->cfg.ntree_mod; ++j) gm += tree_bn(s->t_mod[j], xmod_new + (size_t)i * s->p_mod, &s->xi_m)->mu;
                const double tau = (s->bscale1 - s->bscale0) * gm, mu = s->mscale * gc;
                tau_mean[i] += tau; mu_mean[i] += mu;
                if (tau_draws) tau_draws[(size_t)saved * n_new + i] = tau;
            }
            ++saved;
        }
    }
    if (saved > 0) for (int i = 0; i < n_new; ++i) { tau_mean[i] /= saved; mu_mean[i] /= saved; }
    return saved;
}

/* ---- state getters (parity tests) ---- */
ORC_API void orc_get_scalars(const orc_sampler *s, double out[8])
{
    out[0] = s->sigma; out[1] = s->mscale; out[2] = s->bscale1; out[3] = s->bscale0;
    out[4] = s->delta_con; out[5] = s->tau_con; out[6] = s->tau_mod; out[7] = (double)s->sweep;
}
ORC_API void orc_get_fits(const orc_sampler *s, double *allfit_con, double *allfit_mod)
{
    if (allfit_con) memcpy(allfit_con, s->allfit_con, sizeof(double) * s->n);
    if (allfit_mod) memcpy(allfit_mod, s->allfit_mod, sizeof(double) * s->n);
}
ORC_API void orc_get_counters(const orc_sampler *s, int64_t out[4])
{
    out[0] = s->n_birth_prop; out[1] = s->n_birth_acc; out[2] = s->n_death_prop; out[3] = s->n_death_acc;
}
ORC_API void orc_get_trace(const orc_sampler *s, int32_t *trace5, double *logr)
{
    const int T = s->cfg.ntree_con + s->cfg.ntree_mod;
    memcpy(trace5, s->trace, sizeof(int32_t) * 5 * (size_t)T);
    memcpy(logr, s->trace_logr, sizeof(double) * (size_t)T);
}
static node *get_tree(const orc_sampler *s, int tree_id)
{
    return tree_id < s->cfg.ntree_con ? s->t_con[tree_id] : s->t_mod[tree_id - s->cfg.ntree_con];
}
ORC_API int orc_tree_size(const orc_sampler *s, int tree_id) { return tree_size(get_tree(s, tree_id)); }
static void ser(const node *n, uint32_t id, uint32_t *heap, int32_t *var, int32_t *cut, double *mu, int *k)
{
    heap[*k] = id; var[*k] = n->v; cut[*k] = n->c; mu[*k] = n->mu; ++*k;
    if (n->l) { ser(n->l, 2 * id, heap, var, cut, mu, k); ser(n->r, 2 * id + 1, heap, var, cut, mu, k); }
}
/* preorder serialisation: heap id, split var (-1 leaf), cut index, mu */
ORC_API int orc_get_tree(const orc_sampler *s, int tree_id, uint32_t *heap, int32_t *var, int32_t *cut, double *mu)
{
    int k = 0; ser(get_tree(s, tree_id), 1u, heap, var, cut, mu, &k); return k;
}
/* replace a tree by a serialised one (heap ids need not be in order) */
ORC_API int orc_set_tree(orc_sampler *s, int tree_id, int nnodes, const uint32_t *heap, const int32_t *var,
                         const int32_t *cut, const double *mu)
{
    node **slot = tree_id < s->cfg.ntree_con ? &s->t_con[tree_id] : &s->t_mod[tree_id - s->cfg.ntree_con];
    /* build by heap id */
    node *root = NULL;
    for (int pass = 0; pass < 32; ++pass) {
        for (int k = 0; k < nnodes; ++k) {
            const uint32_t id = heap[k];
            int depth = 0; for (uint32_t t = id; t > 1; t >>= 1) ++depth;
            if (depth != pass) continue;
            if (id == 1) { root = node_new(NULL, mu[k]); root->v = var[k]; root->c = cut[k]; continue; }
            if (!root) return -1;
            node *cur = root;
            for (int b = depth - 1; b >= 1; --b) { cur = ((id >> b) & 1u) ? cur->r : cur->l; if (!cur) return -1; }
            node *nn = node_new(cur, mu[k]); nn->v = var[k]; nn->c = cut[k];
            if (id & 1u) cur->r = nn; else cur->l = nn;
        }
    }
    if (!root) return -1;
    tree_free(*slot); *slot = root;
    return 0;
}

/* rebuild allfit_con / allfit_mod / allfit from the current trees (after orc_set_tree) */
ORC_API void orc_refit(orc_sampler *s)
{
    const int n = s->n;
    for (int k = 0; k < n; ++k) { s->allfit_con[k] = 0; s->allfit_mod[k] = 0; }
    for (int j = 0; j < s->cfg.ntree_con; ++j) {
        fit_tree(s, s->t_con[j], s->xc, s->p_con, &s->xi_c, s->ftemp);
        for (int k = 0; k < n; ++k) s->allfit_con[k] += s->mscale * s->ftemp[k];
    }
    for (int j = 0; j < s->cfg.ntree_mod; ++j) {
        fit_tree(s, s->t_mod[j], s->xm, s->p_mod, &s->xi_m, s->ftemp);
        for (int k = 0; k < n; ++k) s->allfit_mod[k] += (k < s->ntrt ? s->bscale1 : s->bscale0) * s->ftemp[k];
    }
    for (int k = 0; k < n; ++k) s->allfit[k] = s->allfit_con[k] + s->allfit_mod[k];
}

/* ---- standalone unit functions on a sampler's data (kernel-level parity) ---- */
/* leaf heap id each observation falls in, for the given tree of forest `forest` (0 con, 1 mod) */
ORC_API void orc_assign_leaves(orc_sampler *s, int tree_id, uint32_t *leaf_heap_id)
{
    const int con = tree_id < s->cfg.ntree_con;
    const double *x = con ? s->xc : s->xm; const int p = con ? s->p_con : s->p_mod;
    const xinfo *xi = con ? &s->xi_c : &s->xi_m;
    node *t = get_tree(s, tree_id);
    for (int i = 0; i < s->n; ++i) leaf_heap_id[i] = node_heap_id(tree_bn(t, x + (size_t)i * p, xi));
}
/* per-leaf (count, sum phi, sum phi*R) in DFS leaf order for arbitrary phi and R; returns #leaves */
ORC_API int orc_allsuff(orc_sampler *s, int tree_id, const double *phi, const double *R, uint32_t *heap,
                        double *n0, double *sphi, double *sphir)
{
    const int con = tree_id < s->cfg.ntree_con;
    const double *x = con ? s->xc : s->xm; const int p = con ? s->p_con : s->p_mod;
    const xinfo *xi = con ? &s->xi_c : &s->xi_m;
    node *t = get_tree(s, tree_id);
    nvec bots = {0}; tree_bots(t, &bots);
    sinfo *sv = (sinfo *)malloc(sizeof(sinfo) * (size_t)bots.n);
    const int keep = s->cfg.nthreads; s->cfg.nthreads = 1;
    allsuff(s, t, &bots, x, p, xi, phi, R, sv);
    s->cfg.nthreads = keep;
    for (int b = 0; b < bots.n; ++b) { heap[b] = node_heap_id(bots.a[b]); n0[b] = sv[b].n0; sphi[b] = sv[b].n; sphir[b] = sv[b].sy; }
    const int nb = bots.n; free(sv); free(bots.a); return nb;
}
/* birth-form getsuff for the leaf with heap id `leaf_heap` under rule (v,c): out = {l.n0,l.n,l.sy,r.n0,r.n,r.sy} */
ORC_API int orc_getsuff_birth(orc_sampler *s, int tree_id, uint32_t leaf_heap, int v, int c, const double *phi,
                              const double *R, double out[6])
{
    const int con = tree_id < s->cfg.ntree_con;
    const double *x = con ? s->xc : s->xm; const int p = con ? s->p_con : s->p_mod;
    const xinfo *xi = con ? &s->xi_c : &s->xi_m;
    node *t = get_tree(s, tree_id);
    nvec bots = {0}; tree_bots(t, &bots);
    node *nx = NULL;
    for (int b = 0; b < bots.n; ++b) if (node_heap_id(bots.a[b]) == leaf_heap) nx = bots.a[b];
    free(bots.a);
    if (!nx) return -1;
    sinfo sl, sr;
    const int keep = s->cfg.nthreads; s->cfg.nthreads = 1;
    getsuff_birth(s, t, nx, v, c, x, p, xi, phi, R, &sl, &sr);
    s->cfg.nthreads = keep;
    out[0] = sl.n0; out[1] = sl.n; out[2] = sl.sy; out[3] = sr.n0; out[4] = sr.n; out[5] = sr.sy;
    return 0;
}
ORC_API double orc_lil(double sum_phi, double sum_phi_r, double tau) { return lil(sum_phi, sum_phi_r, tau); }

/* log MH ratio of a NAMED move on the tree as it stands, for caller-supplied weights phi and partial residual R
 * (tests: reversibility, birth o death = 1).  move 1: birth at the leaf with heap id `heap` under rule (v, c);
 * move 2: death of the nog node `heap`.  out2 = {log alpha1 (structure), log likelihood ratio}.  Returns 0, or -1 when the
 * node does not exist / is not a leaf (birth) / not a nog (death). */
static node *find_heap(node *n, uint32_t id, uint32_t heap)
{
    if (!n) return NULL;
    if (id == heap) return n;
    node *f = find_heap(n->l, 2 * id, heap);
    return f ? f : find_heap(n->r, 2 * id + 1, heap);
}
ORC_API int orc_mh_logratio(orc_sampler *s, int tree_id, int move, uint32_t heap, int v, int c, const double *phi,
                            const double *R, double out2[2])
{
    const int con = tree_id < s->cfg.ntree_con;
    const double *x = con ? s->xc : s->xm; const int p = con ? s->p_con : s->p_mod;
    const xinfo *xi = con ? &s->xi_c : &s->xi_m;
    const double alpha = con ? s->cfg.alpha_con : s->cfg.alpha_mod, beta = con ? s->cfg.beta_con : s->cfg.beta_mod;
    const double tau = con ? s->tau_con : s->tau_mod;
    node *t = get_tree(s, tree_id);
    node *nx = find_heap(t, 1u, heap);
    if (!nx) return -1;
    sinfo sl, sr;
    const int keep = s->cfg.nthreads; s->cfg.nthreads = 1;
    if (move == 1) {
        if (nx->l) { s->cfg.nthreads = keep; return -1; }
        getsuff_birth(s, t, nx, v, c, x, p, xi, phi, R, &sl, &sr);
        out2[0] = log(birth_alpha1(s, t, nx, v, c, xi, alpha, beta));
        out2[1] = lil(sl.n, sl.sy, tau) + lil(sr.n, sr.sy, tau) - lil(sl.n + sr.n, sl.sy + sr.sy, tau);
    } else {
        if (!node_is_nog(nx)) { s->cfg.nthreads = keep; return -1; }
        getsuff_death(s, t, nx->l, nx->r, x, p, xi, phi, R, &sl, &sr);
        out2[0] = log(death_alpha1(s, t, nx, xi, alpha, beta));
        out2[1] = lil(sl.n + sr.n, sl.sy + sr.sy, tau) - lil(sl.n, sl.sy, tau) - lil(sr.n, sr.sy, tau);
    }
    s->cfg.nthreads = keep;
    return 0;
}
/* replace the outcome vector (joint-distribution tests redraw y given the parameters); the fits are untouched */
ORC_API void orc_set_y(orc_sampler *s, const double *y) { memcpy(s->y, y, sizeof(double) * (size_t)s->n); }
ORC_API void orc_get_allfit(const orc_sampler *s, double *allfit) { memcpy(allfit, s->allfit, sizeof(double) * (size_t)s->n); }

/* homoskedastic lil of BART/bcf 1.x (n, sum y, sum y^2, sigma, tau), for the lil == lilhet identity test */
ORC_API double orc_lil_homosked(double n, double sy, double sy2, double sigma, double tau)
{
    const double s2 = sigma * sigma, t2 = tau * tau, k = n * t2 + s2;
    const double suff = (sy2 - (t2 / k) * sy * sy) / s2;
    return -0.5 * n * log(2 * 3.14159265358979323846 * s2) + 0.5 * log(s2 / k) - 0.5 * suff;
}

/* all-candidate histogram for one forest's X: for leaf-slot labels `slot` (0..nleaf-1, or <0 to skip),
 * for every variable v and bin b = #{cut <= x}: count and sum of r.  Oracle for K2' (SURVEY 8a4).
 * cnt/sum layout: [leaf][var_bin_offset[v] + b], with var_bin_offset[v] = off[v] + v (ncut+1 bins per var). */
ORC_API void orc_hist_all(orc_sampler *s, int forest, const int32_t *slot, int nleaf, const double *r,
                          int64_t *cnt, double *sum)
{
    const double *x = forest == 0 ? s->xc : s->xm; const int p = forest == 0 ? s->p_con : s->p_mod;
    const xinfo *xi = forest == 0 ? &s->xi_c : &s->xi_m;
    const size_t nb = (size_t)xi->off[p] + p;
    memset(cnt, 0, sizeof(int64_t) * nb * nleaf); memset(sum, 0, sizeof(double) * nb * nleaf);
    for (int i = 0; i < s->n; ++i) {
        if (slot[i] < 0) continue;
        for (int v = 0; v < p; ++v) {
            const int nc = xi_ncut(xi, v); const double xv = x[(size_t)i * p + v];
            int b = 0; while (b < nc && xi_cut(xi, v, b) <= xv) ++b;
            const size_t k = (size_t)slot[i] * nb + (size_t)xi->off[v] + v + b;
            cnt[k] += 1; sum[k] += r[i];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Probit BART: the sampler behind bartCause::bartc for a 0/1 response, which the   */
/* reference calls at R/bcf_iv.R:78-80 (complier proportion: response w) and at      */
/* R/bcf_iv.R:198-202, R/bcf_itt.R:166-170 (binary = TRUE: response y).  bartCause   */
/* and dbarts are third-party, un-vendored and un-pinned like bcf (DESCRIPTION:22-24)*/
/* -> PARITY UNPINNED.  Restated: the published algorithm -- Chipman, George &       */
/* McCulloch (2010) BART with the Albert & Chib (1993) data augmentation:            */
/*   y_i = 1{ystar_i > 0},  ystar_i = offset + f(x_i) + eps_i,  eps ~ N(0, 1),       */
/*   f = sum of T trees, leaf values N(0, (3 / (k sqrt T))^2), k = 2,                */
/*   tree prior alpha (1 + d)^-beta = .95, 2;  sigma = 1 is not drawn.               */
/* One sweep = redraw every latent from its truncated normal given f, then the usual */
/* backfitting pass (birth/death + leaf draws: the SAME bd()/drmu() as above with     */
/* one forest, no scales).  bartc is an S-learner: the forest sees (x, pscore, z) and */
/* the individual effect of a draw is  Phi(offset + f(x, 1)) - Phi(offset + f(x, 0)).*/
/* ------------------------------------------------------------------------- */
/* standard normal cdf, accurate in both tails */
static double norm_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
/* its inverse: Wichura's AS 241 (PPND16), relative error ~1e-16 */
static double norm_inv(double p)
{
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r +
                        4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r +
                      1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
               (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r +
                    2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r +
                  4.2313330701600911252e1) * r + 1.0);
    }
    double r = q < 0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    double v;
    if (r <= 5.0) {
        r -= 1.6;
        v = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
                 1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r +
               4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
            (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
                 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r +
               2.05319162663775882187e0) * r + 1.0);
    } else {
        r -= 5.0;
        v = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
                 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r +
               5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
            (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
                 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
               5.99832206555887937690e-1) * r + 1.0);
    }
    return q < 0 ? -v : v;
}
ORC_API double orc_norm_cdf(double x) { return norm_cdf(x); }
ORC_API double orc_norm_inv(double p) { return norm_inv(p); }
/* one latent: ystar ~ N(m, 1) truncated to (0, inf) when y = 1, to (-inf, 0] when y = 0, by inversion of ONE uniform
 * (written on the side where the cdf is small, so it stays accurate however improbable the observed response is) */
static double probit_latent(double m, int y, double u)
{
    if (y) return m - norm_inv(u * norm_cdf(m));     /* eps = -eps', eps' < m */
    return m + norm_inv(u * norm_cdf(-m));           /* eps < -m */
}
ORC_API double orc_probit_latent_probe(double m, int y, double u) { return probit_latent(m, y, u); }

/* Switch a sampler (created with ntree_mod = 0, use_muscale = 0 and y = 0) to probit BART.  ybin / rowid: the 0/1
 * response and the caller's row number of every row, internal order. */
ORC_API void orc_probit_enable(orc_sampler *s, const uint8_t *ybin, const int32_t *rowid, double offset, int trt_col)
{
    s->probit = 1; s->offset = offset; s->trt_col = trt_col;
    s->ybin = (uint8_t *)malloc((size_t)s->n); memcpy(s->ybin, ybin, (size_t)s->n);
    s->rowid = (int32_t *)malloc(sizeof(int32_t) * (size_t)s->n); memcpy(s->rowid, rowid, sizeof(int32_t) * (size_t)s->n);
    s->sigma = 1.0;
}
/* the latent draw that opens sweep s->sweep: the working response is ystar - offset */
static void probit_draw_latents(orc_sampler *s)
{
    for (int i = 0; i < s->n; ++i) {
        double ua, ub;
        rng_u2(&s->rng, s->sweep, TREE_SWEEP_LEVEL, PUR_LATENT, (uint32_t)s->rowid[i], &ua, &ub);
        s->y[i] = probit_latent(s->offset + s->allfit[i], s->ybin[i], ua) - s->offset;
    }
}
/* f(x) of every row with the treatment column set to `trt` (0 / 1), by walking every tree over the raw doubles */
static void probit_fit_with_treatment(orc_sampler *s, double trt, double *out)
{
    const int p = s->p_con;
    double *xx = (double *)malloc(sizeof(double) * (size_t)p);
    for (int i = 0; i < s->n; ++i) {
        memcpy(xx, s->xc + (size_t)i * p, sizeof(double) * (size_t)p);
        if (s->trt_col >= 0) xx[s->trt_col] = trt;
        double f = 0;
        for (int j = 0; j < s->cfg.ntree_con; ++j) f += tree_bn(s->t_con[j], xx, &s->xi_c)->mu;
        out[i] = f;
    }
    free(xx);
}
/* nburn + nsim * nthin sweeps; saved sweeps accumulate, per row (internal order):
 *   ite_mean  = mean of Phi(offset + f(x, 1)) - Phi(offset + f(x, 0))   (NULL or trt_col < 0: skipped)
 *   prob_mean = mean of Phi(offset + f(x)) at the observed covariates */
ORC_API int orc_probit_run(orc_sampler *s, int nburn, int nsim, int nthin, double *ite_mean, double *prob_mean)
{
    const int n = s->n, total = nburn + nsim * nthin;
    int saved = 0;
    double *f1 = (double *)malloc(sizeof(double) * (size_t)n), *f0 = (double *)malloc(sizeof(double) * (size_t)n);
    if (ite_mean) memset(ite_mean, 0, sizeof(double) * (size_t)n);
    if (prob_mean) memset(prob_mean, 0, sizeof(double) * (size_t)n);
    for (int it = 0; it < total; ++it) {
        probit_draw_latents(s);
        sweep_once(s);
        if (it >= nburn && it % nthin == 0 && saved < nsim) {
            if (ite_mean && s->trt_col >= 0) {
                probit_fit_with_treatment(s, 1.0, f1);
                probit_fit_with_treatment(s, 0.0, f0);
                for (int i = 0; i < n; ++i) ite_mean[i] += norm_cdf(s->offset + f1[i]) - norm_cdf(s->offset + f0[i]);
            }
            if (prob_mean) for (int i = 0; i < n; ++i) prob_mean[i] += norm_cdf(s->offset + s->allfit[i]);
            ++saved;
        }
    }
    if (saved > 0) for (int i = 0; i < n; ++i) { if (ite_mean) ite_mean[i] /= saved; if (prob_mean) prob_mean[i] /= saved; }
    free(f1); free(f0);
    return saved;
}
/* one sweep (latent draw + backfitting), for move-by-move parity */
ORC_API void orc_probit_sweeps(orc_sampler *s, int k) { for (int i = 0; i < k; ++i) { probit_draw_latents(s); sweep_once(s); } }
ORC_API void orc_get_y(const orc_sampler *s, double *y) { memcpy(y, s->y, sizeof(double) * (size_t)s->n); }

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API int orc_config_size(void) { return (int)sizeof(orc_config); }
