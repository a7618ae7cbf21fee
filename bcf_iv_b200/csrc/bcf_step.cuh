// bcf_step.cuh -- the per-tree step of the BCF Gibbs sweep: propose -> (fused apply + sufficient statistics)
// -> decide, written as CTA-cooperative device functions that every engine uses:
//   * GRAPH engine: one kernel per tree (k_tree_step), the kernel boundary is the grid-wide dependency;
//   * PERSISTENT engines: the same functions inside one cooperative kernel with a grid barrier per tree.
// Every CTA replays the (cheap, scalar) Metropolis-Hastings decision from the same reduced bits and the same
// Philox counters, so no broadcast of the decision is needed.
//
// Everything that does not depend on the data -- the proposal, the structural part of the MH ratio, log(u) of the
// accept uniform and the standard normals of every leaf the move can produce -- is computed in propose(), which
// the persistent engines run while the grid barrier of the previous tree is in flight.  decide() is then only:
// read the reduced statistics, one log, a handful of divisions.
#pragma once
#include "bcf_device.cuh"

namespace bcf {

struct ApplyInfo {
    int32_t active;      // there is a previous tree whose update must be applied
    int32_t changed;     // structure changed (accepted move): leaf slots must be rewritten
    int32_t birth;       // accepted birth: observations of slot sx with bin > cut move to right_slot
    int32_t sx, cut, right_slot;
    int32_t is16;
    int32_t forest;      // forest of the previous tree
    const void* bins;    // column of the birth variable
    double dlt_right;
    double s1, s0;       // scale of the forest's fit for treated / control rows (mscale or bscale1/bscale0)
};

struct CurInfo {
    int32_t active, K;   // K = number of keys = nleaves (+1 for a birth proposal)
    int32_t birth, sx, cut, is16, forest, new_slot;
    const void* bins;
    double s1, s0;       // scale of this forest's fit for treated / control rows
};

// A tree in flight: its structure, its proposal, the leaf normals and the statistics-pass set-up.  Two stages, so
// that tree j+1 can be proposed while tree j still waits for its reduced statistics.
struct TreeStage {
    TreeRec tr;
    Proposal prop;
    double z[KEYS];                // standard normal of each (current, then new) leaf slot
    CurInfo ci;
};

struct StepShared {
    TreeStage st[2];
    double cnt[KEYS], sphi[KEYS], sphir[KEYS];
    double mu_old[MAXL], mu_new[MAXL];
    double dlt[KEYS];
    uint8_t remap[MAXL];
    int32_t ngood[MAXL];
    uint32_t lkey[MAXL];
    uint32_t nkey[MAXN];
    uint8_t nogflag[MAXN];
    ApplyInfo ai;
    int32_t accepted;
    int32_t bad;
    long long tab[KEYS][2][4];
};

__device__ inline void load_tree(TreeRec* dst, const TreeRec* src)
{
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(TreeRec) / 16); i += blockDim.x) d[i] = __ldcg(s + i);   // L2: written by another CTA
}
__device__ inline void store_tree(TreeRec* dst, const TreeRec* src)
{
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(TreeRec) / 16); i += blockDim.x) d[i] = s[i];
}

// ---------------------------------------------------------------------------------------------
// propose(): one birth/death proposal for the tree staged in sm.tr (bcf's bd(), structural half)
// [bcf-recall, SURVEY A.3 step 2].  All threads of the CTA take part; result in sm.prop, sm.z.
// ---------------------------------------------------------------------------------------------
__device__ inline void propose(StepShared& sm, int stage, const Dev& d, const ForestDev& F, int tree, uint32_t sweep)
{
    TreeStage& S = sm.st[stage];
    TreeRec& tr = S.tr;
    const int t = threadIdx.x;
    const int L = tr.nleaves, NN = tr.nnodes;
    const Rng rng{d.key0, d.key1};

    int my_good = 0, my_nog = 0;
    if (t < L) {
        const int node = tr.slot_node[t];
        sm.lkey[t] = dfs_key(tr.heap[node], tr.depth[node]);
        my_good = tree_ngoodvars(tr, F, node);
        sm.ngood[t] = my_good;
    }
    if (t < NN) {
        const int l = tr.left[t];
        my_nog = (l >= 0 && tr.left[l] < 0 && tr.left[tr.right[t]] < 0) ? 1 : 0;
        sm.nogflag[t] = (uint8_t)my_nog;
        sm.nkey[t] = dfs_key(tr.heap[t], tr.depth[t]);
    }
    if (t == 0) {
        Proposal& P = S.prop;
        P.move = 0; P.log_prior = 0.0; P.log_u = 0.0; P.var = -1; P.cut = -1; P.heap = 0; P.bins = nullptr; P.is16 = 0;
        P.slot = 0; P.slot_b = 0; P.node = 0;
    }
    const int ngoodbots = __syncthreads_count(t < L && my_good > 0);
    const int nnog = __syncthreads_count(t < NN && my_nog);

    // leaf normals of the current leaves, in parallel, off the critical path (one Philox block each)
    if (t >= 32 && t - 32 < L) {
        const int s = t - 32;
        S.z[s] = rng_normal(rng, sweep, (uint32_t)tree, PUR_LEAF, tr.heap[tr.slot_node[s]]);
    }

    const bool capped = (L >= d.max_leaves);
    const int ngood_x = capped ? 0 : ngoodbots;
    const bool single = (NN == 1);
    const double PBx = (ngood_x == 0) ? 0.0 : (single ? 1.0 : 0.5);
    double u_move, u_acc, ua, ub;
    rng_u2(rng, sweep, (uint32_t)tree, PUR_MOVE, 0, u_move, u_acc);
    rng_u2(rng, sweep, (uint32_t)tree, PUR_NODE, 0, ua, ub);

    if (u_move < PBx) {
        // ---- birth ----
        const int ni = min((int)floor(ua * ngood_x), ngood_x - 1);
        if (t < L && my_good > 0) {
            int grank = 0;
            for (int m = 0; m < L; ++m) grank += (sm.ngood[m] > 0 && sm.lkey[m] < sm.lkey[t]) ? 1 : 0;
            if (grank == ni) {
                const int nx = tr.slot_node[t];
                const int ngv = my_good;
                const int vi = min((int)floor(ub * ngv), ngv - 1);
                const int v = tree_pick_var(tr, F, nx, vi);
                int lo, hi;
                tree_rg(tr, F, nx, v, lo, hi);
                double uc, ud;
                rng_u2(rng, sweep, (uint32_t)tree, PUR_CUT, 0, uc, ud);
                const int c = lo + min((int)floor(uc * (hi - lo + 1)), hi - lo);
                const int dnx = tr.depth[nx];
                const double PGnx = F.pg[dnx];
                const bool y_capped = (L + 1 >= d.max_leaves);
                const bool child_ok = (dnx + 1 < MAX_DEPTH) && !y_capped;
                double PGly, PGry;
                if (!child_ok) { PGly = PGry = 0.0; }
                else if (ngv > 1) { PGly = PGry = F.pg[dnx + 1]; }
                else {
                    PGly = (c - 1 < lo) ? 0.0 : F.pg[dnx + 1];
                    PGry = (hi < c + 1) ? 0.0 : F.pg[dnx + 1];
                }
                double PDy;
                if (y_capped) PDy = 1.0;
                else if (ngoodbots > 1) PDy = 0.5;
                else PDy = (PGly == 0.0 && PGry == 0.0) ? 1.0 : 0.5;
                double Pnogy;
                const int par = tr.parent[nx];
                if (par < 0) Pnogy = 1.0;
                else Pnogy = sm.nogflag[par] ? 1.0 / nnog : 1.0 / (nnog + 1.0);
                const double Pbotx = 1.0 / ngoodbots;
                const double alpha1 = (PGnx * (1.0 - PGly) * (1.0 - PGry) * PDy * Pnogy) / ((1.0 - PGnx) * PBx * Pbotx);
                Proposal& P = S.prop;
                P.move = 1; P.node = nx; P.slot = t; P.slot_b = L; P.var = v; P.cut = c; P.heap = tr.heap[nx];
                P.bins = bin_column(F, v, d.n_pad, P.is16);
                P.log_prior = log(alpha1); P.log_u = log(u_acc);
                P.z_extra[0] = rng_normal(rng, sweep, (uint32_t)tree, PUR_LEAF, 2u * tr.heap[nx]);
                P.z_extra[1] = rng_normal(rng, sweep, (uint32_t)tree, PUR_LEAF, 2u * tr.heap[nx] + 1u);
            }
        }
    } else if (!single) {
        // ---- death ----
        const int ni = min((int)floor(ua * nnog), nnog - 1);
        if (t < NN && my_nog) {
            int nrank = 0;
            for (int m = 0; m < NN; ++m) nrank += (sm.nogflag[m] && sm.nkey[m] < sm.nkey[t]) ? 1 : 0;
            if (nrank == ni) {
                const int nx = t;
                const int a = tr.left[nx], b = tr.right[nx];
                const int sa = tr.node_slot[a], sb = tr.node_slot[b];
                const int dny = tr.depth[nx];
                const double PGny = F.pg[dny];
                const int rl = sm.ngood[sa] > 0, rr = sm.ngood[sb] > 0;
                const double PGlx = (rl && !capped) ? F.pg[dny + 1] : 0.0;
                const double PGrx = (rr && !capped) ? F.pg[dny + 1] : 0.0;
                const double PBy = tr.parent[nx] >= 0 ? 0.5 : 1.0;
                const int ngood_y = ngoodbots - rl - rr + 1;
                const double Pboty = 1.0 / ngood_y;
                const double PDx = 1.0 - PBx;
                const double Pnogx = 1.0 / nnog;
                const double alpha1 = ((1.0 - PGny) * PBy * Pboty) / (PGny * (1.0 - PGlx) * (1.0 - PGrx) * PDx * Pnogx);
                Proposal& P = S.prop;
                P.move = 2; P.node = nx; P.slot = sa; P.slot_b = sb; P.var = tr.var[nx]; P.cut = tr.cut[nx];
                P.heap = tr.heap[nx]; P.log_prior = log(alpha1); P.log_u = log(u_acc);
                P.z_extra[0] = rng_normal(rng, sweep, (uint32_t)tree, PUR_LEAF, tr.heap[nx]);
                P.z_extra[1] = 0.0;
            }
        }
    }
    __syncthreads();
}

// lil(l) + lil(r) - lil(l+r) with one logarithm (bcf's three lil calls fused) [SURVEY A.3]
__device__ __forceinline__ double lil_split_gain(double pl, double rl, double pr, double rr, double tau, double log_tau)
{
    const double a = 1.0 / (tau * tau);
    const double dl = a + pl, dr = a + pr, dt = a + (pl + pr);
    const double st = rl + rr;
    return -log_tau - 0.5 * log(dl * dr / dt) + 0.5 * (rl * rl / dl + rr * rr / dr - st * st / dt);
}

// ---------------------------------------------------------------------------------------------
// decide(): given the reduced statistics of tree `tree` (d.totals) and its proposal (sm.prop), take the MH
// decision, update the structure in sm.tr, redraw every leaf value and build the apply tables.
// bcf's bd() likelihood half + drmu() [bcf-recall, SURVEY A.3 steps 2-3].  `writer` CTA records the trace.
// ---------------------------------------------------------------------------------------------
__device__ inline void decide(StepShared& sm, int stage, const Dev& d, const ForestDev& F, int tree, const Scalars& sc,
                              bool writer)
{
    TreeStage& S = sm.st[stage];
    TreeRec& tr = S.tr;
    const Proposal P = S.prop;
    const int t = threadIdx.x;
    const int L = tr.nleaves;
    const int K = L + (P.move == 1 ? 1 : 0);
    const bool is_con = F.is_con != 0;
    const double s1 = is_con ? sc.mscale : sc.bscale1, s0 = is_con ? sc.mscale : sc.bscale0;
    const double s2 = sc.sigma * sc.sigma;
    const double phi1 = s1 * s1 / s2, phi0 = s0 * s0 / s2;
    const double tau = is_con ? sc.tau_con : sc.tau_mod;

    if (t < K) {
        const long long* q = d.totals + ((size_t)tree * KEYS + t) * 8;
        // reduced by integer atomics in L2: read there, never from a stale L1 line
        const longlong2 c0 = __ldcg(reinterpret_cast<const longlong2*>(q));        // control: count, limb0
        const longlong2 c1 = __ldcg(reinterpret_cast<const longlong2*>(q) + 1);    //          limb1, limb2
        const longlong2 t0 = __ldcg(reinterpret_cast<const longlong2*>(q) + 2);    // treated
        const longlong2 t1 = __ldcg(reinterpret_cast<const longlong2*>(q) + 3);
        const double n1 = (double)t0.x, n0 = (double)c0.x;
        const double S1 = limbs_to_double(t0.y, t1.x, t1.y);
        const double S0 = limbs_to_double(c0.y, c1.x, c1.y);
        const double mo = (t < L) ? tr.mu[t] : tr.mu[P.slot];
        sm.cnt[t] = n1 + n0;
        sm.sphi[t] = n1 * phi1 + n0 * phi0;
        sm.sphir[t] = phi1 * (S1 / s1 + n1 * mo) + phi0 * (S0 / s0 + n0 * mo);
    }
    if (t < MAXL) { sm.remap[t] = (uint8_t)t; if (t < L) sm.mu_old[t] = tr.mu[t]; }
    __syncthreads();

    if (t == 0) {
        int accepted = 0;
        double logr = nan("");
        const double log_tau = log(tau);
        if (P.move == 1) {
            const int sx = P.slot, kr = L;
            if (sm.cnt[sx] >= d.min_leaf && sm.cnt[kr] >= d.min_leaf) {
                logr = P.log_prior + lil_split_gain(sm.sphi[sx], sm.sphir[sx], sm.sphi[kr], sm.sphir[kr], tau, log_tau);
                accepted = (P.log_u < logr) ? 1 : 0;
            }
            if (accepted) {
                const int nx = P.node, a = tr.nnodes, b = tr.nnodes + 1;
                tr.left[nx] = (int16_t)a; tr.right[nx] = (int16_t)b;
                tr.var[nx] = (uint16_t)P.var; tr.cut[nx] = (uint16_t)P.cut;
                for (int q = 0; q < 2; ++q) {
                    const int ch = q ? b : a;
                    tr.parent[ch] = (int16_t)nx; tr.left[ch] = tr.right[ch] = -1;
                    tr.var[ch] = tr.cut[ch] = 0xFFFF;
                    tr.depth[ch] = tr.depth[nx] + 1;
                    tr.heap[ch] = 2u * tr.heap[nx] + (uint32_t)q;
                }
                tr.node_slot[a] = (uint8_t)sx; tr.slot_node[sx] = (int16_t)a;
                tr.node_slot[b] = (uint8_t)kr; tr.slot_node[kr] = (int16_t)b;
                tr.nnodes += 2; tr.nleaves = L + 1;
                S.z[sx] = P.z_extra[0]; S.z[kr] = P.z_extra[1];
            } else {   // the leaf stays whole: fold the would-be right child back in
                sm.cnt[sx] += sm.cnt[kr]; sm.sphi[sx] += sm.sphi[kr]; sm.sphir[sx] += sm.sphir[kr];
            }
        } else if (P.move == 2) {
            const int sa = P.slot, sb = P.slot_b;
            logr = P.log_prior - lil_split_gain(sm.sphi[sa], sm.sphir[sa], sm.sphi[sb], sm.sphir[sb], tau, log_tau);
            accepted = (P.log_u < logr) ? 1 : 0;
            if (accepted) {
                const int nx = P.node, na = tr.left[nx], nb = tr.right[nx];
                tr.left[nx] = tr.right[nx] = -1; tr.var[nx] = tr.cut[nx] = 0xFFFF;
                tr.node_slot[nx] = (uint8_t)sa; tr.slot_node[sa] = (int16_t)nx;
                sm.cnt[sa] += sm.cnt[sb]; sm.sphi[sa] += sm.sphi[sb]; sm.sphir[sa] += sm.sphir[sb];
                S.z[sa] = P.z_extra[0];
                sm.remap[sb] = (uint8_t)sa;
                const int last = L - 1;
                if (sb != last) {   // keep slots compact: the last slot moves into the hole
                    const int nd = tr.slot_node[last];
                    tr.slot_node[sb] = (int16_t)nd; tr.node_slot[nd] = (uint8_t)sb;
                    sm.cnt[sb] = sm.cnt[last]; sm.sphi[sb] = sm.sphi[last]; sm.sphir[sb] = sm.sphir[last];
                    S.z[sb] = S.z[last];
                    for (int s = 0; s < L; ++s) if (sm.remap[s] == last) sm.remap[s] = (uint8_t)sb;
                }
                tr.nleaves = L - 1;
                tree_remove_node(tr, max(na, nb));
                tree_remove_node(tr, min(na, nb));
            }
        }
        sm.accepted = accepted;
        if (writer) {
            int32_t* trc = d.trace + 5 * tree;
            trc[0] = P.move; trc[1] = accepted; trc[2] = (int32_t)P.heap; trc[3] = P.var; trc[4] = P.cut;
            d.trace_logr[tree] = logr;
            if (P.move == 1) { atomicAdd((unsigned long long*)&d.counters[0], 1ull); if (accepted) atomicAdd((unsigned long long*)&d.counters[1], 1ull); }
            if (P.move == 2) { atomicAdd((unsigned long long*)&d.counters[2], 1ull); if (accepted) atomicAdd((unsigned long long*)&d.counters[3], 1ull); }
        }
    }
    __syncthreads();

    const int Lnew = tr.nleaves;
    if (t < Lnew) {
        const double prec = 1.0 / (tau * tau) + sm.sphi[t];
        const double m = sm.sphir[t] / prec + S.z[t] / sqrt(prec);
        if (!(m == m)) sm.bad = ERR_NAN;
        sm.mu_new[t] = m;
        tr.mu[t] = m;
    }
    __syncthreads();
    if (t < L) sm.dlt[t] = sm.mu_new[sm.remap[t]] - sm.mu_old[t];
    if (t == 0) {
        ApplyInfo& ai = sm.ai;
        ai.active = 1; ai.changed = sm.accepted; ai.birth = (sm.accepted && P.move == 1) ? 1 : 0;
        ai.sx = P.slot; ai.cut = P.cut; ai.bins = P.bins; ai.is16 = P.is16; ai.right_slot = L; ai.forest = is_con ? 0 : 1;
        ai.dlt_right = ai.birth ? (sm.mu_new[L] - sm.mu_old[P.slot]) : 0.0;
        ai.s1 = s1; ai.s0 = s0;
    }
    __syncthreads();
}

// Set up the statistics pass for the proposal in sm.prop (tree staged in sm.tr) and clear the CTA table.
__device__ inline void begin_stats(StepShared& sm, int stage, int forest, const Scalars& sc)
{
    const int t = threadIdx.x;
    TreeStage& S = sm.st[stage];
    const int K = S.tr.nleaves + (S.prop.move == 1 ? 1 : 0);
    if (t == 0) {
        CurInfo& ci = S.ci;
        ci.active = 1; ci.K = K; ci.birth = S.prop.move == 1; ci.sx = S.prop.slot; ci.cut = S.prop.cut;
        ci.bins = S.prop.bins; ci.is16 = S.prop.is16; ci.forest = forest; ci.new_slot = S.tr.nleaves;
        ci.s1 = forest == 0 ? sc.mscale : sc.bscale1; ci.s0 = forest == 0 ? sc.mscale : sc.bscale0;
    }
    long long* tb = &sm.tab[0][0][0];
    for (int i = t; i < K * 8; i += blockDim.x) tb[i] = 0;
}

// flush the CTA table into the global totals of `tree` (integer REDs: order-independent)
__device__ inline void flush_stats(StepShared& sm, int K, const Dev& d, int tree)
{
    const long long* tb = &sm.tab[0][0][0];
    long long* q = d.totals + (size_t)tree * KEYS * 8;
    for (int i = threadIdx.x; i < K * 8; i += blockDim.x) {
        const long long v = tb[i];
        if (v != 0) atomicAdd((unsigned long long*)&q[i], (unsigned long long)v);
    }
}

// apply the decided update of the previous tree to 8 observations: new slot and the change of the leaf value
__device__ __forceinline__ void apply8(const StepShared& sm, const ApplyInfo& ai, int (&sp)[8], const int (&bp)[8],
                                       double (&dl)[8])
{
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int s = sp[o];
        dl[o] = 0.0;
        if (s != PAD_SLOT) {
            double x = sm.dlt[s];
            int ns = sm.remap[s];
            if (ai.birth && s == ai.sx && bp[o] > ai.cut) { x = ai.dlt_right; ns = ai.right_slot; }
            dl[o] = x;
            sp[o] = ns;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One warp tile of the fused pass (state in HBM/L2): apply the previous tree's update to r (and its leaf slots),
// then accumulate the current tree's statistics.  r is the FULL residual y - allfit.  All loads are issued
// before the first dependent use.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void step_tile(StepShared& sm, const CurInfo& ci, const Dev& d, int prev, int cur, int tile,
                                          int lane, int& bad)
{
    const int64_t base = (int64_t)tile * TILE + lane * 8;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    double r[8];
    int sp[8], bp[8], key[8], bc[8];
    uint8_t* sl = d.slots + (int64_t)prev * d.n_pad + base;
    ld8_f64(d.r + base, r);
    if (ai.active) {
        ld8_u8(sl, sp);
        if (ai.birth) ld8_bins(ai.bins, ai.is16, base, bp);
    }
    if (ci.active) {
        ld8_u8(d.slots + (int64_t)cur * d.n_pad + base, key);
        if (ci.birth) ld8_bins(ci.bins, ci.is16, base, bc);
    }
    if (ai.active) {
        double dl[8];
        apply8(sm, ai, sp, bp, dl);
        const double sg = g ? ai.s1 : ai.s0;
#pragma unroll
        for (int o = 0; o < 8; ++o) r[o] -= sg * dl[o];
        if (ai.changed) st8_u8(sl, sp);
        st8_f64(d.r + base, r);
    }
    if (ci.active) {
        if (ci.birth) {
#pragma unroll
            for (int o = 0; o < 8; ++o) if (key[o] == ci.sx && bc[o] > ci.cut) key[o] = ci.new_slot;
        }
        accum_tile(sm.tab, key, r, g, ci.K, lane, bad);
    }
}

// moments of 8 observations into tab[m][g][1..3] (limbs); padding rows have y = G = 0
__device__ __forceinline__ void moments8(long long (*mt)[2][4], const double (&y)[8], const double (&gc)[8],
                                         const double (&gm)[8], int g, int lane, int& bad)
{
#pragma unroll
    for (int m = 0; m < NMOM; ++m) {
        int l0 = 0, l1 = 0, l2 = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const double a = (m < 3) ? y[o] : (m < 5 ? gc[o] : gm[o]);
            const double b = (m == 0) ? y[o] : (m == 1 || m == 3) ? gc[o] : gm[o];
            const long long v = to_fixed(a * b, bad);
            l0 += (int)(v & 0x1FFFFF); l1 += (int)((v >> LIMB_BITS) & 0x1FFFFF); l2 += (int)(v >> (2 * LIMB_BITS));
        }
        l0 = __reduce_add_sync(0xffffffffu, l0);
        l1 = __reduce_add_sync(0xffffffffu, l1);
        l2 = __reduce_add_sync(0xffffffffu, l2);
        if (lane < 3) {
            const long long v = lane == 0 ? l0 : lane == 1 ? l1 : l2;
            if (v != 0) atomicAdd((unsigned long long*)&mt[m][g][1 + lane], (unsigned long long)v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sweep epilogue, observation pass (state in HBM): finish the last tree (slots only), re-derive both forest fits
// from the trees (Gc = sum of mu-tree leaves, Gm = sum of tau-tree leaves) and accumulate the six second moments
// per treatment group that the scale and sigma draws need [SURVEY A.4].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void refit_tile(StepShared& sm, const Dev& d, int prev, const TreeRec* newtrees,
                                           int tile, int lane, int& bad)
{
    const int64_t base = (int64_t)tile * TILE + lane * 8;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    int sp[8];
    if (ai.active) {
        uint8_t* sl = d.slots + (int64_t)prev * d.n_pad + base;
        ld8_u8(sl, sp);
        if (ai.changed) {
            int bp[8];
            double dl[8];
            if (ai.birth) ld8_bins(ai.bins, ai.is16, base, bp);
            apply8(sm, ai, sp, bp, dl);
            st8_u8(sl, sp);
        }
    }
    double G[2][8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { G[0][o] = 0.0; G[1][o] = 0.0; }
    for (int f = 0; f < 2; ++f) {
        const int t0 = d.f[f].tree0, nt = d.f[f].ntree;
        for (int j = t0; j < t0 + nt; ++j) {
            int s[8];
            const double* mu;
            if (j == prev) {
#pragma unroll
                for (int o = 0; o < 8; ++o) s[o] = sp[o];
                mu = sm.mu_new;
            } else {
                ld8_u8(d.slots + (int64_t)j * d.n_pad + base, s);
                mu = newtrees[j].mu;
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) if (s[o] != PAD_SLOT) G[f][o] += mu[s[o]];
        }
    }
    st8_f64(d.Gc + base, G[0]);
    st8_f64(d.Gm + base, G[1]);
    double y[8];
    ld8_f64(d.y + base, y);
    moments8(sm.tab, y, G[0], G[1], g, lane, bad);
}

}  // namespace bcf
