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
    double dlt_right;    // change of the leaf value for observations moving to the new right child
    long long dfx_right[2];   // the same, scaled by the forest's scale and in fixed point, control / treated rows
    int32_t uniform;     // the tree had one leaf and keeps it: every observation gets dlt[0] / dfx[g][0]
    int32_t pad_;
};

struct CurInfo {
    int32_t active, K;   // K = number of keys = nleaves (+1 for a birth proposal)
    int32_t birth, sx, cut, is16, forest, new_slot;
    const void* bins;
};

// A tree in flight: its structure, its proposal, the leaf normals and the statistics-pass set-up.  Two stages, so
// that tree j+1 can be proposed while tree j still waits for its reduced statistics.
template <class Tree, int NZ>
struct TreeStageT {
    Tree tr;
    Proposal prop;
    double z[NZ];                  // standard normal of each (current, then new) leaf slot
    CurInfo ci;
};
using TreeStage = TreeStageT<TreeRec, KEYS>;

struct StepShared {
    TreeStage st[2];
    double cnt[KEYS + 1], sphi[KEYS + 1], sphir[KEYS + 1];   // entry KEYS: the pair a move merges / splits, pooled
    double pinv[KEYS + 1], plog[KEYS + 1], pmean[KEYS + 1], psd[KEYS + 1];   // 1/prec, log prec, posterior mean, sd
    double log_tau;
    double mu_old[MAXL], mu_new[MAXL];
    double dlt[KEYS];              // change of the leaf value per OLD slot
    long long dfx[2][KEYS];        // scale * dlt in fixed point, control / treated rows: what the residual loses
    double nc1[KEYS + 1], nc0[KEYS + 1];   // counts per key, treated / control
    uint8_t remap[MAXL];
    int32_t ngood[MAXL];
    uint32_t lkey[MAXL];
    uint32_t nkey[MAXN];
    uint8_t nogflag[MAXN];
    ApplyInfo ai;
    int32_t accepted;
    int32_t bad;
    StatTab tab;
    long long rowsum[ROWSUM_WORDS];
    long long tot[KEYS * 8];       // reduced statistics of the tree being decided, staged from L2
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
template <class SM, class Stage>
__device__ inline void propose_stage(SM& sm, Stage& S, const Dev& d, const ForestDev& F, int tree, uint32_t sweep)
{
    auto& tr = S.tr;
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

__device__ inline void propose(StepShared& sm, int stage, const Dev& d, const ForestDev& F, int tree, uint32_t sweep)
{
    propose_stage(sm, sm.st[stage], d, F, tree, sweep);
}

// publish a staged proposal (and leaf normals) to HBM / fetch it back into a stage
template <class Stage>
__device__ inline void publish_stage(const Stage& S, const Dev& d, int tree)
{
    PropRec* R = d.proprec + tree;
    const int t = threadIdx.x;
    if (t == 0) R->P = S.prop;
    if (t >= 32 && t - 32 < S.tr.nleaves) R->z[t - 32] = S.z[t - 32];
}
__device__ inline void publish_proposal(const StepShared& sm, int stage, const Dev& d, int tree)
{
    publish_stage(sm.st[stage], d, tree);
}
template <class Stage>
__device__ inline void fetch_stage(Stage& S, const Dev& d, int tree, int nleaves)
{
    const PropRec* R = d.proprec + tree;
    const int t = threadIdx.x;
    if (t < (int)(sizeof(Proposal) / 8)) reinterpret_cast<long long*>(&S.prop)[t] = __ldcg(reinterpret_cast<const long long*>(&R->P) + t);
    if (t >= 32 && t - 32 < nleaves) S.z[t - 32] = __ldcg(&R->z[t - 32]);
}
__device__ inline void fetch_proposal(StepShared& sm, int stage, const Dev& d, int tree, int nleaves)
{
    fetch_stage(sm.st[stage], d, tree, nleaves);
}

// ---------------------------------------------------------------------------------------------
// decide(): given the reduced statistics of tree `tree` (d.totals) and its proposal (stage.prop), take the MH
// decision, update the structure in stage.tr, redraw every leaf value and build the apply tables.
// bcf's bd() likelihood half + drmu() [bcf-recall, SURVEY A.3 steps 2-3].  `writer` CTA records the trace.
//
// Phase 1 (one thread per key, plus one for the pooled pair): statistics -> precision, its log and reciprocal,
// posterior mean and sd.  Phase 2 (thread 0): compare, update the structure.  Phase 3 (one thread per leaf):
// mu = mean + z * sd.  The critical path holds one L2 round trip, one log, one division and one sqrt.
// ---------------------------------------------------------------------------------------------
// statistics of one key: counts (cached in the tree, or reduced for the would-be right child of a birth) and the
// reduced fixed-point residual sums -> (count, sum phi, sum phi * R) of bcf's sinfo
__device__ __forceinline__ void key_raw(const long long* tot, int key, double& S1, double& S0, double& c1, double& c0)
{
    const long long* q = tot + key * 8;   // [control: count, limb0, limb1, limb2][treated: ...]
    S1 = limbs_to_double(q[5], q[6], q[7]);
    S0 = limbs_to_double(q[1], q[2], q[3]);
    c1 = (double)q[4]; c0 = (double)q[0];
}

// decide_core(): the statistics of the tree are already in sm.tot (K x [control: count, limb0..2][treated: ...])
// WARP: run by ONE warp (lanes = threads 0..31 of the CTA) with warp-level synchronisation only; needs K <= 16.
template <class SM, class Stage, bool WARP = false>
__device__ inline void decide_core(SM& sm, Stage& S, const Dev& d, const ForestDev& F, int tree, const Scalars& sc,
                                   bool writer)
{
    auto& tr = S.tr;
    const Proposal P = S.prop;
    const int t = threadIdx.x;
    const int L = tr.nleaves;
    const int K = L + (P.move == 1 ? 1 : 0);
    const bool is_con = F.is_con != 0;
    const double s1 = is_con ? sc.mscale : sc.bscale1, s0 = is_con ? sc.mscale : sc.bscale0;
    const double tau = is_con ? sc.tau_con : sc.tau_mod;
    // the two keys a move pools: birth (leaf part staying left, would-be right child), death (the two children)
    const int pa = P.slot, pb = (P.move == 1) ? L : P.slot_b;

    constexpr int TPOOL = WARP ? 16 : KEYS;          // thread of the pooled pair
    constexpr int TLOG = WARP ? 17 : KEYS + 1;
    for (int i = t; i < (WARP ? 2 * 16 : MAXL); i += (WARP ? 32 : MAXL)) { if (i < MAXL) sm.remap[i] = (uint8_t)i; }
    if (t < L) sm.mu_old[t] = tr.mu[t];
    if (WARP) __syncwarp(); else __syncthreads();

    // Phase 1.  Keys: leaf slots 0..L-1 (for a birth the proposed leaf's key holds the part that stays LEFT) and
    // key L = the would-be right child.  Thread KEYS handles the pooled pair of the move (sum of its two keys).
    if (t < K || (t == TPOOL && P.move != 0)) {
        const double s2 = sc.sigma * sc.sigma;
        const double phi1 = s1 * s1 / s2, phi0 = s0 * s0 / s2;
        const double a = 1.0 / (tau * tau);
        const double rs1 = 1.0 / s1, rs0 = 1.0 / s0;   // one reciprocal per forest and sweep, shared by every leaf
        double n1 = 0.0, n0 = 0.0, sphi = 0.0, sphir = 0.0;
        const int nk = (t < K) ? 1 : 2;
        for (int q = 0; q < nk; ++q) {
            const int key = (t < K) ? t : (q == 0 ? pa : pb);
            double S1, S0, c1, c0, k1, k0, mo;
            key_raw(sm.tot, key, S1, S0, c1, c0);
            if (key < L) {
                k1 = (double)tr.cnt1[key]; k0 = (double)tr.cnt0[key]; mo = tr.mu[key];
                if (P.move == 1 && key == P.slot) {      // left part = leaf - right part
                    double R1, R0, r1, r0;
                    key_raw(sm.tot, L, R1, R0, r1, r0);
                    k1 -= r1; k0 -= r0;
                }
            } else { k1 = c1; k0 = c0; mo = tr.mu[P.slot]; }
            n1 += k1; n0 += k0;
            sphi += k1 * phi1 + k0 * phi0;
            sphir += phi1 * (S1 * rs1 + k1 * mo) + phi0 * (S0 * rs0 + k0 * mo);
        }
        const double prec = a + sphi, inv = 1.0 / prec;
        const int w = (t < K) ? t : KEYS;
        sm.nc1[w] = n1; sm.nc0[w] = n0;
        sm.cnt[w] = n1 + n0; sm.sphi[w] = sphi; sm.sphir[w] = sphir;
        sm.pinv[w] = inv; sm.plog[w] = log(prec); sm.pmean[w] = sphir * inv; sm.psd[w] = sqrt(inv);
    }
    if (t == TLOG) sm.log_tau = log(tau);
    if (WARP) __syncwarp(); else __syncthreads();

    if (t == 0) {
        int accepted = 0;
        double logr = nan("");
        if (P.move != 0) {
            // lil(a) + lil(b) - lil(a+b), weighted form [SURVEY A.3]
            const double gain = -sm.log_tau - 0.5 * (sm.plog[pa] + sm.plog[pb] - sm.plog[KEYS])
                              + 0.5 * (sm.sphir[pa] * sm.sphir[pa] * sm.pinv[pa] + sm.sphir[pb] * sm.sphir[pb] * sm.pinv[pb]
                                       - sm.sphir[KEYS] * sm.sphir[KEYS] * sm.pinv[KEYS]);
            if (P.move == 1) {
                if (sm.cnt[pa] >= d.min_leaf && sm.cnt[pb] >= d.min_leaf) {
                    logr = P.log_prior + gain;
                    accepted = (P.log_u < logr) ? 1 : 0;
                }
            } else {
                logr = P.log_prior - gain;
                accepted = (P.log_u < logr) ? 1 : 0;
            }
        }
        if (P.move == 1) {
            const int sx = pa, kr = pb;
            if (accepted) {
                const int nx = P.node, na = tr.nnodes, nb = tr.nnodes + 1;
                tr.left[nx] = (int16_t)na; tr.right[nx] = (int16_t)nb;
                tr.var[nx] = (uint16_t)P.var; tr.cut[nx] = (uint16_t)P.cut;
                for (int q = 0; q < 2; ++q) {
                    const int ch = q ? nb : na;
                    tr.parent[ch] = (int16_t)nx; tr.left[ch] = tr.right[ch] = -1;
                    tr.var[ch] = tr.cut[ch] = 0xFFFF;
                    tr.depth[ch] = tr.depth[nx] + 1;
                    tr.heap[ch] = 2u * tr.heap[nx] + (uint32_t)q;
                }
                tr.node_slot[na] = (uint8_t)sx; tr.slot_node[sx] = (int16_t)na;
                tr.node_slot[nb] = (uint8_t)kr; tr.slot_node[kr] = (int16_t)nb;
                tr.nnodes += 2; tr.nleaves = L + 1;
                S.z[sx] = P.z_extra[0]; S.z[kr] = P.z_extra[1];
                // cached counts: entries sx (left part) and kr (right child) are already what phase 1 computed
            } else {   // the leaf stays whole: it is the pooled pair
                sm.nc1[sx] = sm.nc1[KEYS]; sm.nc0[sx] = sm.nc0[KEYS];
                sm.pmean[sx] = sm.pmean[KEYS]; sm.psd[sx] = sm.psd[KEYS];
            }
        } else if (P.move == 2 && accepted) {
            const int sa = pa, sb = pb;
            const int nx = P.node, na = tr.left[nx], nb = tr.right[nx];
            tr.left[nx] = tr.right[nx] = -1; tr.var[nx] = tr.cut[nx] = 0xFFFF;
            tr.node_slot[nx] = (uint8_t)sa; tr.slot_node[sa] = (int16_t)nx;
            sm.pmean[sa] = sm.pmean[KEYS]; sm.psd[sa] = sm.psd[KEYS];
            sm.nc1[sa] = sm.nc1[KEYS]; sm.nc0[sa] = sm.nc0[KEYS];
            S.z[sa] = P.z_extra[0];
            sm.remap[sb] = (uint8_t)sa;
            const int last = L - 1;
            if (sb != last) {   // keep slots compact: the last slot moves into the hole
                const int nd = tr.slot_node[last];
                tr.slot_node[sb] = (int16_t)nd; tr.node_slot[nd] = (uint8_t)sb;
                sm.pmean[sb] = sm.pmean[last]; sm.psd[sb] = sm.psd[last];
                sm.nc1[sb] = sm.nc1[last]; sm.nc0[sb] = sm.nc0[last];
                S.z[sb] = S.z[last];
                for (int s = 0; s < L; ++s) if (sm.remap[s] == last) sm.remap[s] = (uint8_t)sb;
            }
            tr.nleaves = L - 1;
            tree_remove_node(tr, max(na, nb));
            tree_remove_node(tr, min(na, nb));
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
    if (WARP) __syncwarp(); else __syncthreads();

    const int Lnew = tr.nleaves;
    if (t < Lnew) {
        const double m = sm.pmean[t] + S.z[t] * sm.psd[t];
        if (!(m == m)) sm.bad = ERR_NAN;
        sm.mu_new[t] = m;
        tr.mu[t] = m;
        tr.cnt1[t] = (int32_t)sm.nc1[t]; tr.cnt0[t] = (int32_t)sm.nc0[t];
    }
    if (WARP) __syncwarp(); else __syncthreads();
    int fxbad = 0;
    if (t < L) {
        const double dl = sm.mu_new[sm.remap[t]] - sm.mu_old[t];
        sm.dlt[t] = dl;
        sm.dfx[1][t] = to_fixed(s1 * dl, fxbad); sm.dfx[0][t] = to_fixed(s0 * dl, fxbad);
    }
    if (t == 0) {
        ApplyInfo& ai = sm.ai;
        ai.active = 1; ai.changed = sm.accepted; ai.birth = (sm.accepted && P.move == 1) ? 1 : 0;
        ai.sx = P.slot; ai.cut = P.cut; ai.bins = P.bins; ai.is16 = P.is16; ai.right_slot = L; ai.forest = is_con ? 0 : 1;
        ai.dlt_right = ai.birth ? (sm.mu_new[L] - sm.mu_old[P.slot]) : 0.0;
        ai.dfx_right[1] = to_fixed(s1 * ai.dlt_right, fxbad); ai.dfx_right[0] = to_fixed(s0 * ai.dlt_right, fxbad);
        ai.uniform = (L == 1 && !ai.birth) ? 1 : 0;
    }
    if (fxbad) sm.bad = ERR_FX_RANGE;
    if (WARP) __syncwarp(); else __syncthreads();
}

// `totals`: the statistics buffer of the sweep in progress (see sweep_totals)
__device__ inline void decide(StepShared& sm, int stage, const Dev& d, const ForestDev& F, int tree, const Scalars& sc,
                              bool writer, const long long* totals)
{
    TreeStage& S = sm.st[stage];
    const int K = S.tr.nleaves + (S.prop.move == 1 ? 1 : 0);
    // reduced by integer atomics in L2: stage them with ONE round trip, straight from L2 (never a stale L1 line)
    const long long* q = totals + (size_t)tree * KEYS * 8;
    for (int i = threadIdx.x; i < K * 8; i += blockDim.x) sm.tot[i] = __ldcg(q + i);
    decide_core(sm, S, d, F, tree, sc, writer);
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
    }
    long long* tb = reinterpret_cast<long long*>(&sm.tab);
    for (int i = t; i < STATTAB_WORDS64; i += blockDim.x) tb[i] = 0;
}

// flush the CTA table into the global totals of `tree` (integer REDs: order-independent).  Must be entered by every
// thread of the CTA after a __syncthreads() that ends the tile pass.
__device__ inline void flush_stats(StepShared& sm, int K, long long* totals, int tree)
{
    long long* q = totals + (size_t)tree * KEYS * 8;
    const int t = threadIdx.x;
    if (K <= KREG) {
        // sum the warps' private tables: 16 lanes per (key, group, word) entry, then a shuffle tree
        const int nwarps = blockDim.x >> 5;
        const int nent = K * 8;
        for (int e0 = 0; e0 < nent; e0 += blockDim.x / 16) {
            const int e = e0 + (t >> 4), w = t & 15;
            long long v = 0;
            if (e < nent) for (int ww = w; ww < nwarps; ww += 16) v += sm.tab.w[ww][e >> 3][(e >> 2) & 1][e & 3];
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (e < nent && w == 0) sm.rowsum[e] = v;
        }
        __syncthreads();
        if (t < nent) {
            const int k = t >> 3, gq = t & 7;
            long long v = sm.rowsum[t];
            if (k == 0) {          // row 0 holds the total over all keys: key 0 = total - the others (exact)
                if ((gq & 3) == 0) v = 0;
                else for (int kk = 1; kk < K; ++kk) v -= sm.rowsum[kk * 8 + gq];
            }
            if (v != 0) atomicAdd((unsigned long long*)&q[t], (unsigned long long)v);
        }
    } else {
        for (int i = t; i < K * 8; i += blockDim.x) {
            const long long v = stat_word(sm.tab, K, 0, i >> 3, (i >> 2) & 1, i & 3);
            if (v != 0) atomicAdd((unsigned long long*)&q[i], (unsigned long long)v);
        }
    }
}

// apply the decided update of the previous tree to 8 observations: new slot, the change of the leaf value (dl, goes
// into the forest fit G) and what the fixed-point residual loses (df = scale_g * dl, rounded once per leaf)
__device__ __forceinline__ void apply8(const StepShared& sm, const ApplyInfo& ai, int g, int (&sp)[8], const int (&bp)[8],
                                       double (&G)[8], long long (&R)[8])
{
    if (ai.uniform) {
        const double d0 = sm.dlt[0];
        const long long f0 = sm.dfx[g][0];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const bool real = sp[o] != PAD_SLOT;
            G[o] += real ? d0 : 0.0;
            R[o] -= real ? f0 : 0ll;
        }
        return;
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int s = sp[o];
        if (s != PAD_SLOT) {
            double x = sm.dlt[s];
            long long f = sm.dfx[g][s];
            int ns = sm.remap[s];
            if (ai.birth && s == ai.sx && bp[o] > ai.cut) { x = ai.dlt_right; f = ai.dfx_right[g]; ns = ai.right_slot; }
            G[o] += x;
            R[o] -= f;
            sp[o] = ns;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One warp tile of the fused pass (state in HBM/L2): apply the previous tree's update to the residual, to its
// forest's fit and to its leaf slots, then accumulate the current tree's statistics.  All loads are issued
// before the first dependent use.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void step_tile(StepShared& sm, const CurInfo& ci, const Dev& d, int prev, int cur, int tile,
                                          int lane)
{
    const int64_t base = (int64_t)tile * TILE + lane * 8;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    long long R[8];
    double G[8];
    int sp[8], bp[8], key[8], bc[8];
    uint8_t* sl = d.slots + (int64_t)prev * d.n_pad + base;
    double* Gp = (ai.forest ? d.Gm : d.Gc) + base;
    ld8_i64(d.rfx + base, R);
    if (ai.active) {
        ld8_f64(Gp, G);
        ld8_u8(sl, sp);
        if (ai.birth) ld8_bins(ai.bins, ai.is16, base, bp);
    }
    if (ci.active) {
        ld8_u8(d.slots + (int64_t)cur * d.n_pad + base, key);
        if (ci.birth) ld8_bins(ci.bins, ci.is16, base, bc);
    }
    if (ai.active) {
        apply8(sm, ai, g, sp, bp, G, R);
        if (ai.changed) st8_u8(sl, sp);
        st8_f64(Gp, G);
        st8_i64(d.rfx + base, R);
    }
    if (ci.active) {
        if (ci.birth) {
#pragma unroll
            for (int o = 0; o < 8; ++o) if (key[o] == ci.sx && bc[o] > ci.cut) key[o] = ci.new_slot;
        }
        accum_tile(sm.tab, key, R, g, ci.K, ci.birth != 0, lane, (int)(threadIdx.x >> 5));
    }
}

// moments of 8 observations into the warp's table rows 0..5 (words 1..3 = limbs); padding rows have y = G = 0
template <class Tab>
__device__ __forceinline__ void moments8(Tab& tab, const double (&y)[8], const double (&gc)[8],
                                         const double (&gm)[8], int g, int lane, int& bad)
{
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < NMOM; ++m) {
        long long s = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const double a = (m < 3) ? y[o] : (m < 5 ? gc[o] : gm[o]);
            const double b = (m == 0) ? y[o] : (m == 1 || m == 3) ? gc[o] : gm[o];
            s += to_fixed(a * b, bad);
        }
        warp_add_limbs(tab, warp, m, g, lane, s, 0);
    }
    __syncwarp();
}
// clear the table before a moments pass / flush the six moments of both groups into `mom`
__device__ inline void moments_begin(StepShared& sm)
{
    long long* tb = reinterpret_cast<long long*>(&sm.tab);
    for (int i = threadIdx.x; i < STATTAB_WORDS64; i += blockDim.x) tb[i] = 0;
}
__device__ inline void moments_flush(const StepShared& sm, long long* mom)
{
    const int nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < NMOM * 8; i += blockDim.x) {
        const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
        if (q == 0) continue;
        const long long v = table_word(sm.tab, nwarps, m, g, q);
        if (v != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------
// Sweep epilogue, observation pass (state in HBM): apply the last tree, then accumulate the six second moments of
// (y, Gc, Gm) per treatment group that the scale and sigma draws need [SURVEY A.4].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void last_tile(StepShared& sm, const Dev& d, int prev, int tile, int lane, int& bad)
{
    const int64_t base = (int64_t)tile * TILE + lane * 8;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    double gc[8], gm[8], y[8];
    ld8_f64(d.Gc + base, gc); ld8_f64(d.Gm + base, gm); ld8_f64(d.y + base, y);
    if (ai.active) {
        int sp[8], bp[8];
        long long R[8];
        uint8_t* sl = d.slots + (int64_t)prev * d.n_pad + base;
        ld8_u8(sl, sp);
        if (ai.birth) ld8_bins(ai.bins, ai.is16, base, bp);
#pragma unroll
        for (int o = 0; o < 8; ++o) R[o] = 0;            // the residual is re-derived after the scale draws
        if (ai.forest) { apply8(sm, ai, g, sp, bp, gm, R); st8_f64(d.Gm + base, gm); }
        else { apply8(sm, ai, g, sp, bp, gc, R); st8_f64(d.Gc + base, gc); }
        if (ai.changed) st8_u8(sl, sp);
    }
    moments8(sm.tab, y, gc, gm, g, lane, bad);
}

// Make (and publish to HBM) the proposals of sweep `sweep` for the trees j = first, first + stride, ... < T.
// `trees` holds their current structure; tree `smem_tree` (if >= 0) is taken from stage `smem_stage` instead
// (the tree a CTA has just decided and whose HBM copy may not be visible yet).  Uses the other stage as scratch.
__device__ inline void propose_range(StepShared& sm, const Dev& d, const TreeRec* trees, uint32_t sweep, int first,
                                     int stride, int smem_tree, int smem_stage)
{
    const int scratch = smem_stage ^ 1;
    for (int j = first; j < d.T; j += stride) {
        const ForestDev& F = d.f[j >= d.f[1].tree0 ? 1 : 0];
        int st = smem_stage;
        if (j != smem_tree) {
            st = scratch;
            load_tree(&sm.st[st].tr, &trees[j]);
            __syncthreads();
        }
        propose(sm, st, d, F, j, sweep);
        publish_proposal(sm, st, d, j);
        __syncthreads();
    }
}

}  // namespace bcf
