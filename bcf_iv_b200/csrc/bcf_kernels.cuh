// bcf_kernels.cuh -- __global__ entry points (GRAPH engine and stand-alone operators).
#pragma once
#include "bcf_step.cuh"

namespace bcf {

// StepShared is larger than the 48 KB static limit: the sampler kernels take it as dynamic shared memory
// (launch with STEP_SMEM_BYTES, after cudaFuncSetAttribute(MaxDynamicSharedMemorySize)).
constexpr size_t STEP_SMEM_BYTES = (sizeof(StepShared) + 15) / 16 * 16;
#define BCF_STEP_SMEM(sm)                                         \
    extern __shared__ __align__(16) unsigned char bcf_dyn_smem[]; \
    StepShared& sm = *reinterpret_cast<StepShared*>(bcf_dyn_smem)

// ---------------------------------------------------------------------------------------------
// K0: bin one column.  out = #{c : cuts[c] <= x}  (upper bound), so  x < cuts[c]  <=>  out <= c.
// Written at the observation's internal (treated-first, padded) position.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_cut(const double* __restrict__ cuts, int ncut, double x)
{
    int lo = 0, hi = ncut;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cuts[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// grid = (blocks of BIN_ROWS rows, groups of BIN_COLS columns of the chunk).  The rows of a block land in two runs of
// the internal order -- its treated rows and its control rows, each contiguous because bcf's permutation is stable -- so the
// block works out ONCE where each of its rows goes, then per column compacts the bins per run in shared memory and writes
// them as aligned 32-bit words; the next column's values are already in flight while one is flushed.  x is read once with
// coalesced 8-byte loads and every 32-byte sector of a binned column is written once.  (The first versions -- scattered byte
// stores, then one column per block with a binary search even for one-cutpoint columns -- ran at a third of the copy
// bandwidth: instruction-bound, 66 % issue utilisation for 72 MB.)
constexpr int BIN_ROWS = 2048, BIN_COLS = 4;
template <typename B>
__device__ __forceinline__ void bin_flush_run(B* __restrict__ dst, int64_t g0, int len, const B* __restrict__ src, int t)
{
    constexpr int PER = 4 / (int)sizeof(B);                       // elements per 32-bit word
    const int head = min(len, (int)((PER - (g0 & (PER - 1))) & (PER - 1)));
    if (t < head) dst[g0 + t] = src[t];
    const int nwords = (len - head) / PER;
    uint32_t* out = reinterpret_cast<uint32_t*>(dst + g0 + head);
    for (int w = t; w < nwords; w += 256) {
        const B* q = src + head + PER * w;
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < PER; ++e) v |= (uint32_t)q[e] << (8 * (int)sizeof(B) * e);
        out[w] = v;
    }
    const int tail0 = head + PER * nwords;
    if (t < len - tail0) dst[g0 + tail0 + t] = src[tail0 + t];
}
// XT = double: the caller's values; XT = uint8_t: a run of ONE-cutpoint columns that crossed the bus already compared with
// their cutpoint (bin = (cut <= x) as a byte: set_data's host staging pass writes that instead of copying 8 bytes).
template <typename XT>
__global__ void __launch_bounds__(256, 4) k_bin(const XT* __restrict__ x, int64_t n, int ncols, int col0,
                                             const double* __restrict__ cuts, const int32_t* __restrict__ off,
                                             const int32_t* __restrict__ col_index, const uint8_t* __restrict__ col_is16,
                                             const int32_t* __restrict__ pos, int64_t n_pad, int trt_pad, uint8_t* bins8,
                                             uint16_t* bins16)
{
    constexpr int PT = BIN_ROWS / 256;                            // rows per thread
    __shared__ __align__(16) uint16_t sb[BIN_ROWS];               // the block's bins of one column in run order (uint8 columns: as bytes)
    __shared__ int s_min[2], s_cnt;
    const int t = threadIdx.x;
    const int cg0 = blockIdx.y * BIN_COLS, ncg = min(BIN_COLS, ncols - cg0);
    const int64_t r0 = (int64_t)blockIdx.x * BIN_ROWS;
    if (ncg <= 0 || r0 >= n) return;
    const int rows = (int)min((int64_t)BIN_ROWS, n - r0);
    // ---- where the block's rows go: treated positions lie below trt_pad and increase with the row, control ones above ----
    if (t == 0) { s_min[0] = s_min[1] = 0x7fffffff; s_cnt = 0; }
    __syncthreads();
    int li[PT];
    {
        int q[PT], mn[2] = {0x7fffffff, 0x7fffffff}, cnt_t = 0;
#pragma unroll
        for (int k = 0; k < PT; ++k) {
            const int r = k * 256 + t;
            q[k] = r < rows ? pos[r0 + r] : -1;
            if (q[k] >= 0) { const int g = q[k] < trt_pad ? 1 : 0; mn[g] = min(mn[g], q[k]); cnt_t += g; }
        }
        cnt_t = __reduce_add_sync(0xffffffffu, cnt_t);
        mn[0] = __reduce_min_sync(0xffffffffu, mn[0]);
        mn[1] = __reduce_min_sync(0xffffffffu, mn[1]);
        if ((t & 31) == 0) { atomicAdd(&s_cnt, cnt_t); atomicMin(&s_min[0], mn[0]); atomicMin(&s_min[1], mn[1]); }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PT; ++k) li[k] = q[k] < 0 ? -1 : (q[k] < trt_pad ? q[k] - s_min[1] : s_cnt + (q[k] - s_min[0]));
    }
    // the block's treated rows form the run [p0, p0 + n_same) of every column, its control rows [o0, o0 + n_other)
    const int n_same = s_cnt, n_other = rows - n_same, p0 = s_min[1], o0 = s_min[0];
    uint8_t* sb8 = reinterpret_cast<uint8_t*>(sb);
    constexpr bool PREBINNED = sizeof(XT) == 1;
    XT xv[PT];
#pragma unroll
    for (int k = 0; k < PT; ++k) { const int r = k * 256 + t; xv[k] = r < rows ? x[(int64_t)cg0 * n + r0 + r] : (XT)0; }
    for (int cc = 0; cc < ncg; ++cc) {
        const int c = cg0 + cc, v = col0 + c;
        XT xn[PT];
        if (cc + 1 < ncg) {                                        // the next column's values fly while this one is binned and flushed
#pragma unroll
            for (int k = 0; k < PT; ++k) { const int r = k * 256 + t; xn[k] = r < rows ? x[(int64_t)(c + 1) * n + r0 + r] : (XT)0; }
        }
        const double* cv = cuts + off[v];
        const int nc = off[v + 1] - off[v];
        const bool w16 = col_is16[v] != 0;
        if (PREBINNED) {
#pragma unroll
            for (int k = 0; k < PT; ++k) if (li[k] >= 0) sb8[li[k]] = (uint8_t)xv[k];
        } else if (nc == 1) {                                      // two-valued columns (the reference's covariates): one compare
            const double c0v = cv[0];
#pragma unroll
            for (int k = 0; k < PT; ++k) if (li[k] >= 0) sb8[li[k]] = (uint8_t)(c0v <= (double)xv[k] ? 1 : 0);
        } else {
#pragma unroll
            for (int k = 0; k < PT; ++k)
                if (li[k] >= 0) {
                    const int b = upper_bound_cut(cv, nc, (double)xv[k]);
                    if (w16) sb[li[k]] = (uint16_t)b; else sb8[li[k]] = (uint8_t)b;
                }
        }
        __syncthreads();
        const int ci = col_index[v];
        if (w16) {
            uint16_t* dst = bins16 + (int64_t)ci * n_pad;
            if (n_same > 0) bin_flush_run<uint16_t>(dst, p0, n_same, sb, t);
            if (n_other > 0) bin_flush_run<uint16_t>(dst, o0, n_other, sb + n_same, t);
        } else {
            uint8_t* dst = bins8 + (int64_t)ci * n_pad;
            if (n_same > 0) bin_flush_run<uint8_t>(dst, p0, n_same, sb8, t);
            if (n_other > 0) bin_flush_run<uint8_t>(dst, o0, n_other, sb8 + n_same, t);
        }
        __syncthreads();
        if (cc + 1 < ncg) {
#pragma unroll
            for (int k = 0; k < PT; ++k) xv[k] = xn[k];
        }
    }
}
// bcf's perm = order(z, decreasing = TRUE), applied on the device: one CTA per chunk of PLACE_CHUNK rows, whose count of
// treated rows BEFORE the chunk comes from the host's validation pass.  Treated rows go to [0, ntrt) in row order, control
// rows to [trt_pad, trt_pad + nctl) (each group is padded to whole warp tiles).  Writes pos (row -> position), orig
// (position -> row; the padding stays -1) and y in internal order.
__global__ void __launch_bounds__(256) k_place(const double* __restrict__ yraw, const int32_t* __restrict__ zraw,
                                               const int32_t* __restrict__ chunk_trt, int64_t n, int64_t trt_pad,
                                               int32_t* __restrict__ pos, int32_t* __restrict__ orig, double* __restrict__ y)
{
    __shared__ int wsum[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 4096;
    int64_t trt_base = chunk_trt[blockIdx.x];
    int64_t ctl_base = trt_pad + (r0 - trt_base);
    for (int it = 0; it < 4096 / 256 && r0 + it * 256 < n; ++it) {
        const int64_t i = r0 + it * 256 + t;
        const bool valid = i < n;
        const int zi = valid ? zraw[i] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, zi != 0);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int sv = wsum[w]; total += sv; if (w < warp) before += sv; }
        const int rank_t = before + __popc(bal & ((1u << lane) - 1u));
        if (valid) {
            const int64_t q = zi ? trt_base + rank_t : ctl_base + (t - rank_t);
            pos[i] = (int32_t)q;
            orig[q] = (int32_t)i;
            y[q] = yraw[i];
        }
        trt_base += total; ctl_base += 256 - total;
        __syncthreads();
    }
}
// every tree a root [SURVEY A.2]: the mu trees' leaf ybar / ntree_con, the tau trees' trt_init / ntree_mod
__global__ void __launch_bounds__(256) k_init_trees(TreeRec* t0, TreeRec* t1, int ntree_con, double mu_con, double mu_mod,
                                                    int32_t ntrt, int32_t nctl)
{
    __shared__ TreeRec tr;
    const int t = threadIdx.x, j = blockIdx.x;
    for (int i = t; i < (int)(sizeof(TreeRec) / 16); i += blockDim.x) reinterpret_cast<uint4*>(&tr)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = t; i < MAXN; i += blockDim.x) { tr.parent[i] = tr.left[i] = tr.right[i] = -1; tr.var[i] = tr.cut[i] = 0xFFFF; }
    __syncthreads();
    if (t == 0) {
        tr.nnodes = 1; tr.nleaves = 1;
        tr.node_slot[0] = 0; tr.depth[0] = 0; tr.heap[0] = 1; tr.slot_node[0] = 0;
        tr.mu[0] = j < ntree_con ? mu_con : mu_mod;
        tr.cnt1[0] = ntrt; tr.cnt0[0] = nctl;
    }
    __syncthreads();
    store_tree(&t0[j], &tr);
    store_tree(&t1[j], &tr);
}
__global__ void k_bin_simple(const double* __restrict__ x, int64_t n, const double* __restrict__ cuts, int ncut,
                             uint16_t* out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint16_t)upper_bound_cut(cuts, ncut, x[i]);
}

// ---------------------------------------------------------------------------------------------
// K1: assign every observation to its leaf slot by walking the flat tree over the binned columns.
// mode 0: write slots; mode 1: compare with the cached slots and count mismatches.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assign(Dev d, int tree, int forest, const TreeRec* trees, int mode,
                                                unsigned long long* mismatches, const int32_t* __restrict__ orig)
{
    __shared__ TreeRec tr;
    load_tree(&tr, &trees[tree]);
    __syncthreads();
    const ForestDev& F = d.f[forest];
    uint8_t* sl = d.slots + (int64_t)tree * d.n_pad;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        int slot = PAD_SLOT;
        if (orig[i] >= 0) {
            int node = 0;
            while (tr.left[node] >= 0) {
                const int v = tr.var[node];
                const int ci = F.col_index[v];
                const int b = F.col_is16[v] ? (int)F.bins16[(int64_t)ci * d.n_pad + i] : (int)F.bins8[(int64_t)ci * d.n_pad + i];
                node = (b <= (int)tr.cut[node]) ? tr.left[node] : tr.right[node];
            }
            slot = tr.node_slot[node];
        }
        if (mode == 0) sl[i] = (uint8_t)slot;
        else if (sl[i] != (uint8_t)slot) ++bad;
    }
    if (mode == 1 && bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------------------------------------
// GRAPH engine: one kernel per tree.  Kernel j decides tree j-1 (from the totals its predecessor reduced),
// proposes for tree j, then makes ONE pass over the observations that applies tree j-1's update to the
// residual / leaf cache and accumulates tree j's sufficient statistics.
// ---------------------------------------------------------------------------------------------
// proposals of the coming sweep, one CTA per tree (they depend only on the structures and the Philox counters)
__global__ void __launch_bounds__(THREADS) k_propose_all(Dev d)
{
    BCF_STEP_SMEM(sm);
    const Scalars sc = *d.sc;
    propose_range(sm, d, d.trees[sc.parity], sc.sweep, blockIdx.x, gridDim.x, -1, 0);
}

__global__ void __launch_bounds__(THREADS) k_tree_step(Dev d, int prev, int cur)
{
    BCF_STEP_SMEM(sm);
    const Scalars sc = *d.sc;
    const int par = sc.parity;
    const bool writer = (blockIdx.x == 0);
    if (threadIdx.x == 0) { sm.ai.active = 0; sm.st[1].ci.active = 0; sm.bad = 0; }
    __syncthreads();
    if (prev >= 0) {
        const ForestDev& F = d.f[prev >= d.f[1].tree0 ? 1 : 0];
        load_tree(&sm.st[0].tr, &d.trees[par][prev]);
        __syncthreads();
        fetch_proposal(sm, 0, d, prev, sm.st[0].tr.nleaves);
        __syncthreads();
        decide(sm, 0, d, F, prev, sc, writer, d.totals);
        if (writer) store_tree(&d.trees[par ^ 1][prev], &sm.st[0].tr);
        __syncthreads();
    }
    if (cur >= 0) {
        const int fc = cur >= d.f[1].tree0 ? 1 : 0;
        load_tree(&sm.st[1].tr, &d.trees[par][cur]);
        __syncthreads();
        fetch_proposal(sm, 1, d, cur, sm.st[1].tr.nleaves);
        __syncthreads();
        begin_stats(sm, 1, fc, sc);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb)
        step_tile(sm, sm.st[1].ci, d, prev, cur, tile, lane);
    __syncthreads();
    if (cur >= 0) flush_stats(sm, sm.st[1].ci.K, d.totals, cur);
    if (threadIdx.x == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

// Sweep epilogue, observation pass (see last_tile)
__global__ void __launch_bounds__(THREADS) k_sweep_end(Dev d, int prev)
{
    BCF_STEP_SMEM(sm);
    const Scalars sc = *d.sc;
    const int par = sc.parity;
    const bool writer = (blockIdx.x == 0);
    if (threadIdx.x == 0) { sm.ai.active = 0; sm.bad = 0; }
    moments_begin(sm);
    __syncthreads();
    {
        const ForestDev& F = d.f[prev >= d.f[1].tree0 ? 1 : 0];
        load_tree(&sm.st[0].tr, &d.trees[par][prev]);
        __syncthreads();
        fetch_proposal(sm, 0, d, prev, sm.st[0].tr.nleaves);
        __syncthreads();
        decide(sm, 0, d, F, prev, sc, writer, d.totals);
        if (writer) store_tree(&d.trees[par ^ 1][prev], &sm.st[0].tr);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int bad = 0;
    for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb)
        last_tile(sm, d, prev, tile, lane, bad);
    if (bad) sm.bad = ERR_FX_RANGE;
    __syncthreads();
    moments_flush(sm, d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3));   // moment buffer of this sweep
    if (threadIdx.x == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

// ---------------------------------------------------------------------------------------------
// Sweep epilogue, scalar part (one CTA): scale draws, leaf-variance mixing, sigma [SURVEY A.4]; advances the
// sweep counter, flips the tree parity and decides whether this sweep is saved.
// ---------------------------------------------------------------------------------------------
// `scs` is the CTA's shared copy of the scalars (identical in every CTA); thread 0 updates it in place.
// mom_plain: `mom` is a CTA-local (shared memory) copy already summed over the shards -> plain loads
__device__ inline void sweep_scalars(const Dev& d, const TreeRec* newtrees, const long long* mom, double* red /* >= 2*blockDim */,
                                     Scalars* scs, bool write_global, bool mom_plain = false)
{
    const int t = threadIdx.x;
    // sum of mu^2 and leaf count over the mu trees, deterministic (fixed partition, fixed order)
    double ssq = 0.0, cnt = 0.0;
    for (int j = d.f[0].tree0 + t; j < d.f[0].tree0 + d.f[0].ntree; j += blockDim.x) {
        const TreeRec& tr = newtrees[j];
        const int L = __ldcg(&tr.nleaves);
        for (int l = 0; l < L; ++l) { const double m = __ldcg(&tr.mu[l]); ssq += m * m; }
        cnt += L;
    }
    red[t] = ssq; red[blockDim.x + t] = cnt;
    __syncthreads();
    if (t == 0) {
        Scalars sc = *scs;
        const Rng rng{d.key0, d.key1};
        double tssq = 0.0, tcnt = 0.0;
        // fixed order; the entries behind the last tree are zeros (adding them changes no bit): stop there
        const int nred = min((int)blockDim.x, d.f[0].ntree);
        for (int i = 0; i < nred; ++i) { tssq += red[i]; tcnt += red[blockDim.x + i]; }
        double M[2][NMOM];
        for (int g = 0; g < 2; ++g)
            for (int m = 0; m < NMOM; ++m) {
                const long long* q = mom + (g * NMOM + m) * 3;
                M[g][m] = mom_plain ? limbs_to_double(q[0], q[1], q[2]) : limbs_to_double(__ldcg(q), __ldcg(q + 1), __ldcg(q + 2));
            }
        const double s2 = sc.sigma * sc.sigma;
        double ms = sc.mscale, b1 = sc.bscale1, b0 = sc.bscale0;
        if (d.use_tauscale && d.f[1].ntree > 0) {
            // ww_g = sum Gm^2 / s2 ; rw_g = sum (y - ms Gc) Gm / s2   (moments: 0 yy, 1 yGc, 2 yGm, 3 GcGc, 4 GcGm, 5 GmGm)
            double v = 1.0 / (M[1][5] / s2 + 2.0);
            b1 = v * ((M[1][2] - ms * M[1][4]) / s2) + rng_normal(rng, sc.sweep, TREE_SWEEP_LEVEL, PUR_BSCALE1, 0) * sqrt(v);
            v = 1.0 / (M[0][5] / s2 + 2.0);
            b0 = v * ((M[0][2] - ms * M[0][4]) / s2) + rng_normal(rng, sc.sweep, TREE_SWEEP_LEVEL, PUR_BSCALE0, 0) * sqrt(v);
        }
        if (d.use_muscale) {
            const double ww = (M[1][3] + M[0][3]) / s2;
            const double rw = ((M[1][1] - b1 * M[1][4]) + (M[0][1] - b0 * M[0][4])) / s2;
            const double v = 1.0 / (ww + 1.0);
            ms = v * rw + rng_normal(rng, sc.sweep, TREE_SWEEP_LEVEL, PUR_MSCALE, 0) * sqrt(v);
            const double q = tssq / (sc.tau_con * sc.tau_con);
            sc.delta_con = rng_gamma(rng, sc.sweep, PUR_DELTA_CON, 0.5 * (1.0 + tcnt)) / (0.5 * (1.0 + q));
            sc.tau_con = d.con_sd / (sqrt(sc.delta_con) * sqrt((double)d.f[0].ntree));
        }
        double rss = 0.0;
        for (int g = 0; g < 2; ++g) {
            const double b = g ? b1 : b0;
            rss += M[g][0] + ms * ms * M[g][3] + b * b * M[g][5] - 2.0 * ms * M[g][1] - 2.0 * b * M[g][2] + 2.0 * ms * b * M[g][4];
        }
        if (!d.fix_sigma) {
            const double chi2 = 2.0 * rng_gamma(rng, sc.sweep, PUR_SIGMA, 0.5 * (d.nu + (double)d.n_global));
            sc.sigma = sqrt((d.nu * d.lambda + rss) / chi2);
        }
        sc.mscale = ms; sc.bscale1 = b1; sc.bscale0 = b0;
        if (!(sc.sigma == sc.sigma) || !(ms == ms) || !(b1 == b1) || !(b0 == b0)) atomicOr(d.err, ERR_NAN);
        // bcf saves iteration i when i >= burn and i % thin == 0
        sc.save_row = -1;
        if (sc.run_it >= sc.nburn && (sc.run_it % sc.nthin) == 0 && sc.save_ctr < sc.nsim) {
            sc.save_row = sc.save_ctr;
            if (write_global && sc.save_ctr < d.post_cap) {
                d.sigma_post[sc.save_ctr] = sc.sigma;
                d.msd_post[sc.save_ctr] = fabs(ms) * d.con_sd;
                d.bsd_post[sc.save_ctr] = fabs(b1 - b0) * d.mod_sd;
            }
            sc.save_ctr++;
        }
        sc.run_it++;
        sc.sweep++;
        sc.parity ^= 1;
        *scs = sc;
        if (write_global) *d.sc = sc;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(THREADS) k_scalars(Dev d)
{
    __shared__ double red[2 * THREADS];
    __shared__ Scalars scs;
    if (threadIdx.x == 0) scs = *d.sc;
    __syncthreads();
    sweep_scalars(d, d.trees[scs.parity ^ 1], d.mom + (size_t)(scs.sweep & 1u) * (2 * NMOM * 3), red, &scs, true);
}

// ---------------------------------------------------------------------------------------------
// Sweep epilogue, finishing pass: rebuild the full residual with the new scales, accumulate the posterior
// (K5: running sums instead of nsim x n matrices; optional draws), clear the per-sweep totals.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void finish_range(const Dev& d, const Scalars& sc, int64_t i0, int64_t stride)
{
    int bad = 0;
    const double ms = sc.mscale, b1 = sc.bscale1, b0 = sc.bscale0;
    const int64_t trt_end = (int64_t)d.trt_tiles * TILE;
    const int row = sc.save_row;
    for (int64_t i = i0; i < d.n_pad; i += stride) {
        const double gc = d.Gc[i], gm = d.Gm[i];
        const double b = i < trt_end ? b1 : b0;
        const double mu = ms * gc, yh = mu + b * gm;
        d.rfx[i] = resync_fx(d.y[i], ms, gc, b, gm, bad);
        if (row >= 0) {
            const double tau = (b1 - b0) * gm;
            d.tau_sum[i] += tau; d.tau_sq[i] += tau * tau; d.mu_sum[i] += mu; d.yhat_sum[i] += yh;
            if (d.draws_tau) d.draws_tau[(int64_t)row * d.n_pad + i] = tau;
            if (d.draws_mu) d.draws_mu[(int64_t)row * d.n_pad + i] = mu;
            if (d.draws_yhat) d.draws_yhat[(int64_t)row * d.n_pad + i] = yh;
        }
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
}
__global__ void __launch_bounds__(256) k_finish(Dev d)
{
    const Scalars sc = *d.sc;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    finish_range(d, sc, i0, stride);
    const int64_t ntot = (int64_t)d.T * KEYS * 8;
    for (int64_t i = i0; i < ntot; i += stride) d.totals[i] = 0;
    long long* mom_next = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);   // sc.sweep is already the next sweep's index
    for (int64_t i = i0; i < 2 * NMOM * 3; i += stride) mom_next[i] = 0;
}

// initial state [SURVEY A.2]: Gc = ybar (200 root leaves of ybar/200), Gm = 1 (50 root leaves of 1/50)
__global__ void k_init_state(Dev d, const int32_t* __restrict__ orig, double gc0, double gm0)
{
    const Scalars sc = *d.sc;
    const int64_t trt_end = (int64_t)d.trt_tiles * TILE;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        const bool real = orig[i] >= 0;
        const double gc = real ? gc0 : 0.0, gm = real ? gm0 : 0.0;
        d.Gc[i] = gc; d.Gm[i] = gm;
        int bad = 0;
        d.rfx[i] = real ? resync_fx(d.y[i], sc.mscale, gc, i < trt_end ? sc.bscale1 : sc.bscale0, gm, bad) : 0ll;
        if (bad) atomicOr(d.err, ERR_FX_RANGE);
        d.tau_sum[i] = d.tau_sq[i] = d.mu_sum[i] = d.yhat_sum[i] = 0.0;
        for (int t = 0; t < d.T; ++t) d.slots[(int64_t)t * d.n_pad + i] = real ? 0 : PAD_SLOT;
    }
}
// recompute Gc / Gm / r from the current trees (after set_tree)
__global__ void k_refit_all(Dev d, const TreeRec* trees)
{
    const Scalars sc = *d.sc;
    const int64_t trt_end = (int64_t)d.trt_tiles * TILE;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        double G[2] = {0.0, 0.0};
        for (int f = 0; f < 2; ++f)
            for (int j = d.f[f].tree0; j < d.f[f].tree0 + d.f[f].ntree; ++j) {
                const int s = d.slots[(int64_t)j * d.n_pad + i];
                if (s != PAD_SLOT) G[f] += trees[j].mu[s];
            }
        d.Gc[i] = G[0]; d.Gm[i] = G[1];
        int bad = 0;
        d.rfx[i] = resync_fx(d.y[i], sc.mscale, G[0], i < trt_end ? sc.bscale1 : sc.bscale0, G[1], bad);
        if (bad) atomicOr(d.err, ERR_FX_RANGE);
    }
}

// per-leaf observation counts of one tree from its cached slots: out[slot][0 control / 1 treated]
__global__ void k_count_leaves(Dev d, int tree, long long* out)
{
    const int64_t trt_end = (int64_t)d.trt_tiles * TILE;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = d.slots[(int64_t)tree * d.n_pad + i];
        if (s != PAD_SLOT) atomicAdd((unsigned long long*)&out[2 * s + (i < trt_end ? 1 : 0)], 1ull);
    }
}
// gather an internal-order vector back to the caller's row order; which: 0 copy, 1 mean (x/ns), 2 variance
__global__ void k_unpermute(const double* __restrict__ a, const double* __restrict__ b, const int32_t* __restrict__ pos,
                            int64_t n, double ns, int which, double scale_trt, double scale_ctl, int64_t trt_end,
                            double* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = pos[i];
        double v;
        if (which == 0) v = a[q] * (q < trt_end ? scale_trt : scale_ctl);
        else if (which == 1) v = a[q] / ns;
        else { const double m = a[q] / ns; v = ns > 1.0 ? fmax(0.0, (b[q] - ns * m * m) / (ns - 1.0)) : 0.0; }
        out[i] = v;
    }
}
__global__ void k_permute_in(const double* __restrict__ in, const int32_t* __restrict__ pos, int64_t n, double* out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[pos[i]] = in[i];
}
__global__ void k_leaf_heap(Dev d, int tree, const TreeRec* trees, const int32_t* __restrict__ pos, int64_t n,
                            uint32_t* out)
{
    const TreeRec& tr = trees[tree];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = d.slots[(int64_t)tree * d.n_pad + pos[i]];
        out[i] = s == PAD_SLOT ? 0u : tr.heap[tr.slot_node[s]];
    }
}

// ---------------------------------------------------------------------------------------------
// K2 stand-alone: per-leaf (count, sum r) by group for a caller-supplied residual, optional birth split.
// Same accumulation code as the sampler's fused pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) k_op_stats(Dev d, int tree, int forest, const double* __restrict__ r,
                                                      int K, int birth, int sx, int var, int cut, long long* out)
{
    __shared__ StatTab tab;
    for (int i = threadIdx.x; i < STATTAB_WORDS64; i += blockDim.x) reinterpret_cast<long long*>(&tab)[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int bad = 0;
    for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb) {
        const int64_t base = (int64_t)tile * TILE + lane * 8;
        const int g = tile < d.trt_tiles ? 1 : 0;
        double rd[8];
        long long rv[8];
        int key[8];
        ld8_f64(r + base, rd);
#pragma unroll
        for (int o = 0; o < 8; ++o) rv[o] = to_fixed(rd[o], bad);
        ld8_u8(d.slots + (int64_t)tree * d.n_pad + base, key);
        if (birth) {
            int bc[8];
            int w16;
            const void* col = bin_column(d.f[forest], var, d.n_pad, w16);
            ld8_bins(col, w16, base, bc);
#pragma unroll
            for (int o = 0; o < 8; ++o) if (key[o] == sx && bc[o] > cut) key[o] = K - 1;
        }
        accum_tile(tab, key, rv, g, K, true, lane, wib);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 8; i += blockDim.x) {
        const long long v = stat_word(tab, K, wpb, i >> 3, (i >> 2) & 1, i & 3);
        if (v != 0) atomicAdd((unsigned long long*)&out[i], (unsigned long long)v);
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
}

// ---------------------------------------------------------------------------------------------
// K2': all-candidate histogram.  For every variable v of a forest and every bin b, per leaf label:
// (count, sum r) of the observations with that (leaf, bin) -- the sufficient statistics of EVERY candidate
// (variable, cutpoint) split at once (prefix sums over b give the left child of cut c).
// grid.y = variable.  uint8 columns: shared-memory histogram per CTA (leaf x bin x 4 words), flushed with
// integer REDs.  uint16 columns (10 000 cutpoints): global integer atomics directly.
// Loads: 128-bit for the uint8 bins / labels (16 observations per lane), 128-bit pairs for r.
// ---------------------------------------------------------------------------------------------
constexpr int HIST_SMEM_ENTRIES = 1024;   // (leaf, bin) cells of 4 x int64 = 32 KB
__global__ void __launch_bounds__(256) k_hist(Dev d, int forest, const uint8_t* __restrict__ label,
                                              const double* __restrict__ r, int nleaf, const int32_t* __restrict__ off,
                                              long long* out /* [leaf][off[p]+p][4] */, const int32_t* __restrict__ cols)
{
    __shared__ long long h[HIST_SMEM_ENTRIES][4];
    const ForestDev& F = d.f[forest];
    const int v = cols ? cols[blockIdx.y] : blockIdx.y;     // cols: the columns k_hist_small does not take
    const int nb = F.ncut[v] + 1;
    const int64_t nbt = (int64_t)off[F.p] + F.p;
    const int64_t vbase = (int64_t)off[v] + v;
    const bool w16 = F.col_is16[v] != 0;
    const bool use_smem = !w16 && nleaf * nb <= HIST_SMEM_ENTRIES;
    if (use_smem) for (int i = threadIdx.x; i < nleaf * nb * 4; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    int bad = 0;
    const int ci = F.col_index[v];
    const int64_t nchunk = d.n_pad / 16;
    for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunk; ch += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = ch * 16;
        int lab[16], bin[16];
        {
            const uint4 q = *reinterpret_cast<const uint4*>(label + base);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int o = 0; o < 16; ++o) lab[o] = (w[o >> 2] >> (8 * (o & 3))) & 0xFF;
        }
        if (w16) {
            int t8[8];
            ld8_u16(F.bins16 + (int64_t)ci * d.n_pad + base, t8);
#pragma unroll
            for (int o = 0; o < 8; ++o) bin[o] = t8[o];
            ld8_u16(F.bins16 + (int64_t)ci * d.n_pad + base + 8, t8);
#pragma unroll
            for (int o = 0; o < 8; ++o) bin[8 + o] = t8[o];
        } else {
            const uint4 q = *reinterpret_cast<const uint4*>(F.bins8 + (int64_t)ci * d.n_pad + base);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int o = 0; o < 16; ++o) bin[o] = (w[o >> 2] >> (8 * (o & 3))) & 0xFF;
        }
#pragma unroll
        for (int o = 0; o < 16; o += 2) {
            const double2 rr = *reinterpret_cast<const double2*>(r + base + o);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lb = lab[o + e];
                if (lb >= nleaf) continue;
                const long long fx = to_fixed(e ? rr.y : rr.x, bad);
                const long long l0 = fx & 0x1FFFFF, l1 = (fx >> LIMB_BITS) & 0x1FFFFF, l2 = fx >> (2 * LIMB_BITS);
                long long* cell = use_smem ? h[lb * nb + bin[o + e]] : out + ((int64_t)lb * nbt + vbase + bin[o + e]) * 4;
                atomicAdd((unsigned long long*)&cell[0], 1ull);
                atomicAdd((unsigned long long*)&cell[1], (unsigned long long)l0);
                atomicAdd((unsigned long long*)&cell[2], (unsigned long long)l1);
                atomicAdd((unsigned long long*)&cell[3], (unsigned long long)l2);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nleaf * nb * 4; i += blockDim.x) {
            const long long val = (&h[0][0])[i];
            if (val == 0) continue;
            const int cellid = i >> 2, q = i & 3, lb = cellid / nb, b = cellid % nb;
            atomicAdd((unsigned long long*)&out[((int64_t)lb * nbt + vbase + b) * 4 + q], (unsigned long long)val);
        }
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
}

// K2' for columns with few bins (binary / categorical covariates: what BayesIV's designs are).  One pass over the
// observations serves ALL such columns: a warp loads a tile's labels and residuals once (128-bit / 64-bit coalesced
// loads), then per column its 8 bin bytes per lane, and accumulates every (leaf, bin >= 1) cell with a masked sum of
// the fixed-point residual + warp REDUX into a shared-memory table of 32-bit words (count, three 21-bit limbs).  Bin 0
// of a (leaf, column) is the leaf total minus the other bins (exact in integers), formed when the CTA flushes.
// Algorithmic traffic: n * (9 + columns) bytes -- every byte is read once.
// K2' for two-valued columns (the reference's covariates are binary: one cutpoint, bins {0, 1}).  Per observation and
// column the work is ONE predicated fp64 add: every lane keeps, in registers, the sum of r and the count over its
// observations for each (column, leaf) cell of the CTA's column chunk; bin 0 is (leaf total - bin 1), formed by the
// host from the leaf totals the chunk-0 CTAs add.  grid = (tile ranges, column chunks).  The lane sums are fp64 in a
// fixed order (tile order), reduced over the warp by shuffles and over the CTA's warps in warp order, then rounded ONCE
// to the 2^-44 fixed-point grid and added with integer atomics: reproducible for a given launch shape, and within
// 1e-13 relative of the exact sums.
#ifndef HISTB_CELLS_N
#define HISTB_CELLS_N 12
#endif
// bit o of `bits` as the double 0.0 or 1.0 (its high word is 0 or 0x3FF00000): acc = fma(bit, r, acc) is then the
// predicated add `if (bit) acc += r` in three instructions (fma(1, r, acc) and acc + r round identically)
__device__ __forceinline__ double bit_as_double(uint32_t bits, int o)
{
    return __hiloint2double((int)(((bits >> o) & 1u) * 0x3FF00000u), 0);
}
constexpr int HISTB_CELLS = HISTB_CELLS_N;   // (column, leaf) cells per CTA: that many sums + counts in registers (two CTAs per SM)
template <int NLEAF>
__global__ void __launch_bounds__(256, 2) k_hist_bits(Dev d, int forest, const uint8_t* __restrict__ label,
                                                   const double* __restrict__ r, const int32_t* __restrict__ off,
                                                   const int32_t* __restrict__ cols, int ncols, int tiles_per_cta,
                                                   long long* out, long long* leaf_tot_out)
{
    constexpr int CPC = HISTB_CELLS / NLEAF;                     // columns per chunk
    __shared__ double ssum[8][HISTB_CELLS + NLEAF];
    __shared__ int scnt[8][HISTB_CELLS + NLEAF];
    const ForestDev& F = d.f[forest];
    const int t = threadIdx.x, lane = t & 31, wib = t >> 5;
    const int c0 = blockIdx.y * CPC, nc = min(CPC, ncols - c0);
    const bool totals = blockIdx.y == 0;
    double acc[CPC][NLEAF], tot[NLEAF];
    int cnt[CPC][NLEAF], tcnt[NLEAF];
#pragma unroll
    for (int l = 0; l < NLEAF; ++l) {
        tot[l] = 0.0; tcnt[l] = 0;
#pragma unroll
        for (int c = 0; c < CPC; ++c) { acc[c][l] = 0.0; cnt[c][l] = 0; }
    }
    const uint8_t* colp[CPC];
#pragma unroll
    for (int c = 0; c < CPC; ++c) colp[c] = F.bins8 + (int64_t)F.col_index[cols[c0 + min(c, nc - 1)]] * d.n_pad;
    const int tile_lo = blockIdx.x * tiles_per_cta, tile_hi = min(d.n_tiles, tile_lo + tiles_per_cta);
    for (int tile = tile_lo + wib; tile < tile_hi; tile += 8) {
        const int64_t base = (int64_t)tile * TILE + lane * 8;
        const uint2 lab = *reinterpret_cast<const uint2*>(label + base);
        uint2 bn[CPC];
#pragma unroll
        for (int c = 0; c < CPC; ++c) bn[c] = *reinterpret_cast<const uint2*>(colp[c] + base);
        double rd[8];
        ld8_f64(r + base, rd);
        uint32_t lb[NLEAF];                                       // which of the lane's 8 observations sit in leaf l (labels >= NLEAF: none)
#pragma unroll
        for (int l = 0; l < NLEAF; ++l) {
            const uint32_t lv = (uint32_t)l * 0x01010101u;
            lb[l] = bytes_to_bits(__vcmpeq4(lab.x, lv), __vcmpeq4(lab.y, lv));
            if (totals) {
#pragma unroll
                for (int o = 0; o < 8; ++o) tot[l] = fma(bit_as_double(lb[l], o), rd[o], tot[l]);
                tcnt[l] += __popc(lb[l]);
            }
        }
#pragma unroll
        for (int c = 0; c < CPC; ++c) {
            if (c < nc) {
                const uint32_t one = bytes_to_bits(__vcmpeq4(bn[c].x, 0x01010101u), __vcmpeq4(bn[c].y, 0x01010101u));
#pragma unroll
                for (int l = 0; l < NLEAF; ++l) {
                    const uint32_t bits = one & lb[l];
#pragma unroll
                    for (int o = 0; o < 8; ++o) acc[c][l] = fma(bit_as_double(bits, o), rd[o], acc[c][l]);
                    cnt[c][l] += __popc(bits);
                }
            }
        }
    }
    // lanes -> warp (xor tree), warps -> CTA (warp order), then one rounding to fixed point
#pragma unroll
    for (int c = 0; c < CPC; ++c)
#pragma unroll
        for (int l = 0; l < NLEAF; ++l) {
            double v = acc[c][l];
            int k = cnt[c][l];
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, sft); k += __shfl_xor_sync(0xffffffffu, k, sft); }
            if (lane == 0) { ssum[wib][c * NLEAF + l] = v; scnt[wib][c * NLEAF + l] = k; }
        }
#pragma unroll
    for (int l = 0; l < NLEAF; ++l) {
        double v = tot[l];
        int k = tcnt[l];
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, sft); k += __shfl_xor_sync(0xffffffffu, k, sft); }
        if (lane == 0) { ssum[wib][HISTB_CELLS + l] = v; scnt[wib][HISTB_CELLS + l] = k; }
    }
    __syncthreads();
    if (t < CPC * NLEAF || (t >= HISTB_CELLS && t < HISTB_CELLS + NLEAF)) {   // (cells CPC * NLEAF .. HISTB_CELLS - 1 do not exist when NLEAF does not divide HISTB_CELLS)
        double v = 0.0;
        long long k = 0;
        for (int w = 0; w < 8; ++w) { v += ssum[w][t]; k += scnt[w][t]; }
        int bad = 0;
        const long long fx = to_fixed(v, bad);
        long long* o = nullptr;
        if (t < HISTB_CELLS) {
            const int c = t / NLEAF, l = t % NLEAF;
            if (c < nc) {
                const int col = cols[c0 + c];
                const int64_t nbt = (int64_t)off[F.p] + F.p;
                o = out + ((int64_t)l * nbt + (int64_t)off[col] + col + 1) * 4;      // bin 1
            }
        } else if (totals) o = leaf_tot_out + (t - HISTB_CELLS) * 4;
        if (o != nullptr && (k != 0 || fx != 0)) {
            atomicAdd((unsigned long long*)&o[0], (unsigned long long)k);
            atomicAdd((unsigned long long*)&o[1], (unsigned long long)(fx & 0x1FFFFF));
            atomicAdd((unsigned long long*)&o[2], (unsigned long long)((fx >> LIMB_BITS) & 0x1FFFFF));
            atomicAdd((unsigned long long*)&o[3], (unsigned long long)(fx >> (2 * LIMB_BITS)));
        }
        if (bad) atomicOr(d.err, ERR_FX_RANGE);
    }
}

constexpr int HIST2_MAX_BINS = 4;        // bins 0..3 (up to 3 cutpoints)
constexpr int HIST2_MAX_LEAF = 16;
constexpr int HIST2_CELL_WORDS = 49152 / 16;   // shared table: cells of 4 x u32
__global__ void __launch_bounds__(256) k_hist_small(Dev d, int forest, const uint8_t* __restrict__ label,
                                                    const double* __restrict__ r, int nleaf, const int32_t* __restrict__ off,
                                                    const int32_t* __restrict__ cols, int ncols, const int32_t* __restrict__ cellbase,
                                                    int ncells, int tiles_per_cta, long long* out, int per_obs)
{
    extern __shared__ __align__(16) unsigned int h2[];          // [ncells + nleaf][4]: cells, then the leaf totals
    if (per_obs) {
        // Many (leaf, bin) cells per column: one native 32-bit shared atomic per word and observation instead of one
        // masked warp sum per cell.  Cells hold (count, four 16-bit limbs) in 8 words; every bin (0 included) has its own.
        const ForestDev& F = d.f[forest];
        const int t = threadIdx.x, lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
        for (int i = t; i < ncells * 8; i += blockDim.x) h2[i] = 0u;
        __syncthreads();
        int bad = 0;
        const int tile_lo = blockIdx.x * tiles_per_cta, tile_hi = min(d.n_tiles, tile_lo + tiles_per_cta);
        for (int tile = tile_lo + wib; tile < tile_hi; tile += wpb) {
            const int64_t base = (int64_t)tile * TILE + lane * 8;
            int lab[8];
            ld8_u8(label + base, lab);
            double rd[8];
            long long R[8];
            ld8_f64(r + base, rd);
#pragma unroll
            for (int o = 0; o < 8; ++o) R[o] = to_fixed(rd[o], bad);
            uint2 bn_n = *reinterpret_cast<const uint2*>(F.bins8 + (int64_t)F.col_index[cols[0]] * d.n_pad + base);
            for (int ci = 0; ci < ncols; ++ci) {
                const int v = cols[ci];
                const uint2 bn = bn_n;
                if (ci + 1 < ncols) bn_n = *reinterpret_cast<const uint2*>(F.bins8 + (int64_t)F.col_index[cols[ci + 1]] * d.n_pad + base);
                const int nb = F.ncut[v] + 1;
                unsigned int* cell0 = h2 + (size_t)cellbase[ci] * 8;
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    if (lab[o] >= nleaf) continue;
                    const int bin = (int)(((o < 4 ? bn.x : bn.y) >> (8 * (o & 3))) & 0xFFu);
                    unsigned int* e = cell0 + 8 * (lab[o] * nb + bin);
                    const long long vv = R[o];
                    atomicAdd(e, 1u);
                    atomicAdd(e + 1, (unsigned int)(vv & 0xFFFF));
                    atomicAdd(e + 2, (unsigned int)((vv >> 16) & 0xFFFF));
                    atomicAdd(e + 3, (unsigned int)((vv >> 32) & 0xFFFF));
                    atomicAdd(e + 4, (unsigned int)(int)(vv >> 48));
                }
            }
        }
        __syncthreads();
        const int64_t nbt = (int64_t)off[F.p] + F.p;
        for (int i = t; i < ncells; i += blockDim.x) {
            // which (column, leaf, bin) is cell i: columns are few, a linear search is fine
            int ci = 0;
            while (ci + 1 < ncols && cellbase[ci + 1] <= i) ++ci;
            const int v = cols[ci], nb = F.ncut[v] + 1, j = i - cellbase[ci], l = j / nb, b = j % nb;
            const unsigned int* e = h2 + (size_t)i * 8;
            const __int128 V = (__int128)e[1] + ((__int128)e[2] << 16) + ((__int128)e[3] << 32) + ((__int128)(int)e[4] << 48);
            long long* o = out + ((int64_t)l * nbt + (int64_t)off[v] + v + b) * 4;
            const long long w0 = (long long)e[0], w1 = (long long)(V & 0x1FFFFF), w2 = (long long)((V >> LIMB_BITS) & 0x1FFFFF), w3 = (long long)(V >> (2 * LIMB_BITS));
            if (w0) atomicAdd((unsigned long long*)&o[0], (unsigned long long)w0);
            if (w1) atomicAdd((unsigned long long*)&o[1], (unsigned long long)w1);
            if (w2) atomicAdd((unsigned long long*)&o[2], (unsigned long long)w2);
            if (w3) atomicAdd((unsigned long long*)&o[3], (unsigned long long)w3);
        }
        if (bad) atomicOr(d.err, ERR_FX_RANGE);
        return;
    }
    const ForestDev& F = d.f[forest];
    const int t = threadIdx.x, lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    for (int i = t; i < (ncells + nleaf) * 4; i += blockDim.x) h2[i] = 0u;
    __syncthreads();
    unsigned int* leaf_tot = h2 + (size_t)ncells * 4;
    int bad = 0;
    const int tile_lo = blockIdx.x * tiles_per_cta, tile_hi = min(d.n_tiles, tile_lo + tiles_per_cta);
    for (int tile = tile_lo + wib; tile < tile_hi; tile += wpb) {
        const int64_t base = (int64_t)tile * TILE + lane * 8;
        const uint2 lab = *reinterpret_cast<const uint2*>(label + base);
        double rd[8];
        long long R[8];
        ld8_f64(r + base, rd);
#pragma unroll
        for (int o = 0; o < 8; ++o) R[o] = to_fixed(rd[o], bad);
        // which of the lane's 8 observations carry each leaf label (labels >= nleaf: skipped)
        unsigned long long lbA = 0ull, lbB = 0ull;      // 8 bits per leaf: leaves 0..7 / 8..15
        for (int l = 0; l < nleaf; ++l) {
            const uint32_t lv = (uint32_t)l * 0x01010101u;
            const uint32_t bits = bytes_to_bits(__vcmpeq4(lab.x, lv), __vcmpeq4(lab.y, lv));
            if (l < 8) lbA |= (unsigned long long)bits << (8 * l); else lbB |= (unsigned long long)bits << (8 * (l - 8));
            long long s = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) s += ((bits >> o) & 1u) ? R[o] : 0ll;
            warp_add21(leaf_tot + 4 * l, s, lane);
            const int c = __reduce_add_sync(0xffffffffu, __popc(bits));
            if (lane == 0 && c) atomicAdd(leaf_tot + 4 * l, (unsigned int)c);
        }
        uint2 bn_n = *reinterpret_cast<const uint2*>(F.bins8 + (int64_t)F.col_index[cols[0]] * d.n_pad + base);
        for (int ci = 0; ci < ncols; ++ci) {
            const int v = cols[ci];
            const uint2 bn = bn_n;
            if (ci + 1 < ncols) bn_n = *reinterpret_cast<const uint2*>(F.bins8 + (int64_t)F.col_index[cols[ci + 1]] * d.n_pad + base);
            const int nb = F.ncut[v] + 1;
            unsigned int* cell0 = h2 + (size_t)cellbase[ci] * 4;      // cells of this column: [leaf][bin - 1]
            for (int b = 1; b < nb; ++b) {
                const uint32_t bv = (uint32_t)b * 0x01010101u;
                const uint32_t binbits = bytes_to_bits(__vcmpeq4(bn.x, bv), __vcmpeq4(bn.y, bv));
                for (int l = 0; l < nleaf; ++l) {
                    const uint32_t bits = binbits & (uint32_t)((l < 8 ? lbA >> (8 * l) : lbB >> (8 * (l - 8))) & 0xFFull);
                    long long s = 0;
#pragma unroll
                    for (int o = 0; o < 8; ++o) s += ((bits >> o) & 1u) ? R[o] : 0ll;
                    unsigned int* e = cell0 + 4 * (l * (nb - 1) + (b - 1));
                    warp_add21(e, s, lane);
                    const int c = __reduce_add_sync(0xffffffffu, __popc(bits));
                    if (lane == 0 && c) atomicAdd(e, (unsigned int)c);
                }
            }
        }
    }
    __syncthreads();
    // flush: bins >= 1 directly, bin 0 = leaf total - the others
    const int64_t nbt = (int64_t)off[F.p] + F.p;
    for (int i = t; i < ncols * nleaf; i += blockDim.x) {
        const int ci = i / nleaf, l = i % nleaf, v = cols[ci];
        const int nb = F.ncut[v] + 1;
        const int64_t vbase = (int64_t)off[v] + v;
        long long rest[4] = {(long long)leaf_tot[4 * l], acc_word(leaf_tot + 4 * l, 1), acc_word(leaf_tot + 4 * l, 2), acc_word(leaf_tot + 4 * l, 3)};
        for (int b = 1; b < nb; ++b) {
            const unsigned int* e = h2 + (size_t)(cellbase[ci] + l * (nb - 1) + (b - 1)) * 4;
            long long* o = out + ((int64_t)l * nbt + vbase + b) * 4;
            for (int q = 0; q < 4; ++q) {
                const long long w = acc_word(e, q);
                rest[q] -= w;
                if (w != 0) atomicAdd((unsigned long long*)&o[q], (unsigned long long)w);
            }
        }
        long long* o0 = out + ((int64_t)l * nbt + vbase) * 4;
        for (int q = 0; q < 4; ++q) if (rest[q] != 0) atomicAdd((unsigned long long*)&o0[q], (unsigned long long)rest[q]);
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
}

// ---------------------------------------------------------------------------------------------
// Probit BART: the sampler behind bartCause::bartc for a 0/1 response, which the reference calls at
// R/bcf_iv.R:78-80 (complier proportion) and R/bcf_iv.R:198-202, R/bcf_itt.R:166-170 (binary = TRUE).  Chipman-George-
// McCulloch BART with Albert & Chib's latent Gaussian: the forest machinery is the mu forest of the sweep kernels run
// alone (no tau forest, no scales, sigma = 1); what is added is the latent draw that opens every sweep and, per saved
// sweep, the S-learner's individual effect  Phi(offset + f(x, z = 1)) - Phi(offset + f(x, z = 0)).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm_cdf_d(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
// Wichura's AS 241 (PPND16): the expression of the oracle's norm_inv, term for term
__device__ inline double norm_inv_d(double p)
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
// ystar ~ N(m, 1) truncated to (0, inf) when y = 1, to (-inf, 0] when y = 0, by inversion of one uniform on the side
// where the cdf is small
__device__ __forceinline__ double probit_latent_d(double m, int y, double u)
{
    return y ? m - norm_inv_d(u * norm_cdf_d(m)) : m + norm_inv_d(u * norm_cdf_d(-m));
}
constexpr uint32_t PUR_LATENT = 20;       // idx = the caller's row number
// after k_place: the permuted outcome is the 0/1 response; keep it as bytes, the working response starts at 0
__global__ void k_probit_prepare(double* __restrict__ y, uint8_t* __restrict__ ybin, int64_t n_pad)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        ybin[i] = y[i] != 0.0 ? 1 : 0;
        y[i] = 0.0;
    }
}
// the latent draw that opens a sweep: working response y = ystar - offset, full residual re-derived from it
// stall (here and below, may be null): set when a queued sweep kernel left its sweep undone -- everything queued behind it returns
__global__ void __launch_bounds__(256) k_probit_latent(Dev d, const uint8_t* __restrict__ ybin, const int32_t* __restrict__ orig,
                                                       int64_t row0, double offset, double* __restrict__ y, const int* stall)
{
    if (stall != nullptr && __ldcg(stall) != 0) return;
    const Scalars sc = *d.sc;
    const Rng rng{d.key0, d.key1};
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t o = orig[i];
        if (o < 0) continue;                                    // padding: y = 0, residual 0
        double ua, ub;
        rng_u2(rng, sc.sweep, TREE_SWEEP_LEVEL, PUR_LATENT, (uint32_t)(row0 + o), ua, ub);
        const double gc = d.Gc[i];
        const double ys = probit_latent_d(offset + sc.mscale * gc, ybin[i], ua) - offset;
        if (!(ys == ys)) bad = 1;
        y[i] = ys;
        d.rfx[i] = resync_fx(ys, sc.mscale, gc, 0.0, 0.0, bad);
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
}
// which trees split on the treatment column?  one CTA per tree; list[1 + t] = 1 if tree t does (a flag per tree, not a list
// in arrival order: the pass below then adds the trees' contributions in tree order, so the effects are reproducible bit for bit)
__global__ void k_cf_collect(Dev d, const TreeRec* __restrict__ trees, int trt_var, int* list, const int* stall)
{
    if (stall != nullptr && __ldcg(stall) != 0) return;
    const TreeRec& tr = trees[blockIdx.x];
    const int nn = __ldcg(&tr.nnodes);
    int hit = 0;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) hit |= (__ldcg(&tr.left[i]) >= 0 && (int)__ldcg(&tr.var[i]) == trt_var) ? 1 : 0;
    hit = __syncthreads_or(hit);
    if (threadIdx.x == 0) list[1 + blockIdx.x] = hit ? 1 : 0;
}
// One saved sweep of the S-learner: f with the treatment flipped differs from the fitted f only through the trees that
// split on the treatment; those are walked again with the flipped bin.  Accumulates (internal order)
//   tau_sum / tau_sq : Phi(offset + f(x, 1)) - Phi(offset + f(x, 0))  and its square,
//   mu_sum           : Phi(offset + f(x)) at the observed treatment,   yhat_sum: f(x) itself.
constexpr int CF_OBS_PER_THREAD = 4;
__global__ void __launch_bounds__(256) k_cf_accumulate(Dev d, const TreeRec* __restrict__ trees, const int* __restrict__ list,
                                                       int trt_var, double offset, const int32_t* __restrict__ orig, const int* stall)
{
    __shared__ TreeRec tr;
    if (stall != nullptr && __ldcg(stall) != 0) return;
    const ForestDev& F = d.f[0];
    const int nlist = trt_var >= 0 ? d.T : 0;
    const int64_t base = (int64_t)blockIdx.x * (256 * CF_OBS_PER_THREAD) + threadIdx.x;
    double diff[CF_OBS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < CF_OBS_PER_THREAD; ++k) diff[k] = 0.0;
    int tb[CF_OBS_PER_THREAD];                                   // observed treatment (bin of the treatment column: 0 / 1)
#pragma unroll
    for (int k = 0; k < CF_OBS_PER_THREAD; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        tb[k] = 0;
        if (trt_var >= 0 && i < d.n_pad) {
            const int ci = F.col_index[trt_var];
            tb[k] = F.col_is16[trt_var] ? (int)F.bins16[(int64_t)ci * d.n_pad + i] : (int)F.bins8[(int64_t)ci * d.n_pad + i];
        }
    }
    for (int t = 0; t < nlist; ++t) {
        if (list[1 + t] == 0) continue;                          // (the same for the whole CTA)
        __syncthreads();
        load_tree(&tr, &trees[t]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CF_OBS_PER_THREAD; ++k) {
            const int64_t i = base + (int64_t)k * 256;
            if (i >= d.n_pad || orig[i] < 0) continue;
            int node = 0;
            while (tr.left[node] >= 0) {
                const int v = tr.var[node];
                int b;
                if (v == trt_var) b = tb[k] ? 0 : 1;             // the other arm
                else {
                    const int ci = F.col_index[v];
                    b = F.col_is16[v] ? (int)F.bins16[(int64_t)ci * d.n_pad + i] : (int)F.bins8[(int64_t)ci * d.n_pad + i];
                }
                node = (b <= (int)tr.cut[node]) ? tr.left[node] : tr.right[node];
            }
            const int s_fact = d.slots[(int64_t)t * d.n_pad + i];
            diff[k] += tr.mu[tr.node_slot[node]] - tr.mu[s_fact];
        }
    }
#pragma unroll
    for (int k = 0; k < CF_OBS_PER_THREAD; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        if (i >= d.n_pad || orig[i] < 0) continue;
        const double f = d.Gc[i], fcf = f + diff[k];
        const double pf = norm_cdf_d(offset + f);
        double ite = 0.0;
        if (trt_var >= 0) {
            const double pcf = norm_cdf_d(offset + fcf);
            ite = tb[k] ? pf - pcf : pcf - pf;
        }
        d.tau_sum[i] += ite; d.tau_sq[i] += ite * ite; d.mu_sum[i] += pf; d.yhat_sum[i] += f;
    }
}
__global__ void k_probit_probe(const double* m, const int* y, const double* u, int n, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = probit_latent_d(m[i], y[i], u[i]);
}

// ---------------------------------------------------------------------------------------------
// Saved forests and out-of-sample evaluation (bcf's `$tau` contract beyond colMeans, and predict()).
// keep_draws bit 3: after every saved sweep its trees are compacted into a node pool on the device (16 bytes per node:
// split rule or leaf value, children as node indices), together with the sweep's signed scales; bcfgpu_predict() bins new
// covariate rows with the training cutpoints (K0) and walks every saved forest over them.
// ---------------------------------------------------------------------------------------------
struct __align__(16) SavedNode { uint16_t var, cut; int16_t left, right; double mu; };    // leaf: left = -1, mu = leaf value
struct SavedDraw { double mscale, bscale1, bscale0, pad_; };
// one CTA per tree: the draw's tree `blockIdx.x` -> pool (bump allocation: the order of the trees in the pool is arbitrary,
// their offsets are recorded).  off[draw * (T + 1) + tree] = first node, count in cnt[...].
__global__ void k_save_trees(Dev d, const TreeRec* __restrict__ trees, int draw, SavedNode* pool, long long pool_cap,
                             unsigned long long* pool_used, long long* off, int32_t* cnt, SavedDraw* scal, const int* stall)
{
    __shared__ long long base;
    if (stall != nullptr && __ldcg(stall) != 0) return;
    const TreeRec& tr = trees[blockIdx.x];
    const int nn = __ldcg(&tr.nnodes);
    if (threadIdx.x == 0) {
        const unsigned long long b = atomicAdd(pool_used, (unsigned long long)nn);
        base = (b + (unsigned long long)nn <= (unsigned long long)pool_cap) ? (long long)b : -1;
        off[(size_t)draw * d.T + blockIdx.x] = base;
        cnt[(size_t)draw * d.T + blockIdx.x] = nn;
        if (blockIdx.x == 0) { const Scalars sc = *d.sc; scal[draw] = SavedDraw{sc.mscale, sc.bscale1, sc.bscale0, 0.0}; }
    }
    __syncthreads();
    if (base < 0) return;                                          // pool exhausted: predict reports it
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        SavedNode nd;
        nd.left = __ldcg(&tr.left[i]); nd.right = __ldcg(&tr.right[i]);
        nd.var = __ldcg(&tr.var[i]); nd.cut = __ldcg(&tr.cut[i]);
        nd.mu = nd.left < 0 ? __ldcg(&tr.mu[__ldcg(&tr.node_slot[i])]) : 0.0;
        pool[base + i] = nd;
    }
}
// K0 for new rows: no permutation, every column as uint16, [col][n]
__global__ void k_bin_plain(const double* __restrict__ x, int64_t n, int p, const double* __restrict__ cuts,
                            const int32_t* __restrict__ off, uint16_t* __restrict__ out)
{
    const int v = blockIdx.y;
    if (v >= p) return;
    const double* cv = cuts + off[v];
    const int nc = off[v + 1] - off[v];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)v * n + i] = (uint16_t)upper_bound_cut(cv, nc, x[(int64_t)v * n + i]);
}
// one thread per new row, draws [d0, d1): sum over the draw's trees of the leaf the row falls in, per forest; accumulates
// mean / second moment of tau = (bscale1 - bscale0) * g_mod and of mu = mscale * g_con, optionally every draw
__global__ void __launch_bounds__(128) k_predict(int64_t n, int T, int ntree_con, int d0, int d1, const SavedNode* __restrict__ pool,
                                                 const long long* __restrict__ off, const SavedDraw* __restrict__ scal,
                                                 const uint16_t* __restrict__ bc, const uint16_t* __restrict__ bm,
                                                 double* tau_sum, double* tau_sq, double* mu_sum, double* tau_draws, double* mu_draws)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ts = 0.0, tq = 0.0, ms = 0.0;
    for (int dr = d0; dr < d1; ++dr) {
        double g[2] = {0.0, 0.0};
        for (int t = 0; t < T; ++t) {
            const long long b = off[(size_t)dr * T + t];
            const int f = t >= ntree_con ? 1 : 0;
            const uint16_t* bins = f ? bm : bc;
            int node = 0;
            SavedNode nd = pool[b];
            while (nd.left >= 0) {
                node = ((int)bins[(int64_t)nd.var * n + i] <= (int)nd.cut) ? nd.left : nd.right;
                nd = pool[b + node];
            }
            g[f] += nd.mu;
        }
        const SavedDraw sd = scal[dr];
        const double tau = (sd.bscale1 - sd.bscale0) * g[1], mu = sd.mscale * g[0];
        ts += tau; tq += tau * tau; ms += mu;
        if (tau_draws) tau_draws[(int64_t)dr * n + i] = tau;
        if (mu_draws) mu_draws[(int64_t)dr * n + i] = mu;
    }
    tau_sum[i] += ts; tau_sq[i] += tq; mu_sum[i] += ms;
}

__global__ void k_lil(const double* a, const double* b, const double* tau, int m, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = lil(a[i], b[i], tau[i]);
}
__global__ void k_rng_probe(uint32_t k0, uint32_t k1, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx,
                            double* out)
{
    const Rng g{k0, k1};
    rng_u2(g, sweep, tree, purpose, idx, out[0], out[1]);
    out[2] = rng_normal(g, sweep, tree, purpose, idx);
}

}  // namespace bcf
