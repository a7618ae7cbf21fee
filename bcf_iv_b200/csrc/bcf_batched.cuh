// bcf_batched.cuh -- BATCHED engine: up to NB trees of the Gibbs sweep per observation pass and per grid barrier.
//
// Backfitting is sequential in the trees: the sufficient statistics of tree j are sums of the residual AFTER trees
// 0..j-1 have been redrawn.  But the residual is a fixed-point INTEGER and every update is an integer subtraction of
// a per-leaf constant, so the dependency is linear and can be resolved after the fact, exactly:
//
//     S_j[k]  =  S0_j[k]  -  sum_{m<j} sum_{k'} D_m[k'] * N_{m,j}[k'][k]
//
// S0_j[k]      : sum of the residual as it was BEFORE the batch over the observations with key k in tree j,
// D_m[k']      : what an observation with key k' in tree m loses when tree m is redrawn (known once m is decided),
// N_{m,j}[k'][k]: how many observations have key k' in tree m AND key k in tree j (a count: data-independent of the
//                decisions).  "Key" = leaf slot, with the would-be right child of a birth proposal as its own key.
// One pass over the CTA's observations therefore accumulates, for a whole batch, the per-tree key sums and the
// pairwise cross counts; after ONE grid barrier every CTA replays the batch's Metropolis-Hastings decisions in tree
// order from the same reduced integers (so nothing is broadcast), which yields per tree the tables the next pass
// applies to the residual, to the forest fit and to the leaf-slot cache.  The chain is bit-identical to the
// tree-at-a-time engines: same integers, same Philox counters, same fp64 expressions in the same order.
//
// State of the CTA's observations (fixed-point residual R, fit G of the forest being swept, the batch's keys) lives in
// shared memory for the whole launch; per tree the HBM traffic is 1 B/obs of leaf slots + the proposed variable's bin.
// Trees whose key space exceeds KB (or whose cross tables would not fit) are processed alone ("big" batches) with the
// shared-atomics statistics table of the tree-at-a-time engines.
#pragma once
#include "bcf_persistent.cuh"

namespace bcf {

constexpr int NB = 8;              // trees per batch
constexpr int KB = 8;              // key-space bound of a batched tree
constexpr int MAXC = 16;           // "columns" of a batch: its (tree, key >= 1) pairs, sum of (K_q - 1)
constexpr int XCAP = 128;          // cross-count cells of a batch: sum over pairs m<q of (K_m-1)(K_q-1) <= MAXC^2/2
constexpr int APCAP = 136;         // apply-table entries: NB*KB, or the KEYS = 129 of one big tree
static_assert(APCAP >= KEYS && APCAP >= NB * KB, "apply table too small");
using TreeSmall = TreeT<2 * KB, KB>;
using SmallStage = TreeStageT<TreeSmall, KB>;
constexpr int BAT_BYTES_PER_TILE = TILE * (8 + 8) + NB * TILE;   // R, G, keys of the batch (one byte per tree)
constexpr int FLUSH_TILES = 48;    // tiles of a CTA between two flushes of its 32-bit statistics tables (<= 63)

struct BatchStage {
    int32_t first, nb, big, forest;       // trees [first, first + nb); big: ONE tree, staged in BatchSmem::big
    int32_t ncell, ncol;                  // cross-count cells / columns in use
    int32_t K[NB];                        // key-space size of each tree (leaves + 1 for a birth proposal)
    int16_t xoff[NB][NB];                 // xoff[q][m], m < q: first cell of the pair's (K_m-1) x (K_q-1) table
    uint8_t colbase[NB];                  // first column of each tree (column = colbase[q] + key - 1)
    uint8_t cell_c1[XCAP], cell_c2[XCAP]; // the two columns of every cross-count cell
    SmallStage st[NB];
};

// what the next pass applies to the observations of a decided batch, per tree and per (pre-decision) key
struct __align__(16) ApRow {
    long long dr;                         // fixed-point residual loss
    double dg;                            // change of the leaf value (goes into the forest's fit)
};
struct __align__(16) ApplyTab {
    int32_t n, first, forest, pad_;
    int32_t changed[NB], off[NB];
    ApRow T[2][APCAP];                    // control / treated rows: one 16-byte load per observation and tree
    int4 DL[NB * KB][2];                  // the three 21-bit limbs of T[g][row].dr (batched trees): what resolve multiplies
    long long TL[NB][2][4];               // per tree and group: sum over its keys of D * count = what the total of R loses (limbs)
    uint8_t NS[APCAP];                    // leaf slot after the move
};

struct AccTab {                           // CTA-level statistics of the pass (native 32-bit shared atomics)
    unsigned int btab[NB][KB][2][4];      // (count, three 21-bit limbs) per (tree, key >= 1, group)
    unsigned int totr[2][4];              // sum of R over all observations (key 0 of every tree = total - others)
    unsigned int xtab[XCAP][2];           // cross counts per (cell, group)
};
struct ResolveTab {                       // the batch's reduced statistics, staged from L2 with one round trip
    long long tot[NB][KB][2][4];
    long long x[XCAP][2];
};

struct BatchSmem {
    // ---- scratch of propose_stage() / decide_core(): same names as StepShared ----
    double cnt[KEYS + 1], sphi[KEYS + 1], sphir[KEYS + 1];
    double pinv[KEYS + 1], plog[KEYS + 1], pmean[KEYS + 1], psd[KEYS + 1];
    double log_tau;
    double mu_old[MAXL], mu_new[MAXL];
    double dlt[KEYS];
    long long dfx[2][KEYS];
    double nc1[KEYS + 1], nc0[KEYS + 1];
    uint8_t remap[MAXL];
    int32_t ngood[MAXL];
    uint32_t lkey[MAXL];
    uint32_t nkey[MAXN];
    uint8_t nogflag[MAXN];
    ApplyInfo ai;
    int32_t accepted;
    int32_t bad;
    long long tot[KEYS * 8];
    // ---- batched engine ----
    TreeStage big;                        // a big tree in flight / scratch stage of the proposals
    BatchStage bst[2];
    ApplyTab ap;
    union {                               // pass / epilogue tables and the resolve staging never live together
        AccTab acc;
        StatTab tab;
        ResolveTab rs;
        double red[2 * PTHREADS];
    } p;
    int32_t cq[NB][KB][2];                // pre-decision counts per (tree of the batch, key, group)
    long long momsum[MAIL_MOM_WORDS];     // sharded mode: the sweep's moments summed over the shards
};
// what depends only on COUNTS (known before the batch's decisions): precision terms of every key and pooled pair
struct PreKey { double k1, k0, inv, plog, psd, pad_; };
static_assert(sizeof(PreKey) * NB * (KB + 1) <= sizeof(long long) * KEYS * 8, "PreKey table aliases BatchSmem::tot");
constexpr size_t BATCH_SMEM_BYTES = (sizeof(BatchSmem) + 15) / 16 * 16;

struct BatPtrs { uint32_t R, G, keys; };

__device__ __forceinline__ __int128 i128_from(const long long (&p)[2])
{
    return (__int128)(((unsigned __int128)(unsigned long long)p[1] << 64) | (unsigned long long)p[0]);
}
__device__ __forceinline__ void i128_to(long long (&p)[2], __int128 v)
{
    p[0] = (long long)(unsigned long long)v;
    p[1] = (long long)(v >> 64);
}
__device__ __forceinline__ __int128 limbs_to_i128(long long w0, long long w1, long long w2)
{
    return (__int128)w0 + ((__int128)w1 << LIMB_BITS) + ((__int128)w2 << (2 * LIMB_BITS));
}

__device__ __forceinline__ void sts_v2(uint32_t a, uint2 v)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void unpack8(uint2 v, int (&out)[8])
{
    out[0] = v.x & 0xFF; out[1] = (v.x >> 8) & 0xFF; out[2] = (v.x >> 16) & 0xFF; out[3] = v.x >> 24;
    out[4] = v.y & 0xFF; out[5] = (v.y >> 8) & 0xFF; out[6] = (v.y >> 16) & 0xFF; out[7] = v.y >> 24;
}
// keys of one tree for the lane's 8 observations: the cached leaf slot, with the observations a birth proposal would
// send to the right child re-keyed to `new_slot` (padding observations keep 255).  Split in load / compute so that
// the loads of the next tree are in flight while this one is processed.
__device__ __forceinline__ void keys_load(const Dev& d, const CurInfo& ci, int tree, int64_t gb, uint2& sl, uint4& bn)
{
    sl = *reinterpret_cast<const uint2*>(d.slots + (int64_t)tree * d.n_pad + gb);
    if (ci.birth) {
        if (ci.is16) bn = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ci.bins) + gb);
        else {
            const uint2 b = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(ci.bins) + gb);
            bn.x = b.x; bn.y = b.y;
        }
    }
}
__device__ __forceinline__ uint2 keys_make(const CurInfo& ci, uint2 key, uint4 b)
{
    if (ci.birth) {
        uint32_t g0, g1;   // 0xFF where bin > cut
        if (ci.is16) {
            const uint32_t c = (uint32_t)ci.cut;
            g0 = ((b.x & 0xFFFF) > c ? 0xFFu : 0u) | ((b.x >> 16) > c ? 0xFF00u : 0u) | ((b.y & 0xFFFF) > c ? 0xFF0000u : 0u) |
                 ((b.y >> 16) > c ? 0xFF000000u : 0u);
            g1 = ((b.z & 0xFFFF) > c ? 0xFFu : 0u) | ((b.z >> 16) > c ? 0xFF00u : 0u) | ((b.w & 0xFFFF) > c ? 0xFF0000u : 0u) |
                 ((b.w >> 16) > c ? 0xFF000000u : 0u);
        } else {
            const uint32_t cv = (uint32_t)ci.cut * 0x01010101u;
            g0 = __vcmpgtu4(b.x, cv);
            g1 = __vcmpgtu4(b.y, cv);
        }
        const uint32_t sxv = (uint32_t)ci.sx * 0x01010101u, nv = (uint32_t)ci.new_slot * 0x01010101u;
        const uint32_t m0 = __vcmpeq4(key.x, sxv) & g0, m1 = __vcmpeq4(key.y, sxv) & g1;
        key.x = (key.x & ~m0) | (nv & m0);
        key.y = (key.y & ~m1) | (nv & m1);
    }
    return key;
}

// ---------------------------------------------------------------------------------------------------------------
// One warp tile: apply the decided batch `sm.ap` (if do_apply), then accumulate the statistics of batch `cur`.
// PAD: the tile may hold padding observations (key 255, R = G = 0): only the last tile of each treatment group.
// ---------------------------------------------------------------------------------------------------------------
// RES: the state (R, G, keys) of the CTA's tiles is in shared memory; else in HBM / L2 (any n).
template <bool PAD, bool RES>
__device__ __forceinline__ void batch_tile(BatchSmem& sm, const Dev& d, const BatPtrs& rs, int tile0, int lt, int lane,
                                           const BatchStage* cur, bool do_apply, bool fswitch)
{
    const int tile = tile0 + lt;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const int64_t gb = (int64_t)tile * TILE + lane * 8;
    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
    const uint32_t ktile = rs.keys + (uint32_t)lt * (uint32_t)(NB * TILE) + (uint32_t)lane * 8u;
    uint8_t* kglob = RES ? nullptr : d.keys + (int64_t)tile * (NB * TILE) + lane * 8;
    long long R[8];
    if (RES) res_ld8i(rs.R + toff, lane, R); else ld8_i64(d.rfx + gb, R);
    uint2 sl_n = make_uint2(0u, 0u);
    uint4 bn_n = make_uint4(0u, 0u, 0u, 0u);
    if (cur != nullptr) keys_load(d, cur->big ? sm.big.ci : cur->st[0].ci, cur->first, gb, sl_n, bn_n);   // in flight during the apply
    if (do_apply) {
        const ApplyTab& ap = sm.ap;
        double G[8];
        double* Gf = (ap.forest ? d.Gm : d.Gc) + gb;     // HBM mode: the decided batch's forest
        if (RES) res_ld8(rs.G + toff, lane, G); else ld8_f64(Gf, G);
        const int na = ap.n;
        for (int q = 0; q < na; ++q) {
            int key[8];
            unpack8(RES ? lds_v2(ktile + (uint32_t)q * TILE) : *reinterpret_cast<const uint2*>(kglob + q * TILE), key);
            const int off = ap.off[q];
            const ApRow* Tq = &ap.T[g][off];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (PAD && key[o] == PAD_SLOT) continue;
                const ApRow r = Tq[key[o]];
                R[o] -= r.dr;
                G[o] += r.dg;
            }
            if (ap.changed[q]) {
                int ns[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) ns[o] = (PAD && key[o] == PAD_SLOT) ? PAD_SLOT : (int)ap.NS[off + key[o]];
                st8_u8(d.slots + (int64_t)(ap.first + q) * d.n_pad + gb, ns);
            }
        }
        if (RES) {
            if (fswitch) {   // forest switch: park the mu forest's fit in HBM, bring the tau forest's in
                st8_f64(d.Gc + gb, G);
                ld8_f64(d.Gm + gb, G);
            }
            res_st8(rs.G + toff, lane, G);
            res_st8i(rs.R + toff, lane, R);
        } else {
            st8_f64(Gf, G);
            st8_i64(d.rfx + gb, R);
        }
    }
    if (cur == nullptr) return;
    if (cur->big) {
        const CurInfo& ci = sm.big.ci;
        const uint2 kk = keys_make(ci, sl_n, bn_n);
        if (RES) sts_v2(ktile, kk); else *reinterpret_cast<uint2*>(kglob) = kk;
        int key[8];
        unpack8(kk, key);
        accum_tile(sm.p.tab, key, R, g, ci.K, ci.birth != 0, lane, (int)(threadIdx.x >> 5));
        return;
    }
    {   // the total of R: key 0 of every tree of the batch is (total - its other keys), exact in integers
        long long tot = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) tot += R[o];
        warp_add21(sm.p.acc.totr[g], tot, lane);
    }
    // per (tree, key >= 1) "column": the sum of R over its observations, and which of the lane's 8 observations
    // belong to it (one byte per lane and column in the warp's scratch: the cross counts are read off that)
    const int warp = (int)(threadIdx.x >> 5);
    uint8_t* colb = reinterpret_cast<uint8_t*>(sm.tot) + warp * (MAXC * 32);
    const int nb = cur->nb;
    for (int q = 0; q < nb; ++q) {
        const CurInfo& ci = cur->st[q].ci;
        const int K = cur->K[q];
        const uint2 sl = sl_n;
        const uint4 bn = bn_n;
        if (q + 1 < nb) keys_load(d, cur->st[q + 1].ci, cur->first + q + 1, gb, sl_n, bn_n);   // next tree's loads in flight
        const uint2 key = keys_make(ci, sl, bn);
        if (RES) sts_v2(ktile + (uint32_t)q * TILE, key); else *reinterpret_cast<uint2*>(kglob + q * TILE) = key;
        uint8_t* cb = colb + (int)cur->colbase[q] * 32 + lane;
        for (int k = 1; k < K; ++k) {
            const uint32_t kv = (uint32_t)k * 0x01010101u;
            const uint32_t bits = bytes_to_bits(__vcmpeq4(key.x, kv), __vcmpeq4(key.y, kv));
            long long s = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) s += ((bits >> o) & 1u) ? R[o] : 0ll;
            unsigned int* e = sm.p.acc.btab[q][k][g];
            warp_add21(e, s, lane);
            if (ci.birth && k == K - 1) {   // only the would-be right child needs counting: the rest is cached
                const int c = __reduce_add_sync(0xffffffffu, __popc(bits));
                if (lane == 0) atomicAdd(&e[0], (unsigned int)c);
            }
            cb[(k - 1) * 32] = (uint8_t)bits;
        }
    }
    __syncwarp();
    // cross counts: N[c1][c2] = sum over lanes of popc(byte(c1) & byte(c2)); lane l takes cells l, l + 32, ...
    for (int cell = lane; cell < cur->ncell; cell += 32) {
        const uint4* a = reinterpret_cast<const uint4*>(colb + (int)cur->cell_c1[cell] * 32);
        const uint4* b = reinterpret_cast<const uint4*>(colb + (int)cur->cell_c2[cell] * 32);
        const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
        const int c = __popc(a0.x & b0.x) + __popc(a0.y & b0.y) + __popc(a0.z & b0.z) + __popc(a0.w & b0.w) +
                      __popc(a1.x & b1.x) + __popc(a1.y & b1.y) + __popc(a1.z & b1.z) + __popc(a1.w & b1.w);
        if (c) atomicAdd(&sm.p.acc.xtab[cell][g], (unsigned int)c);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Stage the batch that starts at tree `first`: compact copies of up to NB candidate trees with their proposals, then
// (thread 0) how many of them form the batch.  Reads only what was final before the sweep started.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_big(BatchSmem& sm, BatchStage& B, const Dev& d, const TreeRec* trees)
{
    const int t = threadIdx.x;
    load_tree(&sm.big.tr, &trees[B.first]);
    __syncthreads();
    fetch_stage(sm.big, d, B.first, sm.big.tr.nleaves);
    __syncthreads();
    if (t == 0) {
        const int L = sm.big.tr.nleaves;
        CurInfo& ci = sm.big.ci;
        ci.active = 1; ci.K = L + (sm.big.prop.move == 1 ? 1 : 0); ci.birth = sm.big.prop.move == 1;
        ci.sx = sm.big.prop.slot; ci.cut = sm.big.prop.cut; ci.bins = sm.big.prop.bins; ci.is16 = sm.big.prop.is16;
        ci.forest = B.forest; ci.new_slot = L;
        B.K[0] = ci.K;
    }
    __syncthreads();
}
// A big batch keeps its full tree in BatchSmem::big, which the batch in flight may still need: stage_big() is
// called separately, once that stage is free.
__device__ __forceinline__ void stage_batch(BatchSmem& sm, BatchStage& B, const Dev& d, const TreeRec* trees, int first)
{
    const int t = threadIdx.x;
    const int f = (d.f[1].ntree > 0 && first >= d.f[1].tree0) ? 1 : 0;
    const int fend = f ? d.T : (d.f[1].ntree > 0 ? d.f[1].tree0 : d.T);
    const int ncand = min(NB, fend - first);
    {
        const int q = t >> 6, i = t & 63;
        if (q < ncand) {
            const TreeRec* src = &trees[first + q];
            SmallStage& S = B.st[q];
            if (i < 2 * KB) {
                S.tr.parent[i] = __ldcg(&src->parent[i]); S.tr.left[i] = __ldcg(&src->left[i]);
                S.tr.right[i] = __ldcg(&src->right[i]); S.tr.var[i] = __ldcg(&src->var[i]); S.tr.cut[i] = __ldcg(&src->cut[i]);
                S.tr.node_slot[i] = __ldcg(&src->node_slot[i]); S.tr.depth[i] = __ldcg(&src->depth[i]);
                S.tr.heap[i] = __ldcg(&src->heap[i]);
            } else if (i < 3 * KB) {
                const int s = i - 2 * KB;
                S.tr.slot_node[s] = __ldcg(&src->slot_node[s]); S.tr.cnt0[s] = __ldcg(&src->cnt0[s]);
                S.tr.cnt1[s] = __ldcg(&src->cnt1[s]); S.tr.mu[s] = __ldcg(&src->mu[s]);
            } else if (i == 3 * KB) {
                S.tr.nnodes = __ldcg(&src->nnodes); S.tr.nleaves = __ldcg(&src->nleaves);
            } else if (i >= 32 && i < 32 + (int)(sizeof(Proposal) / 8)) {
                reinterpret_cast<long long*>(&S.prop)[i - 32] = __ldcg(reinterpret_cast<const long long*>(&d.proprec[first + q].P) + (i - 32));
            } else if (i >= 48 && i < 48 + KB) {
                S.z[i - 48] = __ldcg(&d.proprec[first + q].z[i - 48]);
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        B.first = first; B.forest = f; B.big = 0; B.ncell = 0; B.ncol = 0;
        int nb = 0, cells = 0, cols = 0;
        for (int q = 0; q < ncand; ++q) {
            const SmallStage& S = B.st[q];
            const int L = S.tr.nleaves, K = L + (S.prop.move == 1 ? 1 : 0);
            if (K > KB) { if (q == 0) { B.big = 1; nb = 1; } break; }
            int add = 0;
            for (int m = 0; m < q; ++m) add += (B.K[m] - 1) * (K - 1);
            if (cells + add > XCAP || cols + K - 1 > MAXC) break;
            B.colbase[q] = (uint8_t)cols;
            for (int m = 0; m < q; ++m) { B.xoff[q][m] = (int16_t)cells; cells += (B.K[m] - 1) * (K - 1); }
            cols += K - 1;
            B.K[q] = K;
            CurInfo& ci = B.st[q].ci;
            ci.active = 1; ci.K = K; ci.birth = S.prop.move == 1; ci.sx = S.prop.slot; ci.cut = S.prop.cut;
            ci.bins = S.prop.bins; ci.is16 = S.prop.is16; ci.forest = f; ci.new_slot = L;
            nb = q + 1;
        }
        B.nb = nb; B.ncell = cells; B.ncol = cols;
    }
    __syncthreads();
    // the two columns of every cross-count cell, in parallel over (q, m, kp, k)
    if (!B.big) {
        for (int i = t; i < NB * NB * (KB - 1) * (KB - 1); i += blockDim.x) {
            const int k = 1 + i % (KB - 1), kp = 1 + (i / (KB - 1)) % (KB - 1);
            const int m = (i / ((KB - 1) * (KB - 1))) % NB, q = i / ((KB - 1) * (KB - 1) * NB);
            if (q < B.nb && m < q && kp < B.K[m] && k < B.K[q]) {
                const int cell = B.xoff[q][m] + (kp - 1) * (B.K[q] - 1) + (k - 1);
                B.cell_c1[cell] = (uint8_t)(B.colbase[m] + kp - 1);
                B.cell_c2[cell] = (uint8_t)(B.colbase[q] + k - 1);
            }
        }
    }
    __syncthreads();
}

// clear the pass tables of the batch about to be accumulated
__device__ __forceinline__ void begin_batch(BatchSmem& sm, const BatchStage& B)
{
    const int t = threadIdx.x;
    if (B.big) {
        long long* tb = reinterpret_cast<long long*>(&sm.p.tab);
        for (int i = t; i < STATTAB_WORDS64; i += blockDim.x) tb[i] = 0;
    } else {
        unsigned int* tb = reinterpret_cast<unsigned int*>(&sm.p.acc);
        const int nw = (int)(offsetof(AccTab, xtab) / 4) + 2 * B.ncell;
        for (int i = t; i < nw; i += blockDim.x) tb[i] = 0u;
    }
}

// CTA tables -> global totals / cross counts (integer REDs: order-independent).  After a __syncthreads().
__device__ __forceinline__ void flush_batch(BatchSmem& sm, const BatchStage& B, long long* totals, long long* cross)
{
    const int t = threadIdx.x;
    if (B.big) {
        const int K = B.K[0];
        long long* q = totals + (size_t)B.first * KEYS * 8;
        for (int i = t; i < K * 8; i += blockDim.x) {
            const long long v = stat_word(sm.p.tab, K, 0, i >> 3, (i >> 2) & 1, i & 3);
            if (v != 0) atomicAdd((unsigned long long*)&q[i], (unsigned long long)v);
        }
        return;
    }
    const int nb = B.nb;
    for (int i = t; i < nb * KB * 8; i += blockDim.x) {
        const int q = i / (KB * 8), k = (i >> 3) & (KB - 1), g = (i >> 2) & 1, w = i & 3;
        long long v = 0;
        if (k == 0) { if (q == 0) v = acc_word(sm.p.acc.totr[g], w); }        // key 0 of the first tree carries the total
        else if (k < B.K[q]) v = acc_word(sm.p.acc.btab[q][k][g], w);
        if (v != 0) atomicAdd((unsigned long long*)&totals[(size_t)(B.first + q) * KEYS * 8 + k * 8 + g * 4 + w], (unsigned long long)v);
    }
    long long* xg = cross + (size_t)B.first * (XCAP * 2);
    for (int i = t; i < 2 * B.ncell; i += blockDim.x) {
        const unsigned int v = sm.p.acc.xtab[i >> 1][i & 1];
        if (v != 0) atomicAdd((unsigned long long*)&xg[i], (unsigned long long)v);
    }
}

__device__ __forceinline__ void store_small(TreeRec* dst, const TreeSmall& tr, int t)   // one warp, t = lane
{
    if (t < 2 * KB) {
        dst->parent[t] = tr.parent[t]; dst->left[t] = tr.left[t]; dst->right[t] = tr.right[t]; dst->var[t] = tr.var[t];
        dst->cut[t] = tr.cut[t]; dst->node_slot[t] = tr.node_slot[t]; dst->depth[t] = tr.depth[t]; dst->heap[t] = tr.heap[t];
    } else if (t < 3 * KB) {
        const int s = t - 2 * KB;
        dst->slot_node[s] = tr.slot_node[s]; dst->cnt0[s] = tr.cnt0[s]; dst->cnt1[s] = tr.cnt1[s]; dst->mu[s] = tr.mu[s];
    } else if (t == 3 * KB) {
        dst->nnodes = tr.nnodes; dst->nleaves = tr.nleaves;
    }
}

// apply-table rows of a decided tree (keys of the PRE-decision key space)
template <class Stage>
__device__ __forceinline__ void fill_apply(BatchSmem& sm, const Stage& S, int q, int off, int L, int K)
{
    const int t = threadIdx.x;
    ApplyTab& ap = sm.ap;
    if (t < K) {
        if (t < L) {
            ap.T[0][off + t].dr = sm.dfx[0][t]; ap.T[1][off + t].dr = sm.dfx[1][t];
            ap.T[0][off + t].dg = ap.T[1][off + t].dg = sm.dlt[t]; ap.NS[off + t] = sm.remap[t];
        } else {   // the part of leaf sx a birth proposal would have sent right
            const int sx = S.prop.slot;
            const bool acc = sm.ai.birth != 0;
            ap.T[0][off + t].dr = acc ? sm.ai.dfx_right[0] : sm.dfx[0][sx];
            ap.T[1][off + t].dr = acc ? sm.ai.dfx_right[1] : sm.dfx[1][sx];
            ap.T[0][off + t].dg = ap.T[1][off + t].dg = acc ? sm.ai.dlt_right : sm.dlt[sx];
            ap.NS[off + t] = acc ? (uint8_t)L : sm.remap[sx];
        }
    }
    if (t == 0) { ap.changed[q] = sm.accepted; ap.off[q] = off; }
}

// ---------------------------------------------------------------------------------------------------------------
// The decision chain of a batch of small trees, run by NB warps (warp w owns tree w; l = lane).  Shared by the batched
// and the pipelined engines.  rs: the batch's reduced statistics (rs.tot[q][k][g][4], rs.x[cell][g]); ap: the apply
// tables being built (T rows, DL limb rows, NS, changed, off); cq: pre-decision counts per (tree, key, group).
// ---------------------------------------------------------------------------------------------------------------
struct ChainConsts { double s1, s0, rs1, rs0, phi1, phi0, log_tau; };
struct ChainPrev {                       // the previous batch, when its update is not in the statistics (n = 0: none)
    int n;
    const int32_t* K;
    const int16_t* xoff;                 // xoff[q * NT + m]: first cross cell of (previous tree m, this batch's tree q)
    const int4 (*DL)[2];
    const int32_t (*cq)[KB][2];
    const long long (*TL)[2][4];         // TL[m][g][limb]: total loss of previous tree m
};
// NT: trees per batch = warps walking the chain (named barriers 1..NT, NT * 32 threads each)
template <int NT, class RS, class AP>
__device__ __forceinline__ void decision_chain(int w, int l, int nb, int first, SmallStage* st, const int32_t* Kq,
                                               const int16_t* xoff, RS& rs, AP& ap, int32_t (*cq)[KB][2],
                                               const PreKey* pre, const ChainConsts& cc, const ChainPrev& pv, const Dev& d,
                                               TreeRec* newtrees, int32_t* bad, long long* trc = nullptr)
{
    {
        const unsigned FULL = 0xffffffffu;
        const int q = w;
#ifndef BCF_PROF
#define CTRACE(e) do { } while (0)
#else
#define CTRACE(e) do { if (trc != nullptr && l == 0) trc[w * 8 + (e)] = clock64(); } while (0)
#endif
        CTRACE(0);
        const bool active = q < nb;
        const int qq = active ? q : 0;
        SmallStage& S = st[qq];
        auto& tr = S.tr;
        const Proposal& P = S.prop;
        const int tree = first + qq;
        const bool writer = active && ((int)blockIdx.x == tree % (int)gridDim.x);
        const int L = tr.nleaves, K = Kq[qq];
        const int move = P.move;
        const bool birth = move == 1;
        const int sx = P.slot;
        const int pa = P.slot, pb = birth ? L : P.slot_b;
        const int off = qq * KB;
        // lane l = key l, both treatment groups.  The key sums stay in the global limb representation
        // (value = w0 + w1 2^21 + w2 2^42, limbs not normalised): a correction is then three 32 x 32 -> 64 bit
        // multiply-adds per term instead of 128-bit arithmetic.  Key 0 is (total - other keys), formed once here.
        long long wa0 = 0, wa1 = 0, wa2 = 0, wb0 = 0, wb1 = 0, wb2 = 0;   // group 0 (control), group 1 (treated)
        int c0 = 0, c1 = 0;
        // everything the decision needs that is known before the chain starts
        const PreKey pk = pre[qq * (KB + 1) + (l < K ? l : KB)];
        const PreKey pkp = pre[qq * (KB + 1) + KB];
        const double mo = tr.mu[(l < L) ? l : sx];            // old value of the leaf the key belongs to
        const double zl = S.z[l < L ? l : sx];
        const double ze0 = P.z_extra[0], ze1 = P.z_extra[1];
        const double log_prior = P.log_prior, log_u = P.log_u;
        double gain_c = 0.0, inv_a = 0.0, inv_b = 0.0;          // count-only part of the likelihood gain
        bool big_enough = true;
        if (active && move != 0) {
            const PreKey pka = pre[qq * (KB + 1) + pa], pkb = pre[qq * (KB + 1) + pb];
            gain_c = -cc.log_tau - 0.5 * (pka.plog + pkb.plog - pkp.plog);
            inv_a = pka.inv; inv_b = pkb.inv;
            big_enough = pka.k1 + pka.k0 >= d.min_leaf && pkb.k1 + pkb.k0 >= d.min_leaf;
        }
        if (active && l < K) {
            // pre-decision counts of the key: written by the parallel count phase, BEFORE the barrier that precedes
            // the chain (the other warps read this tree's counts while they wait for its decision)
            c0 = cq[q][l][0]; c1 = cq[q][l][1];
            if (l >= 1) {
                wa0 = rs.tot[q][l][0][1]; wa1 = rs.tot[q][l][0][2]; wa2 = rs.tot[q][l][0][3];
                wb0 = rs.tot[q][l][1][1]; wb1 = rs.tot[q][l][1][2]; wb2 = rs.tot[q][l][1][3];
            } else {
                wa0 = rs.tot[0][0][0][1]; wa1 = rs.tot[0][0][0][2]; wa2 = rs.tot[0][0][0][3];
                wb0 = rs.tot[0][0][1][1]; wb1 = rs.tot[0][0][1][2]; wb2 = rs.tot[0][0][1][3];
                for (int kx = 1; kx < K; ++kx) {
                    wa0 -= rs.tot[q][kx][0][1]; wa1 -= rs.tot[q][kx][0][2]; wa2 -= rs.tot[q][kx][0][3];
                    wb0 -= rs.tot[q][kx][1][1]; wb1 -= rs.tot[q][kx][1][2]; wb2 -= rs.tot[q][kx][1][3];
                }
            }
        }
        // fold the update of one decided tree (its D rows as limbs, `DLm[kp][g]`) into this tree's sums:
        //     w -= sum over that tree's keys kp of D[kp] * N[kp][l],   N[kp][l] = #{key kp there and key l here}.
        // The counts do not depend on the decision: they are formed BEFORE waiting for it (N[0][l], and the whole of
        // lane 0's row, by subtraction from the marginal counts).  barid > 0: wait for the owner's rows first.
        auto fold = [&](const int4 (*DLm)[2], int Km, int xb, const int32_t (*cqm)[2], int barid) {
            const int stride = K - 1;
            int na0 = 0, na1 = 0, nb0 = 0, nb1 = 0, nc0 = 0, nc1 = 0;   // N[kp][l] for kp = 1, 2, 3, by group
            int nz0 = c0, nz1 = c1;                                       // N[0][l]
            if (l < K) {
                for (int kp = 1; kp < Km; ++kp) {
                    int x0, x1;
                    if (l >= 1) { x0 = (int)rs.x[xb + (kp - 1) * stride + (l - 1)][0]; x1 = (int)rs.x[xb + (kp - 1) * stride + (l - 1)][1]; }
                    else {
                        x0 = cqm[kp][0]; x1 = cqm[kp][1];
                        for (int kx = 1; kx < K; ++kx) { x0 -= (int)rs.x[xb + (kp - 1) * stride + (kx - 1)][0]; x1 -= (int)rs.x[xb + (kp - 1) * stride + (kx - 1)][1]; }
                    }
                    nz0 -= x0; nz1 -= x1;
                    if (kp == 1) { na0 = x0; na1 = x1; } else if (kp == 2) { nb0 = x0; nb1 = x1; } else if (kp == 3) { nc0 = x0; nc1 = x1; }
                }
            }
            if (barid > 0) asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(NT * 32) : "memory");
            if (l < K) {
                const int4 z0 = DLm[0][0], z1 = DLm[0][1];
                wa0 -= (long long)z0.x * nz0; wa1 -= (long long)z0.y * nz0; wa2 -= (long long)z0.z * nz0;
                wb0 -= (long long)z1.x * nz1; wb1 -= (long long)z1.y * nz1; wb2 -= (long long)z1.z * nz1;
                if (Km > 1) {
                    const int4 a0 = DLm[1][0], a1 = DLm[1][1];
                    wa0 -= (long long)a0.x * na0; wa1 -= (long long)a0.y * na0; wa2 -= (long long)a0.z * na0;
                    wb0 -= (long long)a1.x * na1; wb1 -= (long long)a1.y * na1; wb2 -= (long long)a1.z * na1;
                }
                if (Km > 2) {
                    const int4 a0 = DLm[2][0], a1 = DLm[2][1];
                    wa0 -= (long long)a0.x * nb0; wa1 -= (long long)a0.y * nb0; wa2 -= (long long)a0.z * nb0;
                    wb0 -= (long long)a1.x * nb1; wb1 -= (long long)a1.y * nb1; wb2 -= (long long)a1.z * nb1;
                }
                if (Km > 3) {
                    const int4 a0 = DLm[3][0], a1 = DLm[3][1];
                    wa0 -= (long long)a0.x * nc0; wa1 -= (long long)a0.y * nc0; wa2 -= (long long)a0.z * nc0;
                    wb0 -= (long long)a1.x * nc1; wb1 -= (long long)a1.y * nc1; wb2 -= (long long)a1.z * nc1;
                }
                for (int kp = 4; kp < Km; ++kp) {   // rare: wide trees, counts re-read
                    int x0, x1;
                    if (l >= 1) { x0 = (int)rs.x[xb + (kp - 1) * stride + (l - 1)][0]; x1 = (int)rs.x[xb + (kp - 1) * stride + (l - 1)][1]; }
                    else {
                        x0 = cqm[kp][0]; x1 = cqm[kp][1];
                        for (int kx = 1; kx < K; ++kx) { x0 -= (int)rs.x[xb + (kp - 1) * stride + (kx - 1)][0]; x1 -= (int)rs.x[xb + (kp - 1) * stride + (kx - 1)][1]; }
                    }
                    const int4 a0 = DLm[kp][0], a1 = DLm[kp][1];
                    wa0 -= (long long)a0.x * x0; wa1 -= (long long)a0.y * x0; wa2 -= (long long)a0.z * x0;
                    wb0 -= (long long)a1.x * x1; wb1 -= (long long)a1.y * x1; wb2 -= (long long)a1.z * x1;
                }
            }
        };
        // the previous batch (pipelined engine): decided long ago, its update is simply not in the statistics yet
        // All of that batch's D rows are known, so the whole warp works on it: lane = key (l & 7) x quarter of the
        // previous trees (l >> 3); key 0 takes each tree's TOTAL loss and hands back what the other keys took.
        if (pv.n > 0) {
            const int jg = l >> 3, ll = l & 7;
            const int cc0 = __shfl_sync(FULL, c0, ll), cc1 = __shfl_sync(FULL, c1, ll);
            long long a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
            if (active && ll < K) {
                const int stride = K - 1;
                for (int m = jg; m < pv.n; m += 4) {
                    const int Km = pv.K[m];
                    const int4 (*DLm)[2] = &pv.DL[m * KB];
                    if (ll >= 1) {
                        const int xb = pv.xoff[q * NT + m] + (ll - 1);
                        int nz0 = cc0, nz1 = cc1;
                        for (int kp = 1; kp < Km; ++kp) {
                            const int x0 = (int)rs.x[xb + (kp - 1) * stride][0], x1 = (int)rs.x[xb + (kp - 1) * stride][1];
                            const int4 r0 = DLm[kp][0], r1 = DLm[kp][1];
                            nz0 -= x0; nz1 -= x1;
                            a0 += (long long)r0.x * x0; a1 += (long long)r0.y * x0; a2 += (long long)r0.z * x0;
                            b0 += (long long)r1.x * x1; b1 += (long long)r1.y * x1; b2 += (long long)r1.z * x1;
                        }
                        const int4 z0 = DLm[0][0], z1 = DLm[0][1];
                        a0 += (long long)z0.x * nz0; a1 += (long long)z0.y * nz0; a2 += (long long)z0.z * nz0;
                        b0 += (long long)z1.x * nz1; b1 += (long long)z1.y * nz1; b2 += (long long)z1.z * nz1;
                    } else {
                        a0 += pv.TL[m][0][0]; a1 += pv.TL[m][0][1]; a2 += pv.TL[m][0][2];
                        b0 += pv.TL[m][1][0]; b1 += pv.TL[m][1][1]; b2 += pv.TL[m][1][2];
                    }
                }
            }
#pragma unroll
            for (int sft = 8; sft <= 16; sft <<= 1) {   // the four quarters
                a0 += __shfl_xor_sync(FULL, a0, sft); a1 += __shfl_xor_sync(FULL, a1, sft); a2 += __shfl_xor_sync(FULL, a2, sft);
                b0 += __shfl_xor_sync(FULL, b0, sft); b1 += __shfl_xor_sync(FULL, b1, sft); b2 += __shfl_xor_sync(FULL, b2, sft);
            }
            long long s0 = ll >= 1 ? a0 : 0, s1_ = ll >= 1 ? a1 : 0, s2_ = ll >= 1 ? a2 : 0;
            long long u0 = ll >= 1 ? b0 : 0, u1 = ll >= 1 ? b1 : 0, u2 = ll >= 1 ? b2 : 0;
#pragma unroll
            for (int sft = 1; sft <= 4; sft <<= 1) {    // what the keys >= 1 took, over the 8 key lanes
                s0 += __shfl_xor_sync(FULL, s0, sft); s1_ += __shfl_xor_sync(FULL, s1_, sft); s2_ += __shfl_xor_sync(FULL, s2_, sft);
                u0 += __shfl_xor_sync(FULL, u0, sft); u1 += __shfl_xor_sync(FULL, u1, sft); u2 += __shfl_xor_sync(FULL, u2, sft);
            }
            if (active && l < K) {
                if (l >= 1) { wa0 -= a0; wa1 -= a1; wa2 -= a2; wb0 -= b0; wb1 -= b1; wb2 -= b2; }
                else { wa0 -= a0 - s0; wa1 -= a1 - s1_; wa2 -= a2 - s2_; wb0 -= b0 - u0; wb1 -= b1 - u1; wb2 -= b2 - u2; }
            }
        }
        CTRACE(1);
        [[maybe_unused]] const bool cprof = d.prof != nullptr && blockIdx.x == 0 && l == 0;
        long long ct = 0;
#ifndef BCF_PROF
#define CPROF(k) do { } while (0)
#else
#define CPROF(k) do { if (cprof) { const long long now_ = clock64(); atomicAdd((unsigned long long*)&d.prof[8 + (k)], (unsigned long long)(now_ - ct)); ct = now_; } } while (0)
#endif
        for (int m = 0; m < nb; ++m) {
            if (active && q == m) {
                CTRACE(2);
                if (cprof) ct = BCF_PROF_CLOCK();
                // exact integer -> double, as limbs_to_double does
                const __int128 v0 = limbs_to_i128(wa0, wa1, wa2), v1 = limbs_to_i128(wb0, wb1, wb2);
                const double S0 = ((double)(long long)(v0 >> 40) * 1099511627776.0 + (double)(long long)(v0 & (((__int128)1 << 40) - 1))) * (1.0 / 17592186044416.0);
                const double S1 = ((double)(long long)(v1 >> 40) * 1099511627776.0 + (double)(long long)(v1 & (((__int128)1 << 40) - 1))) * (1.0 / 17592186044416.0);
                CPROF(2);
                double Bk = 0.0;
                if (l < K) Bk = cc.phi1 * (S1 * cc.rs1 + pk.k1 * mo) + cc.phi0 * (S0 * cc.rs0 + pk.k0 * mo);
                const double Ba = __shfl_sync(FULL, Bk, pa), Bb = __shfl_sync(FULL, Bk, pb);
                const double Bp = Ba + Bb;
                CPROF(3);
                // ---- the MH decision (every lane computes it from the same values) ----
                int accepted = 0;
                double logr = nan("");
                if (move != 0) {
                    const double gain = gain_c + 0.5 * (Ba * Ba * inv_a + Bb * Bb * inv_b - Bp * Bp * pkp.inv);
                    if (birth) {
                        if (big_enough) { logr = log_prior + gain; accepted = (log_u < logr) ? 1 : 0; }
                    } else {
                        logr = log_prior - gain;
                        accepted = (log_u < logr) ? 1 : 0;
                    }
                }
                // ---- the new leaf of every pre-decision key: posterior, normal, slot ----
                // pooled pair: a rejected birth keeps leaf sx whole, an accepted death merges its two leaves
                const bool pooled = (birth && !accepted && (l == sx || l == L)) || (move == 2 && accepted && (l == pa || l == pb));
                const double pm = pooled ? Bp * pkp.inv : Bk * pk.inv;
                const double ps = pooled ? pkp.psd : pk.psd;
                double z = zl;
                int ns = l;
                if (birth) {
                    if (accepted) z = (l == sx) ? ze0 : (l == L ? ze1 : zl);
                    else if (l == L) ns = sx;
                } else if (move == 2 && accepted) {
                    if (l == pa || l == pb) z = ze0;
                    const int last = L - 1;
                    if (l == pb) ns = pa;
                    if (pb != last && ns == last) ns = pb;      // the last slot moves into the hole
                }
                int fxbad = 0;
                double mnew = 0.0, dl = 0.0;
                if (l < K) {
                    mnew = pm + z * ps;
                    if (!(mnew == mnew)) *bad = ERR_NAN;
                    dl = mnew - mo;
                    ApRow r1, r0;
                    r1.dr = to_fixed(cc.s1 * dl, fxbad); r1.dg = dl; r0.dr = to_fixed(cc.s0 * dl, fxbad); r0.dg = dl;
                    ap.T[1][off + l] = r1; ap.T[0][off + l] = r0;
                    ap.DL[off + l][1] = make_int4((int)(r1.dr & 0x1FFFFF), (int)((r1.dr >> LIMB_BITS) & 0x1FFFFF), (int)(r1.dr >> (2 * LIMB_BITS)), 0);
                    ap.DL[off + l][0] = make_int4((int)(r0.dr & 0x1FFFFF), (int)((r0.dr >> LIMB_BITS) & 0x1FFFFF), (int)(r0.dr >> (2 * LIMB_BITS)), 0);
                }
                if (fxbad) *bad = ERR_FX_RANGE;
                CPROF(4);
                __threadfence_block();
                asm volatile("bar.arrive %0, %1;" ::"r"(1 + m), "r"(NT * 32) : "memory");
                CPROF(5); CTRACE(3);
                // ---- bookkeeping nobody waits for ----
                const double n1 = pooled ? pkp.k1 : pk.k1, n0 = pooled ? pkp.k0 : pk.k0;
                if (l < K) ap.NS[off + l] = (uint8_t)ns;
                if (l == 0) { ap.changed[q] = accepted; ap.off[q] = off; }
                __syncwarp();
                if (l == 0) {   // topology, exactly decide_core's
                    if (birth) {
                        if (accepted) {
                            const int kr = pb, nx = P.node, na = tr.nnodes, nbn = tr.nnodes + 1;
                            tr.left[nx] = (int16_t)na; tr.right[nx] = (int16_t)nbn;
                            tr.var[nx] = (uint16_t)P.var; tr.cut[nx] = (uint16_t)P.cut;
                            for (int ch2 = 0; ch2 < 2; ++ch2) {
                                const int ch = ch2 ? nbn : na;
                                tr.parent[ch] = (int16_t)nx; tr.left[ch] = tr.right[ch] = -1;
                                tr.var[ch] = tr.cut[ch] = 0xFFFF;
                                tr.depth[ch] = tr.depth[nx] + 1;
                                tr.heap[ch] = 2u * tr.heap[nx] + (uint32_t)ch2;
                            }
                            tr.node_slot[na] = (uint8_t)sx; tr.slot_node[sx] = (int16_t)na;
                            tr.node_slot[nbn] = (uint8_t)kr; tr.slot_node[kr] = (int16_t)nbn;
                            tr.nnodes += 2; tr.nleaves = L + 1;
                        }
                    } else if (move == 2 && accepted) {
                        const int sa = pa, sb = pb;
                        const int nx = P.node, na = tr.left[nx], nbn = tr.right[nx];
                        tr.left[nx] = tr.right[nx] = -1; tr.var[nx] = tr.cut[nx] = 0xFFFF;
                        tr.node_slot[nx] = (uint8_t)sa; tr.slot_node[sa] = (int16_t)nx;
                        const int last = L - 1;
                        if (sb != last) {   // keep slots compact: the last slot moves into the hole
                            const int nd = tr.slot_node[last];
                            tr.slot_node[sb] = (int16_t)nd; tr.node_slot[nd] = (uint8_t)sb;
                        }
                        tr.nleaves = L - 1;
                        tree_remove_node(tr, max(na, nbn));
                        tree_remove_node(tr, min(na, nbn));
                    }
                    if (writer) {
                        int32_t* trc = d.trace + 5 * tree;
                        trc[0] = move; trc[1] = accepted; trc[2] = (int32_t)P.heap; trc[3] = P.var; trc[4] = P.cut;
                        d.trace_logr[tree] = logr;
                        if (move == 1) { atomicAdd((unsigned long long*)&d.counters[0], 1ull); if (accepted) atomicAdd((unsigned long long*)&d.counters[1], 1ull); }
                        if (move == 2) { atomicAdd((unsigned long long*)&d.counters[2], 1ull); if (accepted) atomicAdd((unsigned long long*)&d.counters[3], 1ull); }
                    }
                }
                __syncwarp();
                if (l < K) { tr.mu[ns] = mnew; tr.cnt1[ns] = (int32_t)n1; tr.cnt0[ns] = (int32_t)n0; }   // keys of one new leaf agree
                __syncwarp();
                if (writer) store_small(&newtrees[tree], tr, l);
                {   // what the total of R loses to this tree, for the next batch's key 0 (pipelined engine)
                    long long t0 = 0, t1 = 0, t2 = 0, v0 = 0, v1 = 0, v2 = 0;
                    if (l < K) {
                        const int4 r0 = ap.DL[off + l][0], r1 = ap.DL[off + l][1];
                        t0 = (long long)r0.x * c0; t1 = (long long)r0.y * c0; t2 = (long long)r0.z * c0;
                        v0 = (long long)r1.x * c1; v1 = (long long)r1.y * c1; v2 = (long long)r1.z * c1;
                    }
#pragma unroll
                    for (int sft = 1; sft <= 4; sft <<= 1) {
                        t0 += __shfl_xor_sync(FULL, t0, sft); t1 += __shfl_xor_sync(FULL, t1, sft); t2 += __shfl_xor_sync(FULL, t2, sft);
                        v0 += __shfl_xor_sync(FULL, v0, sft); v1 += __shfl_xor_sync(FULL, v1, sft); v2 += __shfl_xor_sync(FULL, v2, sft);
                    }
                    if (l == 0) {
                        ap.TL[q][0][0] = t0; ap.TL[q][0][1] = t1; ap.TL[q][0][2] = t2;
                        ap.TL[q][1][0] = v0; ap.TL[q][1][1] = v1; ap.TL[q][1][2] = v2;
                    }
                }
                CTRACE(4);
            } else if (active && q > m) {
                fold(&ap.DL[m * KB], Kq[m], xoff[q * NT + m], cq[m], 1 + m);
            } else {
                asm volatile("bar.arrive %0, %1;" ::"r"(1 + m), "r"(NT * 32) : "memory");
            }
        }
        CTRACE(5);
    }
#undef CPROF
#undef CTRACE
}

// Observation shards: CTA 0 stores this rank's reduced words of the batch into its mailbox on every peer and releases
// sequence number `seq` there (see PeerDev).  Layout: totals words packed (tree q, word w) -> q * KB * 8 + w (a big
// tree: its K * 8 words), then the cross counts from MAIL_TOT_WORDS on.
__device__ __forceinline__ void peer_push_batch(const BatchStage& B, const Dev& d, const long long* totals, const long long* cross,
                                       int slot, unsigned long long seq)
{
    const int t = threadIdx.x;
    const int nt = B.big ? B.K[0] * 8 : B.nb * KB * 8, nx = B.big ? 0 : 2 * B.ncell;
    for (int i = t; i < nt + nx; i += blockDim.x) {
        long long v;
        int dst;
        if (i < nt) {
            v = __ldcg(&totals[B.big ? (size_t)B.first * KEYS * 8 + i : (size_t)(B.first + i / (KB * 8)) * KEYS * 8 + (i & (KB * 8 - 1))]);
            dst = i;
        } else {
            v = __ldcg(&cross[(size_t)B.first * (XCAP * 2) + (i - nt)]);
            dst = MAIL_TOT_WORDS + (i - nt);
        }
        for (int p = 0; p < d.pr.nranks; ++p)
            if (p != d.pr.rank) d.pr.peer[p][d.pr.rank].batch[slot][dst] = v;
    }
    peer_release(d, slot, seq);
}
// sum of the peers' words (fixed order; integers, so any order gives the same bits)
__device__ __forceinline__ long long peer_sum(const Dev& d, int slot, int idx)
{
    long long v = 0;
    for (int p = 0; p < d.pr.nranks; ++p)
        if (p != d.pr.rank) v += __ldcg(&d.pr.mine[p].batch[slot][idx]);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// After the grid barrier: replay the batch's decisions in tree order from the reduced integers.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void resolve_batch(BatchSmem& sm, BatchStage& B, const Dev& d, const Scalars& sc, TreeRec* newtrees,
                                     const long long* totals, const long long* cross, int slot)
{
    const bool sharded = d.pr.nranks > 1;
    const int t = threadIdx.x;
    const ForestDev& F = d.f[B.forest];
    if (B.big) {
        const int tree = B.first;
        const bool writer = ((int)blockIdx.x == tree % (int)gridDim.x);
        const int L = sm.big.tr.nleaves, K = B.K[0];
        const long long* q = totals + (size_t)tree * KEYS * 8;
        for (int i = t; i < K * 8; i += blockDim.x) sm.tot[i] = __ldcg(q + i) + (sharded ? peer_sum(d, slot, i) : 0ll);
        decide_core(sm, sm.big, d, F, tree, sc, writer);
        fill_apply(sm, sm.big, 0, 0, L, K);
        if (t == 0) { sm.ap.n = 1; sm.ap.first = tree; sm.ap.forest = B.forest; }
        if (writer) store_tree(&newtrees[tree], &sm.big.tr);
        __syncthreads();
        return;
    }
    [[maybe_unused]] const bool rprof = d.prof != nullptr && blockIdx.x == 0 && t == 0;
    long long rp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    [[maybe_unused]] long long rtp = BCF_PROF_CLOCK();
#ifndef BCF_PROF
#define RPROF(k) do { } while (0)
#else
#define RPROF(k) do { if (rprof) { const long long now_ = clock64(); rp[k] += now_ - rtp; rtp = now_; } } while (0)
#endif
    const int nb = B.nb;
    ResolveTab& rs = sm.p.rs;
    for (int i = t; i < nb * KB * 8; i += blockDim.x) {
        const int q = i / (KB * 8), k = (i >> 3) & (KB - 1);
        if (k < B.K[q]) (&rs.tot[0][0][0][0])[i] = __ldcg(&totals[(size_t)(B.first + q) * KEYS * 8 + (i & (KB * 8 - 1))]) + (sharded ? peer_sum(d, slot, i) : 0ll);
    }
    {
        const long long* xg = cross + (size_t)B.first * (XCAP * 2);
        for (int i = t; i < 2 * B.ncell; i += blockDim.x) (&rs.x[0][0])[i] = __ldcg(xg + i) + (sharded ? peer_sum(d, slot, MAIL_TOT_WORDS + i) : 0ll);
    }
    __syncthreads();
    RPROF(0);
    const bool is_con = F.is_con != 0;
    const double s1 = is_con ? sc.mscale : sc.bscale1, s0 = is_con ? sc.mscale : sc.bscale0;
    const double tau = is_con ? sc.tau_con : sc.tau_mod;
    const double s2 = sc.sigma * sc.sigma;
    const double phi1 = s1 * s1 / s2, phi0 = s0 * s0 / s2;
    const double a = 1.0 / (tau * tau);
    const double rs1 = 1.0 / s1, rs0 = 1.0 / s0;
    const double log_tau = log(tau);
    // Everything that depends on counts only -- precision, its log and reciprocal, the posterior sd -- for every key
    // and pooled pair of every tree of the batch, in parallel (the divisions, logs and square roots of the batch).
    PreKey* pre = reinterpret_cast<PreKey*>(sm.tot);
    if (t < nb * (KB + 1)) {
        const int q = t / (KB + 1), e = t % (KB + 1);
        const SmallStage& S = B.st[q];
        const int L = S.tr.nleaves, K = B.K[q], mv = S.prop.move, sx = S.prop.slot;
        if (e < K || (e == KB && mv != 0)) {
            const double cL1 = mv == 1 ? (double)rs.tot[q][L][1][0] : 0.0, cL0 = mv == 1 ? (double)rs.tot[q][L][0][0] : 0.0;
            const int pa = sx, pb = (mv == 1) ? L : S.prop.slot_b;
            double k1 = 0.0, k0 = 0.0, A = 0.0;
            const int nk = e < K ? 1 : 2;
            for (int j = 0; j < nk; ++j) {
                const int key = e < K ? e : (j == 0 ? pa : pb);
                double c1, c0;
                if (key < L) {
                    c1 = (double)S.tr.cnt1[key]; c0 = (double)S.tr.cnt0[key];
                    if (mv == 1 && key == sx) { c1 -= cL1; c0 -= cL0; }
                } else { c1 = cL1; c0 = cL0; }
                k1 += c1; k0 += c0;
                A += c1 * phi1 + c0 * phi0;
            }
            const double prec = a + A, inv = 1.0 / prec;
            PreKey& pk = pre[q * (KB + 1) + e];
            pk.k1 = k1; pk.k0 = k0; pk.inv = inv; pk.plog = log(prec); pk.psd = sqrt(inv);
            if (e < K) { sm.cq[q][e][1] = (int32_t)k1; sm.cq[q][e][0] = (int32_t)k0; }
        }
    }
    __syncthreads();
    RPROF(1);
    // The decisions form a serial chain (tree q needs the residual losses D of the trees before it), but a short
    // one: WARP q owns tree q of the batch.  It keeps the tree's key sums in registers (lane 2k+g = key k, group g),
    // folds in the correction of every earlier tree as soon as that tree's D rows are published (one named barrier
    // per tree: the owner arrives, the later warps wait), takes its decision, publishes its own D rows, and only
    // then does the bookkeeping nobody waits for (tree record, slot map, trace, HBM write-back).
    if ((t >> 5) < NB) {
        ChainConsts cc{s1, s0, rs1, rs0, phi1, phi0, log_tau};
        ChainPrev pv{0, nullptr, nullptr, nullptr, nullptr, nullptr};
        decision_chain<NB>(t >> 5, t & 31, nb, B.first, B.st, B.K, &B.xoff[0][0], rs, sm.ap, sm.cq, pre, cc, pv, d, newtrees, &sm.bad);
    }
    __syncthreads();
    if (t == 0) { sm.ap.n = nb; sm.ap.first = B.first; sm.ap.forest = B.forest; }
    __syncthreads();
    RPROF(6);
    if (rprof) { atomicAdd((unsigned long long*)&d.prof[8], (unsigned long long)rp[0]); atomicAdd((unsigned long long*)&d.prof[9], (unsigned long long)rp[1]); atomicAdd((unsigned long long*)&d.prof[14], (unsigned long long)rp[6]); }
#undef RPROF
}

// proposals of sweep `sweep` for the trees j = blockIdx.x, + gridDim.x, ... from their (visible) HBM records
__device__ __forceinline__ void propose_all_b(BatchSmem& sm, const Dev& d, const TreeRec* trees, uint32_t sweep)
{
    for (int j = blockIdx.x; j < d.T; j += gridDim.x) {
        const ForestDev& F = d.f[(d.f[1].ntree > 0 && j >= d.f[1].tree0) ? 1 : 0];
        load_tree(&sm.big.tr, &trees[j]);
        __syncthreads();
        propose_stage(sm, sm.big, d, F, j, sweep);
        publish_stage(sm.big, d, j);
        __syncthreads();
    }
}

#define BCF_PROF_DECL                                                                                   \
    [[maybe_unused]] const bool prof = (d.prof != nullptr) && blockIdx.x == 0 && t == 0;                                 \
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};                                                         \
    [[maybe_unused]] long long tp = BCF_PROF_CLOCK();
#ifndef BCF_PROF
#define PROF(k) do { } while (0)
#else
#define PROF(k) do { if (prof) { const long long now_ = clock64(); pc[k] += now_ - tp; tp = now_; } } while (0)
#endif

// done_in (may be null): sweeps of this request the pipelined kernel has already run (device memory): the launch queued
// behind k_pipe to take over a sweep that holds a wide tree returns at once when there is nothing left for it
template <bool RES>
__global__ void __launch_bounds__(PTHREADS, 1) k_batched(Dev d, int nsweeps, unsigned int* bar, int ntl_max, const int* done_in)
{
    if (done_in != nullptr) nsweeps -= __ldcg(done_in);
    if (nsweeps <= 0) return;
    extern __shared__ __align__(16) unsigned char bcf_dyn_smem[];
    BatchSmem& sm = *reinterpret_cast<BatchSmem*>(bcf_dyn_smem);
    unsigned char* dyn = bcf_dyn_smem + BATCH_SMEM_BYTES;
    __shared__ Scalars scs;
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const bool writer0 = (blockIdx.x == 0);
    const int lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    unsigned int target = 0;
    unsigned long long bev = 0, mev = 0;   // batch / moments exchanges of this launch (sharded mode)
    int bad = 0;
    BCF_PROF_DECL
    BatPtrs rs;
    rs.R = (uint32_t)__cvta_generic_to_shared(dyn);
    rs.G = rs.R + (uint32_t)ntl_max * (TILE * 8u);
    rs.keys = rs.G + (uint32_t)ntl_max * (TILE * 8u);
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;
    const int tmod = d.f[1].tree0;
    const bool has_mod = d.f[1].ntree > 0;
    const int tile0 = (int)((int64_t)blockIdx.x * d.n_tiles / gridDim.x);
    const int tile1 = (int)((int64_t)(blockIdx.x + 1) * d.n_tiles / gridDim.x);
    const int ntl = tile1 - tile0;
    const int pad_a = d.trt_tiles - 1 - tile0, pad_b = d.n_tiles - 1 - tile0;   // local tiles that may hold padding

    // entry: the residual and the mu forest's fit move into shared memory
    if (RES)
        for (int lt = wib; lt < ntl; lt += wpb) {
            const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
            const uint32_t toff = (uint32_t)lt * (TILE * 8u);
            long long R[8];
            double gc[8];
            ld8_i64(d.rfx + gb, R); ld8_f64(d.Gc + gb, gc);
            res_st8i(rs.R + toff, lane, R);
            res_st8(rs.G + toff, lane, gc);
        }
    __syncthreads();
    // proposals of the first sweep of this launch (later sweeps: made in the previous sweep's epilogue)
    propose_all_b(sm, d, d.trees[scs.parity], scs.sweep);
    grid_arrive(bar, target);
    if (!grid_wait(bar, target, d.err, &bar_ok)) return;

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        long long* totals = sweep_totals(d, sc.sweep);
        long long* cross = d.cross + (size_t)(sc.sweep & 1u) * (size_t)T * (XCAP * 2);
        if (t == 0) sm.ap.n = 0;
        stage_batch(sm, sm.bst[0], d, d.trees[par], 0);
        if (sm.bst[0].big) stage_big(sm, sm.bst[0], d, d.trees[par]);
        begin_batch(sm, sm.bst[0]);
        __syncthreads();
        PROF(6);
        int first = 0, bi = 0;
        while (first < T) {
            BatchStage& B = sm.bst[bi & 1];
            BatchStage& Bn = sm.bst[(bi + 1) & 1];
            const int nb = B.nb;
            const bool do_apply = sm.ap.n > 0;
            const bool fswitch = has_mod && first == tmod && first > 0;
            // the CTA tables hold 32-bit words: at most FLUSH_TILES tiles between two flushes (HBM mode: many tiles)
            for (int lt0 = 0; lt0 < ntl; lt0 += (RES ? (1 << 30) : FLUSH_TILES)) {   // resident: an SM holds < 48 tiles, one chunk
                const int lt1 = RES ? ntl : min(ntl, lt0 + FLUSH_TILES);
                for (int lt = lt0 + wib; lt < lt1; lt += wpb) {
                    if (lt == pad_a || lt == pad_b) batch_tile<true, RES>(sm, d, rs, tile0, lt, lane, &B, do_apply, fswitch);
                    else batch_tile<false, RES>(sm, d, rs, tile0, lt, lane, &B, do_apply, fswitch);
                }
                if (!RES && lt1 < ntl) {   // (compile-time dead in the resident kernel: one call site of the flush there)
                    __syncthreads();
                    flush_batch(sm, B, totals, cross);
                    __syncthreads();
                    begin_batch(sm, B);
                    __syncthreads();
                }
            }
            __syncthreads();
            PROF(0);
            flush_batch(sm, B, totals, cross);
            grid_arrive(bar, target);
            PROF(1);
            if (first + nb < T) stage_batch(sm, Bn, d, d.trees[par], first + nb);   // while the barrier is in flight
            PROF(5);
            if (!grid_wait(bar, target, d.err, &bar_ok)) return;
            if (d.pr.nranks > 1) {   // observation shards: swap the batch's words with the peers
                ++bev;
                if (blockIdx.x == 0) peer_push_batch(B, d, totals, cross, (int)(bev & 1u), d.pr.seq0 + bev);
                if (!peer_wait(d, (int)(bev & 1u), d.pr.seq0 + bev, &bar_ok)) return;
            }
            PROF(2);
            resolve_batch(sm, B, d, sc, d.trees[par ^ 1], totals, cross, (int)(bev & 1u));
            PROF(3);
            first += nb; ++bi;
            if (first < T) {
                if (Bn.big) stage_big(sm, Bn, d, d.trees[par]);   // BatchSmem::big is free only now
                begin_batch(sm, Bn);
            }
            __syncthreads();
            PROF(4);
        }
        // ---- epilogue: apply the last batch, moments of (y, Gc, Gm) ----
        {
            long long* tb = reinterpret_cast<long long*>(&sm.p.tab);
            for (int i = t; i < STATTAB_WORDS64; i += blockDim.x) tb[i] = 0;
        }
        __syncthreads();
        {
            const bool fswitch = false;
            for (int lt = wib; lt < ntl; lt += wpb) {
                if (lt == pad_a || lt == pad_b) batch_tile<true, RES>(sm, d, rs, tile0, lt, lane, nullptr, true, fswitch);
                else batch_tile<false, RES>(sm, d, rs, tile0, lt, lane, nullptr, true, fswitch);
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                double G[8], y[8], other[8];
                if (RES) res_ld8(rs.G + (uint32_t)lt * (TILE * 8u), lane, G); else ld8_f64((has_mod ? d.Gm : d.Gc) + gb, G);
                ld8_f64(d.y + gb, y);
                if (has_mod) { ld8_f64(d.Gc + gb, other); moments8(sm.p.tab, y, other, G, tile < d.trt_tiles ? 1 : 0, lane, bad); }
                else {
#pragma unroll
                    for (int o = 0; o < 8; ++o) other[o] = 0.0;
                    moments8(sm.p.tab, y, G, other, tile < d.trt_tiles ? 1 : 0, lane, bad);
                }
            }
        }
        __syncthreads();
        {
            const int nwarps = blockDim.x >> 5;
            for (int i = t; i < NMOM * 8; i += blockDim.x) {
                const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
                if (q == 0) continue;
                const long long v = table_word(sm.p.tab, nwarps, m, g, q);
                if (v != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)v);
            }
        }
        grid_arrive(bar, target);
        // the next sweep's proposals need only the new trees (each is proposed by the CTA that wrote it) and the Philox
        // counters: they are made while the moments cross the grid barrier
        if (sw + 1 < nsweeps) propose_all_b(sm, d, d.trees[par ^ 1], sc.sweep + 1u);
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        if (d.pr.nranks > 1) {   // the sweep's moments, summed over the shards
            ++mev;
            if (blockIdx.x == 0) {
                for (int i = t; i < MAIL_MOM_WORDS; i += blockDim.x) {
                    const long long v = __ldcg(&mom[i]);
                    for (int p = 0; p < d.pr.nranks; ++p)
                        if (p != d.pr.rank) d.pr.peer[p][d.pr.rank].mom[i] = v;
                }
                peer_release(d, MAIL_FLAG_MOM, d.pr.seq0 + mev);
            }
            if (!peer_wait(d, MAIL_FLAG_MOM, d.pr.seq0 + mev, &bar_ok)) return;
            for (int i = t; i < MAIL_MOM_WORDS; i += blockDim.x) {
                long long v = __ldcg(&mom[i]);
                for (int p = 0; p < d.pr.nranks; ++p)
                    if (p != d.pr.rank) v += __ldcg(&d.pr.mine[p].mom[i]);
                sm.momsum[i] = v;
            }
            __syncthreads();
            sweep_scalars(d, d.trees[par ^ 1], sm.momsum, sm.p.red, &scs, writer0, true);
        } else
            sweep_scalars(d, d.trees[par ^ 1], mom, sm.p.red, &scs, writer0);
        grid_arrive(bar, target);   // the proposals must be visible before the next sweep stages them (wait: below)
        // ---- finishing pass: posterior accumulation, residual re-derived with the new scales, back to the mu forest ----
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            for (int lt = wib; lt < ntl; lt += wpb) {
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                const uint32_t toff = (uint32_t)lt * (TILE * 8u);
                const double b = tile < d.trt_tiles ? b1 : b0;
                double gc[8], gm[8], y[8];
                long long R[8];
                ld8_f64(d.y + gb, y);
                if (!RES) {
                    ld8_f64(d.Gc + gb, gc);
                    if (has_mod) ld8_f64(d.Gm + gb, gm);
                    else {
#pragma unroll
                        for (int o = 0; o < 8; ++o) gm[o] = 0.0;
                    }
                } else if (has_mod) { res_ld8(rs.G + toff, lane, gm); ld8_f64(d.Gc + gb, gc); st8_f64(d.Gm + gb, gm); res_st8(rs.G + toff, lane, gc); }
                else {
                    res_ld8(rs.G + toff, lane, gc);
#pragma unroll
                    for (int o = 0; o < 8; ++o) gm[o] = 0.0;
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) R[o] = resync_fx(y[o], ms, gc[o], b, gm[o], bad);
                if (RES) res_st8i(rs.R + toff, lane, R); else st8_i64(d.rfx + gb, R);
                if (row >= 0) {
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int64_t i = gb + o;
                        const double mu = ms * gc[o], yh = mu + b * gm[o], tau = (b1 - b0) * gm[o];
                        d.tau_sum[i] += tau; d.tau_sq[i] += tau * tau; d.mu_sum[i] += mu; d.yhat_sum[i] += yh;
                        if (d.draws_tau) d.draws_tau[(int64_t)row * d.n_pad + i] = tau;
                        if (d.draws_mu) d.draws_mu[(int64_t)row * d.n_pad + i] = mu;
                        if (d.draws_yhat) d.draws_yhat[(int64_t)row * d.n_pad + i] = yh;
                    }
                }
            }
            const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + t, gstride = (int64_t)gridDim.x * blockDim.x;
            const int64_t ntot = (int64_t)T * KEYS * 8;
            for (int64_t i = gtid; i < ntot; i += gstride) totals[i] = 0;
            const int64_t nx = (int64_t)T * XCAP * 2;
            for (int64_t i = gtid; i < nx; i += gstride) cross[i] = 0;
            long long* mom_next = d.mom + (size_t)((sc.sweep + 1u) & 1u) * (2 * NMOM * 3);
            for (int64_t i = gtid; i < 2 * NMOM * 3; i += gstride) mom_next[i] = 0;
        }
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        PROF(6);
    }
    // exit: park the state in HBM in the form every engine understands (Gm was stored by the finishing pass)
    if (RES)
        for (int lt = wib; lt < ntl; lt += wpb) {
            const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
            const uint32_t toff = (uint32_t)lt * (TILE * 8u);
            double G[8];
            long long R[8];
            res_ld8(rs.G + toff, lane, G);
            res_ld8i(rs.R + toff, lane, R);
            st8_f64(d.Gc + gb, G);
            st8_i64(d.rfx + gb, R);
        }
    if (prof) for (int k = 0; k < 8; ++k) d.prof[k] += pc[k];
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
    if (t == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

#undef PROF
#undef BCF_PROF_DECL

}  // namespace bcf
