// bcf_pipe.cuh -- PIPELINED engine: the batched engine with the decisions taken off the observation pass's critical path.
//
// In the batched engine a batch costs  pass + flush + grid barrier + decision chain, one after the other, and during
// the last three almost every warp of the SM idles.  Here the CTA is split in three roles:
//   * 27 TILE warps stream the SM's observations (27 tiles of 256 at the headline size: one round):
//         STATS(b)  ->  flush, warp 0 arrives at grid barrier b  ->  APPLY(b-1)  ->  STATS(b+1) ...
//   * 4 CHAIN warps (one per tree of a batch) wait for grid barrier b, fetch the reduced statistics, walk the decision
//     chain of batch b and publish its apply tables;
//   * 1 UTILITY warp plans the batches and stages their trees from HBM, two batches ahead of the statistics, paced by
//     a shared-memory progress counter: nothing waits for it in the steady state.
// STATS(b) therefore runs while batch b-1 is still being decided: it sees the residual WITHOUT batch b-1's update.
// That is repaired exactly like the dependencies inside a batch (bcf_batched.cuh): the pass also counts, per pair
// (tree m of batch b-1, tree q of batch b), how many observations fall in each pair of keys, and the chain folds
// D_m * N into tree q's integer sums before deciding it.  Still bit-identical to the tree-at-a-time chain.
// CTA-local hand-shakes: named barriers S(b) / F(b) among the tile warps ("statistics of b complete" / "b flushed"),
// H1(b): tile warp 0 -> chain "batch b flushed (and, transitively, planned and staged)", H2(b): chain -> tile warps
// "batch b's tables are ready"; counters plan_ready / flushed in shared memory (release / acquire) between the utility
// warp and tile warp 0.  The grid barrier is arrived at by tile warp 0 and waited for by the chain warps.
//
// Only sweeps whose trees all have <= KB keys run here: when a wider tree shows up at the start of a sweep the kernel
// parks the chain state and returns; the host runs that sweep with the batched engine and comes back.
#pragma once
#include "bcf_batched.cuh"

namespace bcf {

constexpr int PNB = 4;                               // trees per batch of this engine (its syncs are hidden: small batches
                                                     // mean less cross-counting; and 27 warps hold an SM's 27 tiles in ONE round)
constexpr int PP_WARPS = 28;                         // pass warps: PT_WARPS tile warps + the utility warp
constexpr int PT_WARPS = PP_WARPS - 1;
#ifndef PIPE_CHAIN_FIRST
#define PIPE_CHAIN_FIRST 0
#endif
constexpr int PR_WARPS = PNB;                        // chain warps: one per tree of a batch
constexpr int PP_THREADS = PP_WARPS * 32, PR_THREADS = PR_WARPS * 32;
constexpr int PIPE_THREADS = PP_THREADS + PR_THREADS;   // 1024
constexpr int PMAXC = 12;                            // columns ((tree, key >= 1) pairs) of a batch
constexpr int PXW = 72;                              // cells between trees of one batch: (C^2 - sum c_q^2) / 2 <= 66
constexpr int PXX = PMAXC * PMAXC;                   // cells between two consecutive batches
constexpr int PXCAP = PXW + PXX;
constexpr int PIPE_XSTRIDE = 512;                    // words per batch slot of Dev::cross (>= 2 * PXCAP)
constexpr int PIPE_BYTES_PER_TILE = TILE * (8 + 8) + 2 * (PNB / 2) * TILE;   // R, G, nibble-packed keys of two batches
static_assert(2 * PXCAP <= PIPE_XSTRIDE, "cross-count slot too small");

// registers per thread: 1024 x 64 at launch; the role section re-splits the CTA's OWN allocation (the pool of setmaxnreg
// is per CTA, the instruction per warpgroup of 4 warps): 896 x 56 + 128 x 112 = 64512
constexpr int PIPE_REGS_BASE = 64, PIPE_REGS_PASS = 56, PIPE_REGS_RESOLVE = 112;
enum { BAR_S = 5, BAR_F = 8, BAR_H1 = 9, BAR_H2 = 11, BAR_RES = 14 };   // 1..PNB: the decision chain
constexpr int H1_THREADS = 32 + PR_THREADS;          // tile warp 0 arrives, chain warps wait
constexpr int PT_THREADS = PT_WARPS * 32;
constexpr int H2_THREADS = PT_THREADS + PR_THREADS;  // chain warps arrive, tile warps wait

struct __align__(8) TMeta {                          // what the pass needs to know about a tree, per sweep (8 bytes: the
    uint16_t var, cut;                               // column pointer of the proposed variable is formed by the utility warp)
    uint8_t K, L, birth, sx;
};

struct PipePlan {
    int32_t first, nb, forest, nprev;                // nprev: trees of the previous batch not yet applied (0 or its nb)
    int32_t ncell_w, ncell, ncol, ncolp;
    int32_t K[PNB], Kp[PNB];
    uint8_t colbase[PNB], colbasep[PNB];
    int16_t cbase[PNB], cbasep[PNB];                   // first cell of tree q: inside the batch / against the previous batch
    int16_t xoff[PNB][PNB];                            // [q][m < q]       cells inside the batch
    int16_t xoffp[PNB][PNB];                           // [q][m in prev]   cells against the previous batch
    uint8_t cell_c1[PXCAP], cell_c2[PXCAP];          // columns of a cell: 0.. this batch, PMAXC.. previous batch
    CurInfo ci[PNB];
};

struct __align__(16) PipeApply {
    int32_t n, first, forest, pad_;
    int32_t changed[PNB], off[PNB];
    ApRow T[2][PNB * KB];
    int4 DL[PNB * KB][2];
    long long TL[PNB][2][4];
    uint8_t NS[PNB * KB];
};
struct PipeConsts { ChainConsts cc; double a; };    // what a batch's decisions need from the sweep's scalars
struct PipeRS { long long tot[PNB][KB][2][4]; int x[PXCAP][2]; };
struct PipeAcc { unsigned int btab[PNB][KB][2][4]; unsigned int totr[2][4]; unsigned int xtab[PXCAP][2]; };

struct PipeSmem {
    // scratch of propose_stage() (same names as StepShared; the trees here have <= KB leaves)
    int32_t ngood[KB];
    uint32_t lkey[KB];
    uint32_t nkey[2 * KB];
    uint8_t nogflag[2 * KB];
    int32_t bad;
    int32_t anybig;
    int32_t plan_ready, flushed;                      // batches planned + staged (utility warp) / flushed (tile warps), this sweep
    SmallStage pst;                                  // stage of the proposals
    SmallStage st[4][PNB];                            // the batch's trees (staged two batches ahead)
    PipePlan plan[4];                                 // (planned two batches ahead)
    PipeApply ap[2];
    int32_t cq[2][PNB][KB][2];
    PipeRS rs;
    PreKey pre[PNB * (KB + 1)];
    PipeConsts cons[2];                               // per forest, per sweep
    union { PipeAcc acc[2]; StatTabT<32> tab; double red[2 * PIPE_THREADS]; } p;
    alignas(16) uint8_t colb[PT_WARPS][2 * PMAXC][32];   // per tile warp; read 16 bytes at a time
};
constexpr size_t PIPE_SMEM_BYTES = (sizeof(PipeSmem) + 15) / 16 * 16;

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n)
{
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__device__ __forceinline__ int ld_acquire_cta_s32(const int* p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s32(int* p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// one lane waits for a CTA-local progress counter to reach `target` (bounded: a logic error traps instead of hanging)
__device__ __forceinline__ void pipe_wait_flag(const int* p, int target, int32_t* err, unsigned ns, [[maybe_unused]] int* dbg = nullptr, [[maybe_unused]] int what = 0)
{
    const long long t0 = clock64();
    while (ld_acquire_cta_s32(p) < target) {
        __nanosleep(ns);
        // (a plain trap: a call to give_up() -- or its stores inlined, or the call moved behind the loop -- in this loop and in the
        // chain's grid-barrier poll costs 3.5 % of the sweep, 1 110 against 1 150 sweeps/s on the same box; the waits that only
        // exist in the sharded mode keep their notes)
        if (clock64() - t0 > 8000000000ll) { atomicOr(err, ERR_BARRIER_TIMEOUT); __trap(); }
    }
}

struct PipePtrs { uint32_t R, G, keys; };            // keys: [tile][parity][pair][TILE bytes]

// ---------------------------------------------------------------------------------------------------------------
// pass role
// ---------------------------------------------------------------------------------------------------------------
// Plan of the batch starting at `first`, by ONE warp (the utility warp).  Cell layout (closed form, so nothing is enumerated
// serially): tree q owns the cells  base_q + (K_q - 1) * c1 + (k - 1)  for every EARLIER column c1 of the batch
// (c1 < colbase[q]) and, behind all of those, the cells against the previous batch's columns; hence
//     xoff[q][m]  = base_q  + (K_q - 1) * colbase[m],        xoffp[q][m] = basep_q + (K_q - 1) * colbasep[m].
__device__ inline void pipe_plan(PipePlan& P, const PipePlan* prev, const Dev& d, const TMeta* meta, int first, int lane)
{
    {
        const unsigned FULL = 0xffffffffu;
        const int l = lane;
        const int f = (d.f[1].ntree > 0 && first >= d.f[1].tree0) ? 1 : 0;
        const int fend = f ? d.T : (d.f[1].ntree > 0 ? d.f[1].tree0 : d.T);
        const int ncand = min(PNB, fend - first);
        TMeta m;
        m.K = 1; m.L = 1; m.birth = 0; m.sx = 0; m.cut = 0; m.var = 0;
        if (l < ncand) m = meta[first + l];
        const int c = (l < ncand) ? (int)m.K - 1 : 0;            // own columns
        int cb = c;                                                 // inclusive prefix of the columns
#pragma unroll
        for (int sft = 1; sft < PNB; sft <<= 1) { const int v = __shfl_up_sync(FULL, cb, sft); if (l >= sft) cb += v; }
        const int colbase = cb - c;
        const unsigned okm = __ballot_sync(FULL, l < ncand && cb <= PMAXC);
        const int nb = __ffs(~okm) - 1;                             // leading trees whose columns still fit
        const bool in = l < nb;
        const int wcells = in ? c * colbase : 0;                    // cells against earlier columns of this batch
        int wb = wcells;
#pragma unroll
        for (int sft = 1; sft < PNB; sft <<= 1) { const int v = __shfl_up_sync(FULL, wb, sft); if (l >= sft) wb += v; }
        const int base = wb - wcells;
        const int ncell_w = __shfl_sync(FULL, wb, PNB - 1);
        const int ncol = __shfl_sync(FULL, in ? cb : 0, max(nb - 1, 0));
        const int np = prev ? prev->nb : 0, ncolp = prev ? prev->ncol : 0;
        const int xcells = in ? c * ncolp : 0;                      // cells against the previous batch's columns
        int xb = xcells;
#pragma unroll
        for (int sft = 1; sft < PNB; sft <<= 1) { const int v = __shfl_up_sync(FULL, xb, sft); if (l >= sft) xb += v; }
        const int basep = ncell_w + xb - xcells;
        const int ncell = ncell_w + __shfl_sync(FULL, xb, PNB - 1);
        if (l < PNB) {
            P.K[l] = in ? (int)m.K : 1;
            P.colbase[l] = (uint8_t)colbase;
            P.cbase[l] = (int16_t)base; P.cbasep[l] = (int16_t)basep;
            P.Kp[l] = (l < np) ? prev->K[l] : 1;
            P.colbasep[l] = (l < np) ? prev->colbase[l] : 0;
            CurInfo& ci = P.ci[l];
            int is16 = 0;
            const void* bins = (l < ncand && m.birth) ? bin_column(d.f[f], (int)m.var, d.n_pad, is16) : nullptr;
            ci.active = 1; ci.K = m.K; ci.birth = m.birth; ci.sx = m.sx; ci.cut = m.cut; ci.bins = bins; ci.is16 = is16;
            ci.forest = f; ci.new_slot = m.L;
        }
        if (l == 0) {
            P.first = first; P.nb = nb; P.forest = f; P.ncol = ncol; P.ncell_w = ncell_w; P.nprev = np; P.ncolp = ncolp;
            P.ncell = ncell;
        }
    }
    __syncwarp();
    {   // offsets of every pair and the two columns of every cell, from the closed form
        if (lane < PNB * PNB) {
            const int q = lane / PNB, m = lane % PNB;
            P.xoff[q][m] = (int16_t)(P.cbase[q] + (P.K[q] - 1) * (int)P.colbase[m]);
            P.xoffp[q][m] = (int16_t)(P.cbasep[q] + (P.K[q] - 1) * (int)P.colbasep[m]);
        }
        const int nb = P.nb, ncw = P.ncell_w, nc = P.ncell;
        for (int cell = lane; cell < nc; cell += 32) {
            const bool cross = cell >= ncw;
            int q = 0;
            for (int qq = 1; qq < nb; ++qq) if (cell >= (cross ? P.cbasep[qq] : P.cbase[qq])) q = qq;
            const int j = cell - (cross ? P.cbasep[q] : P.cbase[q]), w = P.K[q] - 1;
            P.cell_c1[cell] = (uint8_t)((cross ? PMAXC : 0) + j / w);
            P.cell_c2[cell] = (uint8_t)(P.colbase[q] + j % w);
        }
    }
    __syncwarp();
}

// statistics of batch P over one warp tile; `kw` / `kr`: key areas of this batch (written) and of the previous one
// keep_prev: this warp has ONE tile (a single round of tiles), so the byte columns it formed for the previous batch are still
// in its column buffer (they live in the halves of `colb` by batch parity) and need not be rebuilt from the stored keys.
__device__ __forceinline__ void pipe_stats_tile(PipeSmem& sm, PipeAcc& acc, const PipePlan& P, const Dev& d, const PipePtrs& rs,
                                                int tile0, int lt, int lane, int warp, int parity, bool keep_prev)
{
    const int tile = tile0 + lt;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const int64_t gb = (int64_t)tile * TILE + lane * 8;
    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
    const uint32_t kbase = rs.keys + (uint32_t)lt * (uint32_t)(2 * (PNB / 2) * TILE) + (uint32_t)lane * 8u;
    const uint32_t kw = kbase + (uint32_t)parity * (uint32_t)((PNB / 2) * TILE);
    const uint32_t kr = kbase + (uint32_t)(parity ^ 1) * (uint32_t)((PNB / 2) * TILE);
    long long R[8];
    uint2 sl_n = make_uint2(0u, 0u);
    uint4 bn_n = make_uint4(0u, 0u, 0u, 0u);
    keys_load(d, P.ci[0], P.first, gb, sl_n, bn_n);
    res_ld8i(rs.R + toff, lane, R);
    {
        long long tot = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) tot += R[o];
        warp_add21(acc.totr[g], tot, lane);
    }
    uint8_t* colb = &sm.colb[warp][parity * PMAXC][0];            // this batch's columns
    uint8_t* colp = &sm.colb[warp][(parity ^ 1) * PMAXC][0];      // the previous batch's
    const int nb = P.nb;
    uint2 held = make_uint2(0u, 0u);
    for (int q = 0; q < nb; ++q) {
        const CurInfo& ci = P.ci[q];
        const int K = P.K[q];
        const uint2 sl = sl_n;
        const uint4 bn = bn_n;
        if (q + 1 < nb) keys_load(d, P.ci[q + 1], P.first + q + 1, gb, sl_n, bn_n);
        const uint2 key = keys_make(ci, sl, bn);
        // two trees share a byte (4 bits each; padding observations read 0xF)
        const uint2 nib = make_uint2(key.x & 0x0F0F0F0Fu, key.y & 0x0F0F0F0Fu);
        if (q & 1) sts_v2(kw + (uint32_t)(q >> 1) * TILE, make_uint2(held.x | (nib.x << 4), held.y | (nib.y << 4)));
        else { held = nib; if (q == nb - 1) sts_v2(kw + (uint32_t)(q >> 1) * TILE, held); }
        uint8_t* cb = colb + (int)P.colbase[q] * 32 + lane;
        for (int k = 1; k < K; ++k) {
            const uint32_t kv = (uint32_t)k * 0x01010101u;
            const uint32_t bits = bytes_to_bits(__vcmpeq4(key.x, kv), __vcmpeq4(key.y, kv));
            long long m[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) m[o] = ((bits >> o) & 1u) ? R[o] : 0ll;
            const long long s = ((m[0] + m[1]) + (m[2] + m[3])) + ((m[4] + m[5]) + (m[6] + m[7]));   // short dependency chain
            unsigned int* e = acc.btab[q][k][g];
            warp_add21(e, s, lane);
            if (ci.birth && k == K - 1) {
                const int c = __reduce_add_sync(0xffffffffu, __popc(bits));
                if (lane == 0) atomicAdd(&e[0], (unsigned int)c);
            }
            cb[(k - 1) * 32] = (uint8_t)bits;
        }
    }
    // the previous batch's columns, from its stored keys
    const int np = keep_prev ? 0 : P.nprev;
    for (int j = 0; 2 * j < np; ++j) {
        const uint2 w = lds_v2(kr + (uint32_t)j * TILE);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = 2 * j + h;
            if (m < np) {
                const uint32_t n0 = (w.x >> (4 * h)) & 0x0F0F0F0Fu, n1 = (w.y >> (4 * h)) & 0x0F0F0F0Fu;
                uint8_t* cb = colp + (int)P.colbasep[m] * 32 + lane;
                const int Km = P.Kp[m];
                for (int kp = 1; kp < Km; ++kp) {
                    const uint32_t kv = (uint32_t)kp * 0x01010101u;
                    cb[(kp - 1) * 32] = (uint8_t)bytes_to_bits(__vcmpeq4(n0, kv), __vcmpeq4(n1, kv));
                }
            }
        }
    }
    __syncwarp();
    for (int cell = lane; cell < P.ncell; cell += 32) {
        const int c1 = (int)P.cell_c1[cell];                      // >= PMAXC: a column of the previous batch
        const uint4* a = reinterpret_cast<const uint4*>((c1 < PMAXC ? colb : colp - PMAXC * 32) + c1 * 32);
        const uint4* b = reinterpret_cast<const uint4*>(colb + (int)P.cell_c2[cell] * 32);
        const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
        const int c = __popc(a0.x & b0.x) + __popc(a0.y & b0.y) + __popc(a0.z & b0.z) + __popc(a0.w & b0.w) +
                      __popc(a1.x & b1.x) + __popc(a1.y & b1.y) + __popc(a1.z & b1.z) + __popc(a1.w & b1.w);
        if (c) atomicAdd(&acc.xtab[cell][g], (unsigned int)c);
    }
    __syncwarp();
}

// apply a decided batch to one warp tile; fswitch: the batch is the first of the tau forest (G changes forest first)
template <bool PAD>
__device__ __forceinline__ void pipe_apply_tile(const PipeApply& ap, const Dev& d, const PipePtrs& rs, int tile0, int lt, int lane,
                                                int parity, bool fswitch)
{
    const int tile = tile0 + lt;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const int64_t gb = (int64_t)tile * TILE + lane * 8;
    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
    const uint32_t kr = rs.keys + (uint32_t)lt * (uint32_t)(2 * (PNB / 2) * TILE) + (uint32_t)lane * 8u +
                        (uint32_t)parity * (uint32_t)((PNB / 2) * TILE);
    // two halves of 4 observations: half the live state (the role has 56 registers; with all 8 the loop spilled)
    const int na = ap.n;
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
        long long R[4];
        double G[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a = (uint32_t)((hf * 4 + k) * 32 + lane) * 8u;
            const uint2 v = lds_v2(rs.R + toff + a);
            R[k] = (long long)(((unsigned long long)v.y << 32) | v.x);
            G[k] = lds_f64(rs.G + toff + a);
        }
        if (fswitch) {   // park the mu forest's fit in HBM, bring the tau forest's in
            double2* gc = reinterpret_cast<double2*>(d.Gc + gb + hf * 4);
            const double2* gm = reinterpret_cast<const double2*>(d.Gm + gb + hf * 4);
            gc[0] = make_double2(G[0], G[1]); gc[1] = make_double2(G[2], G[3]);
            const double2 m0 = gm[0], m1 = gm[1];
            G[0] = m0.x; G[1] = m0.y; G[2] = m1.x; G[3] = m1.y;
        }
        for (int j = 0; 2 * j < na; ++j) {
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(kr + (uint32_t)j * TILE + (uint32_t)hf * 4u) : "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * j + h;
                if (q < na) {
                    const uint32_t kk = (w >> (4 * h)) & 0x0F0F0F0Fu;
                    const int off = ap.off[q];
                    const ApRow* Tq = &ap.T[g][off];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int key = (int)((kk >> (8 * o)) & 0xFFu);
                        if (PAD && key == 15) continue;
                        const ApRow r = Tq[key];
                        R[o] -= r.dr;
                        G[o] += r.dg;
                    }
                    if (ap.changed[q]) {
                        uint32_t nsw = 0u;
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            const int key = (int)((kk >> (8 * o)) & 0xFFu);
                            const uint32_t ns = (PAD && key == 15) ? (uint32_t)PAD_SLOT : (uint32_t)ap.NS[off + key];
                            nsw |= ns << (8 * o);
                        }
                        *reinterpret_cast<uint32_t*>(d.slots + (int64_t)(ap.first + q) * d.n_pad + gb + hf * 4) = nsw;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a = (uint32_t)((hf * 4 + k) * 32 + lane) * 8u;
            sts_f64(rs.G + toff + a, G[k]);
            const unsigned long long u = (unsigned long long)R[k];
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs.R + toff + a), "r"((uint32_t)u), "r"((uint32_t)(u >> 32)) : "memory");
        }
    }
}

// the CTA's tables of one batch -> the grid-wide ones, by the tile warps (one word per thread); leaves the tables zeroed
__device__ inline void pipe_flush(PipeAcc& acc, const PipePlan& P, long long* totals, long long* cross, int tt)
{
    const int nb = P.nb;
    constexpr int NTOT = PNB * KB * 8;
    for (int i = tt; i < NTOT + 2 * PXCAP; i += PT_WARPS * 32) {
        if (i < NTOT) {
            const int q = i / (KB * 8), k = (i >> 3) & (KB - 1), g = (i >> 2) & 1, w = i & 3;
            long long v = 0;
            if (k == 0) { if (q == 0) { v = acc_word(acc.totr[g], w); acc.totr[g][w] = 0u; } }   // key 0 of the first tree carries the total
            else if (q < nb && k < P.K[q]) { v = acc_word(acc.btab[q][k][g], w); acc.btab[q][k][g][w] = 0u; }
            if (v != 0) atomicAdd((unsigned long long*)&totals[(size_t)(P.first + q) * KEYS * 8 + k * 8 + g * 4 + w], (unsigned long long)v);
        } else {
            const int j = i - NTOT;
            if (j < 2 * P.ncell) {
                const unsigned int v = acc.xtab[j >> 1][j & 1];
                acc.xtab[j >> 1][j & 1] = 0u;
                if (v != 0) atomicAdd((unsigned long long*)&cross[(size_t)P.first * PIPE_XSTRIDE + j], (unsigned long long)v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// resolve role
// ---------------------------------------------------------------------------------------------------------------
// compact copies of the (up to PNB) trees starting at `first`, with proposals and leaf normals, by ONE warp.  One tree
// at a time: all of a tree's loads are in flight together (one L2 round trip per tree) and nothing spills.
__device__ inline void pipe_stage_trees(SmallStage* st, const Dev& d, const TreeRec* trees, int first, int l)
{
    const int f = (d.f[1].ntree > 0 && first >= d.f[1].tree0) ? 1 : 0;
    const int fend = f ? d.T : (d.f[1].ntree > 0 ? d.f[1].tree0 : d.T);
    const int nq = min(PNB, fend - first);
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
        const TreeRec* src = &trees[first + q];
        SmallStage& S = st[q];
        long long pr = 0;
        double z = 0.0;
        if (l < (int)(sizeof(Proposal) / 8)) pr = __ldcg(reinterpret_cast<const long long*>(&d.proprec[first + q].P) + l);
        else if (l >= 16 && l < 16 + KB) z = __ldcg(&d.proprec[first + q].z[l - 16]);
        if (l < 2 * KB) {
            const int16_t pa = __ldcg(&src->parent[l]), le = __ldcg(&src->left[l]), ri = __ldcg(&src->right[l]);
            const uint16_t va = __ldcg(&src->var[l]), cu = __ldcg(&src->cut[l]);
            const uint8_t ns = __ldcg(&src->node_slot[l]), de = __ldcg(&src->depth[l]);
            const uint32_t he = __ldcg(&src->heap[l]);
            S.tr.parent[l] = pa; S.tr.left[l] = le; S.tr.right[l] = ri; S.tr.var[l] = va; S.tr.cut[l] = cu;
            S.tr.node_slot[l] = ns; S.tr.depth[l] = de; S.tr.heap[l] = he;
        } else if (l < 3 * KB) {
            const int s = l - 2 * KB;
            const int16_t sn = __ldcg(&src->slot_node[s]);
            const int32_t c0 = __ldcg(&src->cnt0[s]), c1 = __ldcg(&src->cnt1[s]);
            const double mu = __ldcg(&src->mu[s]);
            S.tr.slot_node[s] = sn; S.tr.cnt0[s] = c0; S.tr.cnt1[s] = c1; S.tr.mu[s] = mu;
        } else if (l == 3 * KB) {
            const int32_t nn = __ldcg(&src->nnodes), nl = __ldcg(&src->nleaves);
            S.tr.nnodes = nn; S.tr.nleaves = nl;
        }
        if (l < (int)(sizeof(Proposal) / 8)) reinterpret_cast<long long*>(&S.prop)[l] = pr;
        else if (l >= 16 && l < 16 + KB) S.z[l - 16] = z;
    }
    __syncwarp();
}

#ifndef LL_TIMEOUT
#define LL_TIMEOUT 120000000000ll
#endif
// ---- observation shards (see PeerDev, Mailbox::ll): the exchange of a batch's reduced words ----
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// index of word i of the batch's message: totals words of the trees' live keys first (i < nb * KB * 8), then the cross counts
__device__ __forceinline__ bool pipe_word_live(const PipePlan& P, int i)
{
    return i >= P.nb * KB * 8 || ((i >> 3) & (KB - 1)) < P.K[i / (KB * 8)];
}
// ONE warp (the utility warp of the first CTAs: CTA c serves the c-th peer), after this rank's grid barrier of the batch: this
// rank's reduced words, tagged with the batch's global number g, go into its mailbox on that peer (posted stores over NVLink).
__device__ inline void pipe_peer_push(const PipePlan& P, const Dev& d, const long long* totals, const long long* cross, int peer,
                                      unsigned long long g, int lane)
{
    const int nt = P.nb * KB * 8, nx = 2 * P.ncell;
    unsigned long long* dst = d.pr.peer[peer][d.pr.rank].ll[g & (MAIL_SLOTS - 1)];
    const unsigned long long tag = g & 0xFFFFull;
    for (int i0 = 0; i0 < nt + nx; i0 += 32 * 4) {
        long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                                 // loads first: one L2 round trip per 128 words
            const int i = i0 + u * 32 + lane;
            v[u] = 0;
            if (i < nt) { if (pipe_word_live(P, i)) v[u] = __ldcg(&totals[(size_t)(P.first + i / (KB * 8)) * KEYS * 8 + (i & (KB * 8 - 1))]); }
            else if (i < nt + nx) v[u] = __ldcg(&cross[(size_t)P.first * PIPE_XSTRIDE + (i - nt)]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            if (i < nt + nx && pipe_word_live(P, i)) st_relaxed_sys_u64(dst + i, ((unsigned long long)v[u] << 16) | tag);
        }
    }
}

// everything of a batch's decisions (resolve role, PR_THREADS threads; rt = thread within the role)
// gtag != 0: observation shards -- the batch's global number: the peers' words (Mailbox::ll) are added to this rank's
__device__ inline void pipe_resolve(PipeSmem& sm, SmallStage* st, const PipePlan& P, int b, const Dev& d, const Scalars& sc,
                                    TreeRec* newtrees, const long long* totals, const long long* cross, int rt,
                                    unsigned long long gtag, long long* trc = nullptr)
{
    const int nb = P.nb;
    PipeRS& rs = sm.rs;
    [[maybe_unused]] const bool rprof = d.prof != nullptr && blockIdx.x == 0 && rt == 0;
    [[maybe_unused]] long long rtp = BCF_PROF_CLOCK();
#ifndef BCF_PROF
#define RPROF(k) do { } while (0)
#else
#define RPROF(k) do { if (rprof) { const long long now_ = clock64(); atomicAdd((unsigned long long*)&d.prof[k], (unsigned long long)(now_ - rtp)); rtp = now_; } } while (0)
#endif
    {   // the batch's reduced words: every load is issued before the first store (one L2 round trip)
        constexpr int NT_IT = (PNB * KB * 8 + PR_THREADS - 1) / PR_THREADS, NX_IT = (2 * PXCAP + PR_THREADS - 1) / PR_THREADS;
        const long long* xg = cross + (size_t)P.first * PIPE_XSTRIDE;
        const int nx = 2 * P.ncell;
        long long tv[NT_IT], xv[NX_IT];
#pragma unroll
        for (int it = 0; it < NT_IT; ++it) {
            const int i = rt + it * PR_THREADS;
            const int q = i / (KB * 8), k = (i >> 3) & (KB - 1);
            tv[it] = (i < nb * KB * 8 && k < P.K[q]) ? __ldcg(&totals[(size_t)(P.first + q) * KEYS * 8 + (i & (KB * 8 - 1))]) : 0ll;
        }
#pragma unroll
        for (int it = 0; it < NX_IT; ++it) {
            const int i = rt + it * PR_THREADS;
            xv[it] = i < nx ? __ldcg(xg + i) : 0ll;
        }
        if (gtag != 0ull) {   // observation shards: add the peers' words (integers: any order gives the same bits)
            const unsigned long long tag = gtag & 0xFFFFull;
            const long long t0 = clock64();
            for (int p = 0; p < d.pr.nranks; ++p) {
                if (p == d.pr.rank) continue;
                const unsigned long long* mb = d.pr.mine[p].ll[gtag & (MAIL_SLOTS - 1)];
                unsigned long long pv[NT_IT + NX_IT];
                bool need[NT_IT + NX_IT];
#pragma unroll
                for (int it = 0; it < NT_IT + NX_IT; ++it) {
                    const int i = it < NT_IT ? rt + it * PR_THREADS : nb * KB * 8 + rt + (it - NT_IT) * PR_THREADS;
                    need[it] = it < NT_IT ? (i < nb * KB * 8 && pipe_word_live(P, i)) : (i - nb * KB * 8 < nx);
                    pv[it] = 0ull;
                }
                for (bool all = false; !all;) {        // every word carries its batch's tag: poll until each has arrived
                    all = true;
#pragma unroll
                    for (int it = 0; it < NT_IT + NX_IT; ++it) {
                        const int i = it < NT_IT ? rt + it * PR_THREADS : nb * KB * 8 + rt + (it - NT_IT) * PR_THREADS;
                        if (need[it]) pv[it] = ld_relaxed_sys_u64(mb + i);
                    }
#pragma unroll
                    for (int it = 0; it < NT_IT + NX_IT; ++it)
                        if (need[it]) { if ((pv[it] & 0xFFFFull) == tag) need[it] = false; else all = false; }
                    if (!all && clock64() - t0 > LL_TIMEOUT) { atomicOr(d.err, ERR_BARRIER_TIMEOUT); give_up(d.dbg, 20, p, (int)gtag); }
                }
#pragma unroll
                for (int it = 0; it < NT_IT; ++it) tv[it] += (long long)pv[it] >> 16;       // (words not needed stayed 0)
#pragma unroll
                for (int it = 0; it < NX_IT; ++it) xv[it] += (long long)pv[NT_IT + it] >> 16;
            }
        }
#pragma unroll
        for (int it = 0; it < NT_IT; ++it) {
            const int i = rt + it * PR_THREADS;
            if (i < nb * KB * 8) (&rs.tot[0][0][0][0])[i] = tv[it];
        }
#pragma unroll
        for (int it = 0; it < NX_IT; ++it) {
            const int i = rt + it * PR_THREADS;
            if (i < nx) (&rs.x[0][0])[i] = (int)xv[it];
        }
    }
    named_sync(BAR_RES, PR_THREADS);
    RPROF(8);
    const PipeConsts& pc = sm.cons[P.forest];
    if (rt < 2 * nb * (KB + 1)) {   // count-only terms of every key and pooled pair (as in resolve_batch); two lanes per
        const int part = rt & 1;    // entry: one takes 1/prec and its square root, the other log(prec)
        const int q = (rt >> 1) / (KB + 1), e = (rt >> 1) % (KB + 1);
        const SmallStage& S = st[q];
        const int L = S.tr.nleaves, K = P.K[q], mv = S.prop.move, sx = S.prop.slot;
        if (e < K || (e == KB && mv != 0)) {
            const double phi1 = pc.cc.phi1, phi0 = pc.cc.phi0;
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
            const double prec = pc.a + A;
            PreKey& pk = sm.pre[q * (KB + 1) + e];
            if (part == 0) {
                const double inv = 1.0 / prec;
                pk.k1 = k1; pk.k0 = k0; pk.inv = inv; pk.psd = sqrt(inv);
                if (e < K) { sm.cq[b & 1][q][e][1] = (int32_t)k1; sm.cq[b & 1][q][e][0] = (int32_t)k0; }
            } else {
                pk.plog = log(prec);
            }
        }
    }
    named_sync(BAR_RES, PR_THREADS);
    RPROF(9);
    {
        const ChainConsts cc = pc.cc;
        const PipeApply& app = sm.ap[(b & 1) ^ 1];
        ChainPrev pv{P.nprev, P.Kp, &P.xoffp[0][0], app.DL, sm.cq[(b & 1) ^ 1], app.TL};
        decision_chain<PNB>(rt >> 5, rt & 31, nb, P.first, st, P.K, &P.xoff[0][0], rs, sm.ap[b & 1], sm.cq[b & 1], sm.pre, cc, pv, d,
                       newtrees, &sm.bad, trc);
    }
    named_sync(BAR_RES, PR_THREADS);
    RPROF(14);
#undef RPROF
    if (rt == 0) { PipeApply& ap = sm.ap[b & 1]; ap.n = nb; ap.first = P.first; ap.forest = P.forest; }
}

// proposals of sweep `sweep` (all threads of the CTA)
__device__ __forceinline__ void pipe_propose_all(PipeSmem& sm, const Dev& d, const TreeRec* trees, uint32_t sweep)
{
    const int t = threadIdx.x;
    for (int j = blockIdx.x; j < d.T; j += gridDim.x) {
        const ForestDev& F = d.f[(d.f[1].ntree > 0 && j >= d.f[1].tree0) ? 1 : 0];
        if (t < 32) {
            // only this tree: reuse the staging code with a one-warp mapping
            const TreeRec* src = &trees[j];
            SmallStage& S = sm.pst;
            if (t < 2 * KB) {
                S.tr.parent[t] = __ldcg(&src->parent[t]); S.tr.left[t] = __ldcg(&src->left[t]); S.tr.right[t] = __ldcg(&src->right[t]);
                S.tr.var[t] = __ldcg(&src->var[t]); S.tr.cut[t] = __ldcg(&src->cut[t]); S.tr.node_slot[t] = __ldcg(&src->node_slot[t]);
                S.tr.depth[t] = __ldcg(&src->depth[t]); S.tr.heap[t] = __ldcg(&src->heap[t]);
            } else if (t < 3 * KB) {
                const int s = t - 2 * KB;
                S.tr.slot_node[s] = __ldcg(&src->slot_node[s]); S.tr.cnt0[s] = __ldcg(&src->cnt0[s]);
                S.tr.cnt1[s] = __ldcg(&src->cnt1[s]); S.tr.mu[s] = __ldcg(&src->mu[s]);
            } else if (t == 3 * KB) {
                S.tr.nnodes = __ldcg(&src->nnodes); S.tr.nleaves = __ldcg(&src->nleaves);
            }
        }
        __syncthreads();
        if (sm.pst.tr.nleaves <= KB) {
            propose_stage(sm, sm.pst, d, F, j, sweep);
            publish_stage(sm.pst, d, j);
        }
        __syncthreads();
    }
}

// probit BART: the latent draw that opens the NEXT sweep, for a lane's 8 observations of a tile, made in the finishing pass of
// the sweep before (k_probit_latent's arithmetic and Philox addresses, term for term; its own function so that the
// double-precision inverse normal does not weigh on the register allocation of the sweep itself)
__device__ __noinline__ void pipe_probit_latent8(Rng rng, const uint8_t* ybin, const int32_t* orig, double* yout, double offset, uint32_t sweep,
                                                   double ms, const double (&gc)[8], double (&y)[8], int* bad)
{
    // (scalars, not the Dev: a reference to the kernel's parameter block would make the kernel copy it to its stack)
    const uint2 yb = *reinterpret_cast<const uint2*>(ybin);
    const int4 o0 = *reinterpret_cast<const int4*>(orig), o1 = *reinterpret_cast<const int4*>(orig + 4);
    const int32_t og[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
        if (og[o] < 0) { y[o] = 0.0; continue; }                // padding: y = 0, residual 0
        double ua, ub;
        rng_u2(rng, sweep, TREE_SWEEP_LEVEL, PUR_LATENT, (uint32_t)og[o], ua, ub);
        const int yo = (int)(((o < 4 ? yb.x : yb.y) >> (8 * (o & 3))) & 0xFFu);
        const double ys = probit_latent_d(offset + ms * gc[o], yo, ua) - offset;
        if (!(ys == ys)) *bad = 1;
        y[o] = ys;
    }
    st8_f64(yout, y);
}

// PROBIT: probit BART (the instantiation that draws the latent in its finishing pass; the BCF one carries none of that code)
template <bool PROBIT>
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_pipe(Dev d, int nsweeps, unsigned int* bar, int ntl_max, int* done_out)
{
    extern __shared__ __align__(16) unsigned char bcf_dyn_smem[];
    if (__ldcg(&done_out[2]) != 0) return;            // queued behind a launch that handed its sweep to the batched kernel: the host reruns from there
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(bcf_dyn_smem);
    unsigned char* dyn = bcf_dyn_smem + PIPE_SMEM_BYTES;
    __shared__ Scalars scs;
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    // role of a warp.  setmaxnreg is a WARPGROUP instruction (4 consecutive warps must ask for the same count), so the
    // chain warps are one aligned group of four.
#if PIPE_CHAIN_FIRST
    const bool is_pass = warp >= PR_WARPS;
    const int pw = warp - PR_WARPS, rt = t;                 // pass warp index (0: the utility warp), chain thread index
    const bool is_util = pw == 0;
    const int tw = pw - 1;                                   // tile warp index
#else
    const bool is_pass = warp < PP_WARPS;
    const int pw = warp, rt = t - PP_THREADS;
    const bool is_util = pw == PT_WARPS;
    const int tw = pw;
#endif
    const int pt = pw * 32 + lane;
    const bool writer0 = (blockIdx.x == 0);
    // Grid barrier counters (every thread counts the events itself).  bar[2]: the full barriers.  bar[0] / bar[1]: the
    // batch barriers, by batch parity -- the pass role may arrive at batch b+1 before batch b's barrier has completed
    // (it only waits for b-1's decisions), so consecutive batches must not share a counter; it cannot reach b+2
    // before b has completed, so two counters are enough.
    unsigned int target = 0, tbatch[2] = {0u, 0u};
    unsigned long long gbatch = d.pr.batch0, mev = 0;   // observation shards: global batch number (tags, slots) / moment exchanges of this launch
    unsigned int ubatch[2] = {0u, 0u};                  // the utility warp's own count of the batch barriers ...
    unsigned long long ugbatch = d.pr.batch0;           // ... and of the global batch number
    const bool sharded = d.pr.nranks > 1;
    int bad = 0;
    PipePtrs rs;
    rs.R = (uint32_t)__cvta_generic_to_shared(dyn);
    rs.G = rs.R + (uint32_t)ntl_max * (TILE * 8u);
    rs.keys = rs.G + (uint32_t)ntl_max * (TILE * 8u);
    TMeta* meta = reinterpret_cast<TMeta*>(dyn + (size_t)ntl_max * PIPE_BYTES_PER_TILE);
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;
    const int tmod = d.f[1].tree0;
    const bool has_mod = d.f[1].ntree > 0;
    const int tile0 = (int)((int64_t)blockIdx.x * d.n_tiles / gridDim.x);
    const int tile1 = (int)((int64_t)(blockIdx.x + 1) * d.n_tiles / gridDim.x);
    const int ntl = tile1 - tile0;
    const int pad_a = d.trt_tiles - 1 - tile0, pad_b = d.n_tiles - 1 - tile0;
    auto grid_all = [&]() {   // full grid barrier, every thread of the CTA
        target += gridDim.x;
        __syncthreads();
        if (t == 0) red_release_add_u32(bar + 2, 1u);
    };
    auto grid_all_wait = [&]() -> bool { return grid_wait(bar + 2, target, d.err, &bar_ok); };

    if (is_pass)
        for (int lt = pw; lt < ntl; lt += PP_WARPS) {
            const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
            const uint32_t toff = (uint32_t)lt * (TILE * 8u);
            long long R[8];
            double gc[8];
            ld8_i64(d.rfx + gb, R); ld8_f64(d.Gc + gb, gc);
            res_st8i(rs.R + toff, lane, R);
            res_st8(rs.G + toff, lane, gc);
        }
    __syncthreads();
    if (d.dbg != nullptr && t == 0 && blockIdx.x == 0) ((volatile int*)d.dbg)[8] += 1;      // launches of this kernel that started running
    pipe_propose_all(sm, d, d.trees[scs.parity], scs.sweep);
    grid_all();
    if (!grid_all_wait()) return;
    if (d.dbg != nullptr && t == 0 && blockIdx.x == 0) ((volatile int*)d.dbg)[9] += 1;      // ... and got through their first grid barrier

    int done = 0;
    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        long long* totals = sweep_totals(d, sc.sweep);
        long long* cross = d.cross + (size_t)(sc.sweep & 1u) * (size_t)T * PIPE_XSTRIDE;
        // ---- what the pass needs to know about every tree; a tree too wide for this engine ends the launch ----
        if (t == 0) { sm.anybig = 0; sm.plan_ready = 0; sm.flushed = 0; }
        __syncthreads();
        for (int j = t; j < T; j += PIPE_THREADS) {
            const TreeRec* tr = &d.trees[par][j];
            const Proposal* P = &d.proprec[j].P;
            const int L = __ldcg(&tr->nleaves), mv = __ldcg(&P->move);
            TMeta m;
            m.L = (uint8_t)L; m.birth = mv == 1; m.K = (uint8_t)min(255, L + (mv == 1 ? 1 : 0));
            m.sx = (uint8_t)__ldcg(&P->slot); m.cut = (uint16_t)__ldcg(&P->cut); m.var = (uint16_t)__ldcg(&P->var);
            meta[j] = m;
            if (m.K > KB || L > KB) sm.anybig = 1;
        }
        __syncthreads();
        if (sm.anybig) break;

        // The roles need different register budgets (the decision chain keeps a tree's sums, posterior terms and
        // constants live; the pass is a streaming loop): re-split the SM's register file for the role section.
#ifndef BCF_PROF
#define PTRACE(k) do { } while (0)
#else
#define PTRACE(k) do { if (prof && sw == 0 && b < 64) d.prof[16 + b * 8 + (k)] = clock64(); } while (0)
#endif
        if (is_pass) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PIPE_REGS_PASS));
            if (!is_util) {
                // =================== tile warps ===================
                [[maybe_unused]] const bool prof = d.prof != nullptr && blockIdx.x == 0 && tw == 0 && lane == 0;
                [[maybe_unused]] const bool cprof = d.prof != nullptr && tw == 0 && lane == 0;
                [[maybe_unused]] long long tp = BCF_PROF_CLOCK();
#ifndef BCF_PROF
#define PPROF(k) do { } while (0)
#else
#define PPROF(k) do { if (cprof) { const long long now_ = clock64(); if (prof) atomicAdd((unsigned long long*)&d.prof[k], (unsigned long long)(now_ - tp)); atomicAdd((unsigned long long*)&d.prof[1024 + blockIdx.x * 8 + (k)], (unsigned long long)(now_ - tp)); tp = now_; } } while (0)
#endif
                int first = 0, b = 0;
                if (lane == 0) pipe_wait_flag(&sm.plan_ready, 1, d.err, 20, d.dbg, 1);     // the first plan
                __syncwarp();
                while (first < T) {
                    const PipePlan& P = sm.plan[b & 3];
                    PPROF(5); PTRACE(0);
                    for (int lt = tw; lt < ntl; lt += PT_WARPS) pipe_stats_tile(sm, sm.p.acc[b & 1], P, d, rs, tile0, lt, lane, tw, b & 1, ntl <= PT_WARPS);
                    // the next plan is made two batches ahead: warp 0 checks it (an acquire load: kept off the other
                    // warps, and away from the apply's stores) and the barrier hands the knowledge to everybody
                    if (tw == 0 && lane == 0 && first + P.nb < T) pipe_wait_flag(&sm.plan_ready, b + 2, d.err, 20, d.dbg, 2);
                    named_sync(BAR_S, PT_THREADS);                               // every tile warp is through STATS(b)
                    PPROF(0); PTRACE(1);
                    pipe_flush(sm.p.acc[b & 1], P, totals, cross, tw * 32 + lane);
                    named_sync(BAR_F, PT_THREADS);                               // ... and has flushed its share
                    if (tw == 0) {
                        if (lane == 0) { red_release_add_u32(bar + (b & 1), 1u); st_release_cta_s32(&sm.flushed, b + 1); }   // grid barrier b
                        named_arrive(BAR_H1 + (b & 1), H1_THREADS);              // H1(b): flushed (planned, staged: see plan_ready)
                    }
                    PPROF(1); PTRACE(2);
                    if (b > 0) {
                        named_sync(BAR_H2 + ((b - 1) & 1), H2_THREADS);         // H2(b-1): its tables are ready
                        PPROF(2); PTRACE(3);
                        const PipeApply& ap = sm.ap[(b - 1) & 1];
                        const bool fswitch = has_mod && ap.first == tmod && tmod > 0;
                        for (int lt = tw; lt < ntl; lt += PT_WARPS) {
                            if (lt == pad_a || lt == pad_b) pipe_apply_tile<true>(ap, d, rs, tile0, lt, lane, (b - 1) & 1, fswitch);
                            else pipe_apply_tile<false>(ap, d, rs, tile0, lt, lane, (b - 1) & 1, fswitch);
                        }
                        PPROF(3); PTRACE(4);
                    }
                    first += P.nb; ++b;
                }
#undef PPROF
                {   // drain: the last batch
                    named_sync(BAR_H2 + ((b - 1) & 1), H2_THREADS);
                    const PipeApply& ap = sm.ap[(b - 1) & 1];
                    const bool fswitch = has_mod && ap.first == tmod && tmod > 0;
                    for (int lt = tw; lt < ntl; lt += PT_WARPS) {
                        if (lt == pad_a || lt == pad_b) pipe_apply_tile<true>(ap, d, rs, tile0, lt, lane, (b - 1) & 1, fswitch);
                        else pipe_apply_tile<false>(ap, d, rs, tile0, lt, lane, (b - 1) & 1, fswitch);
                    }
                }
            } else {
                // =================== utility warp ===================
                // Plans the batches and stages their trees, two batches ahead of the statistics and paced only by the
                // `flushed` counter: nothing ever waits for this warp in the steady state (its shared-memory round
                // trips are slow while the tile warps stream).  Buffers of batch pb are those of batch pb - 4, whose
                // decisions every tile warp has seen applied once batch pb - 2 is flushed.
                {
                    unsigned int* tb = reinterpret_cast<unsigned int*>(&sm.p.acc[0]);
                    for (int i = lane; i < (int)(2 * sizeof(PipeAcc) / 4); i += 32) tb[i] = 0u;
                }
                int pfirst = 0, pb = 0;
                // CTA c (c < nranks - 1) pushes to the c-th peer of this rank
                const bool pusher = sharded && (int)blockIdx.x < d.pr.nranks - 1;
                const int push_peer = (int)blockIdx.x < d.pr.rank ? (int)blockIdx.x : (int)blockIdx.x + 1;
                auto push_batch = [&](int ub) {
                    ubatch[ub & 1] += gridDim.x;
                    ++ugbatch;
                    if (lane == 0) {
                        const long long t0 = clock64();
                        while (ld_acquire_u32(bar + (ub & 1)) < ubatch[ub & 1]) {
                            __nanosleep(100);
                            if (clock64() - t0 > 8000000000ll) { atomicOr(d.err, ERR_BARRIER_TIMEOUT); give_up(d.dbg, 30, ub, (int)ubatch[ub & 1]); }
                        }
                    }
                    __syncwarp();
                    pipe_peer_push(sm.plan[ub & 3], d, totals, cross, push_peer, ugbatch, lane);
                };
                [[maybe_unused]] long long up = BCF_PROF_CLOCK();
                [[maybe_unused]] const bool uprof = d.prof != nullptr && lane == 0;
#ifndef BCF_PROF
#define UPROF(k) do { } while (0)
#else
#define UPROF(k) do { if (uprof) { const long long now_ = clock64(); atomicAdd((unsigned long long*)&d.prof[1024 + 2048 + blockIdx.x * 8 + (k)], (unsigned long long)(now_ - up)); up = now_; } } while (0)
#endif
                while (pfirst < T) {
                    if (pb >= 2) {
                        if (lane == 0) pipe_wait_flag(&sm.flushed, pb - 1, d.err, pusher ? 40 : 200, d.dbg, 3);
                        __syncwarp();
                        // observation shards: batch pb - 2 is flushed by this CTA; once every CTA of this rank has flushed
                        // it, its reduced words go to the peer this CTA serves -- first thing, before the planning: the
                        // peers' chain warps are about to look for them
                        if (pusher) push_batch(pb - 2);
                    }
                    UPROF(0);
                    PipePlan& NP = sm.plan[pb & 3];
                    pipe_plan(NP, pb > 0 ? &sm.plan[(pb - 1) & 3] : nullptr, d, meta, pfirst, lane);
                    UPROF(1);
                    pipe_stage_trees(sm.st[pb & 3], d, d.trees[par], pfirst, lane);
                    __syncwarp();
                    UPROF(2);
                    if (lane == 0) st_release_cta_s32(&sm.plan_ready, pb + 1);
                    pfirst += NP.nb; ++pb;
                }
                if (pusher)
                    for (int ub = max(pb - 2, 0); ub < pb; ++ub) {               // the sweep's last two batches
                        if (lane == 0) pipe_wait_flag(&sm.flushed, ub + 1, d.err, 40, d.dbg, 4);
                        __syncwarp();
                        push_batch(ub);
                    }
#undef UPROF
            }
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PIPE_REGS_BASE));
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PIPE_REGS_RESOLVE));
            // =================== chain warps ===================
            [[maybe_unused]] const bool prof = d.prof != nullptr && blockIdx.x == 0 && rt == 0;
            [[maybe_unused]] const bool cprof = d.prof != nullptr && rt == 0;
            [[maybe_unused]] long long tp = BCF_PROF_CLOCK();
#ifndef BCF_PROF
#define QPROF(k) do { } while (0)
#else
#define QPROF(k) do { if (cprof) { const long long now_ = clock64(); if (prof) atomicAdd((unsigned long long*)&d.prof[k], (unsigned long long)(now_ - tp)); atomicAdd((unsigned long long*)&d.prof[1024 + blockIdx.x * 8 + (k)], (unsigned long long)(now_ - tp)); tp = now_; } } while (0)
#endif
            if (rt < 2) {   // the sweep's scalars as the decisions use them, per forest (the expressions of resolve_batch)
                const bool is_con = d.f[rt].is_con != 0;
                const double s1 = is_con ? sc.mscale : sc.bscale1, s0 = is_con ? sc.mscale : sc.bscale0;
                const double tau = is_con ? sc.tau_con : sc.tau_mod;
                const double s2 = sc.sigma * sc.sigma;
                PipeConsts c;
                c.cc = ChainConsts{s1, s0, 1.0 / s1, 1.0 / s0, s1 * s1 / s2, s0 * s0 / s2, log(tau)};
                c.a = 1.0 / (tau * tau);
                sm.cons[rt] = c;
            }
            named_sync(BAR_RES, PR_THREADS);
            int first = 0, b = 0;
            while (first < T) {
                named_sync(BAR_H1 + (b & 1), H1_THREADS);                        // H1(b)
                QPROF(4); PTRACE(5);
                const PipePlan& P = sm.plan[b & 3];
                tbatch[b & 1] += gridDim.x;
                ++gbatch;                                                        // global number of the batch (observation shards: tag and slot)
                if (rt == 0) {
                    const long long t0 = clock64();
                    while (ld_acquire_u32(bar + (b & 1)) < tbatch[b & 1]) {
                        __nanosleep(40);   // the pass warps of this SM need the issue slots
                        if (clock64() - t0 > 8000000000ll) { atomicOr(d.err, ERR_BARRIER_TIMEOUT); __trap(); }   // (no note: see pipe_wait_flag)
                    }
                }
                named_sync(BAR_RES, PR_THREADS);
                QPROF(6); PTRACE(6);
                pipe_resolve(sm, sm.st[b & 3], P, b, d, sc, d.trees[par ^ 1], totals, cross, rt, sharded ? gbatch : 0ull,
                             (d.prof != nullptr && blockIdx.x == 0 && sw == 0 && b >= 8 && b < 16) ? d.prof + 16 + 512 + (b - 8) * 32 : nullptr);
                named_arrive(BAR_H2 + (b & 1), H2_THREADS);                      // H2(b)
                QPROF(7); PTRACE(7);
                first += P.nb; ++b;
            }
#undef QPROF
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PIPE_REGS_BASE));
        }
#undef PTRACE
        __syncthreads();
#ifdef BCF_PROF
        long long te = clock64();
#define EPROF(k) do { if (d.prof != nullptr && t == 0) { const long long now_ = clock64(); atomicAdd((unsigned long long*)&d.prof[1024 + 2048 + blockIdx.x * 8 + (k)], (unsigned long long)(now_ - te)); te = now_; } } while (0)
#else
#define EPROF(k) do { } while (0)
#endif
        // ---- epilogue: moments of (y, Gc, Gm) ----
        if (is_pass) {
            long long* tb = reinterpret_cast<long long*>(&sm.p.tab);
            for (int i = pt; i < (int)(sizeof(sm.p.tab) / 8); i += PP_THREADS) tb[i] = 0;
        }
        __syncthreads();
        if (is_pass)
            for (int lt = pw; lt < ntl; lt += PP_WARPS) {
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                double G[8], y[8], other[8];
                res_ld8(rs.G + (uint32_t)lt * (TILE * 8u), lane, G);
                ld8_f64(d.y + gb, y);
                if (has_mod) { ld8_f64(d.Gc + gb, other); moments8(sm.p.tab, y, other, G, tile < d.trt_tiles ? 1 : 0, lane, bad); }
                else {
#pragma unroll
                    for (int o = 0; o < 8; ++o) other[o] = 0.0;
                    moments8(sm.p.tab, y, G, other, tile < d.trt_tiles ? 1 : 0, lane, bad);
                }
            }
        __syncthreads();
        for (int i = t; i < NMOM * 8; i += PIPE_THREADS) {
            const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
            if (q == 0) continue;
            const long long v = table_word(sm.p.tab, PIPE_THREADS / 32, m, g, q);
            if (v != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)v);
        }
        EPROF(3);
        grid_all();
        // the next sweep's proposals need only the new trees (each is proposed by the CTA that wrote it) and the Philox
        // counters: they are made while the moments cross the grid barrier
        if (sw + 1 < nsweeps) pipe_propose_all(sm, d, d.trees[par ^ 1], sc.sweep + 1u);
        EPROF(6);
        if (!grid_all_wait()) return;
        EPROF(4);
        if (sharded) {   // the sweep's moments, summed over the shards (as in k_batched)
            ++mev;
            if (blockIdx.x == 0) {
                for (int i = t; i < MAIL_MOM_WORDS; i += PIPE_THREADS) {
                    const long long v = __ldcg(&mom[i]);
                    for (int p = 0; p < d.pr.nranks; ++p)
                        if (p != d.pr.rank) d.pr.peer[p][d.pr.rank].mom[i] = v;
                }
                peer_release(d, MAIL_FLAG_MOM, d.pr.seq0 + mev);
            }
            if (!peer_wait(d, MAIL_FLAG_MOM, d.pr.seq0 + mev, &bar_ok)) return;
            long long* momsum = &sm.rs.tot[0][0][0][0];      // (the resolve staging is idle between the role sections)
            for (int i = t; i < MAIL_MOM_WORDS; i += PIPE_THREADS) {
                long long v = __ldcg(&mom[i]);
                for (int p = 0; p < d.pr.nranks; ++p)
                    if (p != d.pr.rank) v += __ldcg(&d.pr.mine[p].mom[i]);
                momsum[i] = v;
            }
            __syncthreads();
            sweep_scalars(d, d.trees[par ^ 1], momsum, sm.p.red, &scs, writer0, true);
        } else
            sweep_scalars(d, d.trees[par ^ 1], mom, sm.p.red, &scs, writer0);
        EPROF(5);
        grid_all();   // proposals visible before the next sweep reads them (wait: below)
        // ---- finishing pass ----
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            if (is_pass)
                for (int lt = pw; lt < ntl; lt += PP_WARPS) {
                    const int tile = tile0 + lt;
                    const int64_t gb = (int64_t)tile * TILE + lane * 8;
                    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
                    const double bb = tile < d.trt_tiles ? b1 : b0;
                    double gc[8], gm[8], y[8];
                    long long R[8];
                    ld8_f64(d.y + gb, y);
                    if (has_mod) { res_ld8(rs.G + toff, lane, gm); ld8_f64(d.Gc + gb, gc); st8_f64(d.Gm + gb, gm); res_st8(rs.G + toff, lane, gc); }
                    else {
                        res_ld8(rs.G + toff, lane, gc);
#pragma unroll
                        for (int o = 0; o < 8; ++o) gm[o] = 0.0;
                    }
                    if (PROBIT)                                                                  // (one forest: gc is the fit)
                        pipe_probit_latent8(Rng{d.key0, d.key1}, d.ybin + gb, d.orig + gb, const_cast<double*>(d.y) + gb, d.probit_offset, sn.sweep, ms, gc, y, &bad);
#pragma unroll
                    for (int o = 0; o < 8; ++o) R[o] = resync_fx(y[o], ms, gc[o], bb, gm[o], bad);
                    res_st8i(rs.R + toff, lane, R);
                    if (row >= 0) {
#pragma unroll
                        for (int o = 0; o < 8; ++o) {
                            const int64_t i = gb + o;
                            const double mu = ms * gc[o], yh = mu + bb * gm[o], tau = (b1 - b0) * gm[o];
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
            const int64_t nx = (int64_t)T * PIPE_XSTRIDE;
            for (int64_t i = gtid; i < nx; i += gstride) cross[i] = 0;
            long long* mom_next = d.mom + (size_t)((sc.sweep + 1u) & 1u) * (2 * NMOM * 3);
            for (int64_t i = gtid; i < 2 * NMOM * 3; i += gstride) mom_next[i] = 0;
        }
        if (!grid_all_wait()) return;
        EPROF(7);
        ++done;
    }
#undef EPROF
    // exit: park the state in HBM in the form every engine understands (Gm was stored by the finishing pass)
    if (is_pass)
        for (int lt = pw; lt < ntl; lt += PP_WARPS) {
            const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
            const uint32_t toff = (uint32_t)lt * (TILE * 8u);
            double G[8];
            long long R[8];
            res_ld8(rs.G + toff, lane, G);
            res_ld8i(rs.R + toff, lane, R);
            st8_f64(d.Gc + gb, G);
            st8_i64(d.rfx + gb, R);
        }
    if (writer0 && t == 0) {
        done_out[0] += done;                          // (the host zeroes it per request; sweep-by-sweep callers queue several launches)
        if (done < nsweeps) done_out[2] = 1;          // a sweep was left to the batched kernel: later queued work must not run ahead of it
    }
    if (writer0 && !is_pass && rt == 0) done_out[1] = (int)(gbatch - d.pr.batch0);   // batches exchanged (observation shards: tags go on)
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
    if (t == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

}  // namespace bcf
