// bcf_persistent.cuh -- PERSISTENT engines: a batch of Gibbs sweeps in ONE cooperative kernel.
//
// One CTA per SM; every CTA owns a fixed, contiguous range of observation tiles for the whole launch, so the
// residual, the leaf-slot cache and the forest fits of a tile are only ever touched by their owner (no cross-CTA
// hazards on observation data).  Per tree there is exactly one grid-wide dependency -- the reduced leaf
// statistics -- carried by integer REDs into L2 plus one split (arrive / wait) grid barrier.  Between arrive and
// wait each CTA stages tree j+1: loads its structure, makes its proposal (every data-independent piece of the MH
// step), and starts the cp.async prefetch of its leaf slots and of the proposed variable's bin column.  After the
// wait every CTA replays the decision and the leaf draws from the same bits, so nothing is broadcast.
// Per sweep: T + 1 barriers (T trees + the moments of the scale / sigma draws).
//
//   k_resident   : the observation state (fixed-point residual R, current forest's fit G) lives in SHARED MEMORY
//                  for the whole launch; per tree the only HBM traffic is 1 B/obs of leaf slots + the proposed
//                  variable's bins (+ the slot write-back of accepted moves).  Used when the CTA's tiles fit.
//   k_persistent : same schedule with the state in HBM/L2 (any n).
#pragma once
#include "bcf_kernels.cuh"

namespace bcf {

constexpr int PTHREADS = 512;
constexpr int PERSIST_CTAS_PER_SM = 1;
constexpr int RES_BYTES_PER_TILE = TILE * (8 + 8) + 2 * TILE + 2 * 2 * TILE;   // R, G, 2 slot stages, 2 bin stages (u16)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Split grid barrier on a monotonically increasing arrival counter (zeroed by the host before each launch).
// arrive: after the CTA's REDs (bar.sync orders them before thread 0's release).  wait: bounded spin -- if a CTA
// never arrives the kernel flags ERR_BARRIER_TIMEOUT and leaves instead of hanging the GPU.
__device__ __forceinline__ void grid_arrive(unsigned int* bar, unsigned int& target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        red_release_add_u32(bar, 1u);
    }
}
__device__ __forceinline__ bool grid_wait(unsigned int* bar, unsigned int target, int32_t* err, int* ok_smem)
{
    if (threadIdx.x == 0) {
        int good = 1;
        const long long t0 = clock64();
        while (ld_acquire_u32(bar) < target) {
            if (clock64() - t0 > 4000000000ll) { good = 0; atomicOr(err, ERR_BARRIER_TIMEOUT); break; }
        }
        if (good && (ld_acquire_u32((const unsigned int*)err) & ERR_BARRIER_TIMEOUT)) good = 0;
        *ok_smem = good;
    }
    __syncthreads();
    return *ok_smem != 0;
}

// ---- peer mailboxes (observation shards): see PeerDev ----
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// after this rank's words are in every peer's mailbox (all threads of the pushing CTA, then a __syncthreads)
__device__ __forceinline__ void peer_release(const Dev& d, int kind, unsigned long long seq)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < d.pr.nranks && (int)threadIdx.x != d.pr.rank)
        st_release_sys_u64(&d.pr.peer[threadIdx.x][d.pr.rank].flag[kind], seq);
}
// every CTA: wait until every peer's words of sequence number `seq` have arrived in this rank's mailboxes
__device__ __forceinline__ bool peer_wait(const Dev& d, int kind, unsigned long long seq, int* ok_smem)
{
    if (threadIdx.x == 0) {
        int good = 1;
        const long long t0 = clock64();
        for (int p = 0; p < d.pr.nranks && good; ++p) {
            if (p == d.pr.rank) continue;
            while (ld_acquire_sys_u64(&d.pr.mine[p].flag[kind]) < seq)
                if (clock64() - t0 > 120000000000ll) { good = 0; atomicOr(d.err, ERR_BARRIER_TIMEOUT); break; }
        }
        *ok_smem = good;
    }
    __syncthreads();
    return *ok_smem != 0;
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

// =================================================================================================
// k_persistent: state in HBM/L2
// =================================================================================================
__global__ void __launch_bounds__(PTHREADS, 1) k_persistent(Dev d, int nsweeps, unsigned int* bar)
{
    BCF_STEP_SMEM(sm);
    __shared__ Scalars scs;
    __shared__ double red[2 * PTHREADS];
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const bool writer0 = (blockIdx.x == 0);
    const int lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    unsigned int target = 0;
    int bad = 0;
    BCF_PROF_DECL
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;
    // contiguous tile range of this CTA
    const int tile0 = (int)((int64_t)blockIdx.x * d.n_tiles / gridDim.x);
    const int tile1 = (int)((int64_t)(blockIdx.x + 1) * d.n_tiles / gridDim.x);

    // proposals of the first sweep of this launch (later sweeps: made in the previous sweep's epilogue)
    propose_range(sm, d, d.trees[scs.parity], scs.sweep, blockIdx.x, gridDim.x, -1, 0);
    grid_arrive(bar, target);
    if (!grid_wait(bar, target, d.err, &bar_ok)) return;

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        long long* totals = sweep_totals(d, sc.sweep);
        if (t == 0) sm.ai.active = 0;
        load_tree(&sm.st[0].tr, &d.trees[par][0]);
        __syncthreads();
        fetch_proposal(sm, 0, d, 0, sm.st[0].tr.nleaves);
        __syncthreads();
        begin_stats(sm, 0, 0, sc);
        __syncthreads();
        for (int j = 0; j < T; ++j) {
            const int s = j & 1;
            const int fj = j >= d.f[1].tree0 ? 1 : 0;
            const bool writer = ((int)blockIdx.x == j % (int)gridDim.x);   // rotate the tree write-back
            for (int tile = tile0 + wib; tile < tile1; tile += wpb)
                step_tile(sm, sm.st[s].ci, d, j - 1, j, tile, lane);
            __syncthreads();
            PROF(0);
            flush_stats(sm, sm.st[s].ci.K, totals, j);
            grid_arrive(bar, target);
            PROF(1);
            if (j + 1 < T) {   // stage tree j+1 while the barrier is in flight
                const int fn = (j + 1) >= d.f[1].tree0 ? 1 : 0;
                load_tree(&sm.st[s ^ 1].tr, &d.trees[par][j + 1]);
                __syncthreads();
                fetch_proposal(sm, s ^ 1, d, j + 1, sm.st[s ^ 1].tr.nleaves);
                __syncthreads();
                begin_stats(sm, s ^ 1, fn, sc);
            }
            PROF(5);
            if (!grid_wait(bar, target, d.err, &bar_ok)) return;
            PROF(2);
            decide(sm, s, d, d.f[fj], j, sc, writer, totals);
            PROF(3);
            if (writer) store_tree(&d.trees[par ^ 1][j], &sm.st[s].tr);
            PROF(4);
        }
        // epilogue: next sweep's proposals, finish the last tree, moments
        if (sw + 1 < nsweeps) propose_range(sm, d, d.trees[par ^ 1], sc.sweep + 1u, blockIdx.x, gridDim.x, T - 1, (T - 1) & 1);
        moments_begin(sm);
        __syncthreads();
        for (int tile = tile0 + wib; tile < tile1; tile += wpb)
            last_tile(sm, d, T - 1, tile, lane, bad);
        __syncthreads();
        moments_flush(sm, mom);
        grid_arrive(bar, target);
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        sweep_scalars(d, d.trees[par ^ 1], mom, red, &scs, writer0);
        // finishing pass over this CTA's own tiles; clear next sweep's accumulators (nobody reads them now)
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            for (int tile = tile0 + wib; tile < tile1; tile += wpb) {
                const int64_t base = (int64_t)tile * TILE + lane * 8;
                const double b = tile < d.trt_tiles ? b1 : b0;
                double gc[8], gm[8], y[8];
                long long R[8];
                ld8_f64(d.Gc + base, gc); ld8_f64(d.Gm + base, gm); ld8_f64(d.y + base, y);
#pragma unroll
                for (int o = 0; o < 8; ++o) R[o] = resync_fx(y[o], ms, gc[o], b, gm[o], bad);
                st8_i64(d.rfx + base, R);
                if (row >= 0) {
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int64_t i = base + o;
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
            long long* mom_next = d.mom + (size_t)((sc.sweep + 1u) & 1u) * (2 * NMOM * 3);
            for (int64_t i = gtid; i < 2 * NMOM * 3; i += gstride) mom_next[i] = 0;
        }
        __syncthreads();
        PROF(6);
    }
    if (prof) for (int k = 0; k < 8; ++k) d.prof[k] += pc[k];
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
    if (t == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

// =================================================================================================
// k_resident: state in shared memory
// =================================================================================================
// Shared layout of the observation state: per tile of 256 observations the doubles are stored TRANSPOSED,
// element k of lane l at [tile*256 + k*32 + l] (conflict-free LDS.64 with 8 observations per lane); the byte
// arrays (leaf slots, bins) keep their natural order (lane l owns bytes l*8 .. l*8+7: one LDS.64).
// All accesses go through 32-bit shared-window addresses and ld/st.shared (pointers kept in a struct decay to
// generic LD/ST, which take the slower global-memory path).
struct ResSmem {
    uint32_t R;          // fixed-point full residual (int64)
    uint32_t G;          // the unscaled fit of the forest being swept (sum of its trees' leaf values)
    uint32_t slot[2];    // leaf slots of the tree in flight, two stages
    uint32_t bins[2];    // bin column of the proposed variable (uint8 or uint16), two stages
};

__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// doubles of one tile: element k of lane l at tile_base + (k*32 + l)*8
__device__ __forceinline__ void res_ld8(uint32_t tile_base, int lane, double (&out)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = lds_f64(tile_base + (uint32_t)(k * 32 + lane) * 8u);
}
__device__ __forceinline__ void res_st8(uint32_t tile_base, int lane, const double (&in)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) sts_f64(tile_base + (uint32_t)(k * 32 + lane) * 8u, in[k]);
}
__device__ __forceinline__ void res_ld8i(uint32_t tile_base, int lane, long long (&out)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint2 v = lds_v2(tile_base + (uint32_t)(k * 32 + lane) * 8u);
        out[k] = (long long)(((unsigned long long)v.y << 32) | v.x);
    }
}
__device__ __forceinline__ void res_st8i(uint32_t tile_base, int lane, const long long (&in)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned long long u = (unsigned long long)in[k];
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tile_base + (uint32_t)(k * 32 + lane) * 8u), "r"((uint32_t)u), "r"((uint32_t)(u >> 32)) : "memory");
    }
}
__device__ __forceinline__ void lds8_u8(uint32_t a, int (&out)[8])
{
    const uint2 v = lds_v2(a);
    out[0] = v.x & 0xFF; out[1] = (v.x >> 8) & 0xFF; out[2] = (v.x >> 16) & 0xFF; out[3] = v.x >> 24;
    out[4] = v.y & 0xFF; out[5] = (v.y >> 8) & 0xFF; out[6] = (v.y >> 16) & 0xFF; out[7] = v.y >> 24;
}
__device__ __forceinline__ void lds8_bins(uint32_t stage, int is16, int lt, int lane, int (&out)[8])
{
    if (is16) {
        const uint4 v = lds_v4(stage + (uint32_t)(lt * TILE + lane * 8) * 2u);
        out[0] = v.x & 0xFFFF; out[1] = v.x >> 16; out[2] = v.y & 0xFFFF; out[3] = v.y >> 16;
        out[4] = v.z & 0xFFFF; out[5] = v.z >> 16; out[6] = v.w & 0xFFFF; out[7] = v.w >> 16;
    } else lds8_u8(stage + (uint32_t)(lt * TILE + lane * 8), out);
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// start the asynchronous copy of tree `tree`'s leaf slots (and of the proposed variable's bins) for this CTA's tiles
__device__ __forceinline__ void res_prefetch(const Dev& d, const ResSmem& rs, int stage, int tree, const Proposal& P,
                                             int tile0, int ntl)
{
    const int64_t o0 = (int64_t)tile0 * TILE;
    const uint8_t* src = d.slots + (int64_t)tree * d.n_pad + o0;
    const int nchunk = ntl * (TILE / 16);
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) cp_async16_s(rs.slot[stage] + (uint32_t)c * 16u, src + c * 16);
    if (P.move == 1) {
        const int w = P.is16 ? 2 : 1;
        const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(P.bins) + o0 * w;
        const int nb = nchunk * w;
        for (int c = threadIdx.x; c < nb; c += blockDim.x) cp_async16_s(rs.bins[stage] + (uint32_t)c * 16u, bsrc + c * 16);
    }
    cp_async_commit();
}

// tile pass with the state in shared memory: apply tree j-1 (exact integer update of the residual, G += change of
// the leaf value, slot write-back of accepted moves), then accumulate tree j's statistics
__device__ __forceinline__ void res_tile(StepShared& sm, const CurInfo& ci, const Dev& d, const ResSmem& rs, int prev,
                                         int ps, int cs, int tile0, int lt, int lane, bool do_apply, bool do_stats)
{
    const int tile = tile0 + lt;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
    long long R[8];
    res_ld8i(rs.R + toff, lane, R);
    if (do_apply) {
        int sp[8], bp[8];
        double G[8];
        res_ld8(rs.G + toff, lane, G);
        lds8_u8(rs.slot[ps] + (uint32_t)(lt * TILE + lane * 8), sp);
        if (ai.birth) lds8_bins(rs.bins[ps], ai.is16, lt, lane, bp);
        apply8(sm, ai, g, sp, bp, G, R);
        res_st8(rs.G + toff, lane, G);
        res_st8i(rs.R + toff, lane, R);
        if (ai.changed) st8_u8(d.slots + (int64_t)prev * d.n_pad + (int64_t)tile * TILE + lane * 8, sp);
    }
    if (do_stats) {
        int key[8];
        lds8_u8(rs.slot[cs] + (uint32_t)(lt * TILE + lane * 8), key);
        if (ci.birth) {
            int bc[8];
            lds8_bins(rs.bins[cs], ci.is16, lt, lane, bc);
#pragma unroll
            for (int o = 0; o < 8; ++o) if (key[o] == ci.sx && bc[o] > ci.cut) key[o] = ci.new_slot;
        }
        accum_tile(sm.tab, key, R, g, ci.K, ci.birth != 0, lane, (int)(threadIdx.x >> 5));
    }
}

__global__ void __launch_bounds__(PTHREADS, 1) k_resident(Dev d, int nsweeps, unsigned int* bar, int ntl_max)
{
    BCF_STEP_SMEM(sm);
    unsigned char* dyn = bcf_dyn_smem + STEP_SMEM_BYTES;   // observation state behind the step scratch
    __shared__ Scalars scs;
    __shared__ double red[2 * PTHREADS];
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const bool writer0 = (blockIdx.x == 0);
    const int lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    unsigned int target = 0;
    int bad = 0;
    BCF_PROF_DECL
    ResSmem rs;
    rs.R = (uint32_t)__cvta_generic_to_shared(dyn);
    rs.G = rs.R + (uint32_t)ntl_max * (TILE * 8u);
    rs.slot[0] = rs.G + (uint32_t)ntl_max * (TILE * 8u);
    rs.slot[1] = rs.slot[0] + (uint32_t)ntl_max * TILE;
    rs.bins[0] = rs.slot[1] + (uint32_t)ntl_max * TILE;
    rs.bins[1] = rs.bins[0] + (uint32_t)ntl_max * TILE * 2u;
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;
    const int tmod = d.f[1].tree0;                 // first tau tree
    const bool has_mod = d.f[1].ntree > 0;
    const int tile0 = (int)((int64_t)blockIdx.x * d.n_tiles / gridDim.x);
    const int tile1 = (int)((int64_t)(blockIdx.x + 1) * d.n_tiles / gridDim.x);
    const int ntl = tile1 - tile0;

    // entry: the residual and the mu forest's fit move into shared memory
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
    propose_range(sm, d, d.trees[scs.parity], scs.sweep, blockIdx.x, gridDim.x, -1, 0);
    grid_arrive(bar, target);
    if (!grid_wait(bar, target, d.err, &bar_ok)) return;

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        long long* totals = sweep_totals(d, sc.sweep);
        if (t == 0) sm.ai.active = 0;
        load_tree(&sm.st[0].tr, &d.trees[par][0]);
        __syncthreads();
        fetch_proposal(sm, 0, d, 0, sm.st[0].tr.nleaves);
        __syncthreads();
        begin_stats(sm, 0, 0, sc);
        res_prefetch(d, rs, 0, 0, sm.st[0].prop, tile0, ntl);
        cp_async_wait_all();
        __syncthreads();
        PROF(6);
        for (int j = 0; j < T; ++j) {
            const int s = j & 1;
            const int fj = j >= tmod ? 1 : 0;
            const bool writer = ((int)blockIdx.x == j % (int)gridDim.x);   // rotate the tree write-back
            bool do_apply = (j > 0);
            if (j == tmod && j > 0) {
                // forest switch: finish the last mu tree, park Gc in HBM, bring Gm in (the residual is untouched)
                for (int lt = wib; lt < ntl; lt += wpb) {
                    res_tile(sm, sm.st[s].ci, d, rs, j - 1, s ^ 1, s, tile0, lt, lane, true, false);
                    const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
                    const uint32_t toff = (uint32_t)lt * (TILE * 8u);
                    double G[8], gm[8];
                    res_ld8(rs.G + toff, lane, G);
                    ld8_f64(d.Gm + gb, gm);
                    st8_f64(d.Gc + gb, G);
                    res_st8(rs.G + toff, lane, gm);
                }
                do_apply = false;
            }
            for (int lt = wib; lt < ntl; lt += wpb)
                res_tile(sm, sm.st[s].ci, d, rs, j - 1, s ^ 1, s, tile0, lt, lane, do_apply, true);
            __syncthreads();
            PROF(0);
            flush_stats(sm, sm.st[s].ci.K, totals, j);
            grid_arrive(bar, target);
            PROF(1);
            if (j + 1 < T) {   // stage tree j+1 while the barrier is in flight
                const int fn = (j + 1) >= tmod ? 1 : 0;
                load_tree(&sm.st[s ^ 1].tr, &d.trees[par][j + 1]);
                __syncthreads();
                fetch_proposal(sm, s ^ 1, d, j + 1, sm.st[s ^ 1].tr.nleaves);
                __syncthreads();
                begin_stats(sm, s ^ 1, fn, sc);
                res_prefetch(d, rs, s ^ 1, j + 1, sm.st[s ^ 1].prop, tile0, ntl);
            }
            PROF(5);
            if (!grid_wait(bar, target, d.err, &bar_ok)) return;
            PROF(2);
            decide(sm, s, d, d.f[fj], j, sc, writer, totals);
            PROF(3);
            if (writer) store_tree(&d.trees[par ^ 1][j], &sm.st[s].tr);
            cp_async_wait_all();
            __syncthreads();
            PROF(4);
        }
        // ---- epilogue: next sweep's proposals, finish the last tree, moments of (y, Gc, Gm) ----
        if (sw + 1 < nsweeps) propose_range(sm, d, d.trees[par ^ 1], sc.sweep + 1u, blockIdx.x, gridDim.x, T - 1, (T - 1) & 1);
        moments_begin(sm);
        __syncthreads();
        {
            const int sl = (T - 1) & 1;
            for (int lt = wib; lt < ntl; lt += wpb) {
                res_tile(sm, sm.st[sl].ci, d, rs, T - 1, sl, sl, tile0, lt, lane, true, false);
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                double G[8], y[8], other[8];
                res_ld8(rs.G + (uint32_t)lt * (TILE * 8u), lane, G);
                ld8_f64(d.y + gb, y);
                if (has_mod) { ld8_f64(d.Gc + gb, other); moments8(sm.tab, y, other, G, tile < d.trt_tiles ? 1 : 0, lane, bad); }
                else {
#pragma unroll
                    for (int o = 0; o < 8; ++o) other[o] = 0.0;
                    moments8(sm.tab, y, G, other, tile < d.trt_tiles ? 1 : 0, lane, bad);
                }
            }
        }
        __syncthreads();
        moments_flush(sm, mom);
        grid_arrive(bar, target);
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        sweep_scalars(d, d.trees[par ^ 1], mom, red, &scs, writer0);
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
                if (has_mod) { res_ld8(rs.G + toff, lane, gm); ld8_f64(d.Gc + gb, gc); st8_f64(d.Gm + gb, gm); res_st8(rs.G + toff, lane, gc); }
                else {
                    res_ld8(rs.G + toff, lane, gc);
#pragma unroll
                    for (int o = 0; o < 8; ++o) gm[o] = 0.0;
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) R[o] = resync_fx(y[o], ms, gc[o], b, gm[o], bad);
                res_st8i(rs.R + toff, lane, R);
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
            long long* mom_next = d.mom + (size_t)((sc.sweep + 1u) & 1u) * (2 * NMOM * 3);
            for (int64_t i = gtid; i < 2 * NMOM * 3; i += gstride) mom_next[i] = 0;
        }
        __syncthreads();
        PROF(6);
    }
    // exit: park the state in HBM in the form every engine understands (Gm was stored by the finishing pass)
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
