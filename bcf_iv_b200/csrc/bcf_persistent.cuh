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
//   k_resident   : the observation state (base = y - other forest, G = this forest's fit) lives in SHARED MEMORY
//                  for the whole launch; per tree the only HBM traffic is 1 B/obs of leaf slots + the proposed
//                  variable's bins (+ the slot write-back of accepted moves).  Used when the CTA's tiles fit.
//   k_persistent : same schedule with the state in HBM/L2 (any n).
#pragma once
#include "bcf_kernels.cuh"

namespace bcf {

constexpr int PTHREADS = 512;
constexpr int PERSIST_CTAS_PER_SM = 1;
constexpr int RES_BYTES_PER_TILE = TILE * (8 + 8) + 2 * TILE + 2 * 2 * TILE;   // base, G, 2 slot stages, 2 bin stages (u16)

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

#define BCF_PROF_DECL                                                                                   \
    const bool prof = (d.prof != nullptr) && blockIdx.x == 0 && t == 0;                                 \
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};                                                         \
    long long tp = clock64();
#define PROF(k) do { if (prof) { const long long now_ = clock64(); pc[k] += now_ - tp; tp = now_; } } while (0)

// =================================================================================================
// k_persistent: state in HBM/L2
// =================================================================================================
__global__ void __launch_bounds__(PTHREADS, 1) k_persistent(Dev d, int nsweeps, unsigned int* bar)
{
    __shared__ StepShared sm;
    __shared__ Scalars scs;
    __shared__ double red[2 * PTHREADS];
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const bool writer = (blockIdx.x == 0);
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

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        if (t == 0) sm.ai.active = 0;
        load_tree(&sm.st[0].tr, &d.trees[par][0]);
        __syncthreads();
        propose(sm, 0, d, d.f[0], 0, sc.sweep);
        begin_stats(sm, 0, 0, sc);
        __syncthreads();
        for (int j = 0; j < T; ++j) {
            const int s = j & 1;
            const int fj = j >= d.f[1].tree0 ? 1 : 0;
            for (int tile = tile0 + wib; tile < tile1; tile += wpb)
                step_tile(sm, sm.st[s].ci, d, j - 1, j, tile, lane, bad);
            __syncthreads();
            PROF(0);
            flush_stats(sm, sm.st[s].ci.K, d, j);
            grid_arrive(bar, target);
            PROF(1);
            if (j + 1 < T) {   // stage tree j+1 while the barrier is in flight
                const int fn = (j + 1) >= d.f[1].tree0 ? 1 : 0;
                load_tree(&sm.st[s ^ 1].tr, &d.trees[par][j + 1]);
                __syncthreads();
                propose(sm, s ^ 1, d, d.f[fn], j + 1, sc.sweep);
                begin_stats(sm, s ^ 1, fn, sc);
            }
            PROF(5);
            if (!grid_wait(bar, target, d.err, &bar_ok)) return;
            PROF(2);
            decide(sm, s, d, d.f[fj], j, sc, writer);
            PROF(3);
            if (writer) store_tree(&d.trees[par ^ 1][j], &sm.st[s].tr);
            PROF(4);
        }
        // epilogue: finish the last tree, rebuild the fits from the trees, moments
        {
            long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) tb[i] = 0;
        }
        __syncthreads();
        for (int tile = tile0 + wib; tile < tile1; tile += wpb)
            refit_tile(sm, d, T - 1, d.trees[par ^ 1], tile, lane, bad);
        __syncthreads();
        {
            const long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) {
                const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
                if (q >= 1 && tb[i] != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)tb[i]);
            }
        }
        grid_arrive(bar, target);
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        sweep_scalars(d, d.trees[par ^ 1], mom, red, &scs, writer);
        // finishing pass over this CTA's own tiles; clear next sweep's accumulators (nobody reads them now)
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            for (int tile = tile0 + wib; tile < tile1; tile += wpb) {
                const int64_t base = (int64_t)tile * TILE + lane * 8;
                const double b = tile < d.trt_tiles ? b1 : b0;
                double gc[8], gm[8], y[8], r[8];
                ld8_f64(d.Gc + base, gc); ld8_f64(d.Gm + base, gm); ld8_f64(d.y + base, y);
#pragma unroll
                for (int o = 0; o < 8; ++o) r[o] = y[o] - (ms * gc[o] + b * gm[o]);
                st8_f64(d.r + base, r);
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
            for (int64_t i = gtid; i < ntot; i += gstride) d.totals[i] = 0;
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
struct ResSmem {
    double* base;        // y - (other forest's scaled fit), fixed while a forest is being swept
    double* G;           // this forest's unscaled fit (sum of its trees' leaf values)
    uint8_t* slot[2];    // leaf slots of the tree in flight, two stages
    uint8_t* bins[2];    // bin column of the proposed variable (uint8 or uint16), two stages
};

__device__ __forceinline__ void res_ld8(const double* s, int lane, double (&out)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = s[k * 32 + lane];
}
__device__ __forceinline__ void res_st8(double* s, int lane, const double (&in)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k * 32 + lane] = in[k];
}
__device__ __forceinline__ void lds8_u8(const uint8_t* p, int (&out)[8])
{
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    out[0] = v.x & 0xFF; out[1] = (v.x >> 8) & 0xFF; out[2] = (v.x >> 16) & 0xFF; out[3] = v.x >> 24;
    out[4] = v.y & 0xFF; out[5] = (v.y >> 8) & 0xFF; out[6] = (v.y >> 16) & 0xFF; out[7] = v.y >> 24;
}
__device__ __forceinline__ void lds8_bins(const uint8_t* stage, int is16, int lt, int lane, int (&out)[8])
{
    if (is16) ld8_u16(reinterpret_cast<const uint16_t*>(stage) + lt * TILE + lane * 8, out);
    else lds8_u8(stage + lt * TILE + lane * 8, out);
}

// start the asynchronous copy of tree `tree`'s leaf slots (and of the proposed variable's bins) for this CTA's tiles
__device__ __forceinline__ void res_prefetch(const Dev& d, const ResSmem& rs, int stage, int tree, const Proposal& P,
                                             int tile0, int ntl)
{
    const int64_t o0 = (int64_t)tile0 * TILE;
    const uint8_t* src = d.slots + (int64_t)tree * d.n_pad + o0;
    const int nchunk = ntl * (TILE / 16);
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) cp_async16(rs.slot[stage] + c * 16, src + c * 16);
    if (P.move == 1) {
        const int w = P.is16 ? 2 : 1;
        const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(P.bins) + o0 * w;
        const int nb = nchunk * w;
        for (int c = threadIdx.x; c < nb; c += blockDim.x) cp_async16(rs.bins[stage] + c * 16, bsrc + c * 16);
    }
    cp_async_commit();
}

// tile pass with the state in shared memory: apply tree j-1 (G += change of its leaf value, slot write-back of
// accepted moves), then accumulate tree j's statistics on r = base - scale * G
__device__ __forceinline__ void res_tile(StepShared& sm, const CurInfo& ci, const Dev& d, const ResSmem& rs, int prev,
                                         int ps, int cs, int tile0, int lt, int lane, bool do_apply, bool do_stats,
                                         int& bad)
{
    const int tile = tile0 + lt;
    const int g = tile < d.trt_tiles ? 1 : 0;
    const ApplyInfo& ai = sm.ai;
    double G[8], B[8];
    res_ld8(rs.G + lt * TILE, lane, G);
    if (do_stats) res_ld8(rs.base + lt * TILE, lane, B);
    if (do_apply) {
        int sp[8], bp[8];
        double dl[8];
        lds8_u8(rs.slot[ps] + lt * TILE + lane * 8, sp);
        if (ai.birth) lds8_bins(rs.bins[ps], ai.is16, lt, lane, bp);
        apply8(sm, ai, sp, bp, dl);
#pragma unroll
        for (int o = 0; o < 8; ++o) G[o] += dl[o];
        res_st8(rs.G + lt * TILE, lane, G);
        if (ai.changed) st8_u8(d.slots + (int64_t)prev * d.n_pad + (int64_t)tile * TILE + lane * 8, sp);
    }
    if (do_stats) {
        int key[8];
        double r[8];
        lds8_u8(rs.slot[cs] + lt * TILE + lane * 8, key);
        if (ci.birth) {
            int bc[8];
            lds8_bins(rs.bins[cs], ci.is16, lt, lane, bc);
#pragma unroll
            for (int o = 0; o < 8; ++o) if (key[o] == ci.sx && bc[o] > ci.cut) key[o] = ci.new_slot;
        }
        const double scale = g ? ci.s1 : ci.s0;
#pragma unroll
        for (int o = 0; o < 8; ++o) r[o] = B[o] - scale * G[o];
        accum_tile(sm.tab, key, r, g, ci.K, lane, bad);
    }
}

__global__ void __launch_bounds__(PTHREADS, 1) k_resident(Dev d, int nsweeps, unsigned int* bar, int ntl_max)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ StepShared sm;
    __shared__ Scalars scs;
    __shared__ double red[2 * PTHREADS];
    __shared__ int bar_ok;
    const int t = threadIdx.x;
    const bool writer = (blockIdx.x == 0);
    const int lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    unsigned int target = 0;
    int bad = 0;
    BCF_PROF_DECL
    ResSmem rs;
    rs.base = reinterpret_cast<double*>(dyn);
    rs.G = rs.base + (size_t)ntl_max * TILE;
    rs.slot[0] = reinterpret_cast<uint8_t*>(rs.G + (size_t)ntl_max * TILE);
    rs.slot[1] = rs.slot[0] + (size_t)ntl_max * TILE;
    rs.bins[0] = rs.slot[1] + (size_t)ntl_max * TILE;
    rs.bins[1] = rs.bins[0] + (size_t)ntl_max * TILE * 2;
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;
    const int tmod = d.f[1].tree0;                 // first tau tree
    const bool has_mod = d.f[1].ntree > 0;
    const int tile0 = (int)((int64_t)blockIdx.x * d.n_tiles / gridDim.x);
    const int tile1 = (int)((int64_t)(blockIdx.x + 1) * d.n_tiles / gridDim.x);
    const int ntl = tile1 - tile0;

    // entry: G = Gc, base = y - bscale_g * Gm
    {
        const Scalars sc = scs;
        for (int lt = wib; lt < ntl; lt += wpb) {
            const int tile = tile0 + lt;
            const int64_t gb = (int64_t)tile * TILE + lane * 8;
            const double b = tile < d.trt_tiles ? sc.bscale1 : sc.bscale0;
            double y[8], gc[8], gm[8], B[8];
            ld8_f64(d.y + gb, y); ld8_f64(d.Gc + gb, gc); ld8_f64(d.Gm + gb, gm);
#pragma unroll
            for (int o = 0; o < 8; ++o) B[o] = y[o] - b * gm[o];
            res_st8(rs.G + lt * TILE, lane, gc);
            res_st8(rs.base + lt * TILE, lane, B);
        }
    }
    __syncthreads();

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        if (t == 0) sm.ai.active = 0;
        load_tree(&sm.st[0].tr, &d.trees[par][0]);
        __syncthreads();
        propose(sm, 0, d, d.f[0], 0, sc.sweep);
        begin_stats(sm, 0, 0, sc);
        res_prefetch(d, rs, 0, 0, sm.st[0].prop, tile0, ntl);
        cp_async_wait_all();
        __syncthreads();
        PROF(6);
        for (int j = 0; j < T; ++j) {
            const int s = j & 1;
            const int fj = j >= tmod ? 1 : 0;
            bool do_apply = (j > 0);
            if (j == tmod && j > 0) {
                // forest switch: finish the last mu tree, park Gc in HBM, base = y - mscale * Gc, G = Gm
                for (int lt = wib; lt < ntl; lt += wpb) {
                    res_tile(sm, sm.st[s].ci, d, rs, j - 1, s ^ 1, s, tile0, lt, lane, true, false, bad);
                    const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
                    double G[8], y[8], gm[8], B[8];
                    res_ld8(rs.G + lt * TILE, lane, G);
                    ld8_f64(d.y + gb, y); ld8_f64(d.Gm + gb, gm);
                    st8_f64(d.Gc + gb, G);
#pragma unroll
                    for (int o = 0; o < 8; ++o) B[o] = y[o] - sc.mscale * G[o];
                    res_st8(rs.base + lt * TILE, lane, B);
                    res_st8(rs.G + lt * TILE, lane, gm);
                }
                do_apply = false;
            }
            for (int lt = wib; lt < ntl; lt += wpb)
                res_tile(sm, sm.st[s].ci, d, rs, j - 1, s ^ 1, s, tile0, lt, lane, do_apply, true, bad);
            __syncthreads();
            PROF(0);
            flush_stats(sm, sm.st[s].ci.K, d, j);
            grid_arrive(bar, target);
            PROF(1);
            if (j + 1 < T) {   // stage tree j+1 while the barrier is in flight
                const int fn = (j + 1) >= tmod ? 1 : 0;
                load_tree(&sm.st[s ^ 1].tr, &d.trees[par][j + 1]);
                __syncthreads();
                propose(sm, s ^ 1, d, d.f[fn], j + 1, sc.sweep);
                begin_stats(sm, s ^ 1, fn, sc);
                res_prefetch(d, rs, s ^ 1, j + 1, sm.st[s ^ 1].prop, tile0, ntl);
            }
            PROF(5);
            if (!grid_wait(bar, target, d.err, &bar_ok)) return;
            PROF(2);
            decide(sm, s, d, d.f[fj], j, sc, writer);
            PROF(3);
            if (writer) store_tree(&d.trees[par ^ 1][j], &sm.st[s].tr);
            cp_async_wait_all();
            __syncthreads();
            PROF(4);
        }
        // ---- epilogue: finish the last tree, moments of (y, Gc, Gm) ----
        {
            long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) tb[i] = 0;
        }
        __syncthreads();
        {
            const int sl = (T - 1) & 1;
            for (int lt = wib; lt < ntl; lt += wpb) {
                res_tile(sm, sm.st[sl].ci, d, rs, T - 1, sl, sl, tile0, lt, lane, true, false, bad);
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                double G[8], y[8], other[8];
                res_ld8(rs.G + lt * TILE, lane, G);
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
        {
            const long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) {
                const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
                if (q >= 1 && tb[i] != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)tb[i]);
            }
        }
        grid_arrive(bar, target);
        if (!grid_wait(bar, target, d.err, &bar_ok)) return;
        sweep_scalars(d, d.trees[par ^ 1], mom, red, &scs, writer);
        // ---- finishing pass: posterior accumulation, back to the mu forest (G = Gc, base = y - bscale' * Gm) ----
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            for (int lt = wib; lt < ntl; lt += wpb) {
                const int tile = tile0 + lt;
                const int64_t gb = (int64_t)tile * TILE + lane * 8;
                const double b = tile < d.trt_tiles ? b1 : b0;
                double gc[8], gm[8], y[8], B[8];
                ld8_f64(d.y + gb, y);
                if (has_mod) { res_ld8(rs.G + lt * TILE, lane, gm); ld8_f64(d.Gc + gb, gc); st8_f64(d.Gm + gb, gm); }
                else {
                    res_ld8(rs.G + lt * TILE, lane, gc);
#pragma unroll
                    for (int o = 0; o < 8; ++o) gm[o] = 0.0;
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) B[o] = y[o] - b * gm[o];
                res_st8(rs.base + lt * TILE, lane, B);
                if (has_mod) res_st8(rs.G + lt * TILE, lane, gc);
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
            for (int64_t i = gtid; i < ntot; i += gstride) d.totals[i] = 0;
            long long* mom_next = d.mom + (size_t)((sc.sweep + 1u) & 1u) * (2 * NMOM * 3);
            for (int64_t i = gtid; i < 2 * NMOM * 3; i += gstride) mom_next[i] = 0;
        }
        __syncthreads();
        PROF(6);
    }
    // exit: park the state in HBM in the form every engine understands (Gc, Gm, full residual r)
    {
        const Scalars sc = scs;
        for (int lt = wib; lt < ntl; lt += wpb) {
            const int64_t gb = (int64_t)(tile0 + lt) * TILE + lane * 8;
            double G[8], B[8], r[8];
            res_ld8(rs.G + lt * TILE, lane, G);
            res_ld8(rs.base + lt * TILE, lane, B);
            st8_f64(d.Gc + gb, G);
#pragma unroll
            for (int o = 0; o < 8; ++o) r[o] = B[o] - sc.mscale * G[o];
            st8_f64(d.r + gb, r);
        }
    }
    if (prof) for (int k = 0; k < 8; ++k) d.prof[k] += pc[k];
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
    if (t == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

#undef PROF
#undef BCF_PROF_DECL

}  // namespace bcf
