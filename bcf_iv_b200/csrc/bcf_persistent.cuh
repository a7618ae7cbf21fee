// bcf_persistent.cuh -- PERSISTENT engine: a batch of Gibbs sweeps in ONE cooperative kernel.
//
// One CTA per SM, every CTA owns a fixed set of observation tiles for the whole launch (so the residual, the
// leaf-slot cache and the forest fits of a tile are only ever touched by their owner: no cross-CTA hazards on
// observation data).  Per tree there is exactly one grid-wide dependency -- the reduced leaf statistics --
// carried by integer REDs into L2 plus one grid barrier; after it every CTA replays the Metropolis-Hastings
// decision and the leaf draws from the same bits and Philox counters, so nothing is broadcast.
// Per sweep: T + 1 barriers (T trees + the moments of the scale / sigma draws).
#pragma once
#include "bcf_kernels.cuh"

namespace bcf {

constexpr int PTHREADS = 512;
constexpr int PERSIST_CTAS_PER_SM = 1;

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

// Grid barrier on a monotonically increasing arrival counter (zeroed by the host before each launch).
// A bounded spin: if a CTA never arrives the kernel flags ERR_BARRIER_TIMEOUT and leaves instead of hanging the GPU.
__device__ __forceinline__ bool grid_barrier(unsigned int* bar, unsigned int& target, int32_t* err)
{
    __shared__ int ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        red_release_add_u32(bar, 1u);
        int good = 1;
        const long long t0 = clock64();
        while (ld_acquire_u32(bar) < target) {
            if (clock64() - t0 > 4000000000ll) { good = 0; atomicOr(err, ERR_BARRIER_TIMEOUT); break; }
            if (ld_acquire_u32((const unsigned int*)err) & ERR_BARRIER_TIMEOUT) { good = 0; break; }
        }
        __threadfence();
        ok = good;
    }
    __syncthreads();
    return ok != 0;
}

__global__ void __launch_bounds__(PTHREADS, 1) k_persistent(Dev d, int nsweeps, unsigned int* bar)
{
    __shared__ StepShared sm;
    __shared__ Scalars scs;
    __shared__ double red[2 * PTHREADS];
    const int t = threadIdx.x;
    const bool writer = (blockIdx.x == 0);
    const int lane = t & 31, wib = t >> 5, wpb = blockDim.x >> 5;
    unsigned int target = 0;
    int bad = 0;
    if (t == 0) { scs = *d.sc; sm.bad = 0; }
    __syncthreads();
    const int T = d.T;

    for (int sw = 0; sw < nsweeps; ++sw) {
        const Scalars sc = scs;
        const int par = sc.parity;
        long long* mom = d.mom + (size_t)(sc.sweep & 1u) * (2 * NMOM * 3);
        if (t == 0) { sm.ai.active = 0; sm.ci.active = 0; }
        __syncthreads();
        // first tree of the sweep
        load_tree(&sm.tr, &d.trees[par][0]);
        __syncthreads();
        propose(sm, d, d.f[0], 0, sc.sweep);
        begin_stats(sm, 0);
        __syncthreads();
        for (int j = 0; j < T; ++j) {
            const int fj = j >= d.f[1].tree0 ? 1 : 0;
            for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb)
                step_tile(sm, d, j - 1, j, tile, lane, bad);
            __syncthreads();
            flush_stats(sm, d, j);
            if (!grid_barrier(bar, target, d.err)) return;
            decide(sm, d, d.f[fj], j, sc, writer);
            if (writer) store_tree(&d.trees[par ^ 1][j], &sm.tr);
            __syncthreads();
            if (j + 1 < T) {
                const int fn = (j + 1) >= d.f[1].tree0 ? 1 : 0;
                load_tree(&sm.tr, &d.trees[par][j + 1]);
                __syncthreads();
                propose(sm, d, d.f[fn], j + 1, sc.sweep);
                begin_stats(sm, fn);
                __syncthreads();
            }
        }
        // epilogue: finish the last tree, rebuild the fits from the trees, moments
        {
            long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) tb[i] = 0;
            if (t == 0) sm.ci.active = 0;
        }
        __syncthreads();
        for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb)
            refit_tile(sm, d, T - 1, d.trees[par ^ 1], tile, lane, bad);
        __syncthreads();
        {
            const long long* tb = &sm.tab[0][0][0];
            for (int i = t; i < NMOM * 8; i += blockDim.x) {
                const int m = i >> 3, g = (i >> 2) & 1, q = i & 3;
                if (q >= 1 && tb[i] != 0) atomicAdd((unsigned long long*)&mom[(g * NMOM + m) * 3 + (q - 1)], (unsigned long long)tb[i]);
            }
        }
        if (!grid_barrier(bar, target, d.err)) return;
        sweep_scalars(d, d.trees[par ^ 1], mom, red, &scs, writer);
        // finishing pass over this CTA's own tiles; clear next sweep's accumulators (nobody reads them now)
        {
            const Scalars sn = scs;
            const double ms = sn.mscale, b1 = sn.bscale1, b0 = sn.bscale0;
            const int row = sn.save_row;
            for (int tile = blockIdx.x * wpb + wib; tile < d.n_tiles; tile += gridDim.x * wpb) {
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
    }
    if (bad) atomicOr(d.err, ERR_FX_RANGE);
    if (t == 0 && sm.bad) atomicOr(d.err, sm.bad);
}

}  // namespace bcf
