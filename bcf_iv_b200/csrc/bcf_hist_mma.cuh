// bcf_hist_mma.cuh -- K2' (all-candidate histogram) for two-valued columns on the 5th-generation tensor cores.
//
// For two-valued (one-cutpoint) covariates -- every covariate of the reference's designs -- the statistics of all candidate
// splits of all leaves are ONE contraction over the observations:
//         D[(leaf, word), column] = sum_i  A[(leaf, word), i] * B[column, i]
//     B[column, i] = the column's bin of observation i (0 / 1): the sampler's uint8 bins exactly as they lie in HBM
//                    (column-major = K-major), brought to shared memory by TMA (128-byte swizzle), never touched by a thread;
//     A[(leaf, word), i] = (label_i == leaf) ? word(r_i) : 0, word = one of nine 7-bit limbs of the residual in 2^-44 fixed point
//                    (7 divides the 21-bit limbs every other kernel sums) or the constant 1 (the count): int8, built by
//                    the builder warps straight into the swizzled K-major tile.
// tcgen05.mma.kind::i8 accumulates int8 x int8 into int32 EXACTLY (a CTA's share of the rows keeps |sum| <= 127 * rows < 2^31),
// the accumulator (128 x N) lives in tensor memory for the whole pass, and the epilogue recombines the limbs into the
// (count, three words of weight 1, 2^21, 2^42) output: bit-identical to summing the fixed-point residuals in any order, and
// independent of the number of leaves (<= 12 per pass: 12 x 10 rows <= M = 128).  One extra row of B holds ones: its
// column of D is the leaf totals (bin 0 = total - bin 1).  Algorithmic traffic n * (9 + columns) bytes, every byte read once.
//
// What bounds it (measured on B200 with the pieces switched off one at a time, BCFGPU_HIST_PROBE, and on CTA 0's timeline,
// BCFGPU_HIST_TRACE; n = 1e6, 100 columns): not the tensor pipe and not the TMA engine but LATENCY -- a stage's round trip (copy
// issued -> landed 2-3.5 k cycles -> A built ~1 k -> MMAs retired ~0.9 k -> stage free) against the bytes its ring holds, and
// the single threads that issue a tile's copies or MMAs (about 100 cycles per asynchronous operation).  The design follows:
//   * only the 1 KB row groups of A that hold live rows exist in shared memory (the MMA reads M = 128 rows regardless; rows
//     beyond the leaves alias whatever follows and only feed rows of D nobody reads);
//   * THREE rings: the bins (B: 28 KB per K-tile at 100 columns, the bulk of the traffic) in as many stages as shared memory
//     holds (6 at 1-4 leaves), held only from "copy issued" to "MMAs retired"; the tiles' residuals and labels (2.3 KB) eight
//     deep; the A tiles built from them in 4 stages, built while B is still in flight;
//   * K-tiles of 256 observations (two 128-byte swizzle atoms side by side): one wait, one fence, two commits per EIGHT MMAs;
//   * every single-thread role twice, over the tiles' parity.
// (An e4m3 variant -- kind::f8f6f4, sixteen 4-bit limbs, 0x01 read as the subnormal 2^-9, fp32 accumulators -- was built and
// is exact bit for bit as well, but needs 17 rows per leaf and is no faster: removed.)
//
// Roles (704 threads, one CTA per SM, a contiguous range of K-tiles per CTA):
//   warps 0, 1  B producers (even / odd K-tiles): two 2-D TMA boxes (128 observations x all columns each) -> B[stage], one
//               mbarrier (complete_tx)
//   warps 2, 3  input producers: two bulk copies (the tile's residuals, labels) -> the input ring
//   warps 4, 5  MMA issuers (an accumulator each: two MMAs in flight never share one): one lane issues 8 x tcgen05.mma
//               (K = 32 each) per tile and commits twice (A stage free, B stage free); warp 4 owns the tensor memory
//   warps 6-21  builders, four per A stage (64 observations each, independent of one another: no barrier anywhere in the loop):
//               residuals -> limbs (two observations per lane) -> masked by leaf into A[stage]; afterwards warps 8-11 read the
//               two accumulators (tcgen05.ld) and everybody adds the words to the output
// mbarrier discipline: a parity wait two phases ahead of a barrier returns at once, so a thread only waits on barriers whose
// every phase it observes itself -- stage counts are even (a stage is always served by the same producer and issuer), the
// input ring is a multiple of the A ring (an input stage is always read by the same builders), builders own their A stage.
#pragma once
#include <cuda.h>
#include "bcf_device.cuh"

namespace bcf {

constexpr int HM_KT = 128;                 // observations per swizzle atom: one 128-byte row of an 8-bit operand
constexpr int HM_SUB = 2;                  // atoms per K-tile
constexpr int HM_TILE = HM_KT * HM_SUB;    // observations per K-tile
constexpr int HM_ROWS = 10;                // rows of A per leaf: nine 7-bit limbs + the count
constexpr int HM_MAX_LEAF = 12;            // 12 * 10 = 120 <= M
constexpr int HM_M = 128;
constexpr int HM_MAX_A_STAGES = 4, HM_MAX_B_STAGES = 16;
constexpr int HM_IN_STAGES = 8;             // ring of the tiles' residuals and labels (a multiple of the A stages: a stage of it is always read by the same builders)
constexpr int HM_IN_BYTES = HM_TILE * 8 + HM_TILE;   // a K-tile's residuals and labels
constexpr int HM_MAX_COLS = 240;           // columns per pass (+ the row of ones, rounded up to 16: N <= 256)
constexpr int HM_HALF = HM_KT / 2;         // observations per builder warp (two per lane)
constexpr int HM_BUILDERS = HM_MAX_A_STAGES * HM_TILE / HM_HALF;   // 16 builder warps: four per A stage
constexpr int HM_THREADS = 192 + 32 * HM_BUILDERS;   // two B producers, two input producers, two MMA issuers, the builders
constexpr int HM_A_BYTES = HM_M * HM_KT;   // what an MMA reads of A: 16 KB

__host__ __device__ constexpr int hm_n(int ncols_box) { return (ncols_box + 1 + 15) / 16 * 16; }
// bytes of one atom of A that exist in shared memory: the 8-row groups (1 KB) that hold a live row
__host__ __device__ constexpr size_t hm_a_bytes(int nleaf) { return (size_t)((nleaf * HM_ROWS + 7) / 8) * 8 * HM_KT; }
__host__ __device__ constexpr size_t hm_b_bytes(int ncols_box) { return (size_t)hm_n(ncols_box) * HM_KT; }
constexpr size_t HM_TAIL_BYTES = HM_BUILDERS * (HM_ROWS + 1) * HM_HALF + 768;   // staging rows of the builder warps, barriers, tensor-memory address
// B stages, A stages, input rows, tail; + 1 KB to align
__host__ __device__ constexpr size_t hm_smem_bytes(int ncols_box, int nleaf, int na, int nb)
{
    return 1024 + nb * HM_SUB * hm_b_bytes(ncols_box) + na * HM_SUB * hm_a_bytes(nleaf) + HM_IN_STAGES * HM_IN_BYTES + HM_TAIL_BYTES;
}

__device__ __forceinline__ uint32_t hm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hm_mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void hm_mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void hm_mbar_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool hm_mbar_try(uint32_t a, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a logic error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void hm_mbar_wait(uint32_t a, uint32_t parity, int32_t* err, int* dbg, int what)
{
    if (hm_mbar_try(a, parity)) return;
    const long long t0 = clock64();
    while (!hm_mbar_try(a, parity))
        if (clock64() - t0 > 4000000000ll) { atomicOr(err, ERR_BARRIER_TIMEOUT); give_up(dbg, 50 + what, (int)parity, (int)a); }
}
__device__ __forceinline__ void hm_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void hm_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hm_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major 8-bit operand tile with the 128-byte swizzle: rows of 128 bytes, groups of 8
// rows 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B); `byte_off` steps along K inside
// the swizzle row (32 bytes per MMA)
__device__ __forceinline__ uint64_t hm_desc(uint32_t tile_addr, uint32_t byte_off)
{
    return (uint64_t)(((tile_addr + byte_off) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(HM_THREADS, 1)
k_hist_mma(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ label, const double* __restrict__ r, int nleaf,
           int ncols_box, int p8, int ktiles, int tiles_per_cta, int na, int nb, const int32_t* __restrict__ colvar,
           const int32_t* __restrict__ off, long long nbt, long long* out, long long* leaf_tot_out, int32_t* err, int* dbg, int probe,
           long long* trace)
{
    // trace (BCFGPU_HIST_TRACE, CTA 0): clock64 stamps [event][K-tile < 64]: 0 bins issued, 1 builder saw its inputs land, 2 A built,
    // 3 MMA thread saw A and B, 4 MMAs committed; [5][0..4]: start, set-up done, accumulator complete, accumulator in smem, end
    const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define HM_TRACE(ev, i) do { if (tr && (i) < 64) trace[(ev) * 64 + (i)] = clock64(); } while (0)
    if (tr && threadIdx.x == 0) trace[5 * 64 + 0] = clock64();
    extern __shared__ unsigned char hm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)hm_raw + 1023) & ~(uintptr_t)1023);
    const int N = hm_n(ncols_box);
    const size_t a_bytes = hm_a_bytes(nleaf), b_bytes = hm_b_bytes(ncols_box);
    unsigned char* a_base = sm + nb * HM_SUB * b_bytes;                         // [A stage][atom][live row groups of A]
    unsigned char* in_base = a_base + na * HM_SUB * a_bytes;                    // [input stage][residuals 2 KB, labels 256 B]
    unsigned char* stg_base = in_base + HM_IN_STAGES * HM_IN_BYTES;                       // [builder][HM_ROWS + 1][HM_HALF]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + HM_BUILDERS * (HM_ROWS + 1) * HM_HALF);
    const uint32_t bar_fullIn = hm_u32(bars), bar_emptyIn = bar_fullIn + 8 * HM_IN_STAGES, bar_fullA = bar_emptyIn + 8 * HM_IN_STAGES,
                   bar_emptyA = bar_fullA + 8 * HM_MAX_A_STAGES, bar_fullB = bar_emptyA + 8 * HM_MAX_A_STAGES,
                   bar_emptyB = bar_fullB + 8 * HM_MAX_B_STAGES, bar_done = bar_emptyB + 8 * HM_MAX_B_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HM_IN_STAGES + 2 * HM_MAX_A_STAGES + 2 * HM_MAX_B_STAGES + 1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int tile_lo = blockIdx.x * tiles_per_cta, ntile = min(ktiles, tile_lo + tiles_per_cta) - tile_lo;
    const int col0 = blockIdx.y * ncols_box;
    uint32_t acc_cols = 32;                                                      // columns of one accumulator (a power of two >= N)
    while ((int)acc_cols < N) acc_cols <<= 1;
    const uint32_t ncols_tmem = 2 * acc_cols;

    // ---- one-time set-up: the row of ones of B (a constant row: the swizzle does not matter), barriers, tensor memory.  Nothing
    //      else is initialised: rows of A beyond the leaves and rows of B beyond the ones only feed rows / columns of D that
    //      nobody reads (rows and columns of an MMA are independent) ----
    for (int i = t; i < nb * HM_SUB * 8; i += HM_THREADS) {
        const uint32_t v = 0x01010101u;
        reinterpret_cast<uint4*>(sm + (i >> 3) * b_bytes + ncols_box * HM_KT)[i & 7] = make_uint4(v, v, v, v);
    }
    hm_fence_async_smem();
    if (t == 0) {
        for (int s = 0; s < HM_IN_STAGES; ++s) {
            hm_mbar_init(bar_fullIn + 8 * s, 1);
            hm_mbar_init(bar_emptyIn + 8 * s, 32 * HM_TILE / HM_HALF);
        }
        for (int s = 0; s < na; ++s) {
            hm_mbar_init(bar_fullA + 8 * s, 32 * HM_TILE / HM_HALF);
            hm_mbar_init(bar_emptyA + 8 * s, 1);
        }
        for (int s = 0; s < nb; ++s) {
            hm_mbar_init(bar_fullB + 8 * s, 1);
            hm_mbar_init(bar_emptyB + 8 * s, 1);
        }
        hm_mbar_init(bar_done, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hm_u32(tmem_slot)), "r"(ncols_tmem) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    hm_tc_fence_before();
    __syncthreads();
    hm_tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (tr && threadIdx.x == 0) trace[5 * 64 + 1] = clock64();

    // (na and nb are even: a thread that takes the tiles of one parity serves the same stages every round, so each of its waits is
    // for the phase right after the last one it saw complete -- a parity wait two phases ahead of a barrier would return at once)
    if (warp < 2) {
        // =================== B producers: warp w takes the K-tiles it = w, w + 2, ... ===================
        if (lane == 0 && !(probe & 2)) {   // (timing probes, BCFGPU_HIST_PROBE: 1 = no MMAs issued, 2 = no bins fetched; results are then wrong)
            for (int it = warp, s = warp, ph = 0; it < ntile; it += 2, s += 2) {
                if (s >= nb) { s -= nb; ph ^= 1; }
                hm_mbar_wait(bar_emptyB + 8 * s, (uint32_t)ph ^ 1u, err, dbg, 0);
                const uint32_t bar = bar_fullB + 8 * s;
                hm_mbar_expect_tx(bar, (uint32_t)(HM_SUB * ncols_box * HM_KT));
#pragma unroll
                for (int j = 0; j < HM_SUB; ++j)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(hm_u32(sm + (s * HM_SUB + j) * b_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar),
                                   "r"((tile_lo + it) * HM_TILE + j * HM_KT), "r"(col0)
                                 : "memory");
                HM_TRACE(0, it);
            }
        }
        __syncwarp();   // (the other lanes must not run ahead into the epilogue's wait: a spinning sibling path would starve lane 0)
    } else if (warp < 4) {
        // =================== input producers: warp 2 + w takes the K-tiles it = w, w + 2, ... ===================
        if (lane == 0) {
            for (int it = warp - 2, s = warp - 2, ph = 0; it < ntile; it += 2, s += 2) {
                if (s >= HM_IN_STAGES) { s -= HM_IN_STAGES; ph ^= 1; }
                hm_mbar_wait(bar_emptyIn + 8 * s, (uint32_t)ph ^ 1u, err, dbg, 5);
                const uint32_t bar = bar_fullIn + 8 * s, din = hm_u32(in_base + s * HM_IN_BYTES);
                const int64_t i0 = (int64_t)(tile_lo + it) * HM_TILE;
                hm_mbar_expect_tx(bar, HM_IN_BYTES);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(din), "l"(r + i0), "r"(HM_TILE * 8), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(din + HM_TILE * 8), "l"(label + i0), "r"(HM_TILE), "r"(bar) : "memory");
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // =================== MMA issuers: warp 4 + w takes the K-tiles it = w, w + 2, ... into accumulator w ===================
        if (lane == 0) {
            const int w = warp - 4;
            const uint32_t tacc = tmem + (uint32_t)w * acc_cols;
            // instruction descriptor (kind::i8): D = s32 (2 at bit 4), A = B = signed 8-bit (1 at bits 7 and 10), both K-major,
            // N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(HM_M >> 4) << 24);
            for (int it = w, sa = w, pa = 0, sb = w, pb = 0; it < ntile; it += 2, sa += 2, sb += 2) {
                if (sa >= na) { sa -= na; pa ^= 1; }
                if (sb >= nb) { sb -= nb; pb ^= 1; }
                hm_mbar_wait(bar_fullA + 8 * sa, (uint32_t)pa, err, dbg, 1);
                if (!(probe & 2)) hm_mbar_wait(bar_fullB + 8 * sb, (uint32_t)pb, err, dbg, 2);
                HM_TRACE(3, it);
                hm_tc_fence_after();
                if (!(probe & 1)) {
#pragma unroll
                    for (int j = 0; j < HM_SUB; ++j) {
                        const uint32_t a_addr = hm_u32(a_base + (sa * HM_SUB + j) * a_bytes), b_addr = hm_u32(sm + (sb * HM_SUB + j) * b_bytes);
#pragma unroll
                        for (int k = 0; k < HM_KT / 32; ++k) {
                            const uint64_t da = hm_desc(a_addr, k * 32), db = hm_desc(b_addr, k * 32);
                            const uint32_t acc = (it >= 2 || j != 0 || k != 0) ? 1u : 0u;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                        }
                    }
                }
                // (tcgen05.commit implies tcgen05.fence::before_thread_sync) both stages are free once these MMAs have read them
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_emptyA + 8 * sa) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_emptyB + 8 * sb) : "memory");
                HM_TRACE(4, it);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
        }
        __syncwarp();
    } else {
        // =================== builders: warps 6 + 4 s .. 6 + 4 s + 3 OWN A stage s, 64 observations each ===================
        const int bw = warp - 6, s = bw >> 2, h = bw & 3, j = h >> 1, half = h & 1;      // atom j of the K-tile, half of the atom
        unsigned char* row = stg_base + (size_t)bw * (HM_ROWS + 1) * HM_HALF;
        int bad = 0;
        // (the input ring has a multiple of na stages: its stage si is always read by this A stage's builders)
        for (int it = s, ph = 0, si = s, pi = 0; s < na && it < ntile; it += na, ph ^= 1, si += na) {
            if (si >= HM_IN_STAGES) { si -= HM_IN_STAGES; pi ^= 1; }
            hm_mbar_wait(bar_fullIn + 8 * si, (uint32_t)pi, err, dbg, 3);        // the tile's residuals and labels have landed
            if (lane == 0 && h == 0) HM_TRACE(1, it);
            const unsigned char* in = in_base + si * HM_IN_BYTES;
            long long fx[2];
            uint32_t labw = 0u;
            {   // this lane's two observations
                const double2 r01 = reinterpret_cast<const double2*>(in)[h * (HM_HALF / 2) + lane];
                const double rv[2] = {r01.x, r01.y};
                const uint32_t lw = reinterpret_cast<const uint16_t*>(in + HM_TILE * 8)[h * (HM_HALF / 2) + lane];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t lab = (lw >> (8 * q)) & 0xFFu;
                    const bool live = lab < (uint32_t)nleaf;
                    fx[q] = live ? to_fixed(rv[q], bad) : 0ll;
                    labw |= (live ? lab : 255u) << (8 * q);
                }
            }
            hm_mbar_arrive(bar_emptyIn + 8 * si);                                // (read: the input stage may be refilled)
            uint16_t* roww = reinterpret_cast<uint16_t*>(row) + lane;            // halfword `lane` of every staging row: this lane's 2 bytes
#pragma unroll
            for (int jj = 0; jj < 9; ++jj) {
                uint32_t w = 0u;
#pragma unroll
                for (int q = 0; q < 2; ++q)                                        // (the top limb keeps the sign)
                    w |= (jj < 8 ? (uint32_t)(fx[q] >> (7 * jj)) & 0x7Fu : (uint32_t)(fx[q] >> 56) & 0xFFu) << (8 * q);
                roww[jj * (HM_HALF / 2)] = (uint16_t)w;
            }
            roww[9 * (HM_HALF / 2)] = 0x0101u;
            roww[HM_ROWS * (HM_HALF / 2)] = (uint16_t)labw;
            __syncwarp();
            hm_mbar_wait(bar_emptyA + 8 * s, (uint32_t)ph ^ 1u, err, dbg, 6);    // the MMAs of the stage's last tile have retired
            unsigned char* A = a_base + (s * HM_SUB + j) * a_bytes;
            for (int c = lane; c < HM_ROWS * 4; c += 32) {   // one (word, 16 observations) chunk at a time, masked once per leaf
                const int jj = c >> 2, kl = c & 3, kc = half * 4 + kl;
                const uint4 v = *reinterpret_cast<const uint4*>(row + jj * HM_HALF + kl * 16);
                const uint4 lb = *reinterpret_cast<const uint4*>(row + HM_ROWS * HM_HALF + kl * 16);
#pragma unroll 4
                for (int l = 0; l < nleaf; ++l) {
                    const int m = l * HM_ROWS + jj;
                    const uint32_t lv = (uint32_t)l * 0x01010101u;
                    *reinterpret_cast<uint4*>(A + m * HM_KT + ((kc ^ (m & 7)) << 4)) =
                        make_uint4(v.x & __vcmpeq4(lb.x, lv), v.y & __vcmpeq4(lb.y, lv), v.z & __vcmpeq4(lb.z, lv), v.w & __vcmpeq4(lb.w, lv));
                }
            }
            hm_fence_async_smem();                                               // generic-proxy stores -> the tensor core's reads
            hm_mbar_arrive(bar_fullA + 8 * s);
            if (lane == 0 && h == 0) HM_TRACE(2, it);
            __syncwarp();                                                        // (the staging rows are rewritten by the next tile)
        }
        if (bad) atomicOr(err, ERR_FX_RANGE);
    }

    // ---- epilogue: accumulator -> shared memory (over the idle stages) -> limbs recombined -> output ----
    hm_mbar_wait(bar_done, 0u, err, dbg, 4);
    hm_tc_fence_after();
    __syncthreads();
    if (tr && threadIdx.x == 0) trace[5 * 64 + 2] = clock64();
    int32_t* D = reinterpret_cast<int32_t*>(sm);                                 // [HM_M][N]
    if (warp >= 8 && warp < 12) {
        const int q = warp & 3;                                                  // the lane quarter of tensor memory this warp may read
        const int m = q * 32 + lane;
        const bool two = ntile >= 2;                                             // (the odd issuer wrote its accumulator)
        for (int c = 0; c < N; c += 16) {
            uint32_t v[16], u[16];
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(ta) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                           "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                         : "r"(ta + acc_cols) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (two) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += u[i];
            }
            if (m < nleaf * HM_ROWS) {
                uint4* dst = reinterpret_cast<uint4*>(D + (size_t)m * N + c);
                dst[0] = make_uint4(v[0], v[1], v[2], v[3]); dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                dst[2] = make_uint4(v[8], v[9], v[10], v[11]); dst[3] = make_uint4(v[12], v[13], v[14], v[15]);
            }
        }
    }
    hm_tc_fence_before();
    __syncthreads();
    if (tr && threadIdx.x == 0) trace[5 * 64 + 3] = clock64();
    for (int idx = t; idx < nleaf * (ncols_box + 1); idx += HM_THREADS) {
        const int l = idx / (ncols_box + 1), c = idx - l * (ncols_box + 1);
        long long* o = nullptr;
        if (c == ncols_box) { if (blockIdx.y == 0) o = leaf_tot_out + l * 4; }
        else if (col0 + c < p8) {
            const int v = colvar[col0 + c];
            if (v >= 0) o = out + (((long long)l * nbt + (long long)off[v] + v + 1) * 4);      // bin 1
        }
        if (o == nullptr) continue;
        const int32_t* d = D + (size_t)(l * HM_ROWS) * N + c;
        long long w[4];
        {
            w[0] = (long long)d[9 * N];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                w[1 + k] = (long long)d[(3 * k) * N] + ((long long)d[(3 * k + 1) * N] << 7) + ((long long)d[(3 * k + 2) * N] << 14);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (w[k] != 0) atomicAdd((unsigned long long*)&o[k], (unsigned long long)w[k]);
    }
    __syncthreads();
    if (tr && threadIdx.x == 0) trace[5 * 64 + 4] = clock64();
#undef HM_TRACE
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols_tmem) : "memory");
}

}  // namespace bcf
