// bcf_device.cuh -- device-side data structures and building blocks of the BCF Gibbs sweep (sm_100a).
//
// What is replaced (third-party `bcf` C++ sampler reached from R/bcf_iv.R:89 and R/bcf_itt.R:71):
//   tree / tree::bn / fit  -> flat TreeRec + cached uint8 leaf slot per (tree, observation)        (a2)
//   getsuff / allsuff      -> one fused pass: per-leaf (count, sum r) by treatment group, exact    (a4,a5)
//   bd / bdhet             -> propose() + decide(): structural prior ratio, integrated likelihood  (a6)
//   drmu / drmuhet         -> leaf draws inside decide()                                            (a7)
// Sums are accumulated in 64-bit FIXED POINT (value * 2^44 split in three 21-bit limbs), so every reduction
// is associative: results do not depend on thread, CTA or GPU partitioning, and every CTA / rank can replay the
// Metropolis-Hastings decision from the same bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// Phase timers / batch timelines of the sweep kernels (BCFGPU_PROFILE, BCFGPU_TRACE) are compiled in only with -DBCF_PROF:
// the clock reads and their bookkeeping cost ~6 % of a sweep.
#ifdef BCF_PROF
#define BCF_PROF_CLOCK() clock64()
#else
#define BCF_PROF_CLOCK() 0ll
#endif

namespace bcf {

constexpr int MAXL = 128;           // BCFGPU_MAX_LEAVES
constexpr int MAXN = 256;           // node capacity (2*MAXL-1 used)
constexpr int KEYS = MAXL + 1;      // leaf slots + the would-be right child of a birth
constexpr int MAX_DEPTH = 30;       // BCFGPU_MAX_DEPTH
constexpr int TILE = 256;           // observations per warp tile (8 per lane)
constexpr int PAD_SLOT = 255;       // slot value of padding observations
constexpr int FX_SHIFT = 44;        // fixed point: value * 2^44
constexpr int LIMB_BITS = 21;
constexpr int KREG = 8;             // key-loop (register + REDUX) path up to this many keys, shared atomics above
constexpr int NMOM = 6;             // moments per group: yy, yGc, yGm, GcGc, GcGm, GmGm
constexpr int THREADS = 256;

enum { PUR_MOVE = 0, PUR_NODE = 1, PUR_CUT = 2, PUR_LEAF = 3, PUR_BSCALE1 = 10, PUR_BSCALE0 = 11,
       PUR_MSCALE = 12, PUR_DELTA_CON = 13, PUR_DELTA_MOD = 14, PUR_SIGMA = 15 };
constexpr uint32_t TREE_SWEEP_LEVEL = 0xFFFFFFFFu;

enum { ERR_NAN = 1, ERR_FX_RANGE = 2, ERR_BARRIER_TIMEOUT = 4 };
// a wait gave up: leave a note the host can still read (code, CTA, two details), then trap
__device__ __noinline__ void give_up(int* dbg, int code, int a, int b)
{
    if (dbg != nullptr) {
        volatile int* q = dbg;
        q[1] = code; q[2] = (int)blockIdx.x; q[3] = a; q[4] = b; q[0] = 1;
        __threadfence_system();
    }
    __trap();
}

// ---------------------------------------------------------------------------------------------
// Flat tree.  Node 0 is the root; nodes are kept compact (swap-with-last on death).
// ---------------------------------------------------------------------------------------------
template <int NN, int NL>
struct __align__(16) TreeT {
    static constexpr int kNodes = NN, kLeaves = NL;
    int32_t nnodes, nleaves;
    int32_t pad_[2];
    int16_t parent[NN], left[NN], right[NN];       // -1 = none
    uint16_t var[NN], cut[NN];                     // split rule of internal nodes
    uint8_t node_slot[NN];                         // leaf slot of leaf nodes
    uint8_t depth[NN];
    uint32_t heap[NN];                             // heap id: root 1, children 2i / 2i+1
    int16_t slot_node[NL];                         // node of each leaf slot
    int32_t cnt0[NL], cnt1[NL];                    // observations in the leaf, control / treated (all shards)
    double mu[NL];                                 // leaf value per slot
};
using TreeRec = TreeT<MAXN, MAXL>;                 // the HBM record of every tree
static_assert(sizeof(TreeRec) % 16 == 0, "TreeRec must be copyable as uint4");

struct Proposal {
    int32_t move;        // 0 none, 1 birth, 2 death
    int32_t node;        // birth: the leaf node; death: the nog node
    int32_t slot;        // birth: slot of the leaf (left child keeps it); death: slot of the left child
    int32_t slot_b;      // birth: slot the right child would get (= nleaves); death: slot of the right child
    int32_t var, cut;
    uint32_t heap;
    int32_t is16;        // width of the proposed variable's bin column
    const void* bins;    // its column in HBM (n_pad entries)
    double log_prior;    // log of the structural part of the MH ratio (alpha1)
    double log_u;        // log of the accept uniform: accept iff log_u < log ratio
    double z_extra[2];   // normals of the leaves a move would create: birth {left, right child}, death {merged leaf}
};

// Proposal of one tree for the coming sweep, made ahead of time (it depends only on the tree's structure and the
// Philox counters) together with the standard normals of its current leaves.
struct PropRec {
    Proposal P;
    double z[MAXL];
};

struct Scalars {
    double sigma, mscale, bscale1, bscale0, delta_con, delta_mod, tau_con, tau_mod;
    uint32_t sweep;      // global sweep index (Philox counter word 0)
    int32_t parity;      // which TreeRec copy is current
    int32_t run_it, nburn, nthin, nsim, save_ctr, save_row;   // save_row: row to write this sweep or -1
    int32_t pad_[2];
};

struct ForestDev {
    int32_t p, ntree, tree0, is_con;
    const uint8_t* bins8;      // [col][n_pad]
    const uint16_t* bins16;    // [col][n_pad]
    const int32_t* col_index;  // [p] index within bins8 / bins16
    const uint8_t* col_is16;   // [p]
    const int32_t* ncut;       // [p]
    double pg[32];             // alpha / (1+depth)^beta, tabulated on the host (bit-identical pow)
};

// ---------------------------------------------------------------------------------------------
// Observation shards over several GPUs: the per-batch leaf statistics are exchanged INSIDE the sweep
// kernel through peer memory (NVLink / NVSwitch): every rank owns one Mailbox per source rank in its own HBM; after
// its local grid barrier a rank stores its batch words into its mailbox on every peer and then releases a sequence
// number; a rank proceeds once all its mailboxes carry the sequence number it expects, and sums the (integer) words in
// a fixed order.  Two batch slots are enough: a rank cannot be two batches ahead of a peer whose words it needs.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 8;
constexpr int MAIL_TOT_WORDS = KEYS * 8;            // totals words of a batch (a big tree: KEYS * 8; 8 small trees: 512)
constexpr int MAIL_X_WORDS = 512;                   // cross-count words of a batch (batched engine: 256; pipelined: 2 * 216)
constexpr int MAIL_MOM_WORDS = 2 * NMOM * 3;
constexpr int MAIL_SLOTS = 4;                       // batch slots of the pipelined engine (the batched engine alternates between
                                                    // two): its statistics pass runs up to two batches ahead of the decisions,
                                                    // so a rank may push batch b + 3 while a peer still reads batch b
constexpr int MAIL_FLAG_MOM = MAIL_SLOTS;           // flag index of the moments exchange
constexpr int MAIL_LL_WORDS = 4 * 8 * 8 + 2 * 216;  // a batch of the pipelined engine: PNB * KB * 8 totals words + 2 * PXCAP cross counts
struct __align__(16) Mailbox {
    unsigned long long flag[MAIL_SLOTS + 4];        // [0..1]: batch slots of the batched engine, [4]: moments -- sequence number of the last complete push
    long long batch[2][MAIL_TOT_WORDS + MAIL_X_WORDS];
    long long mom[MAIL_MOM_WORDS + 4];
    // Pipelined engine: self-validating words, (value << 16) | tag -- a 48-bit value and the low 16 bits of the batch's
    // global number in ONE 8-byte store (single-copy atomic), so a push needs no fence and no flag and a reader that
    // sees the tag it expects has the value: one NVLink crossing per batch instead of store + fence + flag + poll.
    unsigned long long ll[MAIL_SLOTS][MAIL_LL_WORDS];
};
struct PeerDev {
    int32_t rank, nranks;
    unsigned long long seq0;                        // launch id << 32: sequence numbers never repeat across launches
    unsigned long long batch0;                      // batches of the pipelined engine exchanged before this launch (tags and slots go on)
    Mailbox* mine;                                  // [nranks] my mailboxes, indexed by SOURCE rank (local HBM)
    Mailbox* peer[MAX_RANKS];                       // peer[p]: rank p's mailbox array, mapped here (cudaIpc)
};

struct Dev {
    PeerDev pr;
    int64_t n_pad;             // padded local observations (multiple of TILE)
    int64_t n_global;          // real observations over all shards (chi-square degrees of freedom)
    int32_t n_tiles, trt_tiles;  // tiles [0, trt_tiles) are treated, the rest control
    int32_t T, max_leaves, min_leaf;
    int32_t use_muscale, use_tauscale;
    int32_t fix_sigma, pad_;   // probit BART: the latent's variance is 1, sigma is not drawn
    uint32_t key0, key1;       // Philox key
    double nu, lambda, con_sd, mod_sd;
    ForestDev f[2];
    const double* y;
    long long* rfx;            // FULL residual y - allfit as a fixed-point integer (value * 2^44): updated exactly
    double *Gc, *Gm;           // unscaled forest fits: sum of the mu / tau trees' leaf values
    uint8_t* slots;            // [T][n_pad]
    uint8_t* keys;             // [n_tiles][NB][TILE] keys of the batch in flight (batched engine with the state in HBM)
    TreeRec* trees[2];
    PropRec* proprec;          // [T] proposals of the sweep in progress
    long long* totals;         // [2][T][KEYS][2][4]  (count, limb0, limb1, limb2) by (key, group); see sweep_totals
    long long* cross;          // [2][T][XCAP][2] cross counts of the batched engine, indexed by the batch's first tree
    long long* mom;            // [2 buffers][2 groups][NMOM][3]; sweep s uses buffer s & 1
    Scalars* sc;
    int32_t* err;              // device error flags (ERR_*)
    int32_t* trace;            // [T][5]
    double* trace_logr;        // [T]
    long long* counters;       // [4]
    double *tau_sum, *tau_sq, *mu_sum, *yhat_sum;
    double *draws_tau, *draws_mu, *draws_yhat;   // [nsim][n_pad] or null
    double *sigma_post, *msd_post, *bsd_post;    // [cap]
    int32_t post_cap;
    long long* prof;           // optional [16] cycle counters of CTA 0 (persistent engine), null = off
    int* dbg;                  // host-mapped pinned words: where a kernel gave up before it trapped (survives the context's death)
    // probit BART (null / 0 otherwise): the observed 0/1 response and the caller's row numbers in internal order, the offset
    const uint8_t* ybin;
    const int32_t* orig;
    double probit_offset;
    int32_t probit, pad2_;
};

// The persistent engines alternate between two statistics buffers by sweep parity: every sweep zeroes the buffer it
// used in its finishing pass, and a whole sweep (many grid barriers) passes before that buffer is added to again, so a
// CTA that is still zeroing can never erase the first REDs of a CTA that already started the next sweep.  The graph
// engine (kernel boundaries order everything) always uses buffer 0.  Between sweeps both buffers are all zero.
__device__ __forceinline__ long long* sweep_totals(const Dev& d, uint32_t sweep)
{
    return d.totals + (size_t)(sweep & 1u) * (size_t)d.T * KEYS * 8;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, same addressing as the oracle: counter = (sweep, tree, purpose, idx)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng { uint32_t k0, k1; };

__host__ __device__ inline void rng_u2(Rng g, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx,
                                       double& ua, double& ub)
{
    uint32_t o[4];
    philox4x32_10(sweep, tree, purpose, idx, g.k0, g.k1, o);
    const double inv53 = 1.0 / 9007199254740992.0;
    ua = ((double)(((uint64_t)(o[0] >> 5) << 26) | (uint64_t)(o[1] >> 6)) + 0.5) * inv53;
    ub = ((double)(((uint64_t)(o[2] >> 5) << 26) | (uint64_t)(o[3] >> 6)) + 0.5) * inv53;
}

__device__ inline double rng_normal(Rng g, uint32_t sweep, uint32_t tree, uint32_t purpose, uint32_t idx)
{
    double ua, ub;
    rng_u2(g, sweep, tree, purpose, idx, ua, ub);
    return sqrt(-2.0 * log(ua)) * cos(6.283185307179586476925286766559 * ub);
}

// Gamma(a, 1) by Marsaglia-Tsang; attempt t uses idx 2t (normal) and 2t+1 (uniform), as the oracle does.
__device__ inline double rng_gamma(Rng g, uint32_t sweep, uint32_t purpose, double a)
{
    double boost = 1.0;
    if (a < 1.0) {
        double ua, ub;
        rng_u2(g, sweep, TREE_SWEEP_LEVEL, purpose, 0xFFFFFFFFu, ua, ub);
        boost = pow(ua, 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0;; ++t) {
        const double z = rng_normal(g, sweep, TREE_SWEEP_LEVEL, purpose, 2 * t);
        double ua, ub;
        rng_u2(g, sweep, TREE_SWEEP_LEVEL, purpose, 2 * t + 1, ua, ub);
        const double w = 1.0 + c * z;
        if (w <= 0.0) continue;
        const double v = w * w * w;
        if (log(ua) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
        if (t > 1000000u) return boost * d;
    }
}

// ---------------------------------------------------------------------------------------------
// Fixed point
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long to_fixed(double x, int& bad)
{
    // |x| < 2^15 keeps x * 2^44 below 2^59, so eight of them add up inside int64; NaN compares false and trips the flag
    if (!(fabs(x) < 32768.0)) { bad = 1; return 0; }
    return __double2ll_rn(x * 17592186044416.0);
}
// exact value of limb sums (w0 + w1*2^21 + w2*2^42) / 2^44 rounded once to double
__host__ __device__ inline double limbs_to_double(long long w0, long long w1, long long w2)
{
    __int128 v = (__int128)w0 + ((__int128)w1 << LIMB_BITS) + ((__int128)w2 << (2 * LIMB_BITS));
    const long long hi = (long long)(v >> 40);
    const long long lo = (long long)(v & (((__int128)1 << 40) - 1));
    return ((double)hi * 1099511627776.0 + (double)lo) * (1.0 / 17592186044416.0);
}

// CTA-level statistics table.  There is no native 64-bit shared-memory add on sm_100 (atomicAdd on a shared
// long long compiles to a LDS + ATOMS.CAST.SPIN retry loop, which collapses under contention), so:
//   K <= KREG keys : every warp owns a private (key, group, 4-word) table and adds its REDUX results with plain
//                    read-modify-writes; the tables are summed when the CTA flushes;
//   K >  KREG keys : one table of 32-bit words per (key, group): count + four 16-bit limbs, native ATOMS.ADD.32
//                    (addresses are spread when a tree has many leaves).
constexpr int NWARP_MAX = 16;
template <int NW>
union StatTabT {
    long long w[NW][KREG][2][4];            // (count, limb0, limb1, limb2), 21-bit limbs
    unsigned int d[KEYS][2][8];             // (count, m0, m1, m2, m3), 16-bit limbs; 3 words of padding
};
using StatTab = StatTabT<NWARP_MAX>;        // kernels with 16 warps; the pipelined kernel (28 pass warps) uses StatTabT<32>
constexpr int STATTAB_WORDS64 = (int)(sizeof(StatTab) / 8);
constexpr int ROWSUM_WORDS = KREG * 8;   // (key, group, word) sums over the warps' tables

// split a partial sum (|s| < 2^63) into the three words of the global representation and add them, reduced over the
// warp, to row `row` of the warp's private table
template <class Tab>
__device__ __forceinline__ void warp_add_limbs(Tab& tab, int warp, int row, int g, int lane, long long s, int cnt)
{
    int l0 = (int)(s & 0x1FFFFF), l1 = (int)((s >> LIMB_BITS) & 0x1FFFFF), l2 = (int)(s >> (2 * LIMB_BITS));
    l0 = __reduce_add_sync(0xffffffffu, l0);
    l1 = __reduce_add_sync(0xffffffffu, l1);
    l2 = __reduce_add_sync(0xffffffffu, l2);
    if (lane < 4) {
        const long long v = lane == 0 ? cnt : lane == 1 ? l0 : lane == 2 ? l1 : l2;
        tab.w[warp][row][g][lane] += v;
    }
}

// Accumulate one warp tile (8 observations per lane) of (key, fixed-point residual) pairs.  Keys >= K (padding,
// residual 0) are ignored.
//   K <= KREG: row 0 of the warp's table receives the TOTAL over all keys, rows 1..K-1 their own sums (key 0 is
//              recovered by subtraction when the CTA flushes: integer sums are exact).  Only the last key of a
//              birth proposal (the would-be right child) needs its count; every other count is cached in the tree.
//   K >  KREG: native 32-bit shared atomics per observation.
template <class Tab>
__device__ __forceinline__ void accum_tile(Tab& tab, const int (&key)[8], const long long (&R)[8], int g, int K,
                                           bool count_last, int lane, int warp)
{
    if (K <= KREG) {
        long long tot = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) tot += R[o];
        warp_add_limbs(tab, warp, 0, g, lane, tot, 0);
        int clast = 0;
        if (count_last) {
#pragma unroll
            for (int o = 0; o < 8; ++o) clast += (key[o] == K - 1) ? 1 : 0;
            clast = __reduce_add_sync(0xffffffffu, clast);
        }
        for (int k = 1; k < K; ++k) {
            long long s = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) s += (key[o] == k) ? R[o] : 0ll;
            warp_add_limbs(tab, warp, k, g, lane, s, (k == K - 1) ? clast : 0);
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int k = key[o];
            if (k < K) {
                const long long v = R[o];
                atomicAdd(&tab.d[k][g][0], 1u);
                atomicAdd(&tab.d[k][g][1], (unsigned int)(v & 0xFFFF));
                atomicAdd(&tab.d[k][g][2], (unsigned int)((v >> 16) & 0xFFFF));
                atomicAdd(&tab.d[k][g][3], (unsigned int)((v >> 32) & 0xFFFF));
                atomicAdd(&tab.d[k][g][4], (unsigned int)(int)(v >> 48));
            }
        }
    }
}
// word q (0 count, 1..3 the 21-bit limbs of the global representation) of entry (k, g), summed over the CTA
template <class Tab>
__device__ __forceinline__ long long stat_word(const Tab& tab, int K, int nwarps, int k, int g, int q)
{
    if (K <= KREG) {
        long long v = 0;
        if (k == 0) {
            if (q == 0) return 0;
            for (int w = 0; w < nwarps; ++w) {
                v += tab.w[w][0][g][q];
                for (int kk = 1; kk < K; ++kk) v -= tab.w[w][kk][g][q];
            }
        } else {
            for (int w = 0; w < nwarps; ++w) v += tab.w[w][k][g][q];
        }
        return v;
    }
    const unsigned int* e = tab.d[k][g];
    if (q == 0) return (long long)e[0];
    const __int128 V = (__int128)e[1] + ((__int128)e[2] << 16) + ((__int128)e[3] << 32) + ((__int128)(int)e[4] << 48);
    if (q == 1) return (long long)(V & 0x1FFFFF);
    if (q == 2) return (long long)((V >> LIMB_BITS) & 0x1FFFFF);
    return (long long)(V >> (2 * LIMB_BITS));
}
// plain sum over the warps' tables (moments pass: rows are independent quantities)
template <class Tab>
__device__ __forceinline__ long long table_word(const Tab& tab, int nwarps, int row, int g, int q)
{
    long long v = 0;
    for (int w = 0; w < nwarps; ++w) v += tab.w[w][row][g][q];
    return v;
}

// per-lane partial sum (|s| < 2^62) -> the three 21-bit limbs of the global representation, reduced over the warp
// and added to (count, l0, l1, l2) with native 32-bit shared atomics.  A CTA table word holds at most
// tiles * 32 lanes * 2^21, so a flush is due every 63 tiles at the latest (the resident kernel holds < 32).
__device__ __forceinline__ void warp_add21(unsigned int* e, long long s, int lane)
{
    int l0 = (int)(s & 0x1FFFFF), l1 = (int)((s >> LIMB_BITS) & 0x1FFFFF), l2 = (int)(s >> (2 * LIMB_BITS));
    l0 = __reduce_add_sync(0xffffffffu, l0);
    l1 = __reduce_add_sync(0xffffffffu, l1);
    l2 = __reduce_add_sync(0xffffffffu, l2);
    if (lane < 3) {
        const int v = lane == 0 ? l0 : lane == 1 ? l1 : l2;
        atomicAdd(&e[1 + lane], (unsigned int)v);
    }
}
// word q (0 count, 1..3 the limbs; the top limb is signed) of a CTA table entry
__device__ __forceinline__ long long acc_word(const unsigned int* e, int q)
{
    return q == 3 ? (long long)(int)e[3] : (long long)e[q];
}
// one bit per byte of a pair of 0x00 / 0xFF byte masks -> 8 bits (observation o of the lane = bit o)
__device__ __forceinline__ uint32_t bytes_to_bits(uint32_t e0, uint32_t e1)
{
    return (((e0 & 0x08040201u) * 0x01010101u) >> 24) | (((e1 & 0x80402010u) * 0x01010101u) >> 24);
}

// ---------------------------------------------------------------------------------------------
// Loads: 8 consecutive observations per lane
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8_u8(const uint8_t* p, int (&out)[8])
{
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    out[0] = v.x & 0xFF; out[1] = (v.x >> 8) & 0xFF; out[2] = (v.x >> 16) & 0xFF; out[3] = v.x >> 24;
    out[4] = v.y & 0xFF; out[5] = (v.y >> 8) & 0xFF; out[6] = (v.y >> 16) & 0xFF; out[7] = v.y >> 24;
}
__device__ __forceinline__ void st8_u8(uint8_t* p, const int (&in)[8])
{
    uint2 v;
    v.x = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24);
    v.y = (uint32_t)in[4] | ((uint32_t)in[5] << 8) | ((uint32_t)in[6] << 16) | ((uint32_t)in[7] << 24);
    *reinterpret_cast<uint2*>(p) = v;
}
__device__ __forceinline__ void ld8_u16(const uint16_t* p, int (&out)[8])
{
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    out[0] = v.x & 0xFFFF; out[1] = v.x >> 16; out[2] = v.y & 0xFFFF; out[3] = v.y >> 16;
    out[4] = v.z & 0xFFFF; out[5] = v.z >> 16; out[6] = v.w & 0xFFFF; out[7] = v.w >> 16;
}
__device__ __forceinline__ void ld8_f64(const double* p, double (&out)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2 v = reinterpret_cast<const double2*>(p)[j];
        out[2 * j] = v.x; out[2 * j + 1] = v.y;
    }
}
__device__ __forceinline__ void st8_f64(double* p, const double (&in)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<double2*>(p)[j] = make_double2(in[2 * j], in[2 * j + 1]);
}
__device__ __forceinline__ void ld8_i64(const long long* p, long long (&out)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const longlong2 v = reinterpret_cast<const longlong2*>(p)[j];
        out[2 * j] = v.x; out[2 * j + 1] = v.y;
    }
}
__device__ __forceinline__ void st8_i64(long long* p, const long long (&in)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<longlong2*>(p)[j] = make_longlong2(in[2 * j], in[2 * j + 1]);
}
// the residual of one observation, re-derived from the fits: the ONE expression every engine uses
__device__ __forceinline__ long long resync_fx(double y, double ms, double gc, double b, double gm, int& bad)
{
    return to_fixed(y - (ms * gc + b * gm), bad);
}
// bins of one variable (column base pointer resolved at proposal time) for 8 observations starting at `base`
__device__ __forceinline__ void ld8_bins(const void* col, int is16, int64_t base, int (&out)[8])
{
    if (is16) ld8_u16(reinterpret_cast<const uint16_t*>(col) + base, out);
    else ld8_u8(reinterpret_cast<const uint8_t*>(col) + base, out);
}
__device__ __forceinline__ const void* bin_column(const ForestDev& F, int v, int64_t n_pad, int& is16)
{
    const int ci = F.col_index[v];
    is16 = F.col_is16[v];
    return is16 ? (const void*)(F.bins16 + (int64_t)ci * n_pad) : (const void*)(F.bins8 + (int64_t)ci * n_pad);
}

// ---------------------------------------------------------------------------------------------
// Tree structure helpers (run by single threads on a TreeRec staged in shared memory)
// ---------------------------------------------------------------------------------------------
// usable cutpoint range of variable v at `node` (tree::rg)
template <class Tree>
__device__ inline void tree_rg(const Tree& tr, const ForestDev& F, int node, int v, int& lo, int& hi)
{
    lo = 0; hi = F.ncut[v] - 1;
    int child = node;
    for (int a = tr.parent[node]; a >= 0; child = a, a = tr.parent[a]) {
        if (tr.var[a] != v) continue;
        if (tr.left[a] == child) hi = min(hi, (int)tr.cut[a] - 1);
        else lo = max(lo, (int)tr.cut[a] + 1);
    }
}
// Variables exhausted on the path to `node`, ascending; returns how many.  Every variable has >= 1 cutpoint
// (checked by set_data), so the good variables are exactly the others.
template <class Tree>
__device__ inline int tree_exhausted(const Tree& tr, const ForestDev& F, int node, int (&ex)[MAX_DEPTH + 2])
{
    int nex = 0;
    for (int a = tr.parent[node]; a >= 0; a = tr.parent[a]) {
        const int v = tr.var[a];
        bool seen = false;
        for (int b = tr.parent[node]; b != a; b = tr.parent[b]) if (tr.var[b] == v) { seen = true; break; }
        if (seen) continue;
        int lo, hi;
        tree_rg(tr, F, node, v, lo, hi);
        if (hi < lo) {
            int j = nex++;
            while (j > 0 && ex[j - 1] > v) { ex[j] = ex[j - 1]; --j; }
            ex[j] = v;
        }
    }
    return nex;
}
template <class Tree>
__device__ inline int tree_ngoodvars(const Tree& tr, const ForestDev& F, int node)
{
    if (tr.depth[node] >= MAX_DEPTH) return 0;
    int ex[MAX_DEPTH + 2];
    return F.p - tree_exhausted(tr, F, node, ex);
}
// vi-th good variable in increasing order
template <class Tree>
__device__ inline int tree_pick_var(const Tree& tr, const ForestDev& F, int node, int vi)
{
    int ex[MAX_DEPTH + 2];
    const int nex = tree_exhausted(tr, F, node, ex);
    int v = vi;
    for (int j = 0; j < nex; ++j) if (ex[j] <= v) ++v;
    return v;
}
template <class Tree>
__device__ inline void tree_remove_node(Tree& tr, int i)
{
    const int last = tr.nnodes - 1;
    if (i != last) {
        tr.parent[i] = tr.parent[last]; tr.left[i] = tr.left[last]; tr.right[i] = tr.right[last];
        tr.var[i] = tr.var[last]; tr.cut[i] = tr.cut[last]; tr.node_slot[i] = tr.node_slot[last];
        tr.depth[i] = tr.depth[last]; tr.heap[i] = tr.heap[last];
        const int p = tr.parent[i];
        if (p >= 0) { if (tr.left[p] == last) tr.left[p] = (int16_t)i; else tr.right[p] = (int16_t)i; }
        if (tr.left[i] >= 0) { tr.parent[tr.left[i]] = (int16_t)i; tr.parent[tr.right[i]] = (int16_t)i; }
        else tr.slot_node[tr.node_slot[i]] = (int16_t)i;
    }
    tr.nnodes = last;
}
__host__ __device__ inline double lil(double sum_phi, double sum_phi_r, double tau)
{
    const double d = 1.0 / (tau * tau) + sum_phi;
    return -log(tau) - 0.5 * log(d) + 0.5 * sum_phi_r * sum_phi_r / d;
}
// left-aligned heap id: orders prefix-free node sets (leaves; nogs) left-to-right, i.e. in DFS order
__device__ inline uint32_t dfs_key(uint32_t heap, int depth) { return heap << (MAX_DEPTH - depth); }

}  // namespace bcf
