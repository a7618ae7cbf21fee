// bcf_glm.cuh -- the dense pre-steps of bcf_iv() / bcf() on the device: what the reference does on the host
// before the samplers start, and what dominates a call's wall time at n >= 1e6 once the samplers run on a GPU:
//   stats::glm(z ~ x, family = binomial) + predict(type = "link")   R/bcf_iv.R:72-75, R/bcf_itt.R:58-61   -> IRLS
//   summary(lm(yscale ~ z + x_c))$sigma                            bcf's R wrapper           -> normal equations
// Both reduce to the weighted Gram matrix  A' W A  of an n x m column-major matrix (m ~ 100: the one small dense contraction
// on the path) plus O(n m) streaming passes.  fp64 on CUDA cores (the answer feeds a 10 000-point cutpoint grid and a
// convergence test at 1e-8: no reduced precision), deterministic: every block writes its partial sums and a second kernel
// adds them in block order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bcf {

constexpr int GRAM_RC = 32;          // rows per shared-memory tile
constexpr int GRAM_THREADS = 512;
constexpr int GRAM_TB = 4;           // a thread owns 4 x 4 blocks of the Gram matrix
constexpr int GRAM_MAX_M = 352;      // 88 column blocks; a launch takes up to 1024 block pairs, the host loops over the rest

// The augmented matrix A = [1 | z? | X | t]: the intercept, optionally a 0/1 column (lm's z), the p columns of X, then an
// extra column t (IRLS's working response, or lm's y).  Column c of row r:
struct AugMat {
    const double* X;       // n x p column-major
    const int32_t* zcol;   // nullptr: no such column
    const double* t;
    int64_t n;
    int p, m;              // m = 1 + (zcol != nullptr) + p + 1
    __device__ __forceinline__ double at(int c, int64_t r) const
    {
        if (c == 0) return 1.0;
        int k = c - 1;
        if (zcol != nullptr) { if (k == 0) return (double)zcol[r]; --k; }
        return k < p ? X[(int64_t)k * n + r] : t[r];
    }
};

// partial[block][q][16]: sum over the block's rows of w_r * A[r][4 bi + a] * A[r][4 bj + b] for the q-th pair bi <= bj of
// 4-column blocks, for the pairs q0 <= q < q0 + nq of this launch.  Every thread keeps KB block pairs (16 accumulators
// each) in registers; per row it reads 2 x 4 values (two 128-bit shared-memory loads per block) for 16 multiply-adds.
template <int KB>
__global__ void __launch_bounds__(GRAM_THREADS) k_wgram(AugMat A, const double* __restrict__ w, int mpad, int nbp, int q0,
                                                        int nq, const uint16_t* __restrict__ pair_ij, double* __restrict__ partial)
{
    extern __shared__ double As[];                    // [GRAM_RC][mpad] row-major, then [GRAM_RC] weights
    double* ws = As + GRAM_RC * mpad;
    const int tid = threadIdx.x;
    double acc[KB][16];
    int bi[KB], bj[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[k][e] = 0.0;
        const int q = tid + k * GRAM_THREADS;
        bi[k] = q < nq ? 4 * (int)pair_ij[2 * (q0 + q)] : 0;
        bj[k] = q < nq ? 4 * (int)pair_ij[2 * (q0 + q) + 1] : 0;
    }
    const int64_t ntiles = (A.n + GRAM_RC - 1) / GRAM_RC;
    const int64_t t_lo = ntiles * blockIdx.x / gridDim.x, t_hi = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int64_t tl = t_lo; tl < t_hi; ++tl) {
        const int64_t r0 = tl * GRAM_RC;
        __syncthreads();
        for (int e = tid; e < mpad * GRAM_RC; e += GRAM_THREADS) {     // a warp reads 32 consecutive rows of one column
            const int c = e / GRAM_RC, rr = e % GRAM_RC;
            As[rr * mpad + c] = (c < A.m && r0 + rr < A.n) ? A.at(c, r0 + rr) : 0.0;
        }
        if (tid < GRAM_RC) ws[tid] = (r0 + tid < A.n) ? (w != nullptr ? w[r0 + tid] : 1.0) : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (tid + k * GRAM_THREADS >= nq) break;
            for (int rr = 0; rr < GRAM_RC; ++rr) {
                const double* a = As + rr * mpad;
                const double wr = ws[rr];
                const double2 i01 = *(const double2*)(a + bi[k]), i23 = *(const double2*)(a + bi[k] + 2);
                const double2 j01 = *(const double2*)(a + bj[k]), j23 = *(const double2*)(a + bj[k] + 2);
                const double wi[4] = {wr * i01.x, wr * i01.y, wr * i23.x, wr * i23.y};
                const double vj[4] = {j01.x, j01.y, j23.x, j23.y};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[k][4 * x + y] += wi[x] * vj[y];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int q = tid + k * GRAM_THREADS;
        if (q < nq) {
            double* o = partial + ((size_t)blockIdx.x * nbp + q0 + q) * 16;
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = acc[k][e];
        }
    }
}
// out[q] = sum over blocks (in block order) of partial[b][q]
__global__ void k_sum_partials(const double* __restrict__ partial, int nblocks, int nq, double* __restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nq + q];
    out[q] = s;
}
// eta = beta[0] + beta[1] zcol + X beta[..]  (columns whose coefficient is exactly 0 -- aliased, dropped -- are skipped);
// with y != nullptr the block's partial sum of (y - eta)^2 is written as well (lm's residual sum of squares)
__global__ void __launch_bounds__(256) k_linpred(AugMat A, const double* __restrict__ beta, double* __restrict__ eta,
                                                 const double* __restrict__ y, double* __restrict__ rss_partial)
{
    __shared__ double red[256];
    __shared__ double bs[GRAM_MAX_M];
    const int nb = A.m - 1;                                            // coefficients: every column but t
    for (int j = threadIdx.x; j < nb; j += blockDim.x) bs[j] = beta[j];
    __syncthreads();
    const int x0 = 1 + (A.zcol != nullptr);
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.n; r += (int64_t)gridDim.x * blockDim.x) {
        double s = bs[0];
        if (A.zcol != nullptr) s += bs[1] * (double)A.zcol[r];
        for (int j = 0; j < A.p; ++j) { const double b = bs[x0 + j]; if (b != 0.0) s += b * A.X[(int64_t)j * A.n + r]; }
        if (eta != nullptr) eta[r] = s;
        if (y != nullptr) { const double e = y[r] - s; acc += e * e; }
    }
    if (y == nullptr) return;
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) rss_partial[blockIdx.x] = red[0];
}
// One IRLS step's element-wise part, as stats::glm.fit does it for binomial(link = "logit"): mu = linkinv(eta) with eta
// clamped to +-log(1/DBL_EPSILON), mu.eta = max(mu (1 - mu), DBL_EPSILON), working weight w = mu.eta^2 / variance(mu)
// = mu.eta, working response t = eta + (z - mu) / mu.eta, and the deviance of the CURRENT eta (-2 log of the fitted
// probability of the observed class) as per-block partial sums.  first != 0: eta is glm.fit's start, logit((z + 0.5) / 2).
__global__ void __launch_bounds__(256) k_irls_update(double* __restrict__ eta, const int32_t* __restrict__ z, int64_t n, int first,
                                                     double* __restrict__ w, double* __restrict__ t, double* dev_partial)
{
    __shared__ double red[256];
    const double thresh = 36.04365338911715;        // -log(DBL_EPSILON)
    const double eps = 2.220446049250313e-16;
    double dev = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int zi = z[r];
        double e = first ? (zi ? 1.0986122886681098 : -1.0986122886681098) : eta[r];
        if (first) eta[r] = e;
        const double e0 = e;                                             // glm.fit's working response uses eta as it is
        e = fmin(fmax(e, -thresh), thresh);
        const double ex = exp(-fabs(e));                                 // in (0, 1]
        const double p_small = ex / (1.0 + ex), p_big = 1.0 / (1.0 + ex);
        const double mu = e >= 0.0 ? p_big : p_small;
        const double me = fmax(ex / ((1.0 + ex) * (1.0 + ex)), eps);
        w[r] = me;
        t[r] = e0 + ((double)zi - mu) / me;
        const double pobs = (zi != 0) == (e >= 0.0) ? p_big : p_small;   // fitted probability of the observed class
        dev += -2.0 * log(pobs);
    }
    red[threadIdx.x] = dev;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) dev_partial[blockIdx.x] = red[0];
}

// out[c][r] = in[r][c]: a row-major n x p matrix (NumPy's default) to the column-major layout the passes above stream
__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ in, int64_t n, int p, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;             // 32 x 8
    for (int k = ty; k < 32; k += 8)
        if (r0 + k < n && c0 + tx < p) tile[k][tx] = in[(r0 + k) * p + c0 + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
        if (c0 + k < p && r0 + tx < n) out[(int64_t)(c0 + k) * n + r0 + tx] = tile[tx][k];
}

}  // namespace bcf
