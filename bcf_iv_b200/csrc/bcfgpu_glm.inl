// bcfgpu_glm.inl -- host side of the dense pre-steps, included at the end of bcfgpu.cu:
//   bcfgpu_glm_logit   stats::glm(z ~ x, family = binomial) + predict(type = "link")   R/bcf_iv.R:72-75, R/bcf_itt.R:58-61
//   bcfgpu_lm_sigma    summary(lm(yscale ~ z + x_c))$sigma, bcf's default `sighat`       bcf's R wrapper
// The O(n m^2) and O(n m) work runs in the kernels of bcf_glm.cuh; the m x m solve (m ~ 100) is done here on the host.
#include "bcf_glm.cuh"

namespace {

struct GlmWork {
    int device = 0;
    cudaStream_t s = nullptr, cs = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::vector<Block> dev, pin;
    int64_t launches = 0;
    ~GlmWork()
    {
        if (s) cudaStreamSynchronize(s);
        if (cs) cudaStreamSynchronize(cs);
        for (Block b : dev) dev_free(device, b);
        for (Block b : pin) pin_free(b);
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
        if (s) cudaStreamDestroy(s);
        if (cs) cudaStreamDestroy(cs);
    }
    template <typename Tp> int alloc(Tp** p, size_t count)
    {
        Block b;
        cudaError_t e = dev_alloc(device, std::max<size_t>(count * sizeof(Tp), 16), &b);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(BCFGPU_ERR_OOM, std::string("device memory for the dense pre-step: ") + cudaGetErrorString(e)); }
        dev.push_back(b);
        *p = (Tp*)b.p;
        return BCFGPU_OK;
    }
};

// host -> device on the copy stream: straight from pinned memory, through two pinned staging buffers filled by host
// threads from pageable memory (what R and NumPy hand over)
int upload(GlmWork& W, void* dst, const void* src, size_t bytes)
{
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, src) == cudaSuccess && (pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (pinned || bytes <= ((size_t)1 << 20)) { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, W.cs)); return BCFGPU_OK; }
    const size_t chunk = (size_t)32 << 20;
    if (W.pin.empty()) {
        for (int b = 0; b < 2; ++b) {
            Block blk;
            if (pin_alloc(chunk, &blk) != cudaSuccess) { cudaGetLastError(); return fail(BCFGPU_ERR_OOM, "pinned staging buffer"); }
            W.pin.push_back(blk);
        }
    }
    int k = 0;
    for (size_t off = 0; off < bytes; off += chunk, ++k) {
        const int b = k & 1;
        const size_t nb = std::min(chunk, bytes - off);
        if (k >= 2) CK(cudaEventSynchronize(W.ev[b]));
        char* stage = (char*)W.pin[b].p;
        const char* from = (const char*)src + off;
        parallel_ranges((int64_t)nb, 1 << 16, host_threads(), [&](int, int64_t lo, int64_t hi) { memcpy(stage + lo, from + lo, (size_t)(hi - lo)); });
        CK(cudaMemcpyAsync((char*)dst + off, stage, nb, cudaMemcpyHostToDevice, W.cs));
        CK(cudaEventRecord(W.ev[b], W.cs));
    }
    return BCFGPU_OK;
}

int glm_open(GlmWork& W, int32_t device)
{
    const int nd = bcfgpu_device_count();
    if (nd == 0) return fail(BCFGPU_ERR_CUDA, "no CUDA device available (libbcfgpu has no CPU fallback)");
    if (device < 0 || device >= nd) return fail(BCFGPU_ERR_BAD_ARG, "device out of range");
    W.device = device;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&W.s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&W.cs, cudaStreamNonBlocking));
    for (auto& e : W.ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return BCFGPU_OK;
}

// X on the device, column-major (a row-major caller's matrix is transposed there)
int glm_put_matrix(GlmWork& W, int64_t n, int32_t p, const double* X, int row_major, double** dX)
{
    const size_t cnt = (size_t)n * (size_t)std::max(p, 1);
    int rc = W.alloc(dX, cnt);
    if (rc) return rc;
    if (p == 0) return BCFGPU_OK;
    double* raw = *dX;
    if (row_major && p > 1) { rc = W.alloc(&raw, cnt); if (rc) return rc; }
    rc = upload(W, raw, X, cnt * 8);
    if (rc) return rc;
    CK(cudaStreamSynchronize(W.cs));
    if (raw != *dX) {
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((p + 31) / 32));
        k_transpose<<<grid, 256, 0, W.s>>>(raw, n, p, *dX);
        W.launches++;
        CK(cudaGetLastError());
    }
    return BCFGPU_OK;
}

struct GramPlan {
    int m = 0, mpad = 0, mb = 0, nbp = 0, grid = 1;
    size_t smem = 0;
    uint16_t* d_pairs = nullptr;
    double *d_partial = nullptr, *d_G = nullptr;
    std::vector<uint16_t> pairs;
    std::vector<double> hG;       // nbp x 16 as reduced
};

int gram_plan(GlmWork& W, GramPlan& P, int m, int64_t n)
{
    if (m > GRAM_MAX_M) return fail(BCFGPU_ERR_BAD_ARG, "too many columns for the device pre-step (p <= " + std::to_string(GRAM_MAX_M - 3) + ")");
    P.m = m; P.mpad = (m + 3) / 4 * 4; P.mb = P.mpad / 4; P.nbp = P.mb * (P.mb + 1) / 2;
    for (int i = 0; i < P.mb; ++i) for (int j = i; j < P.mb; ++j) { P.pairs.push_back((uint16_t)i); P.pairs.push_back((uint16_t)j); }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, W.device));
    const int64_t ntiles = (n + GRAM_RC - 1) / GRAM_RC;
    P.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, 2 * (int64_t)prop.multiProcessorCount));
    P.smem = ((size_t)GRAM_RC * P.mpad + GRAM_RC) * sizeof(double);
    int rc = W.alloc(&P.d_pairs, P.pairs.size());
    if (!rc) rc = W.alloc(&P.d_partial, (size_t)P.grid * P.nbp * 16);
    if (!rc) rc = W.alloc(&P.d_G, (size_t)P.nbp * 16);
    if (rc) return rc;
    CK(cudaMemcpyAsync(P.d_pairs, P.pairs.data(), P.pairs.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, W.s));
    P.hG.resize((size_t)P.nbp * 16);
    return BCFGPU_OK;
}

template <int KB>
int gram_launch_kb(GlmWork& W, GramPlan& P, const AugMat& A, const double* d_w, int q0, int nq)
{
    CK(cudaFuncSetAttribute(k_wgram<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)GRAM_RC * GRAM_MAX_M + GRAM_RC) * sizeof(double))));
    k_wgram<KB><<<P.grid, GRAM_THREADS, P.smem, W.s>>>(A, d_w, P.mpad, P.nbp, q0, nq, P.d_pairs, P.d_partial);
    W.launches++;
    return BCFGPU_OK;
}
// G (m x m, symmetric, row-major, on the host) = A' diag(w) A
int gram(GlmWork& W, GramPlan& P, const AugMat& A, const double* d_w, std::vector<double>& G)
{
    for (int q0 = 0; q0 < P.nbp; q0 += 2 * GRAM_THREADS) {                // up to 1024 block pairs per launch
        const int nq_l = std::min(2 * GRAM_THREADS, P.nbp - q0);
        const int rc = nq_l <= GRAM_THREADS ? gram_launch_kb<1>(W, P, A, d_w, q0, nq_l) : gram_launch_kb<2>(W, P, A, d_w, q0, nq_l);
        if (rc) return rc;
        CK(cudaGetLastError());
    }
    const int nq = P.nbp * 16;
    k_sum_partials<<<(nq + 127) / 128, 128, 0, W.s>>>(P.d_partial, P.grid, nq, P.d_G);
    CK(cudaGetLastError());
    W.launches++;
    CK(cudaMemcpyAsync(P.hG.data(), P.d_G, (size_t)nq * 8, cudaMemcpyDeviceToHost, W.s));
    CK(cudaStreamSynchronize(W.s));
    G.assign((size_t)P.m * P.m, 0.0);
    for (int q = 0; q < P.nbp; ++q) {
        const int bi = 4 * P.pairs[2 * q], bj = 4 * P.pairs[2 * q + 1];
        for (int x = 0; x < 4; ++x)
            for (int y = 0; y < 4; ++y) {
                const int i = bi + x, j = bj + y;
                if (i >= P.m || j >= P.m || j < i) continue;
                G[(size_t)i * P.m + j] = G[(size_t)j * P.m + i] = P.hG[(size_t)q * 16 + 4 * x + y];
            }
    }
    return BCFGPU_OK;
}

// Least squares from the normal equations, columns taken left to right the way R's dqrdc2 does (limited pivoting): a
// column whose squared residual norm after projection on the kept columns before it is below tol * its squared norm is
// ALIASED -- it gets no coefficient (NaN, R's NA) and does not enter the fit.  G is (k + 1) x (k + 1) with the response as
// its last row / column.  Returns the rank.
int solve_aliased(int k, const std::vector<double>& G, double tol, std::vector<double>& beta)
{
    const int m = k + 1;
    std::vector<double> L((size_t)m * m, 0.0);     // rows: all columns incl. the response; columns: kept order
    std::vector<int> kept;
    for (int j = 0; j < m; ++j) {
        double d = G[(size_t)j * m + j];
        for (size_t a = 0; a < kept.size(); ++a) {
            const int ka = kept[a];
            double v = G[(size_t)j * m + ka];
            for (size_t c = 0; c < a; ++c) v -= L[(size_t)j * m + c] * L[(size_t)ka * m + c];
            v /= L[(size_t)ka * m + a];
            L[(size_t)j * m + a] = v;
            d -= v * v;
        }
        if (j == k) break;                                               // the response's row is the right-hand side
        if (d > tol * G[(size_t)j * m + j] && d > 0.0) { L[(size_t)j * m + kept.size()] = std::sqrt(d); kept.push_back(j); }
    }
    beta.assign(k, std::nan(""));
    const int r = (int)kept.size();
    std::vector<double> x(r);
    for (int a = r - 1; a >= 0; --a) {                                   // L' x = (response's row of L)
        double v = L[(size_t)k * m + a];
        for (int c = a + 1; c < r; ++c) v -= L[(size_t)kept[c] * m + a] * x[c];
        x[a] = v / L[(size_t)kept[a] * m + a];
    }
    for (int a = 0; a < r; ++a) beta[kept[a]] = x[a];
    return r;
}

double ordered_sum(const std::vector<double>& v) { double s = 0.0; for (double x : v) s += x; return s; }

constexpr double ALIAS_TOL = 1e-11;   // on SQUARED norms (R: 1e-11 / 1e-7 on norms; below ~1e-13 the Gram form cannot tell)

}  // namespace

extern "C" int bcfgpu_glm_logit(int32_t device, int64_t n, int32_t p, const double* X, int32_t row_major, const int32_t* z,
                                int32_t max_iter, double epsilon, double* eta_out, double* beta_out, double* deviance_out,
                                int32_t* info)
{
    if (n < 1 || p < 0 || (p > 0 && !X) || !z || !eta_out) return fail(BCFGPU_ERR_BAD_ARG, "glm_logit: bad argument");
    if (max_iter < 1) max_iter = 25;
    if (!(epsilon > 0.0)) epsilon = 1e-8;
    for (int64_t i = 0; i < n; ++i) if (z[i] != 0 && z[i] != 1) return fail(BCFGPU_ERR_BAD_ARG, "glm_logit: z must be 0 / 1");
    GlmWork W;
    int rc = glm_open(W, device);
    if (rc) return rc;
    double *dX = nullptr, *d_eta = nullptr, *d_w = nullptr, *d_t = nullptr, *d_dev = nullptr, *d_beta = nullptr;
    int32_t* d_z = nullptr;
    const int m = p + 2, k = p + 1;
    GramPlan P;
    rc = gram_plan(W, P, m, n);
    if (!rc) rc = W.alloc(&d_z, (size_t)n);
    if (!rc) rc = W.alloc(&d_eta, (size_t)n);
    if (!rc) rc = W.alloc(&d_w, (size_t)n);
    if (!rc) rc = W.alloc(&d_t, (size_t)n);
    if (!rc) rc = W.alloc(&d_beta, (size_t)m);
    if (rc) return rc;
    const int eg = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4 * 148));
    rc = W.alloc(&d_dev, (size_t)eg);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_z, z, (size_t)n * 4, cudaMemcpyHostToDevice, W.s));
    rc = glm_put_matrix(W, n, p, X, row_major, &dX);
    if (rc) return rc;
    const AugMat A{dX, nullptr, d_t, n, p, m};
    std::vector<double> hdev(eg), G, beta;
    auto update = [&](int first, double* dev) -> int {
        k_irls_update<<<eg, 256, 0, W.s>>>(d_eta, d_z, n, first, d_w, d_t, d_dev);
        W.launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hdev.data(), d_dev, (size_t)eg * 8, cudaMemcpyDeviceToHost, W.s));
        CK(cudaStreamSynchronize(W.s));
        *dev = ordered_sum(hdev);
        return BCFGPU_OK;
    };
    double devold = 0.0, dev = 0.0;
    rc = update(1, &devold);
    if (rc) return rc;
    int it = 0, conv = 0, rank = 0;
    std::vector<double> bdev(m, 0.0);
    for (it = 1; it <= max_iter; ++it) {
        rc = gram(W, P, A, d_w, G);
        if (rc) return rc;
        rank = solve_aliased(k, G, ALIAS_TOL, beta);
        for (int j = 0; j < k; ++j) bdev[j] = std::isnan(beta[j]) ? 0.0 : beta[j];
        for (int j = 0; j < k; ++j) if (!std::isfinite(bdev[j])) return fail(BCFGPU_ERR_NUMERIC, "glm_logit: non-finite coefficient");
        CK(cudaMemcpyAsync(d_beta, bdev.data(), (size_t)k * 8, cudaMemcpyHostToDevice, W.s));
        k_linpred<<<eg, 256, 0, W.s>>>(A, d_beta, d_eta, nullptr, nullptr);
        W.launches++;
        CK(cudaGetLastError());
        rc = update(0, &dev);
        if (rc) return rc;
        if (!std::isfinite(dev)) return fail(BCFGPU_ERR_NUMERIC, "glm_logit: the deviance is not finite");
        if (std::fabs(dev - devold) / (std::fabs(dev) + 0.1) < epsilon) { conv = 1; break; }
        devold = dev;
    }
    CK(cudaMemcpyAsync(eta_out, d_eta, (size_t)n * 8, cudaMemcpyDeviceToHost, W.s));
    CK(cudaStreamSynchronize(W.s));
    if (beta_out) for (int j = 0; j < k; ++j) beta_out[j] = beta[j];
    if (deviance_out) *deviance_out = dev;
    if (info) { info[0] = std::min(it, max_iter); info[1] = conv; info[2] = rank; info[3] = (int32_t)W.launches; }
    return BCFGPU_OK;
}

extern "C" int bcfgpu_lm_sigma(int32_t device, int64_t n, int32_t p, const double* X, int32_t row_major, const int32_t* z,
                               const double* y, double* sigma_out, double* beta_out, int32_t* info)
{
    if (n < 2 || p < 0 || (p > 0 && !X) || !y || !sigma_out) return fail(BCFGPU_ERR_BAD_ARG, "lm_sigma: bad argument");
    GlmWork W;
    int rc = glm_open(W, device);
    if (rc) return rc;
    const int nz = z ? 1 : 0, k = 1 + nz + p, m = k + 1;
    double *dX = nullptr, *d_y = nullptr, *d_beta = nullptr, *d_rss = nullptr;
    int32_t* d_z = nullptr;
    GramPlan P;
    rc = gram_plan(W, P, m, n);
    if (!rc) rc = W.alloc(&d_y, (size_t)n);
    if (!rc) rc = W.alloc(&d_beta, (size_t)m);
    if (!rc && z) rc = W.alloc(&d_z, (size_t)n);
    const int eg = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4 * 148));
    if (!rc) rc = W.alloc(&d_rss, (size_t)eg);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_y, y, (size_t)n * 8, cudaMemcpyHostToDevice, W.s));
    if (z) CK(cudaMemcpyAsync(d_z, z, (size_t)n * 4, cudaMemcpyHostToDevice, W.s));
    rc = glm_put_matrix(W, n, p, X, row_major, &dX);
    if (rc) return rc;
    const AugMat A{dX, d_z, d_y, n, p, m};
    std::vector<double> G, beta, bdev(m, 0.0), hrss(eg);
    rc = gram(W, P, A, nullptr, G);
    if (rc) return rc;
    const int rank = solve_aliased(k, G, ALIAS_TOL, beta);
    for (int j = 0; j < k; ++j) bdev[j] = std::isnan(beta[j]) ? 0.0 : beta[j];
    CK(cudaMemcpyAsync(d_beta, bdev.data(), (size_t)k * 8, cudaMemcpyHostToDevice, W.s));
    k_linpred<<<eg, 256, 0, W.s>>>(A, d_beta, nullptr, d_y, d_rss);
    W.launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hrss.data(), d_rss, (size_t)eg * 8, cudaMemcpyDeviceToHost, W.s));
    CK(cudaStreamSynchronize(W.s));
    const double rss = ordered_sum(hrss);
    const int64_t dof = std::max<int64_t>(n - rank, 1);
    *sigma_out = std::sqrt(rss / (double)dof);
    if (beta_out) for (int j = 0; j < k; ++j) beta_out[j] = beta[j];
    if (info) { info[0] = rank; info[1] = (int32_t)W.launches; }
    return BCFGPU_OK;
}
