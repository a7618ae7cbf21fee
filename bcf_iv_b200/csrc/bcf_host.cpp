// bcf_host.cpp -- host-side passes of the library that are worth vector code (compiled by the host compiler alone).
#include "bcf_host.h"

#include <immintrin.h>
#include <cstring>

namespace bcf {

static void scan_scalar(const double* x, int64_t i0, int64_t n, double& lo, double& hi, uint64_t& bad)
{
    for (int64_t i = i0; i < n; ++i) {
        const double v = x[i];
        uint64_t b;
        memcpy(&b, &v, 8);
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
        bad |= (uint64_t)((b & 0x7FF0000000000000ull) == 0x7FF0000000000000ull);
    }
}
static uint64_t other_scalar(const double* x, int64_t i0, int64_t n, double lo, double hi)
{
    uint64_t other = 0;
    for (int64_t i = i0; i < n; ++i) other |= (uint64_t)((x[i] != lo) & (x[i] != hi));
    return other;
}

__attribute__((target("avx2"))) static void scan_avx2(const double* x, int64_t n, double* lo_out, double* hi_out, uint8_t* flag_out)
{
    const __m256i expm = _mm256_set1_epi64x(0x7FF0000000000000ll);
    __m256d lo0 = _mm256_set1_pd(x[0]), lo1 = lo0, hi0 = lo0, hi1 = lo0;
    __m256i bad0 = _mm256_setzero_si256(), bad1 = bad0;
    const int64_t nv = n / 8 * 8;
    for (int64_t i = 0; i < nv; i += 8) {
        const __m256d a = _mm256_loadu_pd(x + i), b = _mm256_loadu_pd(x + i + 4);
        lo0 = _mm256_min_pd(lo0, a); lo1 = _mm256_min_pd(lo1, b);
        hi0 = _mm256_max_pd(hi0, a); hi1 = _mm256_max_pd(hi1, b);
        bad0 = _mm256_or_si256(bad0, _mm256_cmpeq_epi64(_mm256_and_si256(_mm256_castpd_si256(a), expm), expm));
        bad1 = _mm256_or_si256(bad1, _mm256_cmpeq_epi64(_mm256_and_si256(_mm256_castpd_si256(b), expm), expm));
    }
    double l[4], h[4];
    _mm256_storeu_pd(l, _mm256_min_pd(lo0, lo1));
    _mm256_storeu_pd(h, _mm256_max_pd(hi0, hi1));
    double lo = l[0], hi = h[0];
    for (int k = 1; k < 4; ++k) { lo = l[k] < lo ? l[k] : lo; hi = h[k] > hi ? h[k] : hi; }
    uint64_t bad = _mm256_testz_si256(_mm256_or_si256(bad0, bad1), _mm256_set1_epi64x(-1)) ? 0 : 1;
    scan_scalar(x, nv, n, lo, hi, bad);
    // second pass (the column is 4-8 MB: mostly still in cache): is every value lo or hi?
    const __m256d vlo = _mm256_set1_pd(lo), vhi = _mm256_set1_pd(hi);
    __m256d o0 = _mm256_setzero_pd(), o1 = o0;
    for (int64_t i = 0; i < nv; i += 8) {
        const __m256d a = _mm256_loadu_pd(x + i), b = _mm256_loadu_pd(x + i + 4);
        o0 = _mm256_or_pd(o0, _mm256_and_pd(_mm256_cmp_pd(a, vlo, _CMP_NEQ_UQ), _mm256_cmp_pd(a, vhi, _CMP_NEQ_UQ)));
        o1 = _mm256_or_pd(o1, _mm256_and_pd(_mm256_cmp_pd(b, vlo, _CMP_NEQ_UQ), _mm256_cmp_pd(b, vhi, _CMP_NEQ_UQ)));
    }
    uint64_t other = _mm256_movemask_pd(_mm256_or_pd(o0, o1)) ? 1 : 0;
    other |= other_scalar(x, nv, n, lo, hi);
    *lo_out = lo; *hi_out = hi;
    *flag_out = (uint8_t)((other == 0 && !bad ? 1 : 0) | (bad ? 2 : 0));
}

void scan_column(const double* x, int64_t n, double* lo_out, double* hi_out, uint8_t* flag_out)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { scan_avx2(x, n, lo_out, hi_out, flag_out); return; }
    double lo = x[0], hi = x[0];
    uint64_t bad = 0;
    scan_scalar(x, 0, n, lo, hi, bad);
    const uint64_t other = other_scalar(x, 0, n, lo, hi);
    *lo_out = lo; *hi_out = hi;
    *flag_out = (uint8_t)((other == 0 && !bad ? 1 : 0) | (bad ? 2 : 0));
}

}  // namespace bcf
