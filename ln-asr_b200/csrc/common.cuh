// common.cuh -- shared helpers for the lnasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/lnasr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "lnasr_b200 is written for sm_100a (B200) only"
#endif

namespace lnasr {

void set_error(const char* fmt, ...);

#define LNASR_CHECK_ARG(cond, ...)                         \
    do {                                                   \
        if (!(cond)) {                                     \
            ::lnasr::set_error(__VA_ARGS__);               \
            return LNASR_EINVAL;                           \
        }                                                  \
    } while (0)

#define LNASR_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ::lnasr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                \
                               cudaGetErrorString(e__));                                   \
            return LNASR_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

// every kernel launch of the library is followed by this macro: it also counts them (lnasr_launch_count: the benchmark's
// "gpu_launches" is this counter's increase over the timed region, not an estimate)
void count_launch();
#define LNASR_LAUNCH_CHECK()                 \
    do {                                     \
        ::lnasr::count_launch();             \
        LNASR_CUDA(cudaGetLastError());      \
    } while (0)

int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
__device__ __forceinline__ float neg_inf_f() { return __int_as_float(0xff800000); }

// utterance of a global frame index: largest u with off[u] <= f   (off has n+1 entries)
__device__ __forceinline__ int find_utt(const int64_t* __restrict__ off, int n, int64_t f) {
    int lo = 0, hi = n;  // invariant: off[lo] <= f < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= f) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename T>
__device__ __forceinline__ double load_as_double(const void* p, int64_t i);
template <>
__device__ __forceinline__ double load_as_double<float>(const void* p, int64_t i) {
    return (double)__ldg((const float*)p + i);
}
template <>
__device__ __forceinline__ double load_as_double<double>(const void* p, int64_t i) {
    return __ldg((const double*)p + i);
}

__device__ __forceinline__ double shfl_d(double v, int src, int width = 32) {
    return __shfl_sync(0xffffffffu, v, src, width);
}

// log(exp(a)+exp(b)) safe for -inf
__device__ __forceinline__ double logaddexp_d(double a, double b) {
    double mx = fmax(a, b), mn = fmin(a, b);
    if (mx == neg_inf()) return mx;
    return mx + log1p(exp(mn - mx));
}

// ln(x), table driven (used for log-mel energies and the mixture log-sum-exp): x = 2^E m, m in [1,2); table entry i = top 7 mantissa bits holds c = 1/(1+(i+1/2)/128)
// and -ln c; r = m c - 1 has |r| < 2^-8 and ln(1+r) is a degree-6 polynomial (error r^7/7 < 3e-18).  8 DFMA instead
// of the ~25 of log().  Zero, subnormal, inf and NaN go through log().
__device__ __forceinline__ bool fast_log_ok(double x) { return x >= 1e-300 && x <= 1e300; }
__device__ __forceinline__ double fast_log_core(double x, const double2* __restrict__ tab);
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    if (!fast_log_ok(x)) return log(x);
    return fast_log_core(x, tab);
}
// the branch-free part (x normal, positive): callers with several logarithms test fast_log_ok on all of them first, so that the
// compiler can run the dependency chains side by side
__device__ __forceinline__ double fast_log_core(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x);
    const int E = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 ct = tab[(hi >> 13) & 127];
    const double r = fma(m, ct.x, -1.0);
    double q = -1.0 / 6.0;
    q = fma(q, r, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    q = fma(q, r, 1.0);
    return fma(r, q, fma((double)E, 0.69314718055994530942, ct.y));
}


// exp(d) for d <= 0 (the terms of a log-sum-exp after subtracting the maximum): d = (n/64) ln2 + r, |r| <= ln2/128,
// exp(d) = 2^(n>>6) * T[n & 63] * P5(r); relative error ~2e-16, 10 DFMA instead of the ~20 of exp().  T = 2^(i/64).
__device__ __forceinline__ double fast_exp_neg(double d, const double* __restrict__ t64) {
    // branch-free: anything below e^-700 (and -inf, and the NaN of -inf - -inf) gives 0 through the final select
    const bool ok = d > -700.0;
    const double dc = ok ? d : -700.0;
    const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52: the integer lands in the low word
    double nf = fma(dc, 92.33248261689366, MAGIC);                  // 64 / ln 2
    const int n = __double2loint(nf);
    nf -= MAGIC;
    double r = fma(nf, -0.01083042469326756, dc);                   // ln2/64, high part (32 significant bits)
    r = fma(nf, -2.9815858269852933e-12, r);                        //         low part
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = t64[n & 63] * p;
    const double v = __hiloint2double(__double2hiint(res) + ((n >> 6) << 20), __double2loint(res));
    return ok ? v : 0.0;
}

// block-wide set-up of the two tables above: log_tab[128] = (1/c_i, -ln c_i), exp_tab[64] = 2^(i/64)
__device__ __forceinline__ void fill_log_exp_tables(double2* log_tab, double* exp_tab, int tid, int n_threads) {
    for (int i = tid; i < 128; i += n_threads) {
        const double c = 1.0 / (1.0 + ((double)i + 0.5) / 128.0);
        log_tab[i] = make_double2(c, -log(c));
    }
    if (exp_tab)
        for (int i = tid; i < 64; i += n_threads) exp_tab[i] = exp2((double)i / 64.0);
}

}  // namespace lnasr
