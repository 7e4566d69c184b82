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

#define LNASR_LAUNCH_CHECK() LNASR_CUDA(cudaGetLastError())

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

}  // namespace lnasr
