// capi.cu -- error reporting, version and device queries of liblnasr_b200.so (see include/lnasr_b200.h).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace lnasr {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    // per-device cache; cudaGetDevice is cheap, the attribute query is not
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// FP64 peak probes: 8 independent FMA chains per thread / 8 independent m8n8k4 DMMA accumulators per warp
__global__ void fp64_fma_probe_kernel(double* out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void fp64_mma_probe_kernel(double* out, int iters) {
    double c0[8], c1[8];
    for (int i = 0; i < 8; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace lnasr

using namespace lnasr;

extern "C" int64_t lnasr_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

// Measured FP64 peak of the current device in TFLOP/s (2 flops per FMA), for the roofline denominators of the float64
// kernels: kind 0 = DFMA (SIMT), 1 = mma.sync m8n8k4 f64 (tensor).  Synchronises the stream; scratch must hold
// lnasr_debug_fp64_peak_scratch_doubles() doubles.  Not used by the product path.
extern "C" int64_t lnasr_debug_fp64_peak_scratch_doubles(void) { return (int64_t)sm_count() * 512; }

extern "C" int lnasr_debug_fp64_peak(int kind, double* scratch, double* tflops_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(scratch && tflops_out && (kind == 0 || kind == 1), "fp64_peak: bad arguments");
    const int sms = sm_count(), threads = 512, iters = kind == 0 ? 40000 : 10000;
    cudaEvent_t e0, e1;
    LNASR_CUDA(cudaEventCreate(&e0));
    LNASR_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        LNASR_CUDA(cudaEventRecord(e0, stream));
        if (kind == 0) fp64_fma_probe_kernel<<<sms, threads, 0, stream>>>(scratch, iters);
        else fp64_mma_probe_kernel<<<sms, threads, 0, stream>>>(scratch, iters);
        LNASR_CUDA(cudaEventRecord(e1, stream));
        LNASR_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        LNASR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    LNASR_LAUNCH_CHECK();
    const double fma = kind == 0 ? (double)sms * threads * iters * 8.0 : (double)sms * (threads / 32) * iters * 8.0 * 256.0;
    *tflops_out = 2.0 * fma / (best * 1e-3) / 1e12;
    return LNASR_OK;
}

extern "C" const char* lnasr_last_error(void) { return g_err; }

extern "C" int lnasr_version(void) { return 100; /* 0.1.0 */ }

extern "C" int lnasr_device_info(int* sm, int* cc_major, int* cc_minor, size_t* smem_optin_bytes) {
    int dev = 0;
    LNASR_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm) { LNASR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
    if (cc_major) { LNASR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { LNASR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    if (smem_optin_bytes) {
        LNASR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        *smem_optin_bytes = (size_t)v;
    }
    return LNASR_OK;
}
