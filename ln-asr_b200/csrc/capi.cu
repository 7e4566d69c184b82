// capi.cu -- error reporting, version and device queries of liblnasr_b200.so (see include/lnasr_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace lnasr {

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

}  // namespace lnasr

using namespace lnasr;

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
