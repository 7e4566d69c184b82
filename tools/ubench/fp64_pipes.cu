// Microbenchmark (not part of the product): FP64 FMA vs FP64 tensor (mma.sync m8n8k4) throughput on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void dmma_kernel(double* out, int iters) {
    double c0[8], c1[8];
    for (int i = 0; i < 8; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms;
        dfma_kernel<<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e0); dfma_kernel<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fma_s = (double)sms * warps * 32 * iters * 8 / (ms * 1e-3);
        printf("DFMA  %2d warps/SM: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", warps, ms, fma_s / 1e12, fma_s / sms / 1.965e9);
        dmma_kernel<<<sms, warps * 32>>>(out, iters / 4);
        cudaEventRecord(e0); dmma_kernel<<<sms, warps * 32>>>(out, iters / 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double mma_s = (double)sms * warps * (iters / 4) * 8 * 256 / (ms * 1e-3);
        printf("DMMA  %2d warps/SM: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", warps, ms, mma_s / 1e12, mma_s / sms / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
