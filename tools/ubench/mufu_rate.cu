// MUFU.EX2 / MUFU.LG2 issue rate per SM on this GPU (ops / clk / SM), alone and next to FFMA work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu_rate tools/ubench/mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int FMA_PER_MUFU>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* clk) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = 1.0f + 1e-6f * i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = ex2(a[i]) - 1.0f;                 // MUFU + FADD, 8 independent chains per thread
#pragma unroll
            for (int j = 0; j < FMA_PER_MUFU; ++j) f[i] = fmaf(f[i], 0.999999f, 1e-7f);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int F>
void run(const char* name, int threads) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    cudaMalloc(&clk, sizeof(long long) * sms);
    const int iters = 4096;
    k<F><<<sms, threads>>>(out, iters, clk);
    k<F><<<sms, threads>>>(out, iters, clk);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double mufu = (double)iters * 8 * threads;
    printf("%-28s threads/SM %4d: %.2f MUFU.EX2 / clk / SM   (%.2f FFMA / clk / SM alongside)\n", name, threads, mufu / avg, mufu * F / avg);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("ex2 alone", 128);
    run<0>("ex2 alone", 256);
    run<0>("ex2 alone", 512);
    run<2>("ex2 + 2 FFMA each", 512);
    run<6>("ex2 + 6 FFMA each", 512);
    return 0;
}
