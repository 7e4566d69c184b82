// vad.cu -- long-term spectral divergence voice activity detection, the step BEFORE the acoustic path.
//
// Replaces VadLtsd.detect, /root/reference/third/vadlstd/VadLstd.py:44-74: the same framing rule as the MFCC front end
// (utils.split_frames, utils.py:106-130, here after one frame step of leading zeros, VadLstd.py:52), the same Hamming window
// (utils.py:132-138), |rfft| of winsize points (VadLstd.py:58-60), the noise spectrum from the first two frames
// (VadLstd.py:62), the long-term spectral envelope = running maximum over 2 * order frames (VadLstd.py:64-66) and
// LTSD[k] = 10 log10( sum_bins LTSE[k]^2 / noise / winsize ) with the optional noise adaptation (VadLstd.py:68-73).
//
//   vad_amplitude_kernel : one CTA per frame; framing + window into shared memory in bit-reversed order, radix-2 float64 FFT
//                          (winsize a power of two <= 4096: 64 KB of shared memory), amplitude spectrum to HBM.
//   vad_ltsd_kernel      : one CTA per utterance; the envelope and the divergence.  Without adaptation every frame is
//                          independent; with it the noise spectrum is a recurrence over the frames, walked in order.
// A VAD pass handles ~16 frames per second of audio: this is a latency item, not a throughput one.
#include "common.cuh"

namespace lnasr {

constexpr int VAD_THREADS = 256;

__global__ void __launch_bounds__(VAD_THREADS)
vad_amplitude_kernel(const void* __restrict__ pcm, int pcm_dtype, int64_t n_samples, int winsize, int stepsize, int log2n,
                     const double* __restrict__ window, const double2* __restrict__ twiddle, double* __restrict__ amp) {
    extern __shared__ __align__(16) double2 vad_z[];                  // [winsize] complex
    const int64_t f = blockIdx.x;
    const int N = winsize;
    // padded signal = stepsize zeros, then the data, then zeros (VadLstd.py:52, utils.py:121-124)
    for (int n = threadIdx.x; n < N; n += VAD_THREADS) {
        const int64_t s = f * stepsize + n - stepsize;                // index into the data
        double x = 0.0;
        if (s >= 0 && s < n_samples) {
            if (pcm_dtype == LNASR_DT_I16) x = (double)((const int16_t*)pcm)[s];
            else if (pcm_dtype == LNASR_DT_F32) x = (double)((const float*)pcm)[s];
            else x = ((const double*)pcm)[s];
        }
        const int r = (int)(__brev((unsigned)n) >> (32 - log2n));
        vad_z[r] = make_double2(x * window[n], 0.0);
    }
    __syncthreads();
    for (int st = 0; st < log2n; ++st) {
        const int half = 1 << st;
        for (int i = threadIdx.x; i < N / 2; i += VAD_THREADS) {
            const int j = i & (half - 1);
            const int lo = ((i >> st) << (st + 1)) + j, hi = lo + half;
            const double2 w = twiddle[(int64_t)j << (log2n - 1 - st)];   // exp(-2 pi i j / (2 half))
            const double2 a = vad_z[lo], b = vad_z[hi];
            const double2 t = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            vad_z[lo] = make_double2(a.x + t.x, a.y + t.y);
            vad_z[hi] = make_double2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    const int nb = N / 2 + 1;
    for (int k = threadIdx.x; k < nb; k += VAD_THREADS) amp[f * nb + k] = hypot(vad_z[k].x, vad_z[k].y);
}

__device__ __forceinline__ double vad_block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_d(v, lane ^ o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < VAD_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(VAD_THREADS)
vad_ltsd_kernel(const double* __restrict__ amp, int64_t n_frames, int nb, int winsize, int order, double threshold, int adapt,
                double alpha, double* __restrict__ noise, double* __restrict__ ltsd) {
    __shared__ double red[VAD_THREADS / 32];
    // noise = (mean of the first two amplitude spectra)^2   (VadLstd.py:62)
    const int n0 = (int)min((int64_t)2, n_frames);
    for (int k = threadIdx.x; k < nb; k += VAD_THREADS) {
        double s = 0.0;
        for (int r = 0; r < n0; ++r) s += amp[(int64_t)r * nb + k];
        s /= (double)n0;
        noise[k] = s * s;
    }
    for (int64_t f = threadIdx.x; f < n_frames; f += VAD_THREADS) ltsd[f] = 0.0;
    __syncthreads();
    for (int64_t f = order; f < n_frames - order; ++f) {
        double part = 0.0, env = 0.0;
        for (int k = threadIdx.x; k < nb; k += VAD_THREADS) {
            double e = amp[(f - order) * nb + k];
            for (int64_t r = f - order + 1; r < f + order; ++r) e = fmax(e, amp[r * nb + k]);     // frames f-order .. f+order-1
            part += e * e / noise[k];
            env += e;
        }
        const double tot = vad_block_sum(part, red);
        const double v = 10.0 * log10(tot / (double)winsize);
        if (threadIdx.x == 0) ltsd[f] = v;
        if (adapt) {
            const double envsum = vad_block_sum(env, red);
            if (v < threshold) {                                      // VadLstd.py:72-73: scalar mean envelope into every bin
                const double add = (1.0 - alpha) * (envsum / (double)winsize);
                for (int k = threadIdx.x; k < nb; k += VAD_THREADS) noise[k] = alpha * noise[k] + add;
            }
            __syncthreads();
        }
    }
}

}  // namespace lnasr

using namespace lnasr;

extern "C" int64_t lnasr_vad_num_frames(int64_t n_samples, int32_t winsize, int32_t stepsize) {
    return lnasr_mfcc_num_frames(n_samples + stepsize, winsize, stepsize);      // the framing rule of utils.py:118 on the padded signal
}

extern "C" size_t lnasr_vad_workspace_bytes(int64_t n_samples, int32_t winsize, int32_t stepsize) {
    const int64_t F = lnasr_vad_num_frames(n_samples, winsize, stepsize);
    return (size_t)(F + 1) * (winsize / 2 + 1) * sizeof(double) + 256;          // amplitude spectra + the noise spectrum
}

extern "C" int lnasr_vad_ltsd(const void* pcm, int pcm_dtype, int64_t n_samples, int32_t winsize, int32_t stepsize, int32_t order,
                              double threshold, int adapt, double alpha, const double* window, const double* twiddle,
                              double* ltsd_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(pcm && window && twiddle && ltsd_out && workspace, "vad_ltsd: null pointer");
    LNASR_CHECK_ARG(pcm_dtype == LNASR_DT_I16 || pcm_dtype == LNASR_DT_F32 || pcm_dtype == LNASR_DT_F64, "vad_ltsd: pcm dtype must be int16/float32/float64");
    int log2n = 0;
    while ((1 << log2n) < winsize) ++log2n;
    LNASR_CHECK_ARG(winsize >= 4 && winsize <= 4096 && (1 << log2n) == winsize, "vad_ltsd: winsize must be a power of two in [4, 4096]");
    LNASR_CHECK_ARG(stepsize >= 1 && order >= 1 && n_samples >= 0, "vad_ltsd: bad shape");
    const int64_t F = lnasr_vad_num_frames(n_samples, winsize, stepsize);
    if (F == 0) return LNASR_OK;
    LNASR_CHECK_ARG(workspace_bytes >= lnasr_vad_workspace_bytes(n_samples, winsize, stepsize), "vad_ltsd: workspace too small");
    const int nb = winsize / 2 + 1;
    double* amp = reinterpret_cast<double*>(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
    double* noise = amp + F * nb;
    const size_t smem = (size_t)winsize * sizeof(double2);
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(vad_amplitude_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vad_amplitude_kernel<<<(unsigned)F, VAD_THREADS, smem, stream>>>(pcm, pcm_dtype, n_samples, winsize, stepsize, log2n, window,
                                                                    (const double2*)twiddle, amp);
    LNASR_LAUNCH_CHECK();
    vad_ltsd_kernel<<<1, VAD_THREADS, 0, stream>>>(amp, F, nb, winsize, order, threshold, adapt, alpha, noise, ltsd_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}
