// mfcc.cu -- batched MFCC front end (kernel (1) of the north star).
//
// Replaces MFCC.calc_mfcc, /root/reference/lnasr/mfcc.py:108-175 (see include/lnasr_b200.h).
//
// Pass A  (mfcc_frames_kernel): one warp owns one frame at a time and walks a contiguous run of
//   frames.  pre-emphasis + framing + Hamming while loading (mfcc.py:63, utils.py:119-130, mfcc.py:138),
//   512-point real FFT as a 256-point complex radix-4 Stockham FFT in shared memory (mfcc.py:141),
//   power spectrum (mfcc.py:142), triangular mel filters as a banded product (mfcc.py:148; 453 non-zeros
//   instead of a dense 257x40 GEMM), eps substitution + 20 log10 (mfcc.py:149-150), DCT-II ortho
//   (mfcc.py:152), ln energy (mfcc.py:158-160).  Because the DCT is linear, the two per-utterance mean
//   removals (mfcc.py:151,156) commute with it: pass A writes raw cepstra and accumulates per-utterance
//   column sums in float64.
// Pass B  (mfcc_finalize_kernel): subtract the means (+1e-8 constants), delta / delta-delta with the
//   reference's row-0 rule (mfcc.py:162-172) exchanged through warp shuffles, coalesced 39-float rows.
//
// HBM traffic per frame: 320 B pcm (int16) + 64 B raw write + 64 B raw read + 156 B features.
#include "common.cuh"

namespace lnasr {

constexpr int FFT_N = 512;
constexpr int HALF_N = 256;
constexpr int MAX_MEL = 64;
constexpr int RAW_LD = 16;            // raw static row: 12 cepstra + energy, padded to 16 floats
constexpr int WARPS_PER_BLOCK = 8;
constexpr int FRAMES_PER_RUN = 16;    // consecutive frames a warp handles before taking the next run

struct MfccDev {
    int frame_len, frame_step, n_mel, n_ceps, fb_stride;
    float hpf_alpha;
    const float* window;
    const float2* twiddle;
    const int* fb_start;
    const int* fb_len;
    const float* fb_weight;
    const float* dct_t;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <typename PcmT>
__device__ __forceinline__ float load_pcm(const PcmT* p, int64_t i) { return (float)__ldg(p + i); }

// shared memory carve-up (floats)
struct SmemLayout {
    int window, twiddle, fb_w, dct, fb_meta, per_warp, total;
};
__host__ __device__ inline SmemLayout smem_layout(int frame_len, int n_mel, int fb_stride) {
    SmemLayout L;
    int o = 0;
    L.twiddle = o; o += 2 * FFT_N;
    L.window = o;  o += (frame_len + 3) & ~3;
    L.fb_w = o;    o += n_mel * fb_stride;
    L.dct = o;     o += n_mel * n_mel;
    L.fb_meta = o; o += 2 * MAX_MEL;
    o = (o + 3) & ~3;
    L.per_warp = o;
    // per warp: two complex buffers of 256 (1024 floats) + power[260] + logmel[64]
    L.total = o + WARPS_PER_BLOCK * (2 * 2 * HALF_N + 260 + MAX_MEL);
    return L;
}

template <typename PcmT>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
mfcc_frames_kernel(const PcmT* __restrict__ pcm, const int64_t* __restrict__ sample_off,
                   const int64_t* __restrict__ frame_off, int n_utt, int64_t total_frames, MfccDev P,
                   float* __restrict__ raw, double* __restrict__ colsum, float* __restrict__ power_out,
                   float* __restrict__ ceps_raw_out) {
    extern __shared__ __align__(16) float smem[];
    const SmemLayout L = smem_layout(P.frame_len, P.n_mel, P.fb_stride);
    float2* s_tw = (float2*)(smem + L.twiddle);
    float* s_win = smem + L.window;
    float* s_fbw = smem + L.fb_w;
    float* s_dct = smem + L.dct;
    int* s_fbs = (int*)(smem + L.fb_meta);
    int* s_fbl = s_fbs + MAX_MEL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < FFT_N; i += blockDim.x) s_tw[i] = P.twiddle[i];
    for (int i = tid; i < P.frame_len; i += blockDim.x) s_win[i] = P.window[i];
    for (int i = tid; i < P.n_mel * P.fb_stride; i += blockDim.x) s_fbw[i] = P.fb_weight[i];
    for (int i = tid; i < P.n_mel * P.n_mel; i += blockDim.x) s_dct[i] = P.dct_t[i];
    for (int i = tid; i < P.n_mel; i += blockDim.x) { s_fbs[i] = P.fb_start[i]; s_fbl[i] = P.fb_len[i]; }
    __syncthreads();

    float* wbase = smem + L.per_warp + warp * (2 * 2 * HALF_N + 260 + MAX_MEL);
    float2* bufA = (float2*)wbase;
    float2* bufB = bufA + HALF_N;
    float* s_pow = wbase + 4 * HALF_N;
    float* s_lm = s_pow + 260;

    const bool full_ceps = (ceps_raw_out != nullptr);
    const int n_out = full_ceps ? P.n_mel : (P.n_ceps + 1);   // cepstral coefficients 0..n_out-1 computed
    const int64_t n_runs = (total_frames + FRAMES_PER_RUN - 1) / FRAMES_PER_RUN;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    for (int64_t run = warp_global; run < n_runs; run += warp_stride) {
        const int64_t f_begin = run * FRAMES_PER_RUN;
        const int64_t f_end = min(f_begin + (int64_t)FRAMES_PER_RUN, total_frames);
        int u = find_utt(frame_off, n_utt, f_begin);
        int64_t u_f0 = __ldg(frame_off + u), u_f1 = __ldg(frame_off + u + 1);
        int64_t u_s0 = __ldg(sample_off + u), u_len = __ldg(sample_off + u + 1) - u_s0;
        double acc0 = 0.0, acc1 = 0.0;    // running column sums of this utterance: lane k and k+32
        for (int64_t f = f_begin; f < f_end; ++f) {
            if (f >= u_f1) {              // crossed into the next utterance: flush sums
                if (lane < n_out) atomicAdd(colsum + (int64_t)u * MAX_MEL + lane, acc0);
                if (lane + 32 < n_out) atomicAdd(colsum + (int64_t)u * MAX_MEL + lane + 32, acc1);
                acc0 = acc1 = 0.0;
                while (f >= u_f1) { ++u; u_f1 = __ldg(frame_off + u + 1); }
                u_f0 = __ldg(frame_off + u);
                u_s0 = __ldg(sample_off + u);
                u_len = __ldg(sample_off + u + 1) - u_s0;
            }
            const int64_t s_start = (f - u_f0) * P.frame_step;   // first sample of the frame inside the utterance
            const PcmT* x = pcm + u_s0;
            // ---- load + pre-emphasis + window, packed as z[m] = (xw[2m], xw[2m+1])
#pragma unroll
            for (int q = 0; q < HALF_N / 32; ++q) {
                const int m = lane + 32 * q;
                float2 z = make_float2(0.f, 0.f);
                const int n0 = 2 * m;
                if (n0 < P.frame_len) {
                    const int64_t i0 = s_start + n0;
                    float xm1 = (i0 >= 1 && i0 - 1 < u_len) ? load_pcm(x, i0 - 1) : 0.f;
                    float x0 = (i0 < u_len) ? load_pcm(x, i0) : 0.f;
                    float x1 = (i0 + 1 < u_len) ? load_pcm(x, i0 + 1) : 0.f;
                    // y[0] = x[0]; y[t] = x[t] - alpha x[t-1]; samples past the end are zero AFTER filtering
                    float y0 = (i0 < u_len) ? ((i0 == 0) ? x0 : fmaf(-P.hpf_alpha, xm1, x0)) : 0.f;
                    float y1 = (i0 + 1 < u_len) ? fmaf(-P.hpf_alpha, x0, x1) : 0.f;
                    z.x = y0 * s_win[n0];
                    z.y = (n0 + 1 < P.frame_len) ? y1 * s_win[n0 + 1] : 0.f;
                }
                bufA[m] = z;
            }
            __syncwarp();
            // ---- 256-point complex FFT, radix-4 Stockham, 4 stages
            float2* src = bufA;
            float2* dst = bufB;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                const int Ns = 1 << (2 * st);
                const int tw_scale = 128 >> (2 * st);            // 512 / (4 Ns)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = lane + 32 * h;
                    const int k = j & (Ns - 1);
                    float2 v0 = src[j], v1 = src[j + 64], v2 = src[j + 128], v3 = src[j + 192];
                    if (st > 0) {
                        v1 = cmul(v1, s_tw[k * tw_scale]);
                        v2 = cmul(v2, s_tw[2 * k * tw_scale]);
                        v3 = cmul(v3, s_tw[3 * k * tw_scale]);
                    }
                    float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
                    float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
                    float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
                    float2 a3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));   // (v1-v3) * (-i)
                    const int j0 = ((j - k) << 2) + k;
                    dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
                    dst[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
                    dst[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
                    dst[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
                }
                __syncwarp();
                float2* t = src; src = dst; dst = t;
            }
            // after 4 stages the spectrum Z is in `src` (== bufA)
            // ---- real-FFT split + power spectrum |X|^2 / 512
            float esum = 0.f;
#pragma unroll
            for (int q = 0; q <= HALF_N / 32; ++q) {
                const int k = lane + 32 * q;
                if (k <= HALF_N) {
                    const float2 zk = src[k & (HALF_N - 1)];
                    const float2 zr = src[(HALF_N - k) & (HALF_N - 1)];
                    const float2 E = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
                    const float2 O = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));  // -(i/2)(zk - conj zr)
                    const float2 X = make_float2(E.x, E.y);
                    const float2 wo = cmul(s_tw[k], O);
                    const float re = X.x + wo.x, im = X.y + wo.y;
                    const float pw = (re * re + im * im) * (1.0f / FFT_N);
                    s_pow[k] = pw;
                    esum += pw;
                    if (power_out) power_out[f * (HALF_N + 1) + k] = pw;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
            __syncwarp();
            // ---- mel filter bank (banded), eps substitution, 20 log10
            for (int m = lane; m < P.n_mel; m += 32) {
                const int s = s_fbs[m], len = s_fbl[m];
                const float* w = s_fbw + m * P.fb_stride;
                float e = 0.f;
                for (int i = 0; i < len; ++i) e = fmaf(w[i], s_pow[s + i], e);
                if (e == 0.f) e = 2.220446049250313e-16f;          // np.finfo(float).eps, mfcc.py:149
                s_lm[m] = 8.685889638065035f * logf(e);            // 20 log10(e) = 20/ln(10) * ln(e)
            }
            __syncwarp();
            // ---- DCT-II ortho: lane k computes coefficient k (and k+32)
            float c0 = 0.f, c1 = 0.f;
            {
                const int k0 = lane, k1 = lane + 32;
                const bool do1 = k1 < n_out;
                if (k0 < n_out) {
                    for (int n = 0; n < P.n_mel; ++n) {
                        const float lm = s_lm[n];
                        c0 = fmaf(s_dct[n * P.n_mel + k0], lm, c0);
                        if (do1) c1 = fmaf(s_dct[n * P.n_mel + k1], lm, c1);
                    }
                }
            }
            acc0 += (double)c0;
            acc1 += (double)c1;
            // raw static row: cepstra 1..n_ceps then ln(sum power)
            if (lane >= 1 && lane <= P.n_ceps) raw[f * RAW_LD + (lane - 1)] = c0;
            if (lane == 0) raw[f * RAW_LD + P.n_ceps] = logf(esum);
            if (full_ceps) {
                if (lane < P.n_mel) ceps_raw_out[f * P.n_mel + lane] = c0;
                if (lane + 32 < P.n_mel) ceps_raw_out[f * P.n_mel + lane + 32] = c1;
            }
            __syncwarp();
        }
        if (lane < n_out) atomicAdd(colsum + (int64_t)u * MAX_MEL + lane, acc0);
        if (lane + 32 < n_out) atomicAdd(colsum + (int64_t)u * MAX_MEL + lane + 32, acc1);
    }
}

// Pass B.  One warp produces 30 consecutive frames (lanes 0,1 carry the two-frame halo); lane = frame so
// delta / delta-delta are neighbour exchanges by __shfl_up_sync.  The 30x39 output block is contiguous in
// HBM (rows of 39 floats back to back when ld_feats == 39) and is written through a shared-memory
// transpose so that consecutive lanes store consecutive addresses.
constexpr int FIN_WARPS = 8;
constexpr int FIN_FRAMES = 30;

__global__ void __launch_bounds__(FIN_WARPS * 32)
mfcc_finalize_kernel(const float* __restrict__ raw, const double* __restrict__ colsum,
                     const int64_t* __restrict__ frame_off, int n_utt, int64_t total_frames, int n_ceps, int n_mel,
                     float* __restrict__ feats, int ld_feats, float* __restrict__ ceps_io, float dct_dc) {
    __shared__ float tile[FIN_WARPS][32][40];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ns = n_ceps + 1;                                  // 13 static coefficients
    const int64_t n_tiles = (total_frames + FIN_FRAMES - 1) / FIN_FRAMES;
    for (int64_t t = (int64_t)blockIdx.x * FIN_WARPS + warp; t < n_tiles; t += (int64_t)gridDim.x * FIN_WARPS) {
        const int64_t f = t * FIN_FRAMES + lane - 2;            // lanes 0,1: halo frames
        const bool valid = (f >= 0 && f < total_frames);
        int u = 0;
        int64_t f0 = 0, f1 = 1;
        if (valid) {
            u = find_utt(frame_off, n_utt, f);
            f0 = __ldg(frame_off + u);
            f1 = __ldg(frame_off + u + 1);
        }
        const int64_t fl = f - f0;                               // frame index inside the utterance
        const double inv_cnt = 1.0 / (double)(f1 - f0);
        // is the frame one (two) before me in the same utterance?
        const bool has_prev1 = valid && fl >= 1;
        const bool has_prev2 = valid && fl >= 2;
        const bool out_lane = valid && lane >= 2 && lane < 2 + FIN_FRAMES && (t * FIN_FRAMES + lane - 2) < total_frames;
        for (int c = 0; c < ns; ++c) {
            float s = 0.f, s_next = 0.f;
            if (valid) {
                // feature = raw - (mean + 1e-8) for the cepstra (mfcc.py:151,156 folded, see file header); energy as is
                const double mean = (c < n_ceps) ? (__ldg(colsum + (int64_t)u * MAX_MEL + c + 1) * inv_cnt + 1e-8) : 0.0;
                s = (float)((double)__ldg(raw + f * RAW_LD + c) - mean);
                if (fl == 0) s_next = (float)((double)__ldg(raw + (f + 1) * RAW_LD + c) - mean);   // row 1 of the utterance
            }
            float p1 = __shfl_up_sync(0xffffffffu, s, 1);
            float p2 = __shfl_up_sync(0xffffffffu, s, 2);
            // delta[0] = static[1]; delta[t] = static[t] - static[t-1]                       (mfcc.py:162-166)
            const float d = (fl == 0) ? s_next : (s - p1);
            // delta of the previous frame, needed for delta-delta
            //   prev frame fl-1 == 0  -> delta_prev = static[1] = s (me)
            const float d_prev = (fl == 1) ? s : (p1 - p2);
            // ddelta[0] = delta[1] = static[1]-static[0]; ddelta[t] = delta[t]-delta[t-1]   (mfcc.py:168-172)
            const float dd = (fl == 0) ? (s_next - s) : (d - d_prev);
            (void)has_prev1; (void)has_prev2;
            tile[warp][lane][c] = s;
            tile[warp][lane][ns + c] = d;
            tile[warp][lane][2 * ns + c] = dd;
        }
        __syncwarp();
        // coalesced store of the 30 x 39 block
        const int64_t fb = t * FIN_FRAMES;
        const int nf = (int)min((int64_t)FIN_FRAMES, total_frames - fb);
        const int D = 3 * ns;
        for (int i = lane; i < nf * D; i += 32) {
            const int r = i / D, c = i - r * D;
            feats[(fb + r) * ld_feats + c] = tile[warp][r + 2][c];
        }
        (void)out_lane;
        __syncwarp();
        // optional second return value of calc_mfcc: DCT of the mean-removed log-mel, columns 1..n_ceps
        // additionally mean-removed in place (mfcc.py:155-156)
        if (ceps_io) {
            for (int r = 0; r < nf; ++r) {
                const int64_t ff = fb + r;
                const int uu = find_utt(frame_off, n_utt, ff);
                const double ic = 1.0 / (double)(__ldg(frame_off + uu + 1) - __ldg(frame_off + uu));
                for (int k = lane; k < n_mel; k += 32) {
                    double mean = __ldg(colsum + (int64_t)uu * MAX_MEL + k) * ic;
                    double v = (double)ceps_io[ff * n_mel + k] - mean;
                    if (k == 0) v -= 1e-8 * (double)dct_dc;       // DCT of the constant 1e-8 offset
                    if (k >= 1 && k <= n_ceps) v -= 1e-8;
                    ceps_io[ff * n_mel + k] = (float)v;
                }
            }
        }
    }
}

}  // namespace lnasr

using namespace lnasr;

extern "C" int64_t lnasr_mfcc_num_frames(int64_t n_samples, int32_t frame_len, int32_t frame_step) {
    int64_t d = n_samples - (int64_t)(frame_len - frame_step);
    if (d < 0) d = -d;
    return (d + frame_step - 1) / frame_step;
}

extern "C" int lnasr_mfcc_workspace_bytes(int64_t total_frames, int32_t n_utt, size_t* bytes) {
    LNASR_CHECK_ARG(bytes && total_frames >= 0 && n_utt >= 0, "mfcc_workspace_bytes: bad arguments");
    size_t raw = (size_t)total_frames * RAW_LD * sizeof(float);
    raw = (raw + 255) & ~(size_t)255;
    *bytes = raw + (size_t)n_utt * MAX_MEL * sizeof(double) + 256;
    return LNASR_OK;
}

extern "C" int lnasr_mfcc_batch(const void* pcm, int pcm_dtype, const int64_t* sample_offsets,
                                const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                                const lnasr_mfcc_tables* tb, float* feats, int32_t ld_feats, float* power_out,
                                float* ceps_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(pcm && sample_offsets && frame_offsets && tb && feats && workspace, "mfcc_batch: null pointer");
    LNASR_CHECK_ARG(pcm_dtype == LNASR_DT_I16 || pcm_dtype == LNASR_DT_F32, "mfcc_batch: pcm dtype must be int16 or float32");
    LNASR_CHECK_ARG(tb->fft_n == FFT_N, "mfcc_batch: only fft_N=512 is built (got %d)", tb->fft_n);
    LNASR_CHECK_ARG(tb->frame_len >= 2 && tb->frame_len <= FFT_N && tb->frame_step >= 1, "mfcc_batch: frame_len must be in [2,512]");
    LNASR_CHECK_ARG(tb->n_mel >= 2 && tb->n_mel <= MAX_MEL, "mfcc_batch: n_mel must be in [2,64]");
    LNASR_CHECK_ARG(tb->n_ceps >= 1 && tb->n_ceps < tb->n_mel && tb->n_ceps <= 31, "mfcc_batch: n_ceps must be < n_mel and <= 31");
    LNASR_CHECK_ARG(tb->fb_stride >= 1 && tb->fb_stride <= 64, "mfcc_batch: fb_stride must be in [1,64]");
    LNASR_CHECK_ARG(ld_feats >= 3 * (tb->n_ceps + 1), "mfcc_batch: ld_feats too small");
    LNASR_CHECK_ARG(3 * (tb->n_ceps + 1) <= 40 && tb->n_ceps + 1 <= RAW_LD, "mfcc_batch: at most 13 static coefficients are built");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    size_t need = 0;
    lnasr_mfcc_workspace_bytes(total_frames, n_utt, &need);
    if (workspace_bytes < need) {
        set_error("mfcc_batch: workspace %zu < %zu bytes", workspace_bytes, need);
        return LNASR_EWORKSPACE;
    }
    float* raw = (float*)workspace;
    size_t raw_bytes = ((size_t)total_frames * RAW_LD * sizeof(float) + 255) & ~(size_t)255;
    double* colsum = (double*)((char*)workspace + raw_bytes);
    LNASR_CUDA(cudaMemsetAsync(colsum, 0, (size_t)n_utt * MAX_MEL * sizeof(double), stream));

    MfccDev P;
    P.frame_len = tb->frame_len; P.frame_step = tb->frame_step; P.n_mel = tb->n_mel; P.n_ceps = tb->n_ceps;
    P.fb_stride = tb->fb_stride; P.hpf_alpha = tb->hpf_alpha; P.window = tb->window;
    P.twiddle = (const float2*)tb->twiddle; P.fb_start = tb->fb_start; P.fb_len = tb->fb_len;
    P.fb_weight = tb->fb_weight; P.dct_t = tb->dct_t;

    const SmemLayout L = smem_layout(P.frame_len, P.n_mel, P.fb_stride);
    const size_t smem = (size_t)L.total * sizeof(float);
    const int64_t n_runs = ceil_div(total_frames, FRAMES_PER_RUN);
    int grid = (int)std::min<int64_t>(ceil_div(n_runs, WARPS_PER_BLOCK), (int64_t)sm_count() * 3);
    if (grid < 1) grid = 1;
    if (pcm_dtype == LNASR_DT_I16) {
        LNASR_CUDA(cudaFuncSetAttribute(mfcc_frames_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mfcc_frames_kernel<int16_t><<<grid, WARPS_PER_BLOCK * 32, smem, stream>>>(
            (const int16_t*)pcm, sample_offsets, frame_offsets, n_utt, total_frames, P, raw, colsum, power_out, ceps_out);
    } else {
        LNASR_CUDA(cudaFuncSetAttribute(mfcc_frames_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mfcc_frames_kernel<float><<<grid, WARPS_PER_BLOCK * 32, smem, stream>>>(
            (const float*)pcm, sample_offsets, frame_offsets, n_utt, total_frames, P, raw, colsum, power_out, ceps_out);
    }
    LNASR_LAUNCH_CHECK();
    const int64_t n_tiles = ceil_div(total_frames, FIN_FRAMES);
    int grid2 = (int)std::min<int64_t>(ceil_div(n_tiles, FIN_WARPS), (int64_t)sm_count() * 8);
    if (grid2 < 1) grid2 = 1;
    // DCT-II ortho of a constant vector c is c*sqrt(n_mel) in coefficient 0 and 0 elsewhere
    const float dct_dc = sqrtf((float)tb->n_mel);
    mfcc_finalize_kernel<<<grid2, FIN_WARPS * 32, 0, stream>>>(raw, colsum, frame_offsets, n_utt, total_frames,
                                                              tb->n_ceps, tb->n_mel, feats, ld_feats, ceps_out, dct_dc);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

// ------------------------------------------------------------------------------------------------
// MFCC.calc_high_pass_filter alone (mfcc.py:50-64), float64 like the reference.
namespace lnasr {
template <typename T>
__global__ void preemphasis_kernel(const T* __restrict__ x, int64_t n, double alpha, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = (i == 0) ? (double)x[0] : (double)x[i] - alpha * (double)x[i - 1];
}
}  // namespace lnasr

extern "C" int lnasr_preemphasis(const void* pcm, int pcm_dtype, int64_t n, double alpha, double* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(pcm && out && n >= 0, "preemphasis: bad arguments");
    if (n == 0) return LNASR_OK;
    const unsigned grid = (unsigned)ceil_div(n, 256);
    if (pcm_dtype == LNASR_DT_I16) preemphasis_kernel<int16_t><<<grid, 256, 0, stream>>>((const int16_t*)pcm, n, alpha, out);
    else if (pcm_dtype == LNASR_DT_F32) preemphasis_kernel<float><<<grid, 256, 0, stream>>>((const float*)pcm, n, alpha, out);
    else if (pcm_dtype == LNASR_DT_F64) preemphasis_kernel<double><<<grid, 256, 0, stream>>>((const double*)pcm, n, alpha, out);
    else { set_error("preemphasis: pcm dtype must be int16/float32/float64"); return LNASR_EINVAL; }
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}
