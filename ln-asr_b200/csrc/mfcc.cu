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
    double hpf_alpha;
    const double* window;
    const double2* twiddle;
    const int* fb_start;
    const int* fb_len;
    const float* fb_weight;
    const double* dct_t;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// compact per-stage twiddle tables of the radix-4 Stockham FFT: stage st (1..3, Ns = 4^st), factor r (1..3),
// k < Ns  ->  W512^(r * k * 128 / Ns).  k is the fastest index, so a warp's loads are conflict-free.
__host__ __device__ constexpr int tw_stage_off(int st) { return st == 1 ? 0 : (st == 2 ? 12 : 60); }
constexpr int TW_STAGE_TOTAL = 252;              // 3*(4+16+64)
constexpr int TW_SPLIT = HALF_N + 1;             // W512^k, k = 0..256, for the real-FFT split

// shared memory carve-up (in doubles)
struct SmemLayout {
    int tw_stage, tw_split, window, dct, fb_w, fb_meta, per_warp, total;
};
constexpr int PER_WARP_DOUBLES = 2 * 2 * HALF_N + MAX_MEL;   // two complex buffers (power aliases the 2nd), logmel[64]
__host__ __device__ inline SmemLayout smem_layout(int frame_len, int n_mel, int fb_stride, int n_dct_cols) {
    SmemLayout L;
    int o = 0;
    L.tw_stage = o; o += 2 * TW_STAGE_TOTAL;
    L.tw_split = o; o += 2 * TW_SPLIT;
    L.window = o;   o += (frame_len + 1) & ~1;
    L.dct = o;      o += n_mel * n_dct_cols;             // [n][k < n_dct_cols]
    L.fb_w = o;     o += (n_mel * fb_stride + 1) / 2;    // float weights
    L.fb_meta = o;  o += MAX_MEL;                        // 2 * MAX_MEL ints
    o = (o + 1) & ~1;
    L.per_warp = o;
    L.total = o + WARPS_PER_BLOCK * PER_WARP_DOUBLES;
    return L;
}

// where a frame lives: utterance, its sample range, first sample of the frame
struct FrameCtx {
    int u;
    int64_t u_f1;        // first frame of the next utterance
    int64_t u_s0;        // first sample of the utterance in pcm
    int64_t u_len;       // samples in the utterance
    int64_t s_start;     // first sample of the frame, relative to the utterance
};

// The whole per-frame chain runs in float64.  The reference is float64 (mfcc.py), and the features are small
// differences of ~200 dB log-mel values (mean normalisation, deltas) whose low-energy bands sit up to 90 dB below
// the frame's peak: an fp32 FFT leaves ~1e-4 relative error in exactly those bands (measured), i.e. above the
// 1e-4 parity bar after the DCT and the delta-deltas.  On B200 fp64 runs at half the fp32 rate and the kernel is
// latency-, not ALU-bound, so exactness costs little.
template <typename PcmT>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
mfcc_frames_kernel(const PcmT* __restrict__ pcm, const int64_t* __restrict__ sample_off,
                   const int64_t* __restrict__ frame_off, int n_utt, int64_t total_frames, MfccDev P,
                   float* __restrict__ raw, double* __restrict__ colsum, float* __restrict__ power_out,
                   float* __restrict__ ceps_raw_out) {
    extern __shared__ __align__(16) double smem[];
    const bool full_ceps = (ceps_raw_out != nullptr);
    const int n_out = full_ceps ? P.n_mel : (P.n_ceps + 1);   // cepstral coefficients 0..n_out-1 computed
    const SmemLayout L = smem_layout(P.frame_len, P.n_mel, P.fb_stride, n_out);
    double2* s_tws = reinterpret_cast<double2*>(smem + L.tw_stage);
    double2* s_twr = reinterpret_cast<double2*>(smem + L.tw_split);
    double* s_win = smem + L.window;
    double* s_dct = smem + L.dct;
    float* s_fbw = reinterpret_cast<float*>(smem + L.fb_w);
    int* s_fbs = reinterpret_cast<int*>(smem + L.fb_meta);
    int* s_fbl = s_fbs + MAX_MEL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < TW_STAGE_TOTAL; i += blockDim.x) {
        const int st = i < 12 ? 1 : (i < 60 ? 2 : 3);
        const int Ns = 1 << (2 * st);
        const int rel = i - tw_stage_off(st);
        const int r = rel / Ns + 1, k = rel % Ns;
        s_tws[i] = P.twiddle[r * k * (128 >> (2 * st))];
    }
    for (int i = tid; i < TW_SPLIT; i += blockDim.x) s_twr[i] = P.twiddle[i & (FFT_N - 1)];
    for (int i = tid; i < P.frame_len; i += blockDim.x) s_win[i] = P.window[i];
    for (int i = tid; i < P.n_mel * P.fb_stride; i += blockDim.x) s_fbw[i] = P.fb_weight[i];
    for (int i = tid; i < P.n_mel * n_out; i += blockDim.x) {
        const int n = i / n_out, k = i - n * n_out;
        s_dct[i] = P.dct_t[n * P.n_mel + k];
    }
    for (int i = tid; i < P.n_mel; i += blockDim.x) { s_fbs[i] = P.fb_start[i]; s_fbl[i] = P.fb_len[i]; }
    __syncthreads();

    double* wbase = smem + L.per_warp + warp * PER_WARP_DOUBLES;
    double2* bufA = reinterpret_cast<double2*>(wbase);
    double2* bufB = bufA + HALF_N;
    double* s_pow = reinterpret_cast<double*>(bufB);     // the spectrum ends in bufA; bufB is free for the power
    double* s_lm = wbase + 4 * HALF_N;

    const int64_t n_runs = (total_frames + FRAMES_PER_RUN - 1) / FRAMES_PER_RUN;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    constexpr int NQ = HALF_N / 32;                      // sample pairs per lane

    auto make_ctx = [&](int u, int64_t f) {
        FrameCtx c;
        c.u = u;
        c.u_f1 = __ldg(frame_off + u + 1);
        c.u_s0 = __ldg(sample_off + u);
        c.u_len = __ldg(sample_off + u + 1) - c.u_s0;
        c.s_start = (f - __ldg(frame_off + u)) * P.frame_step;
        return c;
    };
    auto next_ctx = [&](const FrameCtx& c, int64_t f) {          // context of frame f = previous frame + 1
        if (f < c.u_f1) { FrameCtx n = c; n.s_start += P.frame_step; return n; }
        int u = c.u;
        int64_t f1 = c.u_f1;
        while (f >= f1) { ++u; f1 = __ldg(frame_off + u + 1); }
        return make_ctx(u, f);
    };
    // raw samples of one frame: lane holds pairs (x[2m], x[2m+1]), m = lane + 32 q; out-of-range samples read as 0
    // (kept in the PCM type: converting here would make the prefetch wait for its own loads)
    auto fetch = [&](const FrameCtx& c, PcmT* xe, PcmT* xo, PcmT& xlead) {
        const PcmT* x = pcm + c.u_s0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int n0 = 2 * (lane + 32 * q);
            const int64_t i0 = c.s_start + n0;
            xe[q] = (n0 < P.frame_len && i0 < c.u_len) ? __ldg(x + i0) : PcmT(0);
            xo[q] = (n0 + 1 < P.frame_len && i0 + 1 < c.u_len) ? __ldg(x + i0 + 1) : PcmT(0);
        }
        // the sample before the frame (needed by lane 0, q = 0)
        xlead = (lane == 0 && c.s_start >= 1 && c.s_start - 1 < c.u_len) ? __ldg(x + c.s_start - 1) : PcmT(0);
    };

    for (int64_t run = warp_global; run < n_runs; run += warp_stride) {
        const int64_t f_begin = run * FRAMES_PER_RUN;
        const int64_t f_end = min(f_begin + (int64_t)FRAMES_PER_RUN, total_frames);
        FrameCtx nxt = make_ctx(find_utt(frame_off, n_utt, f_begin), f_begin);
        PcmT pe[NQ], po[NQ], plead;
        fetch(nxt, pe, po, plead);
        int acc_u = nxt.u;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};   // running column sums of utterance acc_u: coefficient (lane & 15) + 16 i
        for (int64_t f = f_begin; f < f_end; ++f) {
            const FrameCtx cur = nxt;
            float xe[NQ], xo[NQ];                    // int16 and float32 PCM are both exact in float
#pragma unroll
            for (int q = 0; q < NQ; ++q) { xe[q] = (float)pe[q]; xo[q] = (float)po[q]; }
            const float xlead = (float)plead;
            if (f + 1 < f_end) {                      // software prefetch of the next frame's samples
                nxt = next_ctx(cur, f + 1);
                fetch(nxt, pe, po, plead);
            }
            if (cur.u != acc_u) {                     // crossed into another utterance: flush its column sums
                if (lane < 16) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (lane + 16 * i < n_out) atomicAdd(colsum + (int64_t)acc_u * MAX_MEL + lane + 16 * i, acc[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = 0.0;
                acc_u = cur.u;
            }
            // ---- pre-emphasis (mfcc.py:63) + window (mfcc.py:138), packed as z[m] = (xw[2m], xw[2m+1])
            float carry = xlead;                      // x[2m - 1] for lane 0
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int m = lane + 32 * q;
                const int n0 = 2 * m;
                float xm1 = __shfl_up_sync(0xffffffffu, xo[q], 1);
                if (lane == 0) xm1 = carry;
                carry = __shfl_sync(0xffffffffu, xo[q], 31);
                double2 z = make_double2(0.0, 0.0);
                if (n0 < P.frame_len) {
                    const int64_t i0 = cur.s_start + n0;
                    // y[0] = x[0]; y[t] = x[t] - alpha x[t-1]; samples past the end are zero AFTER filtering
                    const double y0 = (i0 < cur.u_len) ? ((i0 == 0) ? (double)xe[q] : (double)xe[q] - P.hpf_alpha * (double)xm1) : 0.0;
                    const double y1 = (i0 + 1 < cur.u_len) ? (double)xo[q] - P.hpf_alpha * (double)xe[q] : 0.0;
                    z.x = y0 * s_win[n0];
                    z.y = (n0 + 1 < P.frame_len) ? y1 * s_win[n0 + 1] : 0.0;
                }
                bufA[m] = z;
            }
            __syncwarp();
            // ---- 256-point complex FFT, radix-4 Stockham, 4 stages
            double2* src = bufA;
            double2* dst = bufB;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                const int Ns = 1 << (2 * st);
                const double2* tw = s_tws + tw_stage_off(st);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = lane + 32 * h;
                    const int k = j & (Ns - 1);
                    double2 v0 = src[j], v1 = src[j + 64], v2 = src[j + 128], v3 = src[j + 192];
                    if (st > 0) {
                        v1 = cmul(v1, tw[k]);
                        v2 = cmul(v2, tw[Ns + k]);
                        v3 = cmul(v3, tw[2 * Ns + k]);
                    }
                    const double2 a0 = make_double2(v0.x + v2.x, v0.y + v2.y);
                    const double2 a1 = make_double2(v0.x - v2.x, v0.y - v2.y);
                    const double2 a2 = make_double2(v1.x + v3.x, v1.y + v3.y);
                    const double2 a3 = make_double2(v1.y - v3.y, -(v1.x - v3.x));   // (v1-v3) * (-i)
                    const int j0 = ((j - k) << 2) + k;
                    dst[j0] = make_double2(a0.x + a2.x, a0.y + a2.y);
                    dst[j0 + Ns] = make_double2(a1.x + a3.x, a1.y + a3.y);
                    dst[j0 + 2 * Ns] = make_double2(a0.x - a2.x, a0.y - a2.y);
                    dst[j0 + 3 * Ns] = make_double2(a1.x - a3.x, a1.y - a3.y);
                }
                __syncwarp();
                double2* t = src; src = dst; dst = t;
            }
            // after 4 stages the spectrum Z is in `src` (== bufA)
            // ---- real-FFT split + power spectrum |X|^2 / 512   (mfcc.py:141-142)
            double esum = 0.0;
#pragma unroll
            for (int q = 0; q <= NQ; ++q) {
                const int k = lane + 32 * q;
                if (k <= HALF_N) {
                    const double2 zk = src[k & (HALF_N - 1)];
                    const double2 zr = src[(HALF_N - k) & (HALF_N - 1)];
                    const double2 E = make_double2(0.5 * (zk.x + zr.x), 0.5 * (zk.y - zr.y));
                    const double2 O = make_double2(0.5 * (zk.y + zr.y), -0.5 * (zk.x - zr.x));  // -(i/2)(zk - conj zr)
                    const double2 wo = cmul(s_twr[k], O);
                    const double re = E.x + wo.x, im = E.y + wo.y;
                    const double pw = (re * re + im * im) * (1.0 / FFT_N);
                    s_pow[k] = pw;
                    esum += pw;
                    if (power_out) power_out[f * (HALF_N + 1) + k] = (float)pw;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
            __syncwarp();
            // ---- mel filter bank (banded, mfcc.py:148): 4 lanes share a filter, eps substitution (mfcc.py:149),
            //      20 log10 (mfcc.py:150) spread over the lanes
            {
                const int grp = lane >> 2, sub = lane & 3;
                const unsigned gmask = 0xFu << (lane & ~3);          // trip counts differ between groups when n_mel % 8 != 0
                for (int i = 0; grp + 8 * i < P.n_mel; ++i) {
                    const int m = grp + 8 * i;
                    const int s = s_fbs[m], len = s_fbl[m];
                    const float* w = s_fbw + m * P.fb_stride;
                    double e = 0.0;
                    for (int t = sub; t < len; t += 4) e = fma((double)w[t], s_pow[s + t], e);
                    e += __shfl_xor_sync(gmask, e, 1);
                    e += __shfl_xor_sync(gmask, e, 2);
                    if (sub == 0) s_lm[m] = e;
                }
                __syncwarp();
                for (int m = lane; m < P.n_mel; m += 32) {                // all lanes take logs: 2 rounds instead of 5
                    double e = s_lm[m];
                    if (e == 0.0) e = 2.220446049250313e-16;              // np.finfo(float).eps
                    s_lm[m] = 8.685889638065035 * log(e);                 // 20 log10(e) = 20/ln(10) * ln(e)
                }
            }
            __syncwarp();
            // ---- DCT-II ortho (mfcc.py:152): lane = (half of n, coefficient k mod 16)
            {
                const int kk = lane & 15, part = lane >> 4;
                const int nh = (P.n_mel + 1) >> 1;
                const int n_lo = part * nh, n_hi = min(P.n_mel, n_lo + nh);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = kk + 16 * i;
                    double c = 0.0;
                    if (k < n_out)
                        for (int n = n_lo; n < n_hi; ++n) c = fma(s_dct[n * n_out + k], s_lm[n], c);
                    c += __shfl_xor_sync(0xffffffffu, c, 16);
                    if (k < n_out && part == 0) {
                        acc[i] += c;
                        // raw static row: cepstra 1..n_ceps then ln(sum power) (mfcc.py:158-160)
                        if (k >= 1 && k <= P.n_ceps) raw[f * RAW_LD + (k - 1)] = (float)c;
                        if (full_ceps) ceps_raw_out[f * P.n_mel + k] = (float)c;
                    }
                }
            }
            if (lane == 31) raw[f * RAW_LD + P.n_ceps] = (float)log(esum);
            __syncwarp();
        }
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (lane + 16 * i < n_out) atomicAdd(colsum + (int64_t)acc_u * MAX_MEL + lane + 16 * i, acc[i]);
        }
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
    P.twiddle = (const double2*)tb->twiddle; P.fb_start = tb->fb_start; P.fb_len = tb->fb_len;
    P.fb_weight = tb->fb_weight; P.dct_t = tb->dct_t;

    const SmemLayout L = smem_layout(P.frame_len, P.n_mel, P.fb_stride, ceps_out ? P.n_mel : P.n_ceps + 1);
    const size_t smem = (size_t)L.total * sizeof(double);
    const int64_t n_runs = ceil_div(total_frames, FRAMES_PER_RUN);
    const int blocks_per_sm = std::max(1, (int)((220 * 1024) / smem));
    int grid = (int)std::min<int64_t>(ceil_div(n_runs, WARPS_PER_BLOCK), (int64_t)sm_count() * blocks_per_sm);
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
