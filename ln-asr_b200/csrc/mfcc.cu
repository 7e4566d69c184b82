// mfcc.cu -- batched MFCC front end (kernel (1) of the north star).
//
// Replaces MFCC.calc_mfcc, /root/reference/lnasr/mfcc.py:108-175 (see include/lnasr_b200.h).
//
// Pass A  (mfcc_frames_kernel): a half-warp (16 threads) owns one frame at a time; the two halves of a warp walk
//   alternate frames of an 8-frame run.  Samples arrive by cp.async one frame ahead; pre-emphasis + framing + Hamming
//   on the way into registers (mfcc.py:63, utils.py:119-130, mfcc.py:138); 512-point real FFT as a 256-point complex
//   FFT done 16 x 16 in registers with one shared-memory transpose (mfcc.py:141); real-FFT split + power spectrum
//   (mfcc.py:142); triangular mel filters as a banded float32 product (mfcc.py:148; 453 non-zeros instead of a dense
//   257x40 GEMM); eps substitution + 20 log10 (mfcc.py:149-150); DCT-II ortho (mfcc.py:152); ln energy
//   (mfcc.py:158-160).  Because the DCT is linear, the two per-utterance mean removals (mfcc.py:151,156) commute
//   with it: pass A writes raw cepstra and accumulates per-utterance column sums in 32.32 fixed point.
// Pass B  (mfcc_finalize_kernel): subtract the means (+1e-8 constants), delta / delta-delta with the
//   reference's row-0 rule (mfcc.py:162-172) exchanged through warp shuffles, coalesced 39-float rows.
//
// HBM traffic per frame: 320 B pcm (int16) + 64 B raw write + 64 B raw read + 156 B features.
#include <type_traits>
#include "common.cuh"

namespace lnasr {

constexpr int FFT_N = 512;
constexpr int HALF_N = 256;
constexpr int MAX_MEL = 64;
constexpr int RAW_LD = 16;            // raw static row: 12 cepstra + energy, padded to 16 floats
constexpr int MAX_THREADS = 512;      // 16 warps = 32 half-warps per SM (one block per SM); one frame per half-warp at a time
constexpr int FRAMES_PER_RUN = 8;     // consecutive frames a warp walks (its two halves take alternate frames); small runs
                                      // keep the static schedule balanced for the few-ten-thousand-frame groups of the e2e path
constexpr int SHORT_RUN = 2;          // run length when the batch cannot give every resident warp a run of 8: one frame pair per warp,
                                      // so a single utterance is one latency-bound step deep instead of four (config 5, batch 1)
constexpr double COLSUM_SCALE = 4294967296.0;      // per-utterance column sums are 32.32 fixed point (|c| < 2^18, < 2^13 frames)
constexpr int TR_LD = 17;             // padded row of the 16x16 complex transpose (conflict-free 16-byte rows)
constexpr int MAX_SLOTS = MAX_MEL / 16;
constexpr int STD_FRAME_LEN = 400;    // MFCC() defaults (mfcc.py:23-48): the kStd instantiation of the frames kernel has
constexpr int STD_N_MEL = 40;         // the trip counts of its per-frame loops as constants
constexpr int STD_N_CEPS = 12;
constexpr int STD_FB_STRIDE = 32;     // widest mel filter the kStd kernel's shared-memory layout has room for (default bank: 32)
constexpr int STD_NA = (STD_FRAME_LEN + 31) >> 5;
// Mel plan (mfcc_setup_kernel -> workspace -> shared memory of the frames kernel).  ints: [0] slots, [1] padded taps of all slots,
// [2] 1 = interval form, [4 + s] taps of slot s (a multiple of 4), [PLAN_K0 + 16 s + b] first power bin lane b reads in slot s,
// [PLAN_J + 16 s + b] interval it sums (PLAN_NONE = none); from PLAN_W on: float2 weights [tap][lane].
constexpr int PLAN_SLOTS = (MAX_MEL + 16) / 16;       // n_mel + 1 intervals over 16 lanes
constexpr int PLAN_POS = PLAN_SLOTS * 16;
constexpr int PLAN_NONE = MAX_MEL + 1;
constexpr int PLAN_K0 = 16, PLAN_J = PLAN_K0 + PLAN_POS, PLAN_W = PLAN_J + PLAN_POS;
constexpr int PLAN_MAX_TAPS = PLAN_SLOTS * (64 + 4);
constexpr size_t PLAN_BYTES = ((size_t)PLAN_W * 4 + (size_t)PLAN_MAX_TAPS * 16 * 8 + 255) & ~(size_t)255;
constexpr int DCT_LD = 20;            // row stride (doubles) of the tensor-core DCT operand: the 4 x 4 doubles a half-warp loads
                                      // per fragment fall into 4 x 8 distinct banks

__device__ __forceinline__ void mfcc_dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct MfccDev {
    int frame_len, frame_step, n_mel, n_ceps, fb_stride;
    int hpf_q;                        // round(100 * hpf_alpha) when that is an integer (exact int16 pre-emphasis), else -1
    int frames_per_run;               // FRAMES_PER_RUN or SHORT_RUN
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
__device__ __forceinline__ double2 csqr(double2 a) { return make_double2(a.x * a.x - a.y * a.y, 2.0 * (a.x * a.y)); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// 4-point forward DFT in place (W4 = -i)
__device__ __forceinline__ void bfly4(double2& a0, double2& a1, double2& a2, double2& a3) {
    const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = make_double2(d02.x + d13.y, d02.y - d13.x);     // d02 - i d13
    a3 = make_double2(d02.x - d13.y, d02.y + d13.x);     // d02 + i d13
}

// 16-point forward DFT of v[0..15] (natural order in), fully in registers: two radix-4 passes with the W16 twiddles as
// immediates.  Output Y[k] is left in v[NAT16(k)].
#define NAT16(k) (4 * ((k) & 3) + ((k) >> 2))
__device__ __forceinline__ void fft16(double2 (&v)[16]) {
    constexpr double C1 = 0.92387953251128673848;    // cos(pi/8)
    constexpr double S1 = 0.38268343236508978178;    // sin(pi/8)
    constexpr double R2 = 0.70710678118654752440;    // sqrt(1/2)
#pragma unroll
    for (int j = 0; j < 4; ++j) bfly4(v[j], v[j + 4], v[j + 8], v[j + 12]);      // v[j + 4q] = T_j[q]
    // v[j + 4q] *= W16^(j q),  W16^n = (cos(2 pi n/16), -sin(2 pi n/16))
    v[5] = cmul(v[5], make_double2(C1, -S1));                                      // j=1,q=1 : n=1
    v[9] = make_double2(R2 * (v[9].x + v[9].y), R2 * (v[9].y - v[9].x));           // j=1,q=2 : n=2
    v[13] = cmul(v[13], make_double2(S1, -C1));                                    // j=1,q=3 : n=3
    v[6] = make_double2(R2 * (v[6].x + v[6].y), R2 * (v[6].y - v[6].x));           // j=2,q=1 : n=2
    v[10] = make_double2(v[10].y, -v[10].x);                                       // j=2,q=2 : n=4  (-i)
    v[14] = make_double2(R2 * (v[14].y - v[14].x), -R2 * (v[14].x + v[14].y));     // j=2,q=3 : n=6  (-r,-r)
    v[7] = cmul(v[7], make_double2(S1, -C1));                                      // j=3,q=1 : n=3
    v[11] = make_double2(R2 * (v[11].y - v[11].x), -R2 * (v[11].x + v[11].y));     // j=3,q=2 : n=6
    v[15] = cmul(v[15], make_double2(-C1, S1));                                    // j=3,q=3 : n=9
#pragma unroll
    for (int q = 0; q < 4; ++q) bfly4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);   // v[4q + p] = Y[q + 4p]
}

__device__ __forceinline__ double2 shfl_c(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// shared memory carve-up (in doubles)
struct SmemLayout {
    int tw_mid, tw_split, window, logtab, dct, fb_w, fb_meta, per_hw, per_hw_doubles, stage_off, total;
};
constexpr int STAGE_SAMPLES = FFT_N + 16;     // x[-1 .. 32 n_a + 1) plus alignment slack (copies start on a 16-byte boundary)
__host__ __device__ constexpr SmemLayout smem_layout(int n_mel, int fb_stride, int n_dct_cols, int pcm_bytes, int n_hw) {
    SmemLayout L{};
    int o = 0;
    L.tw_mid = o;   o += 2 * 256;                         // W256^(b k1) at [k1*16 + b]
    L.tw_split = o; o += 2 * 128;                         // W512^(b + 16 k2) at [k2*16 + b], k2 < 8
    L.window = o;   o += FFT_N;                           // window (x 1/100 on the integer path), zero past frame_len
    L.logtab = o;   o += 2 * 128;                         // (1/c_i, log c_i) of the table-driven log
    // DCT matrix: [n][k < n_dct_cols], or -- when at most 16 coefficients are wanted (the feature path) -- the B operand of
    // the float64 tensor-core DCT: [n < n_mel rounded up to 4][DCT_LD] with zero padding
    L.dct = o;      o += (n_dct_cols <= 16) ? ((n_mel + 3) & ~3) * DCT_LD : n_mel * n_dct_cols;
    L.fb_w = o;     o += ((n_mel + 16) / 16) * ((fb_stride + 3) & ~3) * 16;   // float2 weights [slot][tap][16 lanes], zero padded
    L.fb_meta = o;  o += PLAN_W / 2;                      // the integer part of the mel plan
    o = (o + 1) & ~1;
    L.per_hw = o;
    L.stage_off = 2 * 16 * TR_LD + MAX_MEL;               // transpose buffer | log-mel[64] | sample stage
    L.per_hw_doubles = (L.stage_off + (STAGE_SAMPLES * pcm_bytes + 7) / 8 + 1) & ~1;
    // the two half-warps of a warp read their sample stages with the same instruction: 64 bytes each, in disjoint banks when
    // the two stages are an odd multiple of 64 bytes apart
    while ((L.per_hw_doubles & 15) != 8) L.per_hw_doubles += 2;
    L.total = o + n_hw * L.per_hw_doubles;
    return L;
}

// where a frame lives: utterance, its sample range
struct FrameCtx {
    int u;
    int u_f0, u_f1;      // first frame of this / of the next utterance (a batch has < 2^31 frames: checked by the launcher)
    int64_t u_s0;        // first sample of the utterance in pcm
    int64_t u_len;       // samples in the utterance
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// The whole per-frame chain runs in float64.  The reference is float64 (mfcc.py), and the features are small
// differences of ~200 dB log-mel values (mean normalisation, deltas) whose low-energy bands sit up to 90 dB below
// the frame's peak: an fp32 FFT leaves ~1e-4 relative error in exactly those bands (measured), i.e. above the
// 1e-4 parity bar after the DCT and the delta-deltas.
//
// Mapping: 16 threads (a half-warp) own one frame; the two halves of a warp take alternate frames of a run of 32.
// * Samples reach the SM by cp.async into a per-half-warp stage one frame ahead (no registers held across the
//   latency), and are read back as three 2-byte loads per sample pair.
// * The 512-point real FFT is a 256-point complex FFT of the packed signal z[m] = (xw[2m], xw[2m+1]), done as
//   16 x 16: thread b holds z[16a + b], a = 0..15, in registers, runs a 16-point DFT over a, multiplies by
//   W256^(b k1), transposes once through shared memory (the only shared-memory pass of the FFT: 8 KB per frame
//   instead of the ~70 KB of a stage-by-stage shared-memory FFT), runs the second 16-point DFT over b.  Thread k1
//   then holds Z[k1 + 16 k2]; the real-FFT split pairs bin k with 256-k, which lives in lane 16-k1 -> one shuffle
//   per pair.
// * Mel filters are dealt to the 16 threads longest-first (balanced banded product); each thread takes the log of
//   its own filters; the DCT is one thread per coefficient.
template <typename PcmT, bool kStd>
__global__ void __launch_bounds__(MAX_THREADS, 1)
mfcc_frames_kernel(const PcmT* __restrict__ pcm, const int64_t* __restrict__ sample_off,
                   const int64_t* __restrict__ frame_off, int n_utt, int64_t total_frames, MfccDev P,
                   float* __restrict__ raw, unsigned long long* __restrict__ colsum, const int* __restrict__ run_utt,
                   const int* __restrict__ mel_plan, float* __restrict__ power_out, float* __restrict__ ceps_raw_out) {
    extern __shared__ __align__(16) double smem[];
    constexpr bool kIntPcm = (sizeof(PcmT) == 2);
    const int n_threads = blockDim.x;
    const int n_mel = kStd ? STD_N_MEL : P.n_mel, n_ceps = kStd ? STD_N_CEPS : P.n_ceps;
    const bool full_ceps = !kStd && (ceps_raw_out != nullptr);
    const int n_out = full_ceps ? n_mel : (n_ceps + 1);       // cepstral coefficients 0..n_out-1 computed
    // kStd: every offset of the carve-up is a compile-time constant (one base register instead of a dozen pointers that the
    // compiler either keeps -- the 16 x 16 FFT leaves it ~10 free registers -- or re-derives inside the frame loop)
    constexpr SmemLayout L_STD = smem_layout(STD_N_MEL, STD_FB_STRIDE, STD_N_CEPS + 1, (int)sizeof(PcmT), MAX_THREADS / 16);
    const SmemLayout L = kStd ? L_STD : smem_layout(n_mel, P.fb_stride, n_out, (int)sizeof(PcmT), n_threads >> 4);
    const int fb_stride = kStd ? STD_FB_STRIDE : P.fb_stride;
    double2* s_twm = reinterpret_cast<double2*>(smem + L.tw_mid);
    double2* s_tws = reinterpret_cast<double2*>(smem + L.tw_split);
    double* s_win = smem + L.window;
    double2* s_log = reinterpret_cast<double2*>(smem + L.logtab);
    double* s_dct = smem + L.dct;
    float2* s_fbw = reinterpret_cast<float2*>(smem + L.fb_w);
    int* s_plan = reinterpret_cast<int*>(smem + L.fb_meta);
    const int tid = threadIdx.x, lane = tid & 31, b = lane & 15, half = lane >> 4;
    const bool int_path = kIntPcm && P.hpf_q >= 0;
    const int n_rounds = (n_mel + 16) >> 4;              // slots of the mel product (n_mel + 1 intervals, one per lane and slot) =
                                                         // rounds of logarithms (n_mel filters + the frame energy)
    for (int i = tid; i < 256; i += n_threads) {
        const int k1 = i >> 4, bb = i & 15;
        s_twm[i] = P.twiddle[2 * ((bb * k1) & 255)];                 // W256^n = W512^(2n)
    }
    for (int i = tid; i < 128; i += n_threads) {
        s_tws[i] = P.twiddle[(i & 15) + 16 * (i >> 4)];
        const double c = 1.0 / (1.0 + ((double)i + 0.5) / 128.0);
        s_log[i] = make_double2(c, -log(c));
    }
    for (int i = tid; i < FFT_N; i += n_threads) s_win[i] = (i < P.frame_len) ? P.window[i] * (int_path ? 0.01 : 1.0) : 0.0;
    // DCT on the float64 tensor cores when at most 16 coefficients are wanted (features: 13); the API-parity path that returns
    // all n_mel cepstra keeps the SIMT loop
    const bool dct_mma = kStd || n_out <= 16;
    const int n_mel4 = (n_mel + 3) & ~3;
    if (dct_mma) {
        for (int i = tid; i < n_mel4 * DCT_LD; i += n_threads) {
            const int n = i / DCT_LD, k = i - n * DCT_LD;
            s_dct[i] = (n < n_mel && k < n_out) ? P.dct_t[n * n_mel + k] : 0.0;
        }
    } else {
        for (int i = tid; i < n_mel * n_out; i += n_threads) {
            const int n = i / n_out, k = i - n * n_out;
            s_dct[i] = P.dct_t[n * n_mel + k];
        }
    }
    // the mel plan of mfcc_setup_kernel: which power bins every lane multiplies with which weights
    for (int i = tid; i < PLAN_W; i += n_threads) s_plan[i] = mel_plan[i];
    {
        const int n_w = mel_plan[1] * 16;
        const float2* gw = reinterpret_cast<const float2*>(mel_plan + PLAN_W);
        for (int i = tid; i < n_w; i += n_threads) s_fbw[i] = gw[i];
    }
    __syncthreads();

    double* hbase = smem + L.per_hw + (tid >> 4) * L.per_hw_doubles;
    double2* tr = reinterpret_cast<double2*>(hbase);     // [16][TR_LD] complex
    // power spectra of the warp's TWO frames, interleaved [k][half] (float), in the transpose buffer of half 0 once the second
    // DFT has read it: the two half-warps run the same mel schedule, so lanes b and b + 16 read / write adjacent floats and a
    // warp-wide access is one wavefront instead of one per half-warp
    float* s_pow = reinterpret_cast<float*>(smem + L.per_hw + ((tid >> 5) * 2) * L.per_hw_doubles) + half;
    // (R, F) sums of the mel intervals of the warp's two frames, [interval][half], in the transpose buffer of half 1
    // two float arrays R[interval][half] | F[interval][half]: the intervals of one slot (and the filters of one logarithm round) are
    // consecutive, so 4-byte accesses at 2 j + half fall into 32 distinct banks (one float2 array collided two-fold)
    float* s_rf = reinterpret_cast<float*>(smem + L.per_hw + ((tid >> 5) * 2 + 1) * L.per_hw_doubles) + half;
    constexpr int RF_F = 2 * (PLAN_NONE + 1);             // offset of the F array
    double* s_lm = hbase + 2 * 16 * TR_LD;
    for (int i = b; i < MAX_MEL; i += 16) s_lm[i] = 0.0;                  // entries >= n_mel stay zero (k padding of the DCT)
    PcmT* stage = reinterpret_cast<PcmT*>(hbase + L.stage_off);

    const int fpr = P.frames_per_run, spr = fpr >> 1;
    const int n_frames = (int)total_frames;
    const int n_runs = (n_frames + fpr - 1) / fpr;
    const int warp_global = blockIdx.x * (n_threads >> 5) + (tid >> 5);
    const int warp_stride = gridDim.x * (n_threads >> 5);
    const int partner = ((16 - b) & 15) + 16 * half;     // lane holding Z[256 - k] for my Z[k]
    // rows a with at least one sample inside the frame; kStd = the default front end (25 ms frames at 16 kHz: 400 samples)
    // with the trip counts of the per-frame loops known to the compiler
    const int n_a = kStd ? STD_NA : ((P.frame_len + 31) >> 5);
    constexpr double PSCALE = 1.0 / (4.0 * FFT_N);       // the split below works on 2X
    constexpr int SPQ = 16 / (int)sizeof(PcmT);          // samples per 16-byte chunk
    const int n_words = ((32 * n_a + 2) * (int)sizeof(PcmT) + 3) / 4 + 2;   // 4-byte words covering x[-1 .. 32 n_a] at any alignment
    const int n_quads = (32 * n_a + 2 + 2 * SPQ - 2) / SPQ;   // 16-byte chunks covering them from any position inside the first
    const uintptr_t pcm_end = (uintptr_t)(pcm + __ldg(sample_off + n_utt)) - 1;      // last byte of the batch
    const uintptr_t pcm_lo = (uintptr_t)pcm & ~(uintptr_t)3, pcm_hi = pcm_end & ~(uintptr_t)3;
    const uintptr_t pcm_lo16 = (uintptr_t)pcm & ~(uintptr_t)15, pcm_hi16 = pcm_end & ~(uintptr_t)15;

    auto make_ctx = [&](int u) {
        FrameCtx c;
        c.u = u;
        c.u_f0 = (int)__ldg(frame_off + u);
        c.u_f1 = (int)__ldg(frame_off + u + 1);
        c.u_s0 = __ldg(sample_off + u);
        c.u_len = __ldg(sample_off + u + 1) - c.u_s0;
        return c;
    };
    // step s of this warp = pair of frames  f_pair(s), f_pair(s)+1 ; returns false past the warp's last pair (uniform)
    FrameCtx ctx;
    ctx.u = -1; ctx.u_f0 = 0; ctx.u_f1 = -1; ctx.u_s0 = 0; ctx.u_len = 0;
    int f = 0;
    bool act = false;
    int par = 0;                                         // position (in samples) of x[-1] in the stage
    auto locate = [&](int step) -> bool {
        const int run = warp_global + (step / spr) * warp_stride;
        const int f_pair = run * fpr + 2 * (step % spr);
        if (run >= n_runs || f_pair >= n_frames) return false;
        f = f_pair + half;
        act = f < n_frames;
        if (act) {
            if ((step % spr) == 0 || ctx.u < 0) ctx = make_ctx(__ldg(run_utt + run));   // utterance of the run's first frame
            if (f >= ctx.u_f1) {                         // walked into another utterance
                int u = ctx.u;
                int f1 = ctx.u_f1;
                while (f >= f1) { ++u; f1 = (int)__ldg(frame_off + u + 1); }
                ctx = make_ctx(u);
            }
        }
        return true;
    };
    // asynchronous copy of the samples x[-1 ..] of frame f into the stage: 16-byte chunks from the 16-byte boundary at or below
    // x[-1] (the chunks holding the first and the last byte of the batch are read whole), 4-byte words at the edges of the batch
    auto issue_copy = [&]() {
        if (act) {
            const uintptr_t first = (uintptr_t)(pcm + ctx.u_s0 + (int64_t)(f - ctx.u_f0) * P.frame_step - 1);
            const uintptr_t q0 = first & ~(uintptr_t)15;
            const uintptr_t w0 = first & ~(uintptr_t)3;
            if (q0 >= pcm_lo16 && q0 + 16u * (uintptr_t)(n_quads - 1) <= pcm_hi16) {
                // every chunk lies inside the PCM buffer (all frames but the first and the last few of the whole batch)
                par = (int)((first - q0) / sizeof(PcmT));
                const char* src = reinterpret_cast<const char*>(q0) + 16 * b;
                char* dst = reinterpret_cast<char*>(stage) + 16 * b;
                for (int j = b; j < n_quads; j += 16, src += 256, dst += 256) cp_async16(dst, src);
            } else {
                par = (int)((first - w0) / sizeof(PcmT));
                for (int j = b; j < n_words; j += 16) {
                    uintptr_t src = w0 + 4u * (uintptr_t)j;
                    src = src < pcm_lo ? pcm_lo : (src > pcm_hi ? pcm_hi : src);      // clamped words are never consumed
                    cp_async4(reinterpret_cast<char*>(stage) + 4 * j, reinterpret_cast<const void*>(src));
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    bool more = locate(0);
    if (more) issue_copy();
    int acc_u = -1;
    // running column sums of utterance acc_u (coefficient b + 16 i) in 32.32 fixed point: integer addition is associative, so
    // the per-utterance means -- and with them every feature -- do not depend on the order in which warps flush or on the
    // batch the utterance is scored in (a float64 atomicAdd would make the last bits scheduling-dependent)
    long long acc[MAX_SLOTS] = {0, 0, 0, 0};
    // coefficient whose sum lives in acc[i] of this lane: b + 16 i (SIMT DCT), or 8 (i / 2) + 2 b + i % 2 in lanes b < 4 (the
    // accumulator fragment of the tensor-core DCT)
    auto acc_coef = [&](int i) -> int { return dct_mma ? (b < 4 ? 8 * (i >> 1) + 2 * b + (i & 1) : -1) : b + 16 * i; };
    for (int step = 0; more; ++step) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (act && ctx.u != acc_u) {                      // flush the column sums of the utterance just left
#pragma unroll
            for (int i = 0; i < MAX_SLOTS; ++i) {
                const int k = acc_coef(i);
                if (acc_u >= 0 && k >= 0 && k < n_out)
                    atomicAdd(colsum + (int64_t)acc_u * MAX_MEL + k, (unsigned long long)acc[i]);
                acc[i] = 0;
            }
            acc_u = ctx.u;
        }
        // ---- pre-emphasis (mfcc.py:63) + framing (utils.py:119-130) + window (mfcc.py:138): z[16a+b], a = 0..15
        double2 v[16];
        {
            const int64_t s_start = (int64_t)(f - ctx.u_f0) * P.frame_step;  // first sample of the frame in its utterance
            const int64_t rem64 = act ? (ctx.u_len - s_start) : 0;          // samples from the frame origin on
            const int rem = (int)min(rem64, (int64_t)(2 * FFT_N));
            const bool lead = s_start > 0;                                  // x[-1] exists
            const PcmT* sx = stage + par + 1;                               // sx[n] = x[n]
            // CHK = false: the frame lies inside its utterance with a sample before it and 32 n_a after its origin (all but the
            // first and the last two or three frames of an utterance): none of the per-sample edge tests is needed
            auto load_frame = [&](auto chk) {
                constexpr bool CHK = decltype(chk)::value;
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    v[a] = make_double2(0.0, 0.0);
                    if (a < n_a) {
                        const int n0 = 32 * a + 2 * b;
                        const bool ok0 = !CHK || n0 < rem, ok1 = !CHK || n0 + 1 < rem;
                        PcmT xm, x0, x1;
                        if (kIntPcm) {
                            // x[n0-1], x[n0], x[n0+1] are 6 bytes inside two consecutive 4-byte words: two 32-bit loads and a
                            // funnel shift instead of three 2-byte loads
                            const uint32_t* wp = reinterpret_cast<const uint32_t*>(stage) + ((par + n0) >> 1);
                            const uint32_t w0 = wp[0], w1 = wp[1];
                            const uint32_t lo = __funnelshift_r(w0, w1, 16 * (par & 1));
                            xm = (PcmT)(short)(lo & 0xffffu);
                            x0 = (PcmT)(short)(lo >> 16);
                            x1 = (PcmT)(short)((w1 >> (16 * (par & 1))) & 0xffffu);
                        } else {
                            xm = sx[n0 - 1]; x0 = sx[n0]; x1 = sx[n0 + 1];
                        }
                        if (CHK) {
                            if (!ok0) x0 = PcmT(0);
                            if (!ok0 || !(n0 > 0 || lead)) xm = PcmT(0);
                        }
                        double y0, y1;
                        if (int_path) {
                            // y = x[t] - (q/100) x[t-1] = (100 x[t] - q x[t-1]) / 100, exact in int32; 1/100 is in s_win
                            const int i0 = 100 * (int)x0 - P.hpf_q * (int)xm;
                            const int i1 = ok1 ? 100 * (int)x1 - P.hpf_q * (int)x0 : 0;
                            y0 = (double)i0;
                            y1 = (double)i1;
                        } else {
                            // y[0] = x[0]; y[t] = x[t] - alpha x[t-1]; samples past the end are zero AFTER filtering
                            y0 = (double)x0 - P.hpf_alpha * (double)xm;
                            y1 = ok1 ? (double)x1 - P.hpf_alpha * (double)x0 : 0.0;
                        }
                        v[a] = make_double2(y0 * s_win[n0], y1 * s_win[n0 + 1]);
                    }
                }
            };
            if (kIntPcm && int_path && act && lead && rem >= 32 * n_a) {
                // interior frame of int16 audio: (x[n-1], x[n]) and (x[n], x[n+1]) are 32-bit words of the stage (one funnel
                // shift apart), and 100 x[n] - q x[n-1] is ONE two-way dot product of such a word with the bytes (-q, 100)
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(stage) + (par >> 1) + b;
                const bool odd = (par & 1) != 0;
                const int coef = (100 << 8) | ((-P.hpf_q) & 0xff);
                const double2* wn = reinterpret_cast<const double2*>(s_win) + b;
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    v[a] = make_double2(0.0, 0.0);
                    if (a < n_a) {
                        const uint32_t w0 = wp[16 * a], w1 = wp[16 * a + 1];
                        const uint32_t mid = __funnelshift_r(w0, w1, 16);
                        const int i0 = __dp2a_lo((int)(odd ? mid : w0), coef, 0);
                        const int i1 = __dp2a_lo((int)(odd ? w1 : mid), coef, 0);
                        const double2 wv = wn[16 * a];
                        v[a] = make_double2((double)i0 * wv.x, (double)i1 * wv.y);
                    }
                }
            } else if (act && lead && rem >= 32 * n_a) load_frame(std::false_type{});
            else load_frame(std::true_type{});
        }
        const int f_cur = f;
        const bool act_cur = act;
        __syncwarp();                                                      // the stage has been read by every lane
        more = locate(step + 1);
        if (more) issue_copy();                                            // next frame's samples fly during this frame's math
        // ---- 256-point complex FFT as 16 x 16 (mfcc.py:141)
        fft16(v);                                                          // over a  -> Y_b[k1] in v[NAT16(k1)]
        tr[b] = v[0];                                                      // k1 = 0: twiddle 1
        {
            // W256^(b k1), k1 = 1..15, from TWO table entries (k1 = 1 and 4) and 13 complex products: a 16-byte shared load
            // costs four wavefronts per warp whatever the two half-warps read, and shared-memory bandwidth, not the FP64
            // pipe, is what this kernel runs out of
            const double2 w1 = s_twm[16 + b], w4 = s_twm[64 + b];
            const double2 w2 = csqr(w1), w3 = cmul(w2, w1);
            tr[1 * TR_LD + b] = cmul(v[NAT16(1)], w1);
            tr[2 * TR_LD + b] = cmul(v[NAT16(2)], w2);
            tr[3 * TR_LD + b] = cmul(v[NAT16(3)], w3);
            tr[4 * TR_LD + b] = cmul(v[NAT16(4)], w4);
            tr[5 * TR_LD + b] = cmul(v[NAT16(5)], cmul(w4, w1));
            tr[6 * TR_LD + b] = cmul(v[NAT16(6)], cmul(w4, w2));
            tr[7 * TR_LD + b] = cmul(v[NAT16(7)], cmul(w4, w3));
            const double2 w8 = csqr(w4);
            tr[8 * TR_LD + b] = cmul(v[NAT16(8)], w8);
            tr[9 * TR_LD + b] = cmul(v[NAT16(9)], cmul(w8, w1));
            tr[10 * TR_LD + b] = cmul(v[NAT16(10)], cmul(w8, w2));
            tr[11 * TR_LD + b] = cmul(v[NAT16(11)], cmul(w8, w3));
            const double2 w12 = cmul(w8, w4);
            tr[12 * TR_LD + b] = cmul(v[NAT16(12)], w12);
            tr[13 * TR_LD + b] = cmul(v[NAT16(13)], cmul(w12, w1));
            tr[14 * TR_LD + b] = cmul(v[NAT16(14)], cmul(w12, w2));
            tr[15 * TR_LD + b] = cmul(v[NAT16(15)], cmul(w12, w3));
        }
        __syncwarp();
#pragma unroll
        for (int bb = 0; bb < 16; ++bb) v[bb] = tr[b * TR_LD + bb];
        fft16(v);                                                          // over b  -> Z[b + 16 k2] in v[NAT16(k2)]
        // ---- real-FFT split + power spectrum |X|^2 / 512 (mfcc.py:141-142), bins k = b + 16 k2 and 256 - k
        double esum = 0.0;
        __syncwarp();                                                      // every lane has read its row of tr
        const double2 ws0 = s_tws[b];                                      // W512^b; W512^(b + 16 k2) = W512^b W32^k2 (constants)
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            const double2 zk = v[NAT16(k2)];
            double2 zr = shfl_c(v[NAT16(15 - k2)], partner);
            if (b == 0) zr = v[NAT16((16 - k2) & 15)];                     // lane 0 pairs 16 k2 with 256 - 16 k2, both its own
            const double2 E = make_double2(zk.x + zr.x, zk.y - zr.y);      // 2 E[k]
            const double2 O = make_double2(zk.y + zr.y, zr.x - zk.x);      // 2 O[k]
            constexpr double W32C[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                        0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
            constexpr double W32S[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                        0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
            const double2 T = cmul(k2 == 0 ? ws0 : cmul(ws0, make_double2(W32C[k2], -W32S[k2])), O);
            const double px = E.x + T.x, py = E.y + T.y;                   // 2 X[k]
            const double qx = E.x - T.x, qy = E.y - T.y;                   // conj(2 X[256-k])
            const double pk = px * px + py * py, pr = qx * qx + qy * qy;
            s_pow[2 * (b + 16 * k2)] = (float)pk;
            s_pow[2 * (HALF_N - b - 16 * k2)] = (float)pr;
            esum += pk + pr;
        }
        {
            const double2 z8 = v[NAT16(8)];                                // bin 128 (lane 0): X = conj(Z[128])
            const double p8 = 4.0 * (z8.x * z8.x + z8.y * z8.y);
            if (b == 0) { s_pow[2 * 128] = (float)p8; esum += p8; }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
        // zero-weight taps of the padded mel loop may read up to fb_stride bins past bin 256: stale transpose data there must
        // at least be finite (0 * NaN)
        for (int i = lane; i < fb_stride + 4; i += 32)
            *reinterpret_cast<float2*>(s_pow - half + 2 * (HALF_N + 1 + i)) = make_float2(0.f, 0.f);
        __syncwarp();
        if (power_out && act_cur)
            for (int k = b; k <= HALF_N; k += 16) power_out[(int64_t)f_cur * (HALF_N + 1) + k] = s_pow[2 * k] * (float)PSCALE;
        // ---- mel filter bank (banded, mfcc.py:148), eps substitution (mfcc.py:149), 20 log10 (mfcc.py:150)
        // Interval form (mfcc_setup_kernel): neighbouring triangles overlap, so a bin k between two filter peaks feeds exactly
        // two filters, j = the higher one (weight wr[k]) and j - 1 (weight wf[k]); interval n_mel is the falling side of the last
        // filter.  A lane walks the bins of one interval and keeps R_j = sum wr P and F_j = sum wf P; the energy of filter m is
        // R_m + F_(m+1).  Every power value is loaded once
        // (256 bins instead of the 453 non-zeros of the filter matrix), every slot runs the same padded trip count on all lanes
        // (zero weights), and the plan shifts the lanes' first bins to distinct residues mod 16 -- the interleaved power
        // buffer [k][half] is then read without bank conflicts.  The power spectrum is a set of positive numbers: once the
        // float64 FFT has produced it, sums of <= 64 terms are safe in float32 (relative error ~1e-6, i.e. ~1e-5 dB); the
        // frame energy stays float64.
        {
            int w_base = 0;
#pragma unroll
            for (int s = 0; s < (kStd ? (STD_N_MEL + 16) / 16 : PLAN_SLOTS); ++s) {
                if (s < n_rounds) {
                    const int L = s_plan[4 + s];
                    const float2* w = s_fbw + w_base * 16 + b;
                    w_base += L;
                    const float* pw = s_pow + 2 * s_plan[PLAN_K0 + 16 * s + b];
                    float r = 0.f, fl = 0.f;
                    for (int t = 0; t < L; t += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 wv = w[16 * (t + u)];
                            const float pv = pw[2 * (t + u)];
                            r = fmaf(wv.x, pv, r);
                            fl = fmaf(wv.y, pv, fl);
                        }
                    }
                    const int jw = 2 * s_plan[PLAN_J + 16 * s + b];
                    s_rf[jw] = r;
                    s_rf[RF_F + jw] = fl;
                }
            }
            __syncwarp();
            // the rounds' logarithms are independent dependency chains (8 DFMA each): energies first, then the chains side by
            // side, then the stores
            constexpr int NR = kStd ? (STD_N_MEL + 16) / 16 : PLAN_SLOTS;
            double ev[NR];
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int m = b + 16 * i;
                double e = esum;                                           // m == n_mel: ln of the frame energy (mfcc.py:158-160)
                if (i < n_rounds && m < n_mel) e = (double)(s_rf[2 * m] + s_rf[RF_F + 2 * m + 2]);
                e *= PSCALE;
                if (m < n_mel && e == 0.0) e = 2.220446049250313e-16;      // np.finfo(float).eps
                ev[i] = (i < n_rounds && m <= n_mel) ? e : 1.0;
            }
            bool all_ok = true;
#pragma unroll
            for (int i = 0; i < NR; ++i) all_ok = all_ok && fast_log_ok(ev[i]);
            if (all_ok) {
#pragma unroll
                for (int i = 0; i < NR; ++i) ev[i] = fast_log_core(ev[i], s_log);
            } else {
#pragma unroll
                for (int i = 0; i < NR; ++i) ev[i] = fast_log(ev[i], s_log);
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int m = b + 16 * i;
                if (i < n_rounds) {
                    if (m < n_mel) s_lm[m] = 8.685889638065035 * ev[i];    // 20 log10(e) = 20/ln(10) * ln(e)
                    else if (m == n_mel && act_cur) raw[(int64_t)f_cur * RAW_LD + n_ceps] = (float)ev[i];
                }
            }
        }
        __syncwarp();
        // ---- DCT-II ortho (mfcc.py:152)
        if (dct_mma) {
            // On the float64 tensor cores: D[frame, k] = sum_n lm[frame, n] dct[n, k] as mma.m8n8k4 with the warp's two frames in
            // rows 0 and 4 of the A operand (lanes 0-3 of each half-warp supply their own half's log-mel values, every other
            // lane zero) and the coefficients 0..15 as two n-blocks.  The SIMT loop read one table entry (8 bytes) per
            // multiply-add; here a 4 x 8 block of the table feeds 256 of them, and the DCT stops being a shared-memory item.
            double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
            const double* bp = s_dct + (lane & 3) * DCT_LD + (lane >> 2);
#pragma unroll
            for (int kb = 0; kb < (kStd ? ((STD_N_MEL + 3) & ~3) : n_mel4); kb += 4) {
                const double av = (b < 4) ? s_lm[kb + b] : 0.0;
                mfcc_dmma884(d00, d01, av, bp[kb * DCT_LD]);
                mfcc_dmma884(d10, d11, av, bp[kb * DCT_LD + 8]);
            }
            if (b < 4 && act_cur) {                        // lanes 0-3 / 16-19 hold coefficients 2b, 2b+1 (+8) of frame half
                const double dv[4] = {d00, d01, d10, d11};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = 8 * (i >> 1) + 2 * b + (i & 1);
                    if (k < n_out) {
                        acc[i] += __double2ll_rn(dv[i] * COLSUM_SCALE);
                        if (k >= 1 && k <= n_ceps) raw[(int64_t)f_cur * RAW_LD + (k - 1)] = (float)dv[i];
                    }
                }
            }
        } else
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; ++i) {
            const int k = b + 16 * i;
            if (k < n_out) {
                double c0 = 0.0, c1 = 0.0;
                int n = 0;
                for (; n + 1 < n_mel; n += 2) {
                    c0 = fma(s_dct[n * n_out + k], s_lm[n], c0);
                    c1 = fma(s_dct[(n + 1) * n_out + k], s_lm[n + 1], c1);
                }
                if (n < n_mel) c0 = fma(s_dct[n * n_out + k], s_lm[n], c0);
                const double c = c0 + c1;
                if (act_cur) {
                    acc[i] += __double2ll_rn(c * COLSUM_SCALE);
                    // raw static row: cepstra 1..n_ceps then ln(sum power) (mfcc.py:158-160)
                    if (k >= 1 && k <= n_ceps) raw[(int64_t)f_cur * RAW_LD + (k - 1)] = (float)c;
                    if (full_ceps) ceps_raw_out[(int64_t)f_cur * n_mel + k] = (float)c;
                }
            }
        }
        __syncwarp();                                                      // s_pow / s_lm are rewritten by the next frame
    }
    if (acc_u >= 0) {
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; ++i) {
            const int k = acc_coef(i);
            if (k >= 0 && k < n_out) atomicAdd(colsum + (int64_t)acc_u * MAX_MEL + k, (unsigned long long)acc[i]);
        }
    }
}

// Set-up launch in front of the frames kernel.  Blocks 0 .. gridDim.x - 2: utterance of the first frame of every run (so the
// frames kernel starts a run with one load instead of a binary search).  Last block: the mel plan.
//
// Mel plan.  The caller passes the filter bank as banded rows (fb_start, fb_len, fb_weight).  The triangles of
// create_filter_bank (mfcc.py:66-106) overlap only with their neighbours, so bin k has at most two non-zero weights, in
// filters top(k) and top(k) - 1, and top(k) does not decrease with k: the bins with top(k) = j form interval j, and
// E_m = sum_{top(k) = m} W[m][k] P[k]  +  sum_{top(k) = m + 1} W[m][k] P[k]  =  R_m + F_(m+1).  The block verifies that
// structure; a bank without it is planned with interval j = the whole support of filter j (F = 0), which the frames kernel
// runs with the same code, only with longer trip counts.  Intervals are dealt to 16 lanes x slots longest-first; a slot
// runs max(len) rounded up to 4 taps on every lane; a lane with slack starts up to `slack` bins early (zero weights) on a
// bin whose residue mod 16 no other lane of the slot uses.
__global__ void __launch_bounds__(256)
mfcc_setup_kernel(const int64_t* __restrict__ frame_off, int n_utt, int64_t n_runs, int frames_per_run, int* __restrict__ run_utt,
                  int n_mel, int fb_stride, const int* __restrict__ fb_start, const int* __restrict__ fb_len,
                  const float* __restrict__ fb_weight, int* __restrict__ plan) {
    if (!plan || blockIdx.x + 1 < gridDim.x) {
        const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (r < n_runs) run_utt[r] = find_utt(frame_off, n_utt, r * frames_per_run);
        return;
    }
    __shared__ int s_iv[HALF_N + 1];                  // interval of bin k (-1: no filter covers it)
    __shared__ float s_wr[HALF_N + 1], s_wf[HALF_N + 1];
    __shared__ int s_k0[PLAN_POS], s_len[PLAN_POS], s_ord[PLAN_POS], s_asg[PLAN_POS], s_off[PLAN_POS];
    __shared__ int s_bad, s_L[PLAN_SLOTS], s_base[PLAN_SLOTS + 1];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n_iv = n_mel + 1;                        // intervals 0 .. n_mel
    const int n_slots = (n_iv + 15) >> 4;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // a bin with two non-zero weights, in neighbouring filters m - 1 and m, belongs to interval m (rising side of m, falling side
    // of m - 1); a bin with one, in filter m, to interval m (weight not below the previous bin's: rising side or peak) or to
    // interval m + 1 (falling side); anything else: no interval form
    for (int k = tid; k <= HALF_N; k += nt) {
        int top = -1, second = -1, cnt = 0;
        float wt = 0.f, ws = 0.f, wprev = 0.f;
        for (int m = 0; m < n_mel; ++m) {
            const int t = k - fb_start[m];
            if (t >= 0 && t < fb_len[m]) {
                const float w = fb_weight[m * fb_stride + t];
                if (w != 0.f) { ++cnt; second = top; ws = wt; top = m; wt = w; wprev = t > 0 ? fb_weight[m * fb_stride + t - 1] : 0.f; }
            }
        }
        if (cnt > 2 || (cnt == 2 && second != top - 1)) s_bad = 1;
        int iv = top;
        float wr = wt, wf = (cnt == 2) ? ws : 0.f;
        if (cnt == 1 && wt < wprev) { iv = top + 1; wr = 0.f; wf = wt; }
        s_iv[k] = iv;
        s_wr[k] = wr;
        s_wf[k] = wf;
    }
    __syncthreads();
    const bool compact = !s_bad;
    for (int j = tid; j < PLAN_POS; j += nt) {
        int k0 = 0, len = 0;
        if (j < n_iv) {
            if (compact) {                              // span of the bins of interval j (foreign bins inside it get zero weights)
                int first = -1, last = -1;
                for (int k = 0; k <= HALF_N; ++k)
                    if (s_iv[k] == j) { if (first < 0) first = k; last = k; }
                if (first >= 0) { k0 = first; len = last - first + 1; }
            } else if (j < n_mel) {
                k0 = fb_start[j];
                len = fb_len[j];
            }
        }
        s_k0[j] = k0;
        s_len[j] = len;
        s_asg[j] = -1;
        s_off[j] = 0;
    }
    __syncthreads();
    // longest-processing-time assignment of the intervals to the 16 lanes of a frame.  Step 1 (parallel): rank by descending length.
    if (tid < n_iv) {
        const int len = s_len[tid];
        int r = 0;
        for (int m = 0; m < n_iv; ++m) {
            const int lm = s_len[m];
            r += (lm > len) || (lm == len && m < tid);
        }
        s_ord[r] = tid;
    }
    __syncthreads();
    // Step 2 (one thread, keys = (full?, load, lane) packed in registers): greedy placement.
    if (tid == 0) {
        int key[16], cnt[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) { key[t] = t; cnt[t] = 0; }
        for (int r = 0; r < n_iv; ++r) {
            const int m = s_ord[r];
            int best = key[0];
#pragma unroll
            for (int t = 1; t < 16; ++t) best = min(best, key[t]);
            const int tb = best & 15;
            const int add = s_len[m] << 4;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if (t == tb) {
                    s_asg[cnt[t] * 16 + t] = m;
                    ++cnt[t];
                    key[t] += add;
                    if (cnt[t] >= n_slots) key[t] |= 1 << 28;
                }
        }
    }
    __syncthreads();
    if (tid < n_slots) {
        int L = 0;
        for (int bb = 0; bb < 16; ++bb) {
            const int j = s_asg[tid * 16 + bb];
            if (j >= 0) L = max(L, s_len[j]);
        }
        L = (L + 3) & ~3;
        s_L[tid] = L;
        // early starts: a lane may begin up to `slack` bins before its interval; choose the shifts so that the 16 first bins
        // have distinct residues mod 16 -- a bipartite matching lane <-> residue (augmenting paths, most constrained lanes
        // first); lanes without an interval take whatever residue is left (their weights are zero)
        int k0v[16], slk[16], owner[16];
        for (int bb = 0; bb < 16; ++bb) {
            const int j = s_asg[tid * 16 + bb];
            k0v[bb] = j >= 0 ? s_k0[j] : 15;
            slk[bb] = j >= 0 ? min(15, min(L - s_len[j], s_k0[j])) : 15;
            owner[bb] = -1;
        }
        for (int want = 0; want <= 15; ++want)
            for (int b0 = 0; b0 < 16; ++b0) {
                if (slk[b0] != want) continue;
                int st_lane[16], st_r[16], st_try[16], sp = 0;
                unsigned visited = 0;
                st_lane[0] = b0;
                st_r[0] = 0;
                bool done = false;
                while (sp >= 0 && !done) {
                    const int b = st_lane[sp];
                    bool descended = false;
                    while (st_r[sp] < 16) {
                        const int r = st_r[sp]++;
                        if (((k0v[b] - r) & 15) > slk[b] || ((visited >> r) & 1u)) continue;
                        visited |= 1u << r;
                        st_try[sp] = r;
                        if (owner[r] < 0) {
                            for (int q = sp; q >= 0; --q) owner[st_try[q]] = st_lane[q];
                            done = true;
                            break;
                        }
                        if (sp + 1 < 16) {
                            st_lane[sp + 1] = owner[r];
                            st_r[sp + 1] = 0;
                            ++sp;
                            descended = true;
                            break;
                        }
                    }
                    if (!done && !descended) --sp;
                }
            }
        for (int bb = 0; bb < 16; ++bb) s_off[tid * 16 + bb] = (s_asg[tid * 16 + bb] >= 0) ? 0 : 15;     // unmatched: no shift
        for (int r = 0; r < 16; ++r)
            if (owner[r] >= 0) {
                const int bb = owner[r], c = (k0v[bb] - r) & 15;
                s_off[tid * 16 + bb] = (s_asg[tid * 16 + bb] >= 0) ? c : 15 - c;      // shift of an interval / first bin of an idle lane
            }
    }
    __syncthreads();
    if (tid == 0) {
        s_base[0] = 0;
        for (int sl = 0; sl < n_slots; ++sl) s_base[sl + 1] = s_base[sl] + s_L[sl];
        plan[0] = n_slots;
        plan[1] = s_base[n_slots];
        plan[2] = compact ? 1 : 0;
        plan[3] = 0;
        for (int sl = 0; sl < PLAN_SLOTS; ++sl) plan[4 + sl] = sl < n_slots ? s_L[sl] : 0;
    }
    __syncthreads();
    for (int i = tid; i < PLAN_POS; i += nt) {
        const int j = (i < n_slots * 16) ? s_asg[i] : -1;
        plan[PLAN_K0 + i] = j >= 0 ? s_k0[j] - s_off[i] : ((i < n_slots * 16) ? s_off[i] : 0);
        plan[PLAN_J + i] = j >= 0 ? j : PLAN_NONE;
    }
    float2* pw = reinterpret_cast<float2*>(plan + PLAN_W);
    const int n_w = s_base[n_slots] * 16;
    for (int i = tid; i < n_w; i += nt) {
        const int tt = i >> 4, bb = i & 15;
        int sl = 0;
        while (sl + 1 < n_slots && tt >= s_base[sl + 1]) ++sl;
        const int j = s_asg[sl * 16 + bb];
        float2 w = make_float2(0.f, 0.f);
        if (j >= 0) {
            const int k = s_k0[j] - s_off[sl * 16 + bb] + (tt - s_base[sl]);
            if (k >= s_k0[j] && k < s_k0[j] + s_len[j]) {
                if (compact) { if (s_iv[k] == j) w = make_float2(s_wr[k], s_wf[k]); }
                else w = make_float2(fb_weight[j * fb_stride + (k - fb_start[j])], 0.f);
            }
        }
        pw[i] = w;
    }
}

// Pass B.  One warp produces 30 consecutive frames (lanes 0,1 carry the two-frame halo); lane = frame so
// delta / delta-delta are neighbour exchanges by __shfl_up_sync.  The 30x39 output block is contiguous in
// HBM (rows of 39 floats back to back when ld_feats == 39) and is written through a shared-memory
// transpose so that consecutive lanes store consecutive addresses.
constexpr int FIN_WARPS = 8;
constexpr int FIN_FRAMES = 30;

__global__ void __launch_bounds__(FIN_WARPS * 32)
mfcc_finalize_kernel(const float* __restrict__ raw, const long long* __restrict__ colsum,
                     const int64_t* __restrict__ frame_off, int n_utt, int64_t total_frames, int n_ceps, int n_mel,
                     const int* __restrict__ run_utt, int frames_per_run,
                     float* __restrict__ feats, int ld_feats, float* __restrict__ ceps_io, float dct_dc) {
    __shared__ float tile[FIN_WARPS][32][40];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ns = n_ceps + 1;                                  // 13 static coefficients
    const int64_t n_tiles = (total_frames + FIN_FRAMES - 1) / FIN_FRAMES;
    for (int64_t t = (int64_t)blockIdx.x * FIN_WARPS + warp; t < n_tiles; t += (int64_t)gridDim.x * FIN_WARPS) {
        const int64_t f = t * FIN_FRAMES + lane - 2;            // lanes 0,1: halo frames
        const bool valid = (f >= 0 && f < total_frames);
        // the frame's raw row (16 floats, one 64-byte line) does not depend on the utterance: in flight during the look-up
        float4 r4[4];
        if (valid) {
            const float4* rp = reinterpret_cast<const float4*>(raw + f * RAW_LD);
#pragma unroll
            for (int i = 0; i < 4; ++i) r4[i] = __ldg(rp + i);
        }
        int u = 0;
        int64_t f0 = 0, f1 = 1;
        if (valid) {
            // the set-up launch left the utterance of the first frame of every run: a short walk instead of a binary search
            u = __ldg(run_utt + f / frames_per_run);
            f1 = __ldg(frame_off + u + 1);
            while (f >= f1) { ++u; f1 = __ldg(frame_off + u + 1); }
            f0 = __ldg(frame_off + u);
        }
        const int64_t fl = f - f0;                               // frame index inside the utterance
        const double inv_cnt = 1.0 / (double)(f1 - f0);
        float rn4[4][4];
        if (valid && fl == 0) {                                  // row 1 of the utterance, for the row-0 rule
            const float4* rp = reinterpret_cast<const float4*>(raw + (f + 1) * RAW_LD);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float4 q = __ldg(rp + i); rn4[i][0] = q.x; rn4[i][1] = q.y; rn4[i][2] = q.z; rn4[i][3] = q.w; }
        }
        const float rr[16] = {r4[0].x, r4[0].y, r4[0].z, r4[0].w, r4[1].x, r4[1].y, r4[1].z, r4[1].w,
                              r4[2].x, r4[2].y, r4[2].z, r4[2].w, r4[3].x, r4[3].y, r4[3].z, r4[3].w};
#pragma unroll
        for (int c = 0; c < RAW_LD; ++c) {
            if (c < ns) {
                float s = 0.f, s_next = 0.f;
                if (valid) {
                    // feature = raw - (mean + 1e-8) for the cepstra (mfcc.py:151,156 folded, see file header); energy as is
                    const double mean = (c < n_ceps) ? ((double)__ldg(colsum + (int64_t)u * MAX_MEL + c + 1) * (1.0 / COLSUM_SCALE) * inv_cnt + 1e-8) : 0.0;
                    s = (float)((double)rr[c] - mean);
                    if (fl == 0) s_next = (float)((double)rn4[c >> 2][c & 3] - mean);
                }
                const float p1 = __shfl_up_sync(0xffffffffu, s, 1);
                const float p2 = __shfl_up_sync(0xffffffffu, s, 2);
                // delta[0] = static[1]; delta[t] = static[t] - static[t-1]                       (mfcc.py:162-166)
                const float d = (fl == 0) ? s_next : (s - p1);
                // delta of the previous frame, needed for delta-delta
                //   prev frame fl-1 == 0  -> delta_prev = static[1] = s (me)
                const float d_prev = (fl == 1) ? s : (p1 - p2);
                // ddelta[0] = delta[1] = static[1]-static[0]; ddelta[t] = delta[t]-delta[t-1]   (mfcc.py:168-172)
                const float dd = (fl == 0) ? (s_next - s) : (d - d_prev);
                tile[warp][lane][c] = s;
                tile[warp][lane][ns + c] = d;
                tile[warp][lane][2 * ns + c] = dd;
            }
        }
        __syncwarp();
        // coalesced store of the 30 x 39 block
        const int64_t fb = t * FIN_FRAMES;
        const int nf = (int)min((int64_t)FIN_FRAMES, total_frames - fb);
        const int D = 3 * ns;
        {
            int r = lane / D, c = lane - r * D;                    // element i = lane + 32 n of the block is (row r, column c)
            float* out = feats + fb * ld_feats;
            for (int i = lane; i < nf * D; i += 32) {
                out[(int64_t)r * ld_feats + c] = tile[warp][r + 2][c];
                c += 32;
                while (c >= D) { c -= D; ++r; }
            }
        }
        __syncwarp();
        // optional second return value of calc_mfcc: DCT of the mean-removed log-mel, columns 1..n_ceps
        // additionally mean-removed in place (mfcc.py:155-156)
        if (ceps_io) {
            for (int r = 0; r < nf; ++r) {
                const int64_t ff = fb + r;
                const int uu = find_utt(frame_off, n_utt, ff);
                const double ic = 1.0 / (double)(__ldg(frame_off + uu + 1) - __ldg(frame_off + uu));
                for (int k = lane; k < n_mel; k += 32) {
                    double mean = (double)__ldg(colsum + (int64_t)uu * MAX_MEL + k) * (1.0 / COLSUM_SCALE) * ic;
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

extern "C" size_t lnasr_mfcc_mel_plan_bytes(void) { return PLAN_BYTES; }

extern "C" int lnasr_mfcc_mel_plan(const lnasr_mfcc_tables* tb, int32_t* plan, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(tb && plan && tb->fb_start && tb->fb_len && tb->fb_weight, "mfcc_mel_plan: null pointer");
    LNASR_CHECK_ARG(tb->n_mel >= 2 && tb->n_mel <= MAX_MEL, "mfcc_mel_plan: n_mel must be in [2,64]");
    LNASR_CHECK_ARG(tb->fb_stride >= 1 && tb->fb_stride <= 64, "mfcc_mel_plan: fb_stride must be in [1,64]");
    // the plan block is the last block of the grid; with zero runs it is the only one
    mfcc_setup_kernel<<<1, 256, 0, stream>>>(nullptr, 0, 0, 1, nullptr, tb->n_mel, tb->fb_stride, tb->fb_start, tb->fb_len,
                                              tb->fb_weight, plan);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_mfcc_workspace_bytes(int64_t total_frames, int32_t n_utt, size_t* bytes) {
    LNASR_CHECK_ARG(bytes && total_frames >= 0 && n_utt >= 0, "mfcc_workspace_bytes: bad arguments");
    size_t raw = (size_t)total_frames * RAW_LD * sizeof(float);
    raw = (raw + 255) & ~(size_t)255;
    size_t col = ((size_t)n_utt * MAX_MEL * sizeof(double) + 255) & ~(size_t)255;
    const size_t runs = ((size_t)ceil_div(total_frames, SHORT_RUN) * sizeof(int) + 255) & ~(size_t)255;
    *bytes = raw + col + runs + PLAN_BYTES + 256;
    return LNASR_OK;
}

extern "C" int lnasr_mfcc_batch(const void* pcm, int pcm_dtype, const int64_t* sample_offsets,
                                const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                                const lnasr_mfcc_tables* tb, float* feats, int32_t ld_feats, float* power_out,
                                float* ceps_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(pcm && sample_offsets && frame_offsets && tb && feats && workspace, "mfcc_batch: null pointer");
    LNASR_CHECK_ARG(pcm_dtype == LNASR_DT_I16 || pcm_dtype == LNASR_DT_F32 || pcm_dtype == LNASR_DT_F64,
                    "mfcc_batch: pcm dtype must be int16, float32 or float64");
    LNASR_CHECK_ARG(tb->fft_n == FFT_N, "mfcc_batch: only fft_N=512 is built (got %d)", tb->fft_n);
    LNASR_CHECK_ARG(tb->frame_len >= 2 && tb->frame_len <= FFT_N && tb->frame_step >= 1, "mfcc_batch: frame_len must be in [2,512]");
    LNASR_CHECK_ARG(tb->n_mel >= 2 && tb->n_mel <= MAX_MEL, "mfcc_batch: n_mel must be in [2,64]");
    LNASR_CHECK_ARG(tb->n_ceps >= 1 && tb->n_ceps < tb->n_mel && tb->n_ceps <= 31, "mfcc_batch: n_ceps must be < n_mel and <= 31");
    LNASR_CHECK_ARG(tb->fb_stride >= 1 && tb->fb_stride <= 64, "mfcc_batch: fb_stride must be in [1,64]");
    LNASR_CHECK_ARG(ld_feats >= 3 * (tb->n_ceps + 1), "mfcc_batch: ld_feats too small");
    LNASR_CHECK_ARG(3 * (tb->n_ceps + 1) <= 40 && tb->n_ceps + 1 <= RAW_LD, "mfcc_batch: at most 13 static coefficients are built");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    LNASR_CHECK_ARG(total_frames < ((int64_t)1 << 31) - 64, "mfcc_batch: at most 2^31 frames per call (got %lld)", (long long)total_frames);
    size_t need = 0;
    lnasr_mfcc_workspace_bytes(total_frames, n_utt, &need);
    if (workspace_bytes < need) {
        set_error("mfcc_batch: workspace %zu < %zu bytes", workspace_bytes, need);
        return LNASR_EWORKSPACE;
    }
    float* raw = (float*)workspace;
    size_t raw_bytes = ((size_t)total_frames * RAW_LD * sizeof(float) + 255) & ~(size_t)255;
    unsigned long long* colsum = (unsigned long long*)((char*)workspace + raw_bytes);
    LNASR_CUDA(cudaMemsetAsync(colsum, 0, (size_t)n_utt * MAX_MEL * sizeof(double), stream));
    const size_t col_bytes = ((size_t)n_utt * MAX_MEL * sizeof(double) + 255) & ~(size_t)255;
    int* run_utt = (int*)((char*)workspace + raw_bytes + col_bytes);
    // runs of 8 frames per warp once there are enough frames to give every resident warp (16 per SM) such a run, else frame pairs
    const int fpr = (total_frames >= (int64_t)FRAMES_PER_RUN * 16 * sm_count()) ? FRAMES_PER_RUN : SHORT_RUN;
    // the mel plan: the caller's (built once per filter bank), else planned here by one more block of the set-up launch
    const int* mel_plan = tb->mel_plan;
    int* own_plan = nullptr;
    if (!mel_plan) mel_plan = own_plan = (int*)((char*)run_utt + (((size_t)ceil_div(total_frames, SHORT_RUN) * sizeof(int) + 255) & ~(size_t)255));
    {
        const int64_t nr = ceil_div(total_frames, fpr);
        mfcc_setup_kernel<<<(unsigned)ceil_div(nr, 256) + (own_plan ? 1 : 0), 256, 0, stream>>>(
            frame_offsets, n_utt, nr, fpr, run_utt, tb->n_mel, tb->fb_stride, tb->fb_start, tb->fb_len, tb->fb_weight, own_plan);
        LNASR_LAUNCH_CHECK();
    }

    MfccDev P;
    P.frame_len = tb->frame_len; P.frame_step = tb->frame_step; P.n_mel = tb->n_mel; P.n_ceps = tb->n_ceps;
    P.fb_stride = tb->fb_stride; P.hpf_alpha = tb->hpf_alpha; P.window = tb->window; P.frames_per_run = fpr;
    {
        const double q = tb->hpf_alpha * 100.0;
        const long qi = lround(q);
        P.hpf_q = (fabs(q - (double)qi) < 1e-9 && qi >= 0 && qi <= 100) ? (int)qi : -1;
    }
    P.twiddle = (const double2*)tb->twiddle; P.fb_start = tb->fb_start; P.fb_len = tb->fb_len;
    P.fb_weight = tb->fb_weight; P.dct_t = tb->dct_t;

    // one block per SM, as many half-warps (frames in flight) as shared memory allows, at most 32
    // the kStd instantiation: MFCC() defaults on the feature path (compile-time trip counts and shared-memory offsets)
    const bool std_shape = (P.frame_len == STD_FRAME_LEN && P.n_mel == STD_N_MEL && P.n_ceps == STD_N_CEPS && P.fb_stride <= STD_FB_STRIDE &&
                            !ceps_out && pcm_dtype != LNASR_DT_F64);
    const int pcm_bytes = (pcm_dtype == LNASR_DT_I16) ? 2 : (pcm_dtype == LNASR_DT_F32 ? 4 : 8);
    const int n_cols = ceps_out ? P.n_mel : P.n_ceps + 1;
    int threads = MAX_THREADS;
    size_t smem = 0;
    for (;; threads -= 64) {
        smem = (size_t)smem_layout(P.n_mel, std_shape ? STD_FB_STRIDE : P.fb_stride, n_cols, pcm_bytes, threads / 16).total * sizeof(double);
        if (smem <= 226 * 1024 || threads <= 64) break;
    }
    const int64_t n_runs = ceil_div(total_frames, fpr);
    int grid = (int)std::min<int64_t>(ceil_div(n_runs, threads / 32), (int64_t)sm_count());
    if (grid < 1) grid = 1;
    auto launch = [&](auto kern, auto* typed_pcm) -> int {
        LNASR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, stream>>>(typed_pcm, sample_offsets, frame_offsets, n_utt, total_frames, P, raw, colsum, run_utt,
                                              mel_plan, power_out, ceps_out);
        return LNASR_OK;
    };
    int rc;
    if (pcm_dtype == LNASR_DT_I16)
        rc = std_shape ? launch(mfcc_frames_kernel<int16_t, true>, (const int16_t*)pcm) : launch(mfcc_frames_kernel<int16_t, false>, (const int16_t*)pcm);
    else if (pcm_dtype == LNASR_DT_F32)
        rc = std_shape ? launch(mfcc_frames_kernel<float, true>, (const float*)pcm) : launch(mfcc_frames_kernel<float, false>, (const float*)pcm);
    else   // float64 samples (what the reference's callers pass, mfcc.py:108): staged as doubles, fewer frames in flight per SM
        rc = launch(mfcc_frames_kernel<double, false>, (const double*)pcm);
    if (rc != LNASR_OK) return rc;
    LNASR_LAUNCH_CHECK();
    const int64_t n_tiles = ceil_div(total_frames, FIN_FRAMES);
    int grid2 = (int)std::min<int64_t>(ceil_div(n_tiles, FIN_WARPS), (int64_t)sm_count() * 8);
    if (grid2 < 1) grid2 = 1;
    // DCT-II ortho of a constant vector c is c*sqrt(n_mel) in coefficient 0 and 0 elsewhere
    const float dct_dc = sqrtf((float)tb->n_mel);
    mfcc_finalize_kernel<<<grid2, FIN_WARPS * 32, 0, stream>>>(raw, (const long long*)colsum, frame_offsets, n_utt, total_frames,
                                                              tb->n_ceps, tb->n_mel, run_utt, fpr, feats, ld_feats, ceps_out, dct_dc);
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
