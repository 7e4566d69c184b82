/*
 * lnasr_b200.h -- C ABI of liblnasr_b200.so: the ln-asr acoustic hot path on one B200 (sm_100a).
 *
 * The reference (yehuohan/ln-asr) has NO FFI for this path: the boundary is the Python API of
 * lnasr/mfcc.py, lnasr/gmm.py, lnasr/hmm.py and lnasr/gmmhmm.py.  Each entry
 * point below names the reference method whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes stubs a maintainer of the reference would add to bind them.
 *
 * Conventions
 *   - every function returns 0 on success, a negative LNASR_E* code otherwise; the message of the
 *     last failure on the calling thread is returned by lnasr_last_error();
 *   - all data pointers are DEVICE pointers unless the parameter comment says "host";
 *   - the caller owns every buffer; the library allocates nothing that outlives a call;
 *   - every launch goes to the CUDA stream passed as `stream` (a cudaStream_t cast to void*); no call
 *     synchronises the device;
 *   - utterances are ragged and concatenated: frames of utterance u are rows
 *     [frame_offsets[u], frame_offsets[u+1]) of every per-frame matrix;
 *   - per-frame matrices are FRAME-major ([T_total, n_models*N]); the reference keeps b state-major
 *     [N, T] (lnasr/hmm.py:59) -- the Python mirror transposes on the way out;
 *   - log quantities are natural logs; "model" k of a bank owns states [k*N, (k+1)*N).
 */
#ifndef LNASR_B200_H
#define LNASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNASR_OK            0
#define LNASR_EINVAL       -1   /* bad argument / unsupported shape */
#define LNASR_ECUDA        -2   /* a CUDA runtime call or launch failed */
#define LNASR_EWORKSPACE   -3   /* workspace too small */

#define LNASR_DT_I16 0
#define LNASR_DT_F32 1
#define LNASR_DT_F64 2

/* ------------------------------------------------------------------------------------------------ */
const char* lnasr_last_error(void);
int lnasr_version(void);
/* host out-params; reports the current device */
int lnasr_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes);

/* ================================ (1) MFCC front end ============================================
 * Replaces MFCC.calc_mfcc (lnasr/mfcc.py:108-175) including calc_high_pass_filter (mfcc.py:50-64),
 * utils.split_frames (utils.py:106-130), utils.create_hamming (utils.py:132-138), the 512-point rFFT
 * power spectrum (mfcc.py:141-142), the mel filter bank product (mfcc.py:145-148), 20*log10 + per
 * utterance mean removal (mfcc.py:149-151), DCT-II ortho (mfcc.py:152), c1..c12 + ln energy
 * (mfcc.py:155-160) and delta / delta-delta (mfcc.py:162-172), for a batch of utterances.
 *
 * The constant tables are built by the caller with the reference's formulas (create_hamming,
 * create_filter_bank, the DCT-II ortho matrix) and passed as device arrays. */
typedef struct lnasr_mfcc_tables {
    int32_t frame_len;        /* MFCC.frame_L   (400) ; must be <= fft_n */
    int32_t frame_step;       /* MFCC.frame_step (160) */
    int32_t fft_n;            /* MFCC.fft_N; this build supports 512 only */
    int32_t n_mel;            /* MFCC.mfbank_num (40) ; <= 64 */
    int32_t n_ceps;           /* MFCC.cepstrum_num (12) ; < n_mel, <= 31 */
    double  hpf_alpha;        /* MFCC.hpf_alpha (0.97) */
    int32_t fb_stride;        /* row stride of fb_weight (>= max filter width, <= 64) */
    const double*  window;    /* [frame_len]  Hamming window (float64) */
    const double*  twiddle;   /* [fft_n] pairs (cos, -sin)(2 pi m / fft_n) as double2 */
    const int32_t* fb_start;  /* [n_mel] first FFT bin with non-zero weight */
    const int32_t* fb_len;    /* [n_mel] number of consecutive non-zero bins */
    const float*   fb_weight; /* [n_mel, fb_stride] the non-zero weights of each triangular filter */
    const double*  dct_t;     /* [n_mel(n), n_mel(k)] DCT-II ortho matrix (float64), transposed: dct_t[n*n_mel+k] */
    const int32_t* mel_plan;  /* optional: device buffer of lnasr_mfcc_mel_plan_bytes() bytes filled ONCE by lnasr_mfcc_mel_plan()
                                 for this filter bank (how the kernel walks the banded product); NULL = lnasr_mfcc_batch
                                 derives it in its workspace on every call (~25 us) */
} lnasr_mfcc_tables;

/* MFCC.calc_high_pass_filter (lnasr/mfcc.py:50-64) alone: y[0]=x[0]; y[t]=x[t]-alpha*x[t-1]; float64 out */
int lnasr_preemphasis(const void* pcm, int pcm_dtype /*I16|F32|F64*/, int64_t n_samples, double alpha,
                      double* out, void* stream);
/* frames of one utterance: ceil(|S-(L-step)|/step), utils.py:118 (host helper, no device work) */
int64_t lnasr_mfcc_num_frames(int64_t n_samples, int32_t frame_len, int32_t frame_step);
/* the mel-product plan of a filter bank (see lnasr_mfcc_tables.mel_plan): reads n_mel, fb_stride, fb_start, fb_len, fb_weight.
 * One small launch on `stream`; like every buffer the library fills, the plan is ordered only with respect to that stream --
 * synchronise (or record / wait an event) before lnasr_mfcc_batch uses it from another stream */
size_t lnasr_mfcc_mel_plan_bytes(void);
int lnasr_mfcc_mel_plan(const lnasr_mfcc_tables* tables, int32_t* plan /*device out*/, void* stream);
/* workspace bytes needed by lnasr_mfcc_batch */
int lnasr_mfcc_workspace_bytes(int64_t total_frames, int32_t n_utt, size_t* bytes /*host out*/);
/*
 * pcm            : concatenated samples, int16 (LNASR_DT_I16), float32 (LNASR_DT_F32) or float64 (LNASR_DT_F64)
 * sample_offsets : [n_utt+1] int64, device
 * frame_offsets  : [n_utt+1] int64, device; frame_offsets[u+1]-frame_offsets[u] must equal
 *                  lnasr_mfcc_num_frames(samples of u) and be >= 2 (the reference needs row 1, mfcc.py:164)
 * total_frames   : < 2^31 per call (frame positions are 32-bit inside the kernels; sample offsets are 64-bit)
 * feats          : [total_frames, ld_feats] float32 out, 3*(n_ceps+1) = 39 columns used
 * power_out      : [total_frames, fft_n/2+1] float32 or NULL   (1st return value of calc_mfcc)
 * ceps_out       : [total_frames, n_mel] float32 or NULL       (2nd return value, incl. its in-place quirk)
 */
int lnasr_mfcc_batch(const void* pcm, int pcm_dtype, const int64_t* sample_offsets,
                     const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                     const lnasr_mfcc_tables* tables /*host struct of device pointers*/,
                     float* feats, int32_t ld_feats, float* power_out, float* ceps_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ================================ (2) GMM emission log-likelihoods ==============================
 * Replaces GMMHMM._map_b (lnasr/gmmhmm.py:46-67) and gmm.gaussian_multi_distribution_log
 * (lnasr/gmm.py:59-68) for diagonal covariances: bm[t,j,m] = -1/2 (D ln 2pi + sum ln var + sum (x-mu)^2/var),
 * b[t,j] = lse_m (w[j,m] + bm[t,j,m]).
 *
 * Two kernels behind one packed-model idea:
 *   exact : float64 SIMT evaluation of (x-mu)^2/var -- small models (configs 1-3), optional bm output;
 *   tc    : [x^2, x, 1] . [-1/(2 var), mu/var, c]^T as a 3xTF32 tcgen05 contraction with the log-sum-exp
 *           over mixtures fused into the TMEM epilogue -- large models (config 4).
 */
/* packed exact model: lnasr_gmm_exact_packed_doubles() float64 values -- the SIMT image (chunks of 4 mixtures of one
 * state, n_mix padded with log w = -inf, (dim+1) rows of [a_0..a_3 | b_0..b_3] per chunk) followed, when n_mix <= 8 and
 * dim <= 39, by the FP64 tensor-core image ([groups*40][84] rows of [-1/(2 var) , mu/var interleaved | c | log w]);
 * lnasr_gmm_loglik_exact runs the DMMA kernel on the latter where it exists (see csrc/gmm_exact.cu) */
int64_t lnasr_gmm_exact_packed_doubles(int32_t n_states, int32_t n_mix, int32_t dim);
int lnasr_gmm_pack_exact(const double* logw /*[n_states,n_mix]*/, const double* mu /*[n_states,n_mix,dim]*/,
                         const double* var /*[n_states,n_mix,dim]*/, int32_t n_states, int32_t n_mix, int32_t dim,
                         double* packed, void* stream);
int lnasr_gmm_loglik_exact(const void* feats, int feats_dtype /*F32|F64*/, int32_t ld_feats, int64_t n_frames,
                           const double* packed, int32_t n_states, int32_t n_mix, int32_t dim,
                           void* b_out /*[n_frames, ld_b]*/, int b_dtype /*F32|F64*/, int64_t ld_b,
                           double* bm_out /*[n_frames, n_states*n_mix] or NULL*/, void* stream);

/* full covariance (the reference's Sigma is a full D x D matrix, lnasr/gmm.py:59-68): the caller factorises
 * Sigma = L L^T once and passes linv = L^-1 [n_states*n_mix, dim, dim] (lower triangular, row-major),
 * nu = L^-1 mu [.., dim], cst = -1/2 (D ln 2pi + 2 sum ln L_ii) [..], logw [..]. */
int lnasr_gmm_loglik_full(const void* feats, int feats_dtype, int32_t ld_feats, int64_t n_frames, const double* linv,
                          const double* nu, const double* cst, const double* logw, int32_t n_states, int32_t n_mix,
                          int32_t dim, void* b_out, int b_dtype, int64_t ld_b, double* bm_out, void* stream);

/* tensor-core path (csrc/gmm_tc.cu).  n_mix <= 16 (padded to a power of two inside the packed image), dim <= 39.
 * The packed image holds, per tile of states, the hi and lo TF32 halves of [-1/(2 var) | mu/var | c] * log2(e) in
 * the UMMA K-major shared-memory layout, plus the optional feature centre `shift` (host semantics: the kernel
 * evaluates the same density on x - shift and mu - shift, which halves the fp32 cancellation error). */
int lnasr_gmm_tc_packed_bytes(int32_t n_states, int32_t n_mix, int32_t dim, size_t* bytes /*host out*/);
int lnasr_gmm_pack_tc(const double* logw, const double* mu, const double* var, int32_t n_states, int32_t n_mix,
                      int32_t dim, const double* shift /*[dim] device, or NULL*/, void* packed, void* stream);
/* feats [n_frames, ld_feats] float32; b_out [n_frames, ld_b] float32; workspace: lnasr_gmm_tc_workspace_bytes(n_frames)
 * bytes for the tile-transposed, centred copy of the features the main kernel streams */
int lnasr_gmm_tc_workspace_bytes(int64_t n_frames, size_t* bytes /*host out*/);
int lnasr_gmm_loglik_tc(const float* feats, int32_t ld_feats, int64_t n_frames, const void* packed,
                        int32_t n_states, int32_t n_mix, int32_t dim, float* b_out, int64_t ld_b,
                        void* workspace, size_t workspace_bytes, void* stream);

/* FP16-slice variant of the tensor-core path (csrc/gmm_tc16.cu): the same contraction with every operand split into two FP16
 * slices (22 significant bits, like TF32 hi + lo) and tcgen05 kind::f16 -- K = 16 per instruction, half the MMA count.
 * FP16 has no exponent range to spare, so `scale` [dim] (device, powers of two; NULL = 1) brings the centred features to
 * |x'| <~ 32 and every Gaussian's weights carry their own power-of-two scale (undone in the epilogue).  Same workspace as
 * lnasr_gmm_loglik_tc. */
int lnasr_gmm_tc16_packed_bytes(int32_t n_states, int32_t n_mix, int32_t dim, size_t* bytes /*host out*/);
int lnasr_gmm_pack_tc16(const double* logw, const double* mu, const double* var, int32_t n_states, int32_t n_mix,
                        int32_t dim, const double* shift /*[dim] or NULL*/, const double* scale /*[dim] or NULL*/,
                        void* packed, void* stream);
int lnasr_gmm_loglik_tc16(const float* feats, int32_t ld_feats, int64_t n_frames, const void* packed,
                          int32_t n_states, int32_t n_mix, int32_t dim, float* b_out, int64_t ld_b,
                          void* workspace, size_t workspace_bytes, void* stream);

/* discrete emissions: HMM._map_b (lnasr/hmm.py:47-61): b[t, j] = B[j, obs[t]] */
int lnasr_emit_gather(const double* logB /*[n_states, n_symbols]*/, const int32_t* obs /*[n_frames]*/,
                      int64_t n_frames, int32_t n_states, int32_t n_symbols, double* b_out /*[n_frames,n_states]*/,
                      void* stream);

/* ================================ (3) HMM log-trellis ===========================================
 * Replaces HMM._forward (lnasr/hmm.py:63-100), HMM._backward (hmm.py:102-131), HMM._viterbi incl. the
 * backtrace (hmm.py:133-172), HMM.calc_prob (hmm.py:268-274) and HMM.decode (hmm.py:276-283) for a batch of
 * utterances x a bank of n_models models that share (N) but not parameters.
 *
 * Transitions: dense log A [n_models,N,N] always; for N > 32 the kernels walk predecessor / successor
 * lists of the FINITE entries (ELL layout, built once per model by lnasr_trans_pack).
 * All trellis state is float64; emissions b may be float32 (tc path) or float64 (exact path).
 */
/* structure of the finite transitions of a bank -> device int32[4] {max_in, max_out, max(j - i), max(i - j)}: the largest
 * number of finite entries in any column / row of any model, and the longest forward / backward jump of a finite
 * transition i -> j (max(i - j) = 0 and max(j - i) <= 2: left-to-right band, lnasr_hmm_bank.band = 1 + max(j - i)) */
int lnasr_trans_degree(const double* logA, int32_t n_models, int32_t N, int32_t* deg_out /*device [4]*/, void* stream);
/* pred_idx/val: [n_models, max_in, N]  (k-th finite predecessor i of state j, ascending i; pad idx=0,val=-inf)
 * succ_idx/val: [n_models, max_out, N] (k-th finite successor j of state i) */
int lnasr_trans_pack(const double* logA, int32_t n_models, int32_t N, int32_t max_in, int32_t max_out,
                     int32_t* pred_idx, double* pred_val, int32_t* succ_idx, double* succ_val, void* stream);

typedef struct lnasr_hmm_bank {
    int32_t n_models, N;
    const double* logA;      /* [n_models, N, N] */
    const double* logpi;     /* [n_models, N] */
    int32_t max_in, max_out; /* 0 when the ELL arrays are absent (N <= 32) */
    const int32_t* pred_idx; const double* pred_val;
    const int32_t* succ_idx; const double* succ_val;
    int32_t band;            /* 0: general; else 1 + the largest j - i over the finite transitions i -> j of a bank in which all
                              * of them satisfy 0 <= j - i <= 2 (left-to-right: self, next, skip) -- lnasr_trans_degree reports
                              * it; selects the register-resident banded Viterbi kernel */
} lnasr_hmm_bank;

size_t lnasr_viterbi_workspace_bytes(int64_t total_frames, int32_t n_utt, int32_t n_models, int32_t N);
/*
 * b          : [total_frames, ld_b] emissions (b_dtype F32|F64), columns k*N+j
 * path_out   : [n_models, total_frames] uint32 (hmm.py:168 dtype) or NULL
 * score_out  : [n_utt, n_models] float64 = max_j v[T-1, j]          (NULL allowed)
 * v_out      : [total_frames, n_models*N] float64 or NULL ; bt_out same shape uint32 or NULL (hmm.py:157-158)
 */
int lnasr_viterbi_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                        int64_t total_frames, const lnasr_hmm_bank* bank /*host struct*/,
                        uint32_t* path_out, double* score_out, double* v_out, uint32_t* bt_out,
                        void* workspace, size_t workspace_bytes, void* stream);
/* alpha_out [total_frames, n_models*N] or NULL ; logprob_out [n_utt, n_models] = lse_j alpha[T-1,j] */
int lnasr_forward_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                        int64_t total_frames, const lnasr_hmm_bank* bank, double* alpha_out, double* logprob_out,
                        void* stream);
int lnasr_backward_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                         int64_t total_frames, const lnasr_hmm_bank* bank, double* beta_out, void* stream);

/* isolated-word decision: arg_out[u] = argmax_k score[u,k], lowest index on ties; best_out optional */
int lnasr_argmax_rows(const double* score /*[n_rows,n_cols]*/, int32_t n_rows, int32_t n_cols, int32_t* arg_out,
                      double* best_out, void* stream);

/* ================================ (3b) the decode path in one call ==============================
 * PCM -> MFCC (mfcc.py:108-175) -> emissions (gmmhmm.py:46-67) -> Viterbi (hmm.py:133-172) -> word = argmax over the bank's
 * models of max_j v[T-1,j], on one workspace and one stream: no allocation, no synchronisation.
 * gmm_kind 0: gmm_packed is the lnasr_gmm_pack_exact image (float64 emissions); 1: the lnasr_gmm_pack_tc image (float32);
 * 2: the lnasr_gmm_pack_tc16 image (float32).
 * n_mix / dim describe the Gaussians; the bank's n_models*N states are the GMM's states, model k owning [k*N,(k+1)*N). */
typedef struct lnasr_recognizer {
    const lnasr_mfcc_tables* mfcc;   /* host struct of device pointers */
    const lnasr_hmm_bank* bank;      /* host struct of device pointers */
    const void* gmm_packed;          /* device */
    int32_t gmm_kind, n_mix, dim;
} lnasr_recognizer;
int lnasr_recognize_workspace_bytes(int64_t total_frames, int32_t n_utt, const lnasr_recognizer* rec, int want_paths,
                                    size_t* bytes /*host out*/);
/* words_out [n_utt] int32, best_out [n_utt] float64 or NULL, paths_out [n_models, total_frames] uint32 or NULL (all device) */
int lnasr_recognize_batch(const void* pcm, int pcm_dtype, const int64_t* sample_offsets, const int64_t* frame_offsets,
                          int32_t n_utt, int64_t total_frames, const lnasr_recognizer* rec, int32_t* words_out,
                          double* best_out, uint32_t* paths_out, void* workspace, size_t workspace_bytes, void* stream);

/* ================================ (4) Baum-Welch accumulation + M-step ==========================
 * Replaces HMM._baumwelch (lnasr/hmm.py:174-200) and GMMHMM._estimate (lnasr/gmmhmm.py:69-156) in
 * sufficient-statistic form, never materialising xi [T,N,N] or the
 * [N,M,T,D,D] scatter.  The backward recursion and the accumulation are ONE kernel.
 *
 * Accumulator layout (float64, lnasr_em_acc_doubles() values), summed over utterances and -- across
 * GPUs -- by one all-reduce:
 *   S0[N*M] | S1[N*M*D] | S2c[N*M*D] (scatter about the OLD mean) | xi_sum[N*N] | pi_occ[N] | loglik | n_utt
 */
int64_t lnasr_em_acc_doubles(int32_t N, int32_t n_mix, int32_t dim);
/* alpha must come from lnasr_forward_batch, b/bm (float64) from lnasr_gmm_loglik_exact; single model (n_models=1).
 * workspace = lnasr_fb_workspace_bytes() bytes selects the two-kernel form (backward recursion with G lanes per utterance
 * writing log gamma to the workspace, then a frame-parallel statistics kernel) that fills the machine when there are
 * fewer utterances than a few per SM x warps; workspace = NULL runs the single fused kernel (one CTA per utterance). */
size_t lnasr_fb_workspace_bytes(int64_t total_frames, int32_t N);
int lnasr_fb_accumulate(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                        const double* alpha, const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                        const double* logA, const double* logw, const double* mu, int32_t N, int32_t n_mix,
                        int32_t dim, double* acc /*in/out, caller zeroes*/, double* gamma_out /*[T,N] log or NULL*/,
                        double* beta_out /*[T,N] log or NULL*/, void* workspace /*or NULL*/, size_t workspace_bytes,
                        void* stream);
/* The same statistics from alpha AND beta (lnasr_forward_batch, lnasr_backward_batch: two independent trellis passes that
 * can run concurrently on two streams): everything that is not a recursion -- Z_t, gamma, xi / pi sums, mixture occupancies,
 * S0 / S1 / S2c -- in one frame-parallel kernel.  Same accumulator layout; gamma_out [T,N] optional (row T-1 = -inf). */
int lnasr_em_statistics(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                        const double* alpha, const double* beta, const int64_t* frame_offsets, int32_t n_utt,
                        int64_t total_frames, const double* logA, const double* logw, const double* mu, int32_t N,
                        int32_t n_mix, int32_t dim, double* acc /*in/out*/, double* gamma_out /*or NULL*/, void* stream);
/* The same statistics for ANY number of states (GMMHMM._estimate gmmhmm.py:69-156 / HMM._baumwelch hmm.py:174-200 at the model
 * sizes the trellis kernels decode): N <= 32 forwards to lnasr_em_statistics; larger models run frame-parallel kernels over a
 * compact list of the finite transitions with the accumulators in global memory.  alpha / beta come from lnasr_forward_batch /
 * lnasr_backward_batch.  workspace: lnasr_em_statistics_workspace_bytes(total_frames, N). */
size_t lnasr_em_statistics_workspace_bytes(int64_t total_frames, int32_t N);
int lnasr_em_statistics_ws(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                           const double* alpha, const double* beta, const int64_t* frame_offsets, int32_t n_utt,
                           int64_t total_frames, const double* logA, const double* logw, const double* mu, int32_t N,
                           int32_t n_mix, int32_t dim, double* acc, double* gamma_out /*[T,N] log or NULL*/, void* workspace,
                           size_t workspace_bytes, void* stream);
/* discrete emissions, any N (HMM._estimate hmm.py:202-225): acc layout of lnasr_hmm_fb_accumulate */
int lnasr_hmm_em_statistics_ws(const double* b, const double* alpha, const double* beta, const int32_t* obs,
                               const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames, const double* logA,
                               int32_t N, int32_t n_symbols, double* acc, double* gamma_out, void* workspace,
                               size_t workspace_bytes, void* stream);

/* M-step from (all-reduced) statistics; writes the new model in place of the old device arrays */
int lnasr_mstep_finalize(const double* acc, int32_t N, int32_t n_mix, int32_t dim, double min_std,
                         double* logA, double* logpi, double* logw, double* mu, double* var, void* stream);

/* discrete-emission statistics for plain HMM._estimate (lnasr/hmm.py:202-225):
 * acc layout: xi_sum[N*N] | gamma_sum[N] (t<T-1) | B_sum[N*n_symbols] | pi_occ[N] | loglik | n_utt */
int64_t lnasr_hmm_acc_doubles(int32_t N, int32_t n_symbols);
int lnasr_hmm_fb_accumulate(const double* b, const double* alpha, const int32_t* obs, const int64_t* frame_offsets,
                            int32_t n_utt, int64_t total_frames, const double* logA, int32_t N, int32_t n_symbols,
                            double* acc, double* gamma_out, double* beta_out /*[T,N] log or NULL*/, void* stream);
int lnasr_hmm_mstep_finalize(const double* acc, int32_t N, int32_t n_symbols, double* logA, double* logB,
                             double* logpi, void* stream);

/* HMM._baumwelch (lnasr/hmm.py:174-200) with its return values materialised: xi [T,N,N] and gamma [T,N] (log),
 * row T-1 of every utterance = -inf.  API parity only; train() uses the fused accumulation above. */
int lnasr_baumwelch_xi(const double* alpha, const double* beta, const double* b, const double* logA,
                       const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames, int32_t N,
                       double* xi_out, double* gamma_out, void* stream);

/* ================================ before the path: voice activity detection =====================
 * VadLtsd.detect (third/vadlstd/VadLstd.py:44-74): frames of `winsize` samples every `stepsize` after one step of leading zeros
 * (utils.split_frames rule), Hamming window, |rfft|, long-term spectral envelope over 2*order frames, LTSD in dB per frame.
 * window [winsize] and twiddle [winsize/2] (cos, -sin of 2 pi j / winsize) are caller-provided float64 device tables;
 * adapt != 0 enables the noise update of VadLstd.py:72-73 with factor alpha.  ltsd_out [lnasr_vad_num_frames] float64. */
int64_t lnasr_vad_num_frames(int64_t n_samples, int32_t winsize, int32_t stepsize);
size_t lnasr_vad_workspace_bytes(int64_t n_samples, int32_t winsize, int32_t stepsize);
int lnasr_vad_ltsd(const void* pcm, int pcm_dtype, int64_t n_samples, int32_t winsize, int32_t stepsize, int32_t order,
                   double threshold, int adapt, double alpha, const double* window, const double* twiddle,
                   double* ltsd_out, void* workspace, size_t workspace_bytes, void* stream);

/* kernels this library has launched so far in this process (all threads, all streams) */
int64_t lnasr_launch_count(void);

/* ================================ debug / bring-up ==============================================
 * One tcgen05.mma (M=128, N=128, K=8, kind::tf32) on caller-provided shared-memory images with caller-chosen
 * descriptor fields; used by tests/test_umma_probe.py to pin the descriptor semantics on hardware. */
int lnasr_debug_umma_probe(const float* a_image /*[a_bytes/4]*/, int32_t a_bytes, const float* b_image, int32_t b_bytes,
                           uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t layout_type,
                           uint32_t idesc, float* d_out /*[128,128]*/, void* stream);
/* measured FP64 peak of the current device in TFLOP/s: kind 0 = DFMA, 1 = mma.sync m8n8k4 f64 (the denominators of the
 * float64 kernels' rooflines; bench.py measures them in the same run).  tflops_out is a HOST pointer; synchronises. */
int64_t lnasr_debug_fp64_peak_scratch_doubles(void);
int lnasr_debug_fp64_peak(int kind, double* scratch, double* tflops_out /*host*/, void* stream);
/* bring-up / test switch for lnasr_gmm_loglik_exact: 0 = default choice (persistent DMMA kernel when the model image fits in
 * shared memory, else streaming DMMA kernel, else SIMT), 1 = always SIMT, 2 = streaming DMMA instead of persistent */
int lnasr_debug_set_gmm_simt(int on);
/* bring-up switch: exchange the LBO / SBO fields of the descriptors lnasr_gmm_loglik_tc builds (0 = as documented) */
int lnasr_debug_set_umma_swap(int swap);
/* bring-up / test switch for the banded Viterbi decode: 2 = trellis_band2_kernel (barrier-free, 2-bit backpointers, fused
 * backtrace; default), 1 = trellis_band_kernel + backtrace_kernel */
int lnasr_debug_set_band_kernel(int version);
/* lnasr_gmm_loglik_tc16: share of the mixture exponentials computed by polynomial on the FMA pipe instead of the MUFU
 * (0 = none (default), 1 = 2 of 8 pairs, 2 = 3 of 8, 3 = 4 of 8; environment LNASR_TC16_POLY) */
int lnasr_debug_set_tc16_poly(int level);
/* lnasr_gmm_loglik_tc16: 1 = all sixteen epilogue warps on every tile, 2 = two groups of eight on alternate tiles (default;
 * environment LNASR_TC16_GROUPS) */
int lnasr_debug_set_tc16_groups(int groups);
/* lnasr_debug_umma_probe issues a second, accumulating MMA on the K chunks `byte_offset` further into both images
 * (0 = single MMA): measures how the fp32 TMEM accumulator rounds (tools/probe_umma_rounding.py) */
int lnasr_debug_set_probe_second_mma(uint32_t byte_offset);
/* lnasr_em_statistics_ws (N > 32): 1 = the SIMT moments kernel also where the FP64 tensor-core one would run (dim <= 39) */
int lnasr_debug_set_em_moments_simt(int on);

#ifdef __cplusplus
}
#endif
#endif /* LNASR_B200_H */
