// em.cu -- Baum-Welch: backward recursion fused with occupancy / sufficient-statistic accumulation, and the
// M-step finalise (kernel (4) of the north star).
//
// Replaces, from /root/reference/lnasr: HMM._backward (hmm.py:102-131), HMM._baumwelch (hmm.py:174-200),
// GMMHMM._estimate (gmmhmm.py:69-156) and HMM._estimate (hmm.py:202-225), in sufficient-statistic form
// (SURVEY.md section 8(a) rows E1-E3): the reference materialises xi [T,N,N] and an [N,M,T,D,D] scatter; here
// nothing larger than one trellis column per utterance ever exists.
//
// fb_accum_kernel<G, GMM>: one CTA walks whole utterances backwards in chunks of EM_TB frames.
//   stage A (all threads) : coalesced copy of the chunk's alpha, b, bm rows and feature rows into smem;
//   stage B (warp 0)      : lane = state.  beta[t] from beta[t+1] (hmm.py:126-130); Z_t = lse_i(alpha+beta);
//                           xi[t,i,j] = alpha[t,i] + A[i,j] + b[t+1,j] + beta[t+1,j] - Z_t (hmm.py:190-197),
//                           exp-accumulated into registers (G per lane); mixture occupancy
//                           occ[t,j,m] = exp(alpha+beta-Z_t + w[j,m] + bm[t,j,m] - b[t,j]) (gmmhmm.py:89-106) -> smem;
//   stage C (all threads) : S0 += occ, S1 += occ*x, S2c += occ*(x-mu_old)^2 (gmmhmm.py:117-147), every thread
//                           owning fixed (gaussian, dim) entries of the smem accumulators.
//   The CTA's accumulators are flushed with float64 atomics once, after its last utterance.
// Quirks kept: gamma[T-1] = -inf for the transition / pi / discrete-B statistics (hmm.py:199),
// but the mixture statistics use alpha+beta for ALL t including T-1 (gmmhmm.py:98-100); the scatter is about the
// OLD mean; min_std is added to the variance (gmmhmm.py:147).
#include <algorithm>
#include "common.cuh"

namespace lnasr {

constexpr int EM_THREADS = 128;
constexpr int EM_TB = 16;

struct FbArgs {
    // emissions / trellis inputs
    const void* feats; int feats_f64; int ld_feats;
    const double* b; const double* bm; const double* alpha;
    const int32_t* obs;                 // discrete mode
    const int64_t* frame_off; int n_utt;
    const double* logA; const double* logw; const double* mu;
    int N, M, D, n_symbols;
    double* acc; double* gamma_out; double* beta_out;
};

// group-wide (G lanes, G a power of two, groups aligned) log-sum-exp
template <int G>
__device__ __forceinline__ double group_lse(double v) {
    double mx = v;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o, G));
    double s = (v == neg_inf()) ? 0.0 : exp(v - mx);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, G);
    return (mx == neg_inf()) ? mx : mx + log(s);
}

template <int G, bool GMM>
__global__ void __launch_bounds__(EM_THREADS)
fb_accum_kernel(FbArgs a) {
    extern __shared__ __align__(16) unsigned char em_smem_raw[];
    double* sm = reinterpret_cast<double*>(em_smem_raw);
    const int N = a.N, M = a.M, D = a.D, NM = N * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // carve
    double* acc_s = sm;                                   // GMM: S0[NM] | S1[NM*D] | S2c[NM*D]
    double* mu_s = acc_s + (GMM ? NM * (2 * D + 1) : 0);
    double* logw_s = mu_s + (GMM ? NM * D : 0);
    double* alpha_s = logw_s + (GMM ? NM : 0);            // [TB][N]
    double* b_s = alpha_s + EM_TB * N;                    // [TB][N]
    double* bm_s = b_s + EM_TB * N;                       // [TB][NM]
    double* occ_s = bm_s + (GMM ? EM_TB * NM : 0);        // [TB][NM]
    double* x_s = occ_s + (GMM ? EM_TB * NM : 0);         // [TB][D]

    if (GMM) {
        for (int i = tid; i < NM * (2 * D + 1); i += EM_THREADS) acc_s[i] = 0.0;
        for (int i = tid; i < NM * D; i += EM_THREADS) mu_s[i] = a.mu[i];
        for (int i = tid; i < NM; i += EM_THREADS) logw_s[i] = a.logw[i];
    }
    // warp 0, lane i < N: row i of log A in registers, xi accumulators
    const bool rec_lane = (warp == 0) && (lane < G);
    const bool act = rec_lane && lane < N;
    double tr[G], xi_acc[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
        tr[q] = (act && q < N) ? __ldg(a.logA + (int64_t)lane * N + q) : neg_inf();
        xi_acc[q] = 0.0;
    }
    double pi_acc = 0.0, gsum_acc = 0.0, loglik_acc = 0.0;
    int utt_cnt = 0;
    __syncthreads();

    for (int u = blockIdx.x; u < a.n_utt; u += gridDim.x) {
        const int64_t f0 = __ldg(a.frame_off + u);
        const int T = (int)(__ldg(a.frame_off + u + 1) - f0);
        if (T <= 0) continue;
        double cur = 0.0;          // beta[t+1, lane]
        double e_next = 0.0;       // b[t+1, lane]
        double L = 0.0;
        if (warp == 0) {
            // log P(O) = lse_j alpha[T-1, j]    (hmm.py:274)
            double v = act ? __ldg(a.alpha + (f0 + T - 1) * N + lane) : neg_inf();
            if (G < 32 && lane >= G) v = neg_inf();
            L = group_lse<G>(v);
            if (lane == 0) { loglik_acc += L; ++utt_cnt; }
        }
        for (int t_hi = T; t_hi > 0; t_hi -= EM_TB) {
            const int t_lo = max(0, t_hi - EM_TB);
            const int nt = t_hi - t_lo;
            // ---- stage A: stage the chunk rows
            for (int i = tid; i < nt * N; i += EM_THREADS) {
                alpha_s[i] = __ldg(a.alpha + (f0 + t_lo) * N + i);
                b_s[i] = __ldg(a.b + (f0 + t_lo) * N + i);
            }
            if (GMM) {
                for (int i = tid; i < nt * NM; i += EM_THREADS) bm_s[i] = __ldg(a.bm + (f0 + t_lo) * NM + i);
                for (int i = tid; i < nt * D; i += EM_THREADS) {
                    const int r = i / D, d = i - r * D;
                    const int64_t row = f0 + t_lo + r;
                    x_s[i] = a.feats_f64 ? __ldg((const double*)a.feats + row * a.ld_feats + d)
                                         : (double)__ldg((const float*)a.feats + row * a.ld_feats + d);
                }
            }
            __syncthreads();
            // ---- stage B: recursion (warp 0)
            if (warp == 0) {
                for (int t = t_hi - 1; t >= t_lo; --t) {
                    const int tt = t - t_lo;
                    const bool last = (t == T - 1);
                    const double uj = cur + e_next;                 // beta[t+1,j] + b[t+1,j]  (unused when last)
                    double cand[G];
                    double mx = neg_inf();
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        cand[q] = __shfl_sync(0xffffffffu, uj, q, G) + tr[q];
                        mx = fmax(mx, cand[q]);
                    }
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < G; ++q) s += (cand[q] == neg_inf()) ? 0.0 : exp(cand[q] - mx);
                    double beta = (mx == neg_inf()) ? mx : mx + log(s);
                    if (last) beta = 0.0;                                           // hmm.py:125
                    const double al = act ? alpha_s[tt * N + lane] : neg_inf();
                    const double ab = act ? al + beta : neg_inf();
                    double Z = group_lse<G>((G < 32 && lane >= G) ? neg_inf() : ab);   // == lse_{i,j} of the xi numerator
                    const double g2 = (ab == neg_inf() || Z == neg_inf()) ? neg_inf() : ab - Z;
                    if (act) {
                        if (a.beta_out) a.beta_out[(f0 + t) * N + lane] = beta;
                        if (a.gamma_out) a.gamma_out[(f0 + t) * N + lane] = last ? neg_inf() : g2;   // hmm.py:199
                        if (!last && Z != neg_inf()) {
                            double rs = 0.0;
#pragma unroll
                            for (int q = 0; q < G; ++q) {
                                // xi[t,i,q] = alpha[t,i] + A[i,q] + b[t+1,q] + beta[t+1,q] - Z_t
                                const double x = al + cand[q] - Z;
                                const double ex = (q < N && x != neg_inf()) ? exp(x) : 0.0;
                                xi_acc[q] += ex;
                                rs += ex;
                            }
                            gsum_acc += rs;                       // sum_t exp(gamma[t,i]), t < T-1
                            if (t == 0) pi_acc += rs;             // pi' = gamma[0]   (gmmhmm.py:86 / hmm.py:216)
                            if (!GMM) {
                                const int o = __ldg(a.obs + f0 + t);
                                if (o >= 0 && o < a.n_symbols)    // B numerator: gamma over t with obs[t]==k (hmm.py:211-214)
                                    atomicAdd(a.acc + (int64_t)N * N + N + (int64_t)lane * a.n_symbols + o, rs);
                            }
                        }
                    }
                    if (GMM) {
                        // mixture occupancies for ALL t (gmmhmm.py:98-100)
                        for (int p0 = 0; p0 < NM; p0 += 32) {
                            const int p = p0 + lane;
                            const int j = min(p / M, N - 1);
                            const double gfull = __shfl_sync(0xffffffffu, g2, j);   // lane j < N of warp 0 holds state j
                            if (p < NM) {
                                const double bj = b_s[tt * N + j];
                                double o = 0.0;
                                if (gfull != neg_inf() && bj != neg_inf()) o = exp(gfull + logw_s[p] + bm_s[tt * NM + p] - bj);
                                occ_s[tt * NM + p] = o;
                            }
                        }
                    }
                    cur = beta;
                    e_next = act ? b_s[tt * N + lane] : neg_inf();
                }
            }
            __syncthreads();
            // ---- stage C: sufficient statistics
            if (GMM) {
                for (int e = tid; e < NM * D; e += EM_THREADS) {
                    const int jm = e / D, d = e - jm * D;
                    double s1 = 0.0, s2 = 0.0;
                    const double m0 = mu_s[e];
                    for (int tt = 0; tt < nt; ++tt) {
                        const double o = occ_s[tt * NM + jm];
                        const double x = x_s[tt * D + d];
                        const double df = x - m0;
                        s1 = fma(o, x, s1);
                        s2 = fma(o, df * df, s2);
                    }
                    acc_s[NM + e] += s1;
                    acc_s[NM + NM * D + e] += s2;
                }
                for (int jm = tid; jm < NM; jm += EM_THREADS) {
                    double s0 = 0.0;
                    for (int tt = 0; tt < nt; ++tt) s0 += occ_s[tt * NM + jm];
                    acc_s[jm] += s0;
                }
            }
            __syncthreads();
        }
    }
    // ---- flush
    if (GMM) {
        for (int i = tid; i < NM * (2 * D + 1); i += EM_THREADS) {
            const double v = acc_s[i];
            if (v != 0.0) atomicAdd(a.acc + i, v);
        }
        double* tail = a.acc + (int64_t)NM * (2 * D + 1);     // xi_sum[N*N] | pi_occ[N] | loglik | n_utt
        if (act) {
#pragma unroll
            for (int q = 0; q < G; ++q) if (q < N) atomicAdd(tail + (int64_t)lane * N + q, xi_acc[q]);
            atomicAdd(tail + (int64_t)N * N + lane, pi_acc);
        }
        if (tid == 0) {
            atomicAdd(tail + (int64_t)N * N + N, loglik_acc);
            atomicAdd(tail + (int64_t)N * N + N + 1, (double)utt_cnt);
        }
    } else {
        // xi_sum[N*N] | gamma_sum[N] | B_sum[N*n_symbols] | pi_occ[N] | loglik | n_utt
        if (act) {
#pragma unroll
            for (int q = 0; q < G; ++q) if (q < N) atomicAdd(a.acc + (int64_t)lane * N + q, xi_acc[q]);
            atomicAdd(a.acc + (int64_t)N * N + lane, gsum_acc);
            atomicAdd(a.acc + (int64_t)N * N + N + (int64_t)N * a.n_symbols + lane, pi_acc);
        }
        if (tid == 0) {
            double* tail = a.acc + (int64_t)N * N + N + (int64_t)N * a.n_symbols + N;
            atomicAdd(tail, loglik_acc);
            atomicAdd(tail + 1, (double)utt_cnt);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// M-step from summed statistics (oracle: em_finalize_diag; reference gmmhmm.py:83-86,103-147 in linear domain)
__global__ void mstep_finalize_kernel(const double* __restrict__ acc, int N, int M, int D, double min_std,
                                      double* __restrict__ logA, double* __restrict__ logpi, double* __restrict__ logw,
                                      double* __restrict__ mu, double* __restrict__ var) {
    const int NM = N * M;
    const double* S0 = acc;
    const double* S1 = acc + NM;
    const double* S2 = S1 + (int64_t)NM * D;
    const double* xi = S2 + (int64_t)NM * D;
    const double* pio = xi + (int64_t)N * N;
    const double n_utt = pio[N + 1];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int i = tid; i < N; i += nthr) {
        double rs = 0.0;
        for (int j = 0; j < N; ++j) rs += xi[(int64_t)i * N + j];
        const double lrs = log(rs);
        for (int j = 0; j < N; ++j) logA[(int64_t)i * N + j] = log(xi[(int64_t)i * N + j]) - lrs;
        logpi[i] = log(pio[i] / n_utt);
        double ws = 0.0;
        for (int m = 0; m < M; ++m) ws += S0[i * M + m];
        const double lws = log(ws);
        for (int m = 0; m < M; ++m) logw[i * M + m] = log(S0[i * M + m]) - lws;
    }
    for (int64_t e = tid; e < (int64_t)NM * D; e += nthr) {
        const double s0 = S0[e / D];
        mu[e] = S1[e] / s0;
        var[e] = S2[e] / s0 + min_std;      // gmmhmm.py:147: min_std is added to the variance itself
    }
}

__global__ void hmm_mstep_finalize_kernel(const double* __restrict__ acc, int N, int S, double* __restrict__ logA,
                                          double* __restrict__ logB, double* __restrict__ logpi) {
    const double* xi = acc;
    const double* gs = acc + (int64_t)N * N;
    const double* Bs = gs + N;
    const double* pio = Bs + (int64_t)N * S;
    const double n_utt = pio[N + 1];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int64_t e = tid; e < (int64_t)N * N; e += nthr) logA[e] = log(xi[e]) - log(gs[e / N]);     // hmm.py:208-209
    for (int64_t e = tid; e < (int64_t)N * S; e += nthr) logB[e] = log(Bs[e]) - log(gs[e / S]);     // hmm.py:211-214
    for (int i = tid; i < N; i += nthr) logpi[i] = log(pio[i] / n_utt);                             // hmm.py:216
}

// ------------------------------------------------------------------------------------------------
// Materialising Baum-Welch for API parity with HMM._baumwelch (hmm.py:174-200): xi [T,N,N], gamma [T,N].
// One CTA per frame; not on the fast path (the reference API returns these arrays, train() does not use them).
constexpr int BW_THREADS = 128;
__global__ void __launch_bounds__(BW_THREADS)
baumwelch_xi_kernel(const double* __restrict__ alpha, const double* __restrict__ beta, const double* __restrict__ b,
                    const double* __restrict__ logA, const int64_t* __restrict__ frame_off, int n_utt, int N,
                    double* __restrict__ xi, double* __restrict__ gamma) {
    __shared__ double red[BW_THREADS / 32];
    __shared__ double bc;
    const int64_t f = blockIdx.x;
    const int u = find_utt(frame_off, n_utt, f);
    const int64_t f1 = __ldg(frame_off + u + 1);
    const bool last = (f == f1 - 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t NN = (int64_t)N * N;
    double* xo = xi + f * NN;
    if (last) {                                                     // hmm.py:191,199: row T-1 stays -inf
        for (int64_t e = tid; e < NN; e += BW_THREADS) xo[e] = neg_inf();
        for (int i = tid; i < N; i += BW_THREADS) gamma[f * N + i] = neg_inf();
        return;
    }
    auto numer = [&](int64_t e) {
        const int i = (int)(e / N), j = (int)(e - (int64_t)i * N);
        return ((alpha[f * N + i] + logA[e]) + b[(f + 1) * N + j]) + beta[(f + 1) * N + j];     // hmm.py:193-196 order
    };
    double mx = neg_inf();
    for (int64_t e = tid; e < NN; e += BW_THREADS) mx = fmax(mx, numer(e));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) { double m = red[0]; for (int w = 1; w < BW_THREADS / 32; ++w) m = fmax(m, red[w]); bc = m; }
    __syncthreads();
    mx = bc;
    double s = 0.0;
    if (mx != neg_inf())
        for (int64_t e = tid; e < NN; e += BW_THREADS) { const double v = numer(e); s += (v == neg_inf()) ? 0.0 : exp(v - mx); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < BW_THREADS / 32; ++w) t += red[w]; bc = (mx == neg_inf()) ? mx : mx + log(t); }
    __syncthreads();
    const double Z = bc;
    for (int64_t e = tid; e < NN; e += BW_THREADS) xo[e] = numer(e) - Z;
    __syncthreads();
    // gamma[t,i] = lse_j xi[t,i,j]
    for (int i = tid; i < N; i += BW_THREADS) {
        double m = neg_inf();
        for (int j = 0; j < N; ++j) m = fmax(m, numer((int64_t)i * N + j));
        double t = 0.0;
        if (m != neg_inf())
            for (int j = 0; j < N; ++j) { const double v = numer((int64_t)i * N + j); t += (v == neg_inf()) ? 0.0 : exp(v - m); }
        gamma[f * N + i] = (m == neg_inf()) ? m : (m + log(t)) - Z;
    }
}


// ------------------------------------------------------------------------------------------------
// Two-kernel form of the same statistics for batches whose utterance count cannot fill the machine with one CTA per
// utterance (config 3: 1000 utterances of ~300 frames on 148 SMs).  The serial part of Baum-Welch is only the backward
// recursion; everything weighted by the occupancies is a sum over frames with no order.
//   em_gamma_kernel<G>  : G lanes per utterance (32/G utterances per warp, like the small trellis kernel): beta, Z_t,
//                         log gamma[t,j] for ALL t -> workspace [T,N]; xi / pi statistics and log P(O) in registers.
//                         alpha and b rows are prefetched PF frames ahead; exp / log are the table versions (common.cuh).
//   em_stats_kernel     : persistent CTAs over 64-frame tiles: occ[t,j,m] = exp(log gamma + log w + bm - b)
//                         (gmmhmm.py:89-106) into shared memory, then S0 += occ, S1 += occ x, S2c += occ (x-mu_old)^2
//                         (gmmhmm.py:117-147) with every thread owning fixed (gaussian, dim) entries; one atomic flush per CTA.
constexpr int EG_WARPS = 4;
constexpr int EG_PF = 8;
constexpr int ES_THREADS = 256;
constexpr int ES_TILE = 64;

template <int G>
__global__ void __launch_bounds__(EG_WARPS * 32)
em_gamma_kernel(FbArgs a, double* __restrict__ lgamma) {
    __shared__ __align__(16) double2 log_tab[128];
    __shared__ double exp_tab[64];
    fill_log_exp_tables(log_tab, exp_tab, threadIdx.x, EG_WARPS * 32);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PPW = 32 / G;
    const int N = a.N;
    const int slot = lane / G, gbase = slot * G, j = lane - gbase;
    const int64_t u = ((int64_t)blockIdx.x * EG_WARPS + warp) * PPW + slot;
    const bool utt_ok = u < a.n_utt;
    const bool act = utt_ok && j < N;
    const int64_t f0 = utt_ok ? __ldg(a.frame_off + u) : 0;
    const int T = utt_ok ? (int)(__ldg(a.frame_off + u + 1) - f0) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    if (Tmax == 0) return;
    auto grp = [&](double v, int q) { return __shfl_sync(0xffffffffu, v, gbase + q); };
    auto glse = [&](double v) {                                   // log-sum-exp over the G lanes of my group
        double mx = v;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sm = fast_exp_neg(v - mx, exp_tab);               // v = mx = -inf: NaN -> 0
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        return (mx == neg_inf()) ? mx : mx + fast_log(sm, log_tab);
    };
    double tr[G], xi_acc[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
        tr[q] = (act && q < N) ? __ldg(a.logA + (int64_t)j * N + q) : neg_inf();     // row j of log A
        xi_acc[q] = 0.0;
    }
    double pi_acc = 0.0;
    // log P(O) = lse_j alpha[T-1, j]    (hmm.py:274)
    const double L = glse((act && T > 0) ? __ldg(a.alpha + (f0 + T - 1) * N + j) : neg_inf());
    // step s handles frame t = T-1-s ; rows prefetched PF steps ahead
    const double* ap = a.alpha + (f0 + T - 1) * N + j;
    const double* bp = a.b + (f0 + T - 1) * N + j;
    double aq[EG_PF], bq[EG_PF];
#pragma unroll
    for (int i = 0; i < EG_PF; ++i) {
        const bool ok = act && i < T;
        aq[i] = ok ? __ldg(ap - (int64_t)i * N) : neg_inf();
        bq[i] = ok ? __ldg(bp - (int64_t)i * N) : neg_inf();
    }
    ap -= (int64_t)EG_PF * N;
    bp -= (int64_t)EG_PF * N;
    double* gp = lgamma + (f0 + T - 1) * N + j;
    double* bo = a.beta_out ? a.beta_out + (f0 + T - 1) * N + j : nullptr;
    double* go = a.gamma_out ? a.gamma_out + (f0 + T - 1) * N + j : nullptr;
    double cur = 0.0, e_next = 0.0;                               // beta[t+1,j], b[t+1,j]
    for (int s0 = 0; s0 < Tmax; s0 += EG_PF) {
#pragma unroll
        for (int i = 0; i < EG_PF; ++i) {
            const int s = s0 + i;
            const double al = aq[i], e = bq[i];
            const bool more = act && s + EG_PF < T;
            aq[i] = more ? __ldg(ap) : neg_inf();
            bq[i] = more ? __ldg(bp) : neg_inf();
            ap -= N;
            bp -= N;
            const bool live = act && s < T;
            const bool last = (s == 0);                           // frame T-1
            // beta[T-1] = 0 ; beta[t,i] = lse_q (A[i,q] + b[t+1,q] + beta[t+1,q])     (hmm.py:126-130)
            const double uj = cur + e_next;
            double cand[G];
            double mx = neg_inf();
#pragma unroll
            for (int q = 0; q < G; ++q) {
                cand[q] = grp(uj, q) + tr[q];
                mx = fmax(mx, cand[q]);
            }
            double sm = 0.0;
#pragma unroll
            for (int q = 0; q < G; ++q) sm += fast_exp_neg(cand[q] - mx, exp_tab);
            double beta = (mx == neg_inf()) ? mx : mx + fast_log(sm, log_tab);
            if (last) beta = 0.0;                                 // hmm.py:125
            const double ab = live ? al + beta : neg_inf();
            const double Z = glse(ab);                            // == lse_{i,q} of the xi numerator
            const double g2 = (ab == neg_inf() || Z == neg_inf()) ? neg_inf() : ab - Z;
            if (live) {
                *gp = g2;                                         // log gamma for ALL t (the mixture statistics use t = T-1 too)
                if (bo) *bo = beta;
                if (go) *go = last ? neg_inf() : g2;              // hmm.py:199
                if (!last && Z != neg_inf()) {
                    double rs = 0.0;
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        // xi[t,i,q] = alpha[t,i] + A[i,q] + b[t+1,q] + beta[t+1,q] - Z_t     (hmm.py:190-197)
                        const double ex = (q < N) ? fast_exp_neg(al + cand[q] - Z, exp_tab) : 0.0;
                        xi_acc[q] += ex;
                        rs += ex;
                    }
                    if (s == T - 1) pi_acc += rs;                 // t = 0: pi' = gamma[0]   (gmmhmm.py:86)
                }
                cur = beta;
                e_next = e;
            }
            gp -= N;
            if (bo) bo -= N;
            if (go) go -= N;
        }
    }
    // xi_sum[N*N] | pi_occ[N] | loglik | n_utt
    double* tail = a.acc + (int64_t)N * a.M * (2 * a.D + 1);
    if (act) {
#pragma unroll
        for (int q = 0; q < G; ++q)
            if (q < N && xi_acc[q] != 0.0) atomicAdd(tail + (int64_t)j * N + q, xi_acc[q]);
        if (pi_acc != 0.0) atomicAdd(tail + (int64_t)N * N + j, pi_acc);
        if (j == 0 && T > 0) {
            atomicAdd(tail + (int64_t)N * N + N, L);
            atomicAdd(tail + (int64_t)N * N + N + 1, 1.0);
        }
    }
}

__global__ void __launch_bounds__(ES_THREADS)
em_stats_kernel(FbArgs a, const double* __restrict__ lgamma, int64_t total_frames) {
    extern __shared__ __align__(16) unsigned char em_smem_raw[];
    double* sm = reinterpret_cast<double*>(em_smem_raw);
    const int N = a.N, M = a.M, D = a.D, NM = N * M;
    const int tid = threadIdx.x;
    double2* log_tab = reinterpret_cast<double2*>(sm);    // tables first: 16-byte aligned whatever the model shape
    double* exp_tab = sm + 256;
    double* acc_s = exp_tab + 64;                         // S0[NM] | S1[NM*D] | S2c[NM*D]
    double* mu_s = acc_s + NM * (2 * D + 1);
    double* logw_s = mu_s + NM * D;
    double* occ_s = logw_s + NM;                          // [ES_TILE][NM]
    double* x_s = occ_s + ES_TILE * NM;                   // [ES_TILE][D]
    for (int i = tid; i < NM * (2 * D + 1); i += ES_THREADS) acc_s[i] = 0.0;
    for (int i = tid; i < NM * D; i += ES_THREADS) mu_s[i] = a.mu[i];
    for (int i = tid; i < NM; i += ES_THREADS) logw_s[i] = a.logw[i];
    fill_log_exp_tables(log_tab, exp_tab, tid, ES_THREADS);
    __syncthreads();
    const int64_t n_tiles = (total_frames + ES_TILE - 1) / ES_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * ES_TILE;
        const int nt = (int)min((int64_t)ES_TILE, total_frames - t0);
        // mixture occupancies for ALL t (gmmhmm.py:98-100): exp(log gamma[t,j] + log w[j,m] + bm[t,j,m] - b[t,j])
        for (int i = tid; i < nt * NM; i += ES_THREADS) {
            const int tt = i / NM, p = i - tt * NM, j = p / M;
            const double g = __ldg(lgamma + (t0 + tt) * N + j), bj = __ldg(a.b + (t0 + tt) * N + j);
            double o = 0.0;
            if (g != neg_inf() && bj != neg_inf()) o = fast_exp_neg(g + logw_s[p] + __ldg(a.bm + (t0 + tt) * NM + p) - bj, exp_tab);
            occ_s[i] = o;
        }
        for (int i = tid; i < nt * D; i += ES_THREADS) {
            const int r = i / D, d = i - r * D;
            const int64_t row = t0 + r;
            x_s[i] = a.feats_f64 ? __ldg((const double*)a.feats + row * a.ld_feats + d)
                                 : (double)__ldg((const float*)a.feats + row * a.ld_feats + d);
        }
        __syncthreads();
        for (int e = tid; e < NM * D; e += ES_THREADS) {
            const int jm = e / D, d = e - jm * D;
            double s1 = 0.0, s2 = 0.0;
            const double m0 = mu_s[e];
#pragma unroll 4
            for (int tt = 0; tt < nt; ++tt) {
                const double o = occ_s[tt * NM + jm];
                const double x = x_s[tt * D + d];
                const double df = x - m0;
                s1 = fma(o, x, s1);
                s2 = fma(o * df, df, s2);
            }
            acc_s[NM + e] += s1;
            acc_s[NM + NM * D + e] += s2;
        }
        for (int jm = tid; jm < NM; jm += ES_THREADS) {
            double s0 = 0.0;
            for (int tt = 0; tt < nt; ++tt) s0 += occ_s[tt * NM + jm];
            acc_s[jm] += s0;
        }
        __syncthreads();
    }
    for (int i = tid; i < NM * (2 * D + 1); i += ES_THREADS) {
        const double v = acc_s[i];
        if (v != 0.0) atomicAdd(a.acc + i, v);
    }
}

// Third form: BOTH recursions are plain trellis passes (lnasr_forward_batch, lnasr_backward_batch -- independent of each
// other, so they can run concurrently on two streams) and EVERYTHING else is frame-parallel here: Z_t, log gamma, the xi /
// pi statistics (hmm.py:190-199, gmmhmm.py:83-86), the mixture occupancies and S0 / S1 / S2c (gmmhmm.py:89-147).
// The serial chain per utterance shrinks to 4 exp + 1 log per step; with ~1000 utterances (config 3) the recursions are
// pure latency, so this is what sets the iteration time.
__global__ void __launch_bounds__(ES_THREADS)
em_stats_ab_kernel(FbArgs a, const double* __restrict__ beta, int64_t total_frames) {
    extern __shared__ __align__(16) unsigned char em_smem_raw[];
    double* sm = reinterpret_cast<double*>(em_smem_raw);
    const int N = a.N, M = a.M, D = a.D, NM = N * M;
    const int tid = threadIdx.x;
    double2* log_tab = reinterpret_cast<double2*>(sm);
    double* exp_tab = sm + 256;
    double* acc_s = exp_tab + 64;                         // S0[NM] | S1[NM*D] | S2c[NM*D] | xi[N*N] | pi[N] | loglik | n_utt
    const int n_acc = NM * (2 * D + 1) + N * N + N + 2;
    double* mu_s = acc_s + n_acc;
    double* logw_s = mu_s + NM * D;
    double* A_s = logw_s + NM;                            // [N][N]
    double* occ_s = A_s + N * N;                          // [ES_TILE][NM], 16-byte aligned for the double2 loads below
    occ_s += (occ_s - sm) & 1;
    double* x_s = occ_s + ES_TILE * NM;                   // [ES_TILE][D]
    double* al_s = x_s + ES_TILE * D;                     // [ES_TILE][N]      alpha[t]
    double* ub_s = al_s + ES_TILE * N;                    // [ES_TILE+1][N]    beta[t] (then b[t]+beta[t] for the xi pass)
    double* z_s = ub_s + (ES_TILE + 1) * N;               // [ES_TILE]
    int* flag_s = reinterpret_cast<int*>(z_s + ES_TILE);  // [ES_TILE] bit0 = first frame of its utterance, bit1 = last
    for (int i = tid; i < n_acc; i += ES_THREADS) acc_s[i] = 0.0;
    for (int i = tid; i < NM * D; i += ES_THREADS) mu_s[i] = a.mu[i];
    for (int i = tid; i < NM; i += ES_THREADS) logw_s[i] = a.logw[i];
    for (int i = tid; i < N * N; i += ES_THREADS) A_s[i] = a.logA[i];
    fill_log_exp_tables(log_tab, exp_tab, tid, ES_THREADS);
    __syncthreads();
    double* xi_s = acc_s + NM * (2 * D + 1);
    double* pi_s = xi_s + N * N;
    const int64_t n_tiles = (total_frames + ES_TILE - 1) / ES_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * ES_TILE;
        const int nt = (int)min((int64_t)ES_TILE, total_frames - t0);
        const int nt1 = (int)min((int64_t)ES_TILE + 1, total_frames - t0);       // one more row for the xi terms
        for (int i = tid; i < nt * N; i += ES_THREADS) al_s[i] = __ldg(a.alpha + t0 * N + i);
        for (int i = tid; i < nt1 * N; i += ES_THREADS) ub_s[i] = __ldg(beta + t0 * N + i);
        for (int i = tid; i < nt; i += ES_THREADS) {
            const int u = find_utt(a.frame_off, a.n_utt, t0 + i);
            flag_s[i] = (t0 + i == __ldg(a.frame_off + u) ? 1 : 0) | (t0 + i == __ldg(a.frame_off + u + 1) - 1 ? 2 : 0);
        }
        for (int i = tid; i < nt * D; i += ES_THREADS) {
            const int r = i / D, d = i - r * D;
            const int64_t row = t0 + r;
            x_s[i] = a.feats_f64 ? __ldg((const double*)a.feats + row * a.ld_feats + d)
                                 : (double)__ldg((const float*)a.feats + row * a.ld_feats + d);
        }
        __syncthreads();
        // Z_t = lse_j (alpha + beta)  -- one thread per frame
        for (int tt = tid; tt < nt; tt += ES_THREADS) {
            double mx = neg_inf();
            for (int j = 0; j < N; ++j) mx = fmax(mx, al_s[tt * N + j] + ub_s[tt * N + j]);
            double sum = 0.0;
            for (int j = 0; j < N; ++j) sum += fast_exp_neg(al_s[tt * N + j] + ub_s[tt * N + j] - mx, exp_tab);
            const double Z = (mx == neg_inf()) ? mx : mx + fast_log(sum, log_tab);
            z_s[tt] = Z;
            if ((flag_s[tt] & 2) && Z == Z) {                      // last frame: beta = 0, so Z = log P(O)   (hmm.py:274)
                atomicAdd(acc_s + n_acc - 2, Z);
                atomicAdd(acc_s + n_acc - 1, 1.0);
            }
        }
        __syncthreads();
        // mixture occupancies for ALL t (gmmhmm.py:98-100); gamma output; pi statistics
        for (int i = tid; i < nt * NM; i += ES_THREADS) {
            const int tt = i / NM, p = i - tt * NM, j = p / M;
            const double ab = al_s[tt * N + j] + ub_s[tt * N + j], Z = z_s[tt];
            const double bj = __ldg(a.b + (t0 + tt) * N + j);
            double o = 0.0;
            if (ab != neg_inf() && Z != neg_inf() && bj != neg_inf())
                o = fast_exp_neg(ab - Z + logw_s[p] + __ldg(a.bm + (t0 + tt) * NM + p) - bj, exp_tab);
            occ_s[i] = o;
        }
        for (int i = tid; i < nt * N; i += ES_THREADS) {
            const int tt = i / N;
            const double ab = al_s[i] + ub_s[i], Z = z_s[tt];
            const double g2 = (ab == neg_inf() || Z == neg_inf()) ? neg_inf() : ab - Z;
            if (a.gamma_out) a.gamma_out[t0 * N + i] = (flag_s[tt] & 2) ? neg_inf() : g2;      // hmm.py:199
            if ((flag_s[tt] & 3) == 1 && g2 != neg_inf()) atomicAdd(pi_s + (i - tt * N), fast_exp_neg(g2, exp_tab));   // pi' = gamma[0]
        }
        __syncthreads();
        // u[t,j] = b[t,j] + beta[t,j] replaces beta in place (rows 0..nt1-1); row tt+1 feeds xi[tt]
        for (int i = tid; i < nt1 * N; i += ES_THREADS) ub_s[i] += __ldg(a.b + t0 * N + i);
        // sufficient statistics: a work item = (4 consecutive Gaussians, one dim): the feature value is read once for four
        // occupancies (two 16-byte loads when N*M is a multiple of 4), 1.25 shared-memory loads per 4 x 4 float64 operations
        {
            const int ngrp = (NM + 3) >> 2;
            for (int w = tid; w < ngrp * D; w += ES_THREADS) {
                const int g4 = w / D, d = w - g4 * D;
                const int jm0 = g4 * 4;
                const int cnt = min(4, NM - jm0);
                double m0[4], s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int q = 0; q < 4; ++q) m0[q] = (q < cnt) ? mu_s[(jm0 + q) * D + d] : 0.0;
                if ((NM & 3) == 0) {
                    for (int tt = 0; tt < nt; ++tt) {
                        const double x = x_s[tt * D + d];
                        const double2 o01 = *reinterpret_cast<const double2*>(occ_s + tt * NM + jm0);
                        const double2 o23 = *reinterpret_cast<const double2*>(occ_s + tt * NM + jm0 + 2);
                        const double o[4] = {o01.x, o01.y, o23.x, o23.y};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double df = x - m0[q];
                            s1[q] = fma(o[q], x, s1[q]);
                            s2[q] = fma(o[q] * df, df, s2[q]);
                        }
                    }
                } else {
                    for (int tt = 0; tt < nt; ++tt) {
                        const double x = x_s[tt * D + d];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double o = (q < cnt) ? occ_s[tt * NM + jm0 + q] : 0.0;
                            const double df = x - m0[q];
                            s1[q] = fma(o, x, s1[q]);
                            s2[q] = fma(o * df, df, s2[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < cnt) {
                        acc_s[NM + (jm0 + q) * D + d] += s1[q];
                        acc_s[NM + NM * D + (jm0 + q) * D + d] += s2[q];
                    }
            }
        }
        for (int jm = tid; jm < NM; jm += ES_THREADS) {
            double s0 = 0.0;
            for (int tt = 0; tt < nt; ++tt) s0 += occ_s[tt * NM + jm];
            acc_s[jm] += s0;
        }
        __syncthreads();
        // xi[t,i,j] = alpha[t,i] + A[i,j] + b[t+1,j] + beta[t+1,j] - Z_t, t not the last frame of its utterance (hmm.py:190-197)
        // (lane = transition, warp = every 8th frame: N*N is 9 for a 3-state model, far fewer than the block has threads)
        {
            const int warp = tid >> 5, lane = tid & 31;
            for (int e = lane; e < N * N; e += 32) {
                const int i = e / N, j = e - i * N;
                const double aij = A_s[e];
                if (aij == neg_inf()) continue;
                double sx = 0.0;
                for (int tt = warp; tt < nt; tt += ES_THREADS / 32) {
                    if (flag_s[tt] & 2) continue;
                    sx += fast_exp_neg(al_s[tt * N + i] + aij + ub_s[(tt + 1) * N + j] - z_s[tt], exp_tab);   // -inf / NaN -> 0
                }
                if (sx != 0.0) atomicAdd(xi_s + e, sx);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n_acc; i += ES_THREADS) {
        const double v = acc_s[i];
        if (v != 0.0) atomicAdd(a.acc + i, v);
    }
}

static int launch_stats_ab(const FbArgs& a, const double* beta, int64_t total_frames, cudaStream_t stream) {
    const int NM = a.N * a.M, N = a.N;
    const size_t dbl = 256 + 64 + (size_t)NM * (2 * a.D + 1) + (size_t)N * N + N + 2 + (size_t)NM * a.D + NM + (size_t)N * N +
                       (size_t)ES_TILE * NM + (size_t)ES_TILE * a.D + (size_t)ES_TILE * N + (size_t)(ES_TILE + 1) * N + ES_TILE + ES_TILE / 2 + 4;
    const size_t smem = dbl * sizeof(double);
    if (smem > 220 * 1024) {
        set_error("em_statistics: N=%d, N*M=%d at dim=%d needs %zu bytes of shared memory (> 220 KB); model too large", N, NM, a.D, smem);
        return LNASR_EINVAL;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / smem));
    const int64_t n_tiles = ceil_div(total_frames, ES_TILE);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sm_count() * per_sm));
    if (smem > 48 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(em_stats_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    em_stats_ab_kernel<<<grid, ES_THREADS, smem, stream>>>(a, beta, total_frames);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

static int launch_fb_split(const FbArgs& a, int64_t total_frames, double* lgamma, cudaStream_t stream) {
    int G = 1;
    while (G < a.N) G <<= 1;
    const int ppw = 32 / G;
    const unsigned grid_g = (unsigned)ceil_div(a.n_utt, (int64_t)EG_WARPS * ppw);
    switch (G) {
        case 1:  em_gamma_kernel<1><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
        case 2:  em_gamma_kernel<2><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
        case 4:  em_gamma_kernel<4><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
        case 8:  em_gamma_kernel<8><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
        case 16: em_gamma_kernel<16><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
        default: em_gamma_kernel<32><<<grid_g, EG_WARPS * 32, 0, stream>>>(a, lgamma); break;
    }
    LNASR_LAUNCH_CHECK();
    const int NM = a.N * a.M;
    const size_t dbl = (size_t)NM * (2 * a.D + 1) + (size_t)NM * a.D + NM + (size_t)ES_TILE * NM + (size_t)ES_TILE * a.D + 256 + 64;
    const size_t smem = dbl * sizeof(double);
    if (smem > 220 * 1024) {
        set_error("fb_accumulate: N*M=%d at dim=%d needs %zu bytes of shared memory (> 220 KB); model too large", NM, a.D, smem);
        return LNASR_EINVAL;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / smem));
    const int64_t n_tiles = ceil_div(total_frames, ES_TILE);
    const int grid_s = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sm_count() * per_sm));
    if (smem > 48 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(em_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    em_stats_kernel<<<grid_s, ES_THREADS, smem, stream>>>(a, lgamma, total_frames);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <bool GMM>
static int launch_fb(const FbArgs& a, cudaStream_t stream) {
    int G = 1;
    while (G < a.N) G <<= 1;
    if (G < 4) G = 4;
    const int NM = a.N * a.M;
    size_t dbl = (size_t)2 * EM_TB * a.N;
    if (GMM) dbl += (size_t)NM * (2 * a.D + 1) + (size_t)NM * a.D + NM + (size_t)2 * EM_TB * NM + (size_t)EM_TB * a.D;
    const size_t smem = dbl * sizeof(double);
    if (smem > 220 * 1024) {
        set_error("fb_accumulate: N*M=%d at dim=%d needs %zu bytes of shared memory (> 220 KB); model too large for the fused kernel", NM, a.D, smem);
        return LNASR_EINVAL;
    }
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
    int grid = (int)std::min<int64_t>(a.n_utt, (int64_t)sm_count() * per_sm);
    if (grid < 1) grid = 1;
#define LNASR_FB_CASE(GG)                                                                                              \
    case GG:                                                                                                           \
        if (smem > 48 * 1024)                                                                                          \
            LNASR_CUDA(cudaFuncSetAttribute(fb_accum_kernel<GG, GMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fb_accum_kernel<GG, GMM><<<grid, EM_THREADS, smem, stream>>>(a);                                               \
        break;
    switch (G) {
        LNASR_FB_CASE(4)
        LNASR_FB_CASE(8)
        LNASR_FB_CASE(16)
        LNASR_FB_CASE(32)
        default: set_error("fb_accumulate: N=%d > 32 is not built yet", a.N); return LNASR_EINVAL;
    }
#undef LNASR_FB_CASE
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

}  // namespace lnasr

using namespace lnasr;

extern "C" int64_t lnasr_em_acc_doubles(int32_t N, int32_t n_mix, int32_t dim) {
    const int64_t NM = (int64_t)N * n_mix;
    return NM * (2 * dim + 1) + (int64_t)N * N + N + 2;
}

extern "C" int64_t lnasr_hmm_acc_doubles(int32_t N, int32_t n_symbols) {
    return (int64_t)N * N + N + (int64_t)N * n_symbols + N + 2;
}

extern "C" size_t lnasr_fb_workspace_bytes(int64_t total_frames, int32_t N) {
    return (size_t)total_frames * N * sizeof(double) + 256;      // log gamma [T, N]
}

extern "C" int lnasr_fb_accumulate(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                                   const double* alpha, const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                                   const double* logA, const double* logw, const double* mu, int32_t N, int32_t n_mix,
                                   int32_t dim, double* acc, double* gamma_out, double* beta_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    LNASR_CHECK_ARG(feats && b && bm && alpha && frame_offsets && logA && logw && mu && acc, "fb_accumulate: null pointer");
    LNASR_CHECK_ARG(feats_dtype == LNASR_DT_F32 || feats_dtype == LNASR_DT_F64, "fb_accumulate: feats must be float32/float64");
    LNASR_CHECK_ARG(N >= 1 && N <= 32, "fb_accumulate: the fused kernel is built for 1 <= N <= 32 (got %d)", N);
    LNASR_CHECK_ARG(n_mix >= 1 && dim >= 1 && ld_feats >= dim && n_utt >= 0, "fb_accumulate: bad shape");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    FbArgs a{};
    a.feats = feats; a.feats_f64 = (feats_dtype == LNASR_DT_F64); a.ld_feats = ld_feats;
    a.b = b; a.bm = bm; a.alpha = alpha; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.logA = logA; a.logw = logw; a.mu = mu; a.N = N; a.M = n_mix; a.D = dim; a.n_symbols = 0;
    a.acc = acc; a.gamma_out = gamma_out; a.beta_out = beta_out;
    if (workspace) {
        if (workspace_bytes < lnasr_fb_workspace_bytes(total_frames, N)) {
            set_error("fb_accumulate: workspace %zu < %zu bytes", workspace_bytes, lnasr_fb_workspace_bytes(total_frames, N));
            return LNASR_EWORKSPACE;
        }
        return launch_fb_split(a, total_frames, (double*)workspace, (cudaStream_t)stream);
    }
    return launch_fb<true>(a, (cudaStream_t)stream);
}

extern "C" int lnasr_em_statistics(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                                   const double* alpha, const double* beta, const int64_t* frame_offsets, int32_t n_utt,
                                   int64_t total_frames, const double* logA, const double* logw, const double* mu, int32_t N,
                                   int32_t n_mix, int32_t dim, double* acc, double* gamma_out, void* stream) {
    LNASR_CHECK_ARG(feats && b && bm && alpha && beta && frame_offsets && logA && logw && mu && acc, "em_statistics: null pointer");
    LNASR_CHECK_ARG(feats_dtype == LNASR_DT_F32 || feats_dtype == LNASR_DT_F64, "em_statistics: feats must be float32/float64");
    LNASR_CHECK_ARG(N >= 1 && N <= 32 && n_mix >= 1 && dim >= 1 && ld_feats >= dim && n_utt >= 0, "em_statistics: bad shape (1 <= N <= 32)");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    FbArgs a{};
    a.feats = feats; a.feats_f64 = (feats_dtype == LNASR_DT_F64); a.ld_feats = ld_feats;
    a.b = b; a.bm = bm; a.alpha = alpha; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.logA = logA; a.logw = logw; a.mu = mu; a.N = N; a.M = n_mix; a.D = dim; a.n_symbols = 0;
    a.acc = acc; a.gamma_out = gamma_out; a.beta_out = nullptr;
    return launch_stats_ab(a, beta, total_frames, (cudaStream_t)stream);
}

extern "C" int lnasr_mstep_finalize(const double* acc, int32_t N, int32_t n_mix, int32_t dim, double min_std,
                                    double* logA, double* logpi, double* logw, double* mu, double* var, void* stream) {
    LNASR_CHECK_ARG(acc && logA && logpi && logw && mu && var, "mstep_finalize: null pointer");
    LNASR_CHECK_ARG(N >= 1 && n_mix >= 1 && dim >= 1, "mstep_finalize: bad shape");
    const int64_t n = (int64_t)N * n_mix * dim;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 128), 1024);
    mstep_finalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(acc, N, n_mix, dim, min_std, logA, logpi, logw, mu, var);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_hmm_fb_accumulate(const double* b, const double* alpha, const int32_t* obs, const int64_t* frame_offsets,
                                       int32_t n_utt, int64_t total_frames, const double* logA, int32_t N, int32_t n_symbols,
                                       double* acc, double* gamma_out, double* beta_out, void* stream) {
    LNASR_CHECK_ARG(b && alpha && obs && frame_offsets && logA && acc, "hmm_fb_accumulate: null pointer");
    LNASR_CHECK_ARG(N >= 1 && N <= 32, "hmm_fb_accumulate: the fused kernel is built for 1 <= N <= 32 (got %d)", N);
    LNASR_CHECK_ARG(n_symbols >= 1 && n_utt >= 0, "hmm_fb_accumulate: bad shape");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    FbArgs a{};
    a.b = b; a.alpha = alpha; a.obs = obs; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.logA = logA; a.N = N; a.M = 1; a.D = 1; a.n_symbols = n_symbols;
    a.acc = acc; a.gamma_out = gamma_out; a.beta_out = beta_out;
    return launch_fb<false>(a, (cudaStream_t)stream);
}

extern "C" int lnasr_hmm_mstep_finalize(const double* acc, int32_t N, int32_t n_symbols, double* logA, double* logB,
                                        double* logpi, void* stream) {
    LNASR_CHECK_ARG(acc && logA && logB && logpi && N >= 1 && n_symbols >= 1, "hmm_mstep_finalize: bad arguments");
    const int64_t n = (int64_t)N * std::max(N, n_symbols);
    const int grid = (int)std::min<int64_t>(ceil_div(n, 128), 1024);
    hmm_mstep_finalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(acc, N, n_symbols, logA, logB, logpi);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_baumwelch_xi(const double* alpha, const double* beta, const double* b, const double* logA,
                                  const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames, int32_t N,
                                  double* xi_out, double* gamma_out, void* stream) {
    LNASR_CHECK_ARG(alpha && beta && b && logA && frame_offsets && xi_out && gamma_out, "baumwelch_xi: null pointer");
    LNASR_CHECK_ARG(N >= 1 && n_utt >= 0 && total_frames < (1LL << 31), "baumwelch_xi: bad shape");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    baumwelch_xi_kernel<<<(unsigned)total_frames, BW_THREADS, 0, (cudaStream_t)stream>>>(alpha, beta, b, logA, frame_offsets,
                                                                                        n_utt, N, xi_out, gamma_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}
