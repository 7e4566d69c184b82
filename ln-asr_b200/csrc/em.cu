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
    if (smem > 32 * 1024)
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
    if (smem > 32 * 1024)
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
        if (smem > 32 * 1024)                                                                                          \
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

// ------------------------------------------------------------------------------------------------
// General form (any N; used for N > 32): the same statistics with NOTHING resident per model -- accumulators in global
// memory (float64 atomics, one flush per block and entry), transitions as a compact list of the finite entries of log A,
// every kernel frame-parallel over alpha and beta from the trellis passes (lnasr_forward_batch / lnasr_backward_batch run
// any N through the ELL lists).  Large models are sparse (left-to-right: <= 3 successors), so the xi pass costs
// T x nnz(A), not T x N^2, and the reference's xi [T,N,N] / [N,M,T,D,D] arrays still never exist.
//   em_translist_kernel : (i, j, log a_ij) of the finite transitions, rows in arbitrary order, sorted by j inside a row
//   em_z_kernel         : warp per frame: Z_t = lse_j(alpha + beta)  (== lse_{i,j} of the xi numerator, hmm.py:190-197);
//                         log P(O) at last frames (hmm.py:274), pi statistics at first frames (gmmhmm.py:86), gamma output
//   em_xi_kernel        : thread = transition, grid-stride over frames: sum_t exp(alpha[t,i] + a_ij + b[t+1,j] + beta[t+1,j] - Z_t)
//   em_moments_kernel   : block = 32 Gaussians x a frame range: occupancies (gmmhmm.py:89-106) staged per 32 frames, S0 / S1 / S2c
//                         (gmmhmm.py:117-147) in registers, flushed once
//   em_discrete_kernel  : discrete emissions (hmm.py:202-225): gamma_sum and B numerators
struct EmGen {
    const double* alpha; const double* beta; const double* b; const int64_t* frame_off; int n_utt; int64_t total; int N;
    double* z;          // [T] Z_t (-inf / NaN possible)
    double* zx;         // [T] Z_t for the transition statistics: +inf at the last frame of an utterance and where Z_t is unusable
    int* count; int* ti; int* tj; double* ta;
};

__global__ void __launch_bounds__(256) em_translist_kernel(const double* __restrict__ logA, int N, int* __restrict__ count,
                                                           int* __restrict__ ti, int* __restrict__ tj, double* __restrict__ ta) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int row_base = 0;                                              // finite entries of row i seen so far
    for (int j0 = 0; j0 < N; j0 += 256) {
        const int j = j0 + tid;
        const double a = (j < N) ? logA[(int64_t)i * N + j] : neg_inf();
        const bool fin = (a != neg_inf()) && (a == a);
        const unsigned m = __ballot_sync(0xffffffffu, fin);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += s_warp[w]; tot += s_warp[w]; }
        if (tid == 0) s_base = tot ? atomicAdd(count, tot) : 0;
        __syncthreads();
        if (fin) {
            const int pos = s_base + before + __popc(m & ((1u << lane) - 1));
            ti[pos] = i; tj[pos] = j; ta[pos] = a;
        }
        row_base += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) em_z_kernel(EmGen g, double* __restrict__ pi_occ, double* __restrict__ loglik,
                                                   double* __restrict__ n_utt_acc, double* __restrict__ gamma_out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= g.total) return;
    const int N = g.N;
    const double* al = g.alpha + t * N;
    const double* be = g.beta + t * N;
    double mx = neg_inf();
    for (int j = lane; j < N; j += 32) mx = fmax(mx, al[j] + be[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_d(mx, lane ^ o));
    double sum = 0.0;
    if (mx != neg_inf())
        for (int j = lane; j < N; j += 32) { const double v = al[j] + be[j]; sum += (v == neg_inf()) ? 0.0 : exp(v - mx); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += shfl_d(sum, lane ^ o);
    const double Z = (mx == neg_inf()) ? mx : mx + log(sum);
    const int u = find_utt(g.frame_off, g.n_utt, t);
    const bool first = (t == __ldg(g.frame_off + u)), last = (t == __ldg(g.frame_off + u + 1) - 1);
    const bool usable = (Z == Z) && (Z != neg_inf());
    if (lane == 0) {
        g.z[t] = Z;
        g.zx[t] = (last || !usable) ? __longlong_as_double(0x7ff0000000000000LL) : Z;
        if (last && Z == Z) { atomicAdd(loglik, Z); atomicAdd(n_utt_acc, 1.0); }      // beta[T-1] = 0: Z = log P(O)
    }
    if (gamma_out || (first && !last && usable))
        for (int j = lane; j < N; j += 32) {
            const double ab = al[j] + be[j];
            const double g2 = (ab == neg_inf() || !usable) ? neg_inf() : ab - Z;
            if (gamma_out) gamma_out[t * N + j] = last ? neg_inf() : g2;               // hmm.py:199
            if (first && !last && g2 != neg_inf()) atomicAdd(pi_occ + j, exp(g2));       // pi' = gamma[0]
        }
}

__global__ void __launch_bounds__(256) em_xi_kernel(EmGen g, double* __restrict__ xi_sum) {
    __shared__ __align__(16) double2 log_tab[128];
    __shared__ double exp_tab[64];
    fill_log_exp_tables(log_tab, exp_tab, threadIdx.x, 256);
    __syncthreads();
    const int e = blockIdx.y * 256 + threadIdx.x;
    if (e >= *g.count) return;
    const int i = g.ti[e], j = g.tj[e], N = g.N;
    const double a = g.ta[e];
    double sx = 0.0;
    for (int64_t t = blockIdx.x; t + 1 < g.total; t += gridDim.x) {
        const double zx = g.zx[t];
        if (zx > 1e308) continue;                                    // last frame of its utterance / impossible utterance
        const double al = g.alpha[t * N + i];
        const double x = al + a + (g.b[(t + 1) * N + j] + g.beta[(t + 1) * N + j]) - zx;
        sx += fast_exp_neg(x, exp_tab);                             // -inf / NaN -> 0
    }
    if (sx != 0.0) atomicAdd(xi_sum + (int64_t)i * N + j, sx);
}

constexpr int EMG_GT = 32;      // Gaussians per block
constexpr int EMG_FT = 32;      // frames staged per round
constexpr int EMG_MAXP = 8;     // (Gaussian, dim) pairs per thread: 32 * dim <= 256 * 8  ->  dim <= 64

__global__ void __launch_bounds__(256) em_moments_kernel(EmGen g, const void* __restrict__ feats, int feats_f64, int ld_feats,
                                                         const double* __restrict__ bm, const double* __restrict__ logw,
                                                         const double* __restrict__ mu, int M, int D, int64_t frames_per_block,
                                                         double* __restrict__ acc) {
    __shared__ __align__(16) double2 log_tab[128];
    __shared__ double exp_tab[64];
    __shared__ double occ_s[EMG_FT][EMG_GT];
    __shared__ double x_s[EMG_FT][64];
    const int tid = threadIdx.x, N = g.N, NM = N * M;
    fill_log_exp_tables(log_tab, exp_tab, tid, 256);
    const int jm0 = blockIdx.y * EMG_GT;
    const int ng = min(EMG_GT, NM - jm0);
    const int64_t t_begin = (int64_t)blockIdx.x * frames_per_block, t_end = min(g.total, t_begin + frames_per_block);
    double s1[EMG_MAXP], s2[EMG_MAXP], m0[EMG_MAXP];
#pragma unroll
    for (int k = 0; k < EMG_MAXP; ++k) {
        const int pr = tid + 256 * k, gq = pr / D, d = pr - gq * D;
        s1[k] = s2[k] = 0.0;
        m0[k] = (gq < ng) ? mu[(int64_t)(jm0 + gq) * D + d] : 0.0;
    }
    double s0 = 0.0;
    __syncthreads();
    for (int64_t t0 = t_begin; t0 < t_end; t0 += EMG_FT) {
        const int nt = (int)min((int64_t)EMG_FT, t_end - t0);
        // mixture occupancies for ALL t (gmmhmm.py:98-100): exp(alpha + beta - Z + log w + bm - b)
        for (int i = tid; i < EMG_FT * EMG_GT; i += 256) {
            const int tt = i / EMG_GT, gq = i - tt * EMG_GT;
            double o = 0.0;
            if (tt < nt && gq < ng) {
                const int64_t t = t0 + tt;
                const int jm = jm0 + gq, j = jm / M;
                const double ab = g.alpha[t * N + j] + g.beta[t * N + j], Z = g.z[t], bj = g.b[t * N + j];
                if (ab != neg_inf() && Z != neg_inf() && Z == Z && bj != neg_inf())
                    o = fast_exp_neg(ab - Z + logw[jm] + bm[t * NM + jm] - bj, exp_tab);
            }
            occ_s[tt][gq] = o;
        }
        for (int i = tid; i < nt * D; i += 256) {
            const int r = i / D, d = i - r * D;
            const int64_t row = t0 + r;
            x_s[r][d] = feats_f64 ? ((const double*)feats)[row * ld_feats + d] : (double)((const float*)feats)[row * ld_feats + d];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EMG_MAXP; ++k) {
            const int pr = tid + 256 * k, gq = pr / D, d = pr - gq * D;
            if (gq < ng) {
                double a1 = 0.0, a2 = 0.0;
                for (int tt = 0; tt < nt; ++tt) {
                    const double o = occ_s[tt][gq], x = x_s[tt][d], df = x - m0[k];
                    a1 = fma(o, x, a1);
                    a2 = fma(o * df, df, a2);
                }
                s1[k] += a1;
                s2[k] += a2;
            }
        }
        if (tid < ng)
            for (int tt = 0; tt < nt; ++tt) s0 += occ_s[tt][tid];
        __syncthreads();
    }
    double* S0 = acc;
    double* S1 = acc + NM;
    double* S2 = S1 + (int64_t)NM * D;
#pragma unroll
    for (int k = 0; k < EMG_MAXP; ++k) {
        const int pr = tid + 256 * k, gq = pr / D, d = pr - gq * D;
        if (gq < ng) {
            if (s1[k] != 0.0) atomicAdd(S1 + (int64_t)(jm0 + gq) * D + d, s1[k]);
            if (s2[k] != 0.0) atomicAdd(S2 + (int64_t)(jm0 + gq) * D + d, s2[k]);
        }
    }
    if (tid < ng && s0 != 0.0) atomicAdd(S0 + jm0 + tid, s0);
}

// The same moments on the FLOAT64 TENSOR CORES (dim <= 39).  S0 / S1 / S2 of a tile of Gaussians are one contraction over
// frames,  S[g, :] = sum_t occ[t, g] . [x'^2 | x' | 1][t, :]  with x' = x - c (c = mu_old of the tile's first Gaussian: raw moments
// about a point inside the tile's cloud instead of the origin), as mma.sync.m8n8k4.f64 with M = Gaussians, K = frames, N = the
// 80 moment columns [x'0^2, x'0, x'1^2, x'1, ... | 1 | 0].  The SIMT kernel above loads two shared-memory operands per
// multiply-add pair (it ran at 4x its FP64 time, bound by shared memory); here 12 fragment loads feed 20 DMMAs = 5120
// multiply-adds.  Block = 8 warps = 2 halves of the 32 Gaussians x 4 slices of the 32 staged frames; every warp keeps its
// 16 x 80 partial sums in registers for its whole frame range.  At the flush a thread holds S2', S1' of one (Gaussian, dim)
// pair, which is all the conversion to the accumulators' form needs (both linear in the partial sums, so every warp converts
// and adds its own):  S1 = S1' + c S0,  S2c = sum occ (x - mu_old)^2 = S2' - 2 (mu_old - c) S1' + (mu_old - c)^2 S0.
constexpr int EMM_LDA = 36;     // occupancy row stride (doubles): 4 frames x 8 Gaussians of a fragment fall into 32 distinct banks
constexpr int EMM_COLS = 80;
constexpr int EMM_LDB = 84;     // moment row stride, same property

__device__ __forceinline__ void em_dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int EMM_THREADS = 256;      // 8 warps = 2 halves of the 32 Gaussians x 4 slices of the round's frames
constexpr int EMM_FT = 32;            // frames per round
constexpr int EMM_KS = EMM_THREADS / 64;                        // frame slices
constexpr int EMM_KSTEPS = EMM_FT / 4 / EMM_KS;                 // k-steps (4 frames) per warp and round
constexpr int EMM_STAGES = 1;         // rounds in flight (measured: one more round ahead = 86 KB of slots, one block per SM -- slower)
constexpr int EMM_OQ = EMM_FT * EMG_GT / EMM_THREADS;           // (frame, Gaussian) pairs per thread and round
constexpr int EMM_RAW_OCC = 4 * EMM_OQ;                         // raw slots per thread: (alpha, beta, b, bm) of each pair
constexpr int EMM_RAW_X = (EMM_FT * 39 + EMM_THREADS - 1) / EMM_THREADS;      // + its feature values
constexpr int EMM_RAW = EMM_RAW_OCC + EMM_RAW_X;
constexpr size_t EMM_RAW_BYTES = (size_t)EMM_STAGES * EMM_RAW * EMM_THREADS * sizeof(double);

// exp(d) for d <= 0 without a table (the shared-memory table of fast_exp_neg is read at random: ~2.5 wavefronts per warp, in
// a kernel that runs out of shared-memory bandwidth first): d = n ln2 + r, |r| <= ln2 / 2, Taylor to r^13 (< 4e-18)
__device__ __forceinline__ double exp_neg_poly(double d) {
    const bool ok = d > -700.0;
    const double dc = ok ? d : -700.0;
    const double MAGIC = 6755399441055744.0;                        // 1.5 * 2^52: the integer lands in the low word
    double nf = fma(dc, 1.4426950408889634, MAGIC);
    const int n = __double2loint(nf);
    nf -= MAGIC;
    double r = fma(nf, -0.6931471803691238, dc);                    // ln2, high part (32 significant bits)
    r = fma(nf, -1.9082146973659064e-10, r);                        //      low part
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return ok ? v : 0.0;
}

__device__ __forceinline__ void em_cp_async(void* smem_dst, const void* gmem_src, int bytes8) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (bytes8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

// The inputs of a round (32 frames) are fetched by cp.async into PRIVATE shared-memory slots (raw[stage][slot][thread]) a
// round ahead, while the tensor cores work: nothing but the thread's own wait_group stands between a value and its use, and
// no register is held across the latency (bm is streamed from HBM once: ~2 us away).  Two blocks per SM.
__global__ void __launch_bounds__(EMM_THREADS, 2) em_moments_mma_kernel(EmGen g, const void* __restrict__ feats, int feats_f64,
                                                                        int ld_feats, const double* __restrict__ bm,
                                                                        const double* __restrict__ logw, const double* __restrict__ mu,
                                                                        int M, int D, int64_t frames_per_block, double* __restrict__ acc) {
    extern __shared__ __align__(16) double emm_raw[];       // [EMM_STAGES][EMM_RAW][EMM_THREADS]
    __shared__ double occ_s[EMM_FT][EMM_LDA];
    __shared__ __align__(16) double x_s[EMM_FT][EMM_LDB];
    __shared__ double c_s[40];
    __shared__ double z_s[EMM_THREADS / 32][EMM_OQ];        // Z_t of the frames of each warp's pairs: fetched by the warp's own lanes
                                                            // 0..3, so cp.async.wait_group + __syncwarp makes them visible to the warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, N = g.N, NM = N * M;
    const int gh = warp & 1, ks = warp >> 1;               // half of the Gaussian tile, slice of the staged frames
    const int fr4 = lane & 3, r8 = lane >> 2;
    const int jm0 = blockIdx.y * EMG_GT;
    const int ng = min(EMG_GT, NM - jm0);
    const int64_t t_begin = (int64_t)blockIdx.x * frames_per_block, t_end = min(g.total, t_begin + frames_per_block);
    if (tid < 40) c_s[tid] = (tid < D) ? mu[(int64_t)jm0 * D + tid] : 0.0;
    for (int i = tid; i < EMM_FT * EMM_LDB; i += EMM_THREADS) {
        const int c = i % EMM_LDB;
        (&x_s[0][0])[i] = (c == 2 * D) ? 1.0 : 0.0;        // the constant column; the padding columns stay zero
    }
    double d[2][EMM_COLS / 8][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < EMM_COLS / 8; ++nb) d[mb][nb][0] = d[mb][nb][1] = 0.0;
    // my (frame, Gaussian) pairs of a round: frames warp + 8 q, Gaussian lane; my feature values: elements tid + 256 q of [32][D]
    const int gq = lane, jm = jm0 + gq, jst = (gq < ng) ? jm / M : 0;
    const double lw = (gq < ng) ? logw[jm] : 0.0;
    const uint32_t d_magic = 0xffffffffu / (uint32_t)D + 1u;      // i / D = umulhi(i, d_magic) for i < 2^16
    // shared-memory slot k of this thread: raw_u + k * SLOT; global sources as running pointers of frame t0 + warp (my first
    // pair of a round; the others are 8, 16, 24 frames on)
    constexpr uint32_t SLOT = EMM_THREADS * 8;
    const uint32_t raw_u = (uint32_t)__cvta_generic_to_shared(emm_raw) + tid * 8;
    const uint32_t z_u = (uint32_t)__cvta_generic_to_shared(&z_s[warp][0]) + lane * 8;
    auto cp8 = [](uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); };
    auto cp4 = [](uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); };
    const int64_t stepN = (int64_t)(EMM_THREADS / 32) * N, stepNM = (int64_t)(EMM_THREADS / 32) * NM;
    const char* fbase = (const char*)feats;
    const int fsz = feats_f64 ? 8 : 4;
    auto prefetch = [&](int64_t t0) {
        if (t0 < t_end) {
            const int nt = (int)min((int64_t)EMM_FT, t_end - t0);
            if (gq < ng) {
                const double* pa = g.alpha + (t0 + warp) * N + jst;
                const double* pb = g.beta + (t0 + warp) * N + jst;
                const double* pe = g.b + (t0 + warp) * N + jst;
                const double* pm = bm + (t0 + warp) * NM + jm;
#pragma unroll
                for (int q = 0; q < EMM_OQ; ++q) {
                    if (warp + (EMM_THREADS / 32) * q < nt) {
                        cp8(raw_u + (4 * q + 0) * SLOT, pa + q * stepN);
                        cp8(raw_u + (4 * q + 1) * SLOT, pb + q * stepN);
                        cp8(raw_u + (4 * q + 2) * SLOT, pe + q * stepN);
                        cp8(raw_u + (4 * q + 3) * SLOT, pm + q * stepNM);
                    }
                }
            }
            if (lane < EMM_OQ && warp + (EMM_THREADS / 32) * lane < nt) cp8(z_u, g.z + t0 + warp + (EMM_THREADS / 32) * lane);
            const char* frow = fbase + t0 * ld_feats * fsz;
#pragma unroll
            for (int q = 0; q < EMM_RAW_X; ++q) {
                const int i = tid + EMM_THREADS * q, r = (D == 1) ? i : (int)__umulhi((uint32_t)i, d_magic), dd = i - r * D;
                if (r < nt) {
                    const char* src = frow + ((int64_t)r * ld_feats + dd) * fsz;
                    if (feats_f64) cp8(raw_u + (EMM_RAW_OCC + q) * SLOT, src);
                    else cp4(raw_u + (EMM_RAW_OCC + q) * SLOT, src);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(t_begin);
    __syncthreads();
    const double* raw = emm_raw + tid;
    for (int64_t t0 = t_begin; t0 < t_end; t0 += EMM_FT) {
        const int nt = (int)min((int64_t)EMM_FT, t_end - t0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");       // this round's inputs have landed
        __syncwarp();                                              // ... those fetched by other lanes of the warp (Z) too
        // mixture occupancies for ALL t (gmmhmm.py:98-100): exp(alpha + beta - Z + log w + bm - b)
#pragma unroll
        for (int q = 0; q < EMM_OQ; ++q) {
            const int tt = warp + (EMM_THREADS / 32) * q;
            double o = 0.0;
            if (tt < nt && gq < ng) {
                const double ab = raw[(4 * q + 0) * EMM_THREADS] + raw[(4 * q + 1) * EMM_THREADS];
                const double bj = raw[(4 * q + 2) * EMM_THREADS], Z = z_s[warp][q];
                if (ab != neg_inf() && Z != neg_inf() && Z == Z && bj != neg_inf())
                    o = exp_neg_poly(ab - Z + lw + raw[(4 * q + 3) * EMM_THREADS] - bj);
            }
            occ_s[tt][gq] = o;
        }
#pragma unroll
        for (int q = 0; q < EMM_RAW_X; ++q) {
            const int i = tid + EMM_THREADS * q, r = (D == 1) ? i : (int)__umulhi((uint32_t)i, d_magic), dd = i - r * D;
            if (r < EMM_FT) {
                double xv = 0.0;                           // rows past the range: zero moments (their occupancies are zero too)
                if (r < nt) {
                    const double* slot = raw + (EMM_RAW_OCC + q) * EMM_THREADS;
                    xv = (feats_f64 ? *slot : (double)*reinterpret_cast<const float*>(slot)) - c_s[dd];
                }
                *reinterpret_cast<double2*>(&x_s[r][2 * dd]) = make_double2(xv * xv, xv);
            }
        }
        __syncthreads();                                   // (also: every lane has read its warp's z_s before the prefetch rewrites it)
        prefetch(t0 + EMM_FT);                             // into the slots just consumed
#pragma unroll
        for (int kk = 0; kk < EMM_KSTEPS; ++kk) {
            const int fr = 4 * (ks * EMM_KSTEPS + kk) + fr4;
            const double a0 = occ_s[fr][16 * gh + r8], a1 = occ_s[fr][16 * gh + 8 + r8];
            const double* xr = &x_s[fr][r8];
#pragma unroll
            for (int nb = 0; nb < EMM_COLS / 8; ++nb) {
                const double bv = xr[8 * nb];
                em_dmma884(d[0][nb][0], d[0][nb][1], a0, bv);
                em_dmma884(d[1][nb][0], d[1][nb][1], a1, bv);
            }
        }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // flush: thread = (Gaussian row r8 of block mb, dims 4 nb + fr4): d[mb][nb] = (S2', S1'); S0 sits in the constant column
    double* S0 = acc;
    double* S1 = acc + NM;
    double* S2 = S1 + (int64_t)NM * D;
    const int nb1 = (2 * D) >> 3, q1 = ((2 * D) & 7) >> 1;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        double s0v = 0.0;
#pragma unroll
        for (int nb = 0; nb < EMM_COLS / 8; ++nb)
            if (nb == nb1) s0v = d[mb][nb][0];
        const double s0 = shfl_d(s0v, (lane & ~3) | q1);
        const int gq = 16 * gh + 8 * mb + r8;
        if (gq < ng) {
            if (fr4 == q1 && s0 != 0.0) atomicAdd(S0 + jm0 + gq, s0);
#pragma unroll
            for (int nb = 0; nb < EMM_COLS / 8; ++nb) {
                const int dd = 4 * nb + fr4;
                if (dd < D) {
                    const double s2p = d[mb][nb][0], s1p = d[mb][nb][1];
                    const double c = c_s[dd], dl = mu[(int64_t)(jm0 + gq) * D + dd] - c;
                    const double s1 = fma(c, s0, s1p);
                    const double s2 = fma(dl, fma(dl, s0, -2.0 * s1p), s2p);
                    if (s1 != 0.0) atomicAdd(S1 + (int64_t)(jm0 + gq) * D + dd, s1);
                    if (s2 != 0.0) atomicAdd(S2 + (int64_t)(jm0 + gq) * D + dd, s2);
                }
            }
        }
    }
}

// discrete emissions: gamma_sum[i] += exp gamma[t,i], B_sum[i, obs[t]] += exp gamma[t,i] for t < T-1 (hmm.py:202-225)
__global__ void __launch_bounds__(256) em_discrete_kernel(EmGen g, const int32_t* __restrict__ obs, int n_symbols,
                                                          double* __restrict__ gamma_sum, double* __restrict__ B_sum) {
    const int N = g.N;
    const int64_t t0 = (int64_t)blockIdx.x * 64;
    const int nt = (int)min((int64_t)64, g.total - t0);
    for (int i = threadIdx.x; i < N; i += 256) {
        double gs = 0.0;
        for (int tt = 0; tt < nt; ++tt) {
            const int64_t t = t0 + tt;
            const double zx = g.zx[t];
            if (zx > 1e308) continue;                                // hmm.py:199: gamma[T-1] = -inf
            const double ab = g.alpha[t * N + i] + g.beta[t * N + i];
            if (ab == neg_inf()) continue;
            const double v = exp(ab - zx);
            gs += v;
            const int o = obs[t];
            if (o >= 0 && o < n_symbols && v != 0.0) atomicAdd(B_sum + (int64_t)i * n_symbols + o, v);
        }
        if (gs != 0.0) atomicAdd(gamma_sum + i, gs);
    }
}

static size_t em_general_ws_bytes(int64_t total_frames, int N) {
    return (size_t)2 * total_frames * sizeof(double) + 256 + (size_t)N * N * (2 * sizeof(int) + sizeof(double)) + 256;
}

static int em_general_prepare(EmGen& g, const double* logA, void* workspace, size_t workspace_bytes, const char* who, cudaStream_t stream) {
    const size_t need = em_general_ws_bytes(g.total, g.N);
    if (!workspace || workspace_bytes < need) {
        set_error("%s: N = %d > 32 needs a workspace of %zu bytes (got %zu)", who, g.N, need, workspace_bytes);
        return LNASR_EWORKSPACE;
    }
    char* w = reinterpret_cast<char*>(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
    g.z = reinterpret_cast<double*>(w); w += g.total * sizeof(double);
    g.zx = reinterpret_cast<double*>(w); w += g.total * sizeof(double);
    g.count = reinterpret_cast<int*>(w); w += 16;
    g.ta = reinterpret_cast<double*>(w); w += (size_t)g.N * g.N * sizeof(double);
    g.ti = reinterpret_cast<int*>(w); w += (size_t)g.N * g.N * sizeof(int);
    g.tj = reinterpret_cast<int*>(w);
    LNASR_CUDA(cudaMemsetAsync(g.count, 0, sizeof(int), stream));
    em_translist_kernel<<<g.N, 256, 0, stream>>>(logA, g.N, g.count, g.ti, g.tj, g.ta);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

// z / pi / loglik, then the transition statistics
static int em_general_common(const EmGen& g, double* xi_sum, double* pi_occ, double* loglik, double* n_utt_acc, double* gamma_out,
                             cudaStream_t stream) {
    em_z_kernel<<<(unsigned)ceil_div(g.total, 8), 256, 0, stream>>>(g, pi_occ, loglik, n_utt_acc, gamma_out);
    LNASR_LAUNCH_CHECK();
    // thread = transition (at most N*N of them; the launch covers the worst case, blocks past the count exit at once),
    // frames grid-strided so that the grid has a few waves whatever the sparsity
    const int64_t chunks = ceil_div((int64_t)g.N * g.N, 256);
    const int64_t chunks_sparse = ceil_div((int64_t)g.N * 4, 256);                 // the usual case: a few successors per state
    int64_t gx = std::max<int64_t>(1, std::min<int64_t>(g.total, (int64_t)sm_count() * 16 / std::max<int64_t>(1, chunks_sparse)));
    gx = std::min<int64_t>(gx, 65535);
    dim3 grid((unsigned)gx, (unsigned)std::min<int64_t>(chunks, 65535));
    em_xi_kernel<<<grid, 256, 0, stream>>>(g, xi_sum);
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
    LNASR_CHECK_ARG(N >= 1 && N <= 32, "fb_accumulate: the backward-fused kernels are built for 1 <= N <= 32 (got %d); larger models run "
                    "lnasr_backward_batch + lnasr_em_statistics_ws", N);
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

static int g_em_moments_simt = 0;
// test / measurement hook: 1 = the SIMT moments kernel also where the tensor-core one would run (tests compare the two)
extern "C" int lnasr_debug_set_em_moments_simt(int on) { g_em_moments_simt = on; return LNASR_OK; }

extern "C" size_t lnasr_em_statistics_workspace_bytes(int64_t total_frames, int32_t N) {
    return N <= 32 ? 256 : em_general_ws_bytes(total_frames, N);
}

extern "C" int lnasr_em_statistics_ws(const void* feats, int feats_dtype, int32_t ld_feats, const double* b, const double* bm,
                                      const double* alpha, const double* beta, const int64_t* frame_offsets, int32_t n_utt,
                                      int64_t total_frames, const double* logA, const double* logw, const double* mu, int32_t N,
                                      int32_t n_mix, int32_t dim, double* acc, double* gamma_out, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    if (N <= 32)
        return lnasr_em_statistics(feats, feats_dtype, ld_feats, b, bm, alpha, beta, frame_offsets, n_utt, total_frames, logA, logw, mu,
                                   N, n_mix, dim, acc, gamma_out, stream_);
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(feats && b && bm && alpha && beta && frame_offsets && logA && logw && mu && acc, "em_statistics_ws: null pointer");
    LNASR_CHECK_ARG(feats_dtype == LNASR_DT_F32 || feats_dtype == LNASR_DT_F64, "em_statistics_ws: feats must be float32/float64");
    LNASR_CHECK_ARG(N <= 46340 && n_mix >= 1 && dim >= 1 && dim <= 64 && ld_feats >= dim && n_utt >= 0, "em_statistics_ws: bad shape (dim <= 64)");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    EmGen g{};
    g.alpha = alpha; g.beta = beta; g.b = b; g.frame_off = frame_offsets; g.n_utt = n_utt; g.total = total_frames; g.N = N;
    int rc = em_general_prepare(g, logA, workspace, workspace_bytes, "em_statistics_ws", stream);
    if (rc) return rc;
    const int64_t NM = (int64_t)N * n_mix;
    double* tail = acc + NM * (2 * dim + 1);                           // xi_sum[N*N] | pi_occ[N] | loglik | n_utt
    rc = em_general_common(g, tail, tail + (int64_t)N * N, tail + (int64_t)N * N + N, tail + (int64_t)N * N + N + 1, gamma_out, stream);
    if (rc) return rc;
    const int64_t gtiles = ceil_div(NM, EMG_GT);
    int64_t fsplit = std::max<int64_t>(1, std::min<int64_t>(ceil_div(total_frames, EMG_FT), (int64_t)sm_count() * 8 / std::max<int64_t>(1, gtiles)));
    const int64_t per = ceil_div(ceil_div(total_frames, fsplit), EMG_FT) * EMG_FT;
    fsplit = ceil_div(total_frames, per);
    LNASR_CHECK_ARG(gtiles <= 65535, "em_statistics_ws: too many Gaussians");
    dim3 grid((unsigned)fsplit, (unsigned)gtiles);
    if (dim <= 39 && !g_em_moments_simt) {
        LNASR_CUDA(cudaFuncSetAttribute(em_moments_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMM_RAW_BYTES));
        em_moments_mma_kernel<<<grid, EMM_THREADS, EMM_RAW_BYTES, stream>>>(g, feats, feats_dtype == LNASR_DT_F64, ld_feats, bm, logw, mu, n_mix, dim, per, acc);
    } else
        em_moments_kernel<<<grid, 256, 0, stream>>>(g, feats, feats_dtype == LNASR_DT_F64, ld_feats, bm, logw, mu, n_mix, dim, per, acc);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_hmm_em_statistics_ws(const double* b, const double* alpha, const double* beta, const int32_t* obs,
                                          const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames, const double* logA,
                                          int32_t N, int32_t n_symbols, double* acc, double* gamma_out, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(b && alpha && beta && obs && frame_offsets && logA && acc, "hmm_em_statistics_ws: null pointer");
    LNASR_CHECK_ARG(N >= 1 && N <= 46340 && n_symbols >= 1 && n_utt >= 0, "hmm_em_statistics_ws: bad shape");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    EmGen g{};
    g.alpha = alpha; g.beta = beta; g.b = b; g.frame_off = frame_offsets; g.n_utt = n_utt; g.total = total_frames; g.N = N;
    int rc = em_general_prepare(g, logA, workspace, workspace_bytes, "hmm_em_statistics_ws", stream);
    if (rc) return rc;
    // xi_sum[N*N] | gamma_sum[N] | B_sum[N*n_symbols] | pi_occ[N] | loglik | n_utt
    double* gs = acc + (int64_t)N * N;
    double* Bs = gs + N;
    double* pio = Bs + (int64_t)N * n_symbols;
    rc = em_general_common(g, acc, pio, pio + N, pio + N + 1, gamma_out, stream);
    if (rc) return rc;
    em_discrete_kernel<<<(unsigned)ceil_div(total_frames, 64), 256, 0, stream>>>(g, obs, n_symbols, gs, Bs);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
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
    LNASR_CHECK_ARG(N >= 1 && N <= 32, "hmm_fb_accumulate: the backward-fused kernel is built for 1 <= N <= 32 (got %d); larger models run "
                    "lnasr_backward_batch + lnasr_hmm_em_statistics_ws", N);
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
