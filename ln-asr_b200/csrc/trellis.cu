// trellis.cu -- HMM log-trellis: forward, backward, Viterbi (+ backtrace) for ragged utterance batches
// scored against a bank of models (kernel (3) of the north star).
//
// Replaces, from /root/reference/lnasr/hmm.py: _forward (63-100), _backward (102-131), _viterbi (133-172),
// calc_prob (268-274), decode (276-283).
//
// Four kernels, all persistent over the time axis of their utterance with the trellis column held on chip in float64:
//   * small   (N <= 32): a group of G lanes (G = N for N <= 8, else 16 / 32) owns one (utterance, model) pair, lane =
//     state; the predecessor column of log A lives in registers, the previous trellis column is exchanged with warp
//     shuffles, max / log-sum-exp over predecessors are register loops, the backtrace is fused.
//   * band    (32 < N <= 2048, Viterbi decode of left-to-right banks: every finite transition i -> j has 0 <= j - i <= 2):
//     one CTA per pair, a thread owns 4 or 8 CONSECUTIVE states with their values and transition weights in registers and
//     gets the two values below its range through a shared-memory halo.
//   * reg     (32 < N <= 2048, <= 4 finite predecessors / successors per state, any topology, all three recursions): one CTA
//     per pair, the transition lists of a thread's states in registers, the column double-buffered in shared memory.
//   * general (any N): as reg, with the ELL lists of the FINITE entries of log A read from memory.
//   The last three read emission rows with cp.async several steps ahead and use one __syncthreads per frame.
// Backpointers go to HBM as uint16 state indices; the backtrace (hmm.py:168-171) is a separate
// one-thread-per-pair pointer chase so that it runs for all pairs concurrently.
//
// Tie rule: np.argmax takes the first maximum (hmm.py:166,169) -> strict '>' while scanning ascending i.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace lnasr {

enum { MODE_VITERBI = 0, MODE_FORWARD = 1, MODE_BACKWARD = 2 };

// ------------------------------------------------------------------------------------------------
// transition packing
__global__ void trans_degree_kernel(const double* __restrict__ logA, int n_models, int N, int* __restrict__ deg) {
    // one thread per (model, state): counts finite entries of its column and of its row
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_models * N) return;
    const int k = (int)(idx / N), s = (int)(idx % N);
    const double* A = logA + (int64_t)k * N * N;
    int cin = 0, cout = 0, fwd = 0, back = 0;
    for (int i = 0; i < N; ++i) {
        const bool in = A[(int64_t)i * N + s] != neg_inf();          // transition i -> s
        cin += in;
        cout += (A[(int64_t)s * N + i] != neg_inf());
        if (in) { fwd = max(fwd, s - i); back = max(back, i - s); }
    }
    atomicMax(deg + 0, cin);
    atomicMax(deg + 1, cout);
    atomicMax(deg + 2, fwd);      // largest j - i of a finite transition i -> j (0 if none goes forward)
    atomicMax(deg + 3, back);     // largest i - j (0 for a left-to-right model)
}

__global__ void trans_pack_kernel(const double* __restrict__ logA, int n_models, int N, int max_in, int max_out,
                                  int* __restrict__ pred_idx, double* __restrict__ pred_val,
                                  int* __restrict__ succ_idx, double* __restrict__ succ_val) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_models * N) return;
    const int k = (int)(idx / N), s = (int)(idx % N);
    const double* A = logA + (int64_t)k * N * N;
    int c = 0;
    for (int i = 0; i < N && c < max_in; ++i) {
        const double a = A[(int64_t)i * N + s];
        if (a != neg_inf()) {
            pred_idx[((int64_t)k * max_in + c) * N + s] = i;
            pred_val[((int64_t)k * max_in + c) * N + s] = a;
            ++c;
        }
    }
    for (; c < max_in; ++c) {
        pred_idx[((int64_t)k * max_in + c) * N + s] = 0;
        pred_val[((int64_t)k * max_in + c) * N + s] = neg_inf();
    }
    c = 0;
    for (int j = 0; j < N && c < max_out; ++j) {
        const double a = A[(int64_t)s * N + j];
        if (a != neg_inf()) {
            succ_idx[((int64_t)k * max_out + c) * N + s] = j;
            succ_val[((int64_t)k * max_out + c) * N + s] = a;
            ++c;
        }
    }
    for (; c < max_out; ++c) {
        succ_idx[((int64_t)k * max_out + c) * N + s] = 0;
        succ_val[((int64_t)k * max_out + c) * N + s] = neg_inf();
    }
}

// ------------------------------------------------------------------------------------------------
// small-N kernel
struct TrellisArgs {
    const void* b; int64_t ld_b;
    const int64_t* frame_off; int n_utt; int n_models; int N;
    const double* logA; const double* logpi;
    double* full_out;        // v / alpha / beta, [T_total, n_models*N] or null
    uint32_t* bt_out;        // viterbi only, optional
    uint16_t* bt_ws;         // viterbi only, required when a path is wanted
    double* score_out;       // [n_utt, n_models]: max v[T-1] (viterbi) / lse alpha[T-1] (forward); unused for backward
    int32_t* last_state;     // [n_utt*n_models], viterbi only
    uint32_t* path_out;      // [n_models, total_frames]; the small kernel backtraces in place when this is set
    int64_t total_frames;
    int max_in, max_out;
    const int* pred_idx; const double* pred_val; const int* succ_idx; const double* succ_val;
};

constexpr int TS_WARPS = 4;
constexpr int TS_BT = 16;    // backpointer rows fetched per round of the in-kernel backtrace

// G lanes per (utterance, model) pair.  G = N exactly for N <= 8 (so 5-state word models pack 6 pairs into a warp and the
// predecessor loops run N, not 8, iterations), else 16 or 32 with the surplus lanes idle.  PF = emission rows in flight per
// lane: a trellis step is ~0.1 us of shuffles and adds, an HBM / L2 miss ~0.8 us.
template <int G, typename BT, int MODE, int PF>
__global__ void __launch_bounds__(TS_WARPS * 32)
trellis_small_kernel(TrellisArgs a) {
    // log-sum-exp modes use the table-driven exp / log of common.cuh (the recursion is one serial chain per pair: the
    // latency of exp() and log() IS the run time)
    __shared__ __align__(16) double2 log_tab[MODE == MODE_VITERBI ? 1 : 128];
    __shared__ double exp_tab[MODE == MODE_VITERBI ? 1 : 64];
    if (MODE != MODE_VITERBI) {
        fill_log_exp_tables(log_tab, exp_tab, threadIdx.x, TS_WARPS * 32);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PPW = 32 / G;                                   // pairs per warp
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const int slot = lane / G;
    const int gbase = slot * G;                                   // first lane of my group
    const int64_t p = ((int64_t)blockIdx.x * TS_WARPS + warp) * PPW + slot;
    const int j = lane - gbase;
    const int N = a.N;
    const bool pair_ok = slot < PPW && p < n_pairs;
    const bool active = pair_ok && j < N;
    const int u = pair_ok ? (int)(p / a.n_models) : 0;
    const int k = pair_ok ? (int)(p % a.n_models) : 0;
    const int64_t f0 = pair_ok ? __ldg(a.frame_off + u) : 0;
    const int T = pair_ok ? (int)(__ldg(a.frame_off + u + 1) - f0) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    if (Tmax == 0) return;
    auto grp = [&](double v, int q) { return __shfl_sync(0xffffffffu, v, gbase + q); };   // value of state q of my pair

    // transition column (forward/viterbi: A[:, j]) or row (backward: A[j, :]) in registers
    double tr[G];
    const double* A = a.logA + (int64_t)k * N * N;
#pragma unroll
    for (int i = 0; i < G; ++i) {
        double v = neg_inf();
        if (active && i < N) v = (MODE == MODE_BACKWARD) ? __ldg(A + (int64_t)j * N + i) : __ldg(A + (int64_t)i * N + j);
        tr[i] = v;
    }
    const BT* bcol = (const BT*)a.b + (int64_t)k * N + j;
    const int64_t ncol = (int64_t)a.n_models * N;
    const int64_t col = (int64_t)k * N + j;
    // emission row used at step s: forward/viterbi -> frame s ; backward -> frame T-s (s>=1), frame T-1 at s=0 (unused)
    auto emis_row = [&](int s) -> int64_t {
        if (MODE == MODE_BACKWARD) return f0 + (s == 0 ? T - 1 : T - s);
        return f0 + s;
    };
    BT bq[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) bq[i] = (active && i < T) ? __ldg(bcol + emis_row(i) * a.ld_b) : BT(0);
    const double pi_j = (MODE != MODE_BACKWARD && active) ? __ldg(a.logpi + (int64_t)k * N + j) : neg_inf();
    // running pointers (one add per step instead of 64-bit multiplies): emission row of step s+PF, output row of step s
    const int64_t bstep = (MODE == MODE_BACKWARD) ? -a.ld_b : a.ld_b;
    const BT* bnext = bcol + emis_row(PF) * a.ld_b;
    const int64_t ostep = (MODE == MODE_BACKWARD) ? -ncol : ncol;
    const int64_t o0 = ((MODE == MODE_BACKWARD) ? (f0 + T - 1) : f0) * ncol + col;
    double* pfull = a.full_out ? a.full_out + o0 : nullptr;
    uint32_t* pbt32 = (MODE == MODE_VITERBI && a.bt_out) ? a.bt_out + o0 : nullptr;
    uint16_t* pbt16 = (MODE == MODE_VITERBI && a.bt_ws) ? a.bt_ws + o0 : nullptr;
    const bool slow_out = (pfull != nullptr) || (pbt32 != nullptr);      // API-parity outputs (v / alpha / beta, uint32 bt)

    double cur = neg_inf();       // trellis value of my state
    for (int s0 = 0; s0 < Tmax; s0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int s = s0 + i;
            const double e = (double)bq[i];
            bq[i] = (active && s + PF < T) ? __ldg(bnext) : BT(0);
            bnext += bstep;
            const bool live = active && s < T;
            double nv;
            int arg = 0;
            if (MODE == MODE_BACKWARD) {
                // beta[T-1] = 0 ; beta[t,i] = lse_j (beta[t+1,j] + A[i,j] + b[j,t+1])     (hmm.py:126-130)
                const double uj = cur + e;                       // beta[t+1,j] + b[t+1,j], my j
                double cand[G];
                double mx = neg_inf();
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    cand[q] = grp(uj, q) + tr[q];
                    mx = fmax(mx, cand[q]);
                }
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < G; ++q) sum += fast_exp_neg(cand[q] - mx, exp_tab);          // -inf (and -inf - -inf) -> 0
                nv = (s == 0) ? 0.0 : ((mx == neg_inf()) ? mx : mx + fast_log(sum, log_tab));
            } else if (MODE == MODE_FORWARD) {
                // alpha[0] = pi + b[:,0] ; alpha[t,j] = lse_i (alpha[t-1,i] + A[i,j]) + b[j,t]   (hmm.py:95-99)
                double cand[G];
                double mx = neg_inf();
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    cand[q] = grp(cur, q) + tr[q];
                    mx = fmax(mx, cand[q]);
                }
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < G; ++q) sum += fast_exp_neg(cand[q] - mx, exp_tab);
                const double l = (mx == neg_inf()) ? mx : mx + fast_log(sum, log_tab);
                nv = (s == 0) ? pi_j + e : l + e;
            } else {
                // v[0] = pi + b[:,0] ; v[t,j] = max_i (v[t-1,i] + A[i,j]) + b[j,t] ; bt = FIRST argmax (hmm.py:160-166):
                // maximum by a max tree, then the lowest q that attains it (all -inf -> 0, like np.argmax)
                double cand[G];
#pragma unroll
                for (int q = 0; q < G; ++q) cand[q] = grp(cur, q) + tr[q];
                double best = cand[0];
#pragma unroll
                for (int q = 1; q < G; ++q) best = fmax(best, cand[q]);
                arg = G - 1;
#pragma unroll
                for (int q = G - 2; q >= 0; --q) arg = (cand[q] == best) ? q : arg;
                nv = (s == 0) ? pi_j + e : best + e;
                if (s == 0) arg = 0;
            }
            if (live) {
                cur = nv;
                if (MODE == MODE_VITERBI && pbt16) *pbt16 = (uint16_t)arg;
                if (slow_out) {
                    if (pfull) *pfull = nv;
                    if (MODE == MODE_VITERBI && pbt32) *pbt32 = (uint32_t)arg;
                }
            }
            if (MODE == MODE_VITERBI && pbt16) pbt16 += ostep;
            if (slow_out) {
                if (pfull) pfull += ostep;
                if (MODE == MODE_VITERBI && pbt32) pbt32 += ostep;
            }
        }
    }
    // termination: max / argmax (viterbi, hmm.py:169) or lse (forward, hmm.py:274) over the last column
    if (MODE == MODE_VITERBI) {
        double best = neg_inf();
        int arg = 0;
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const double c = grp(cur, q);
            if (q < N && c > best) { best = c; arg = q; }
        }
        if (pair_ok && j == 0) {
            if (a.score_out) a.score_out[p] = best;
            if (a.last_state) a.last_state[p] = arg;
        }
        // Backtrace (hmm.py:168-171), fused: path[T-1] = argmax v[T-1]; path[t] = bt[t+1, path[t+1]].  Lane j re-reads
        // ITS OWN column of the uint16 backpointers (its own stores, so program order makes them visible), TS_BT rows per
        // round and all in flight together; the pointer chase itself then runs on registers through shuffles instead
        // of one dependent global load per frame.
        if (a.path_out && a.bt_ws) {
            uint32_t* out = a.path_out + (int64_t)k * a.total_frames + f0;
            int st = arg;                                          // state at frame t_hi (same in every lane of the group)
            if (pair_ok && j == 0 && T > 0) out[T - 1] = (uint32_t)st;
            // rows t+1 = T-1 .. 1 are consumed; round r covers rows hi .. hi-TS_BT+1 with hi = Tmax-1 - r*TS_BT
            for (int hi = Tmax - 1; hi >= 1; hi -= TS_BT) {
                uint16_t q[TS_BT];
                const uint16_t* bp = a.bt_ws + (f0 + hi) * ncol + col;
#pragma unroll
                for (int i = 0; i < TS_BT; ++i) {
                    const int row = hi - i;
                    q[i] = (active && row >= 1 && row < T) ? *bp : (uint16_t)0;
                    bp -= ncol;
                }
                int mine = 0;                                      // state of frame hi-1-j, kept by lane j for a coalesced store
#pragma unroll
                for (int i = 0; i < TS_BT; ++i) {
                    const int row = hi - i;
                    const int nx = __shfl_sync(0xffffffffu, (int)q[i], gbase + st);
                    if (row >= 1 && row < T) st = nx;              // uniform inside the group
                    if ((i % G) == j) mine = st;
                    if ((i % G) == G - 1 || i == TS_BT - 1) {      // flush G results: lane j holds frame (hi-1) - (i - i%G) - j
                        const int t = hi - 1 - (i - (i % G)) - j;
                        if (pair_ok && (i - (i % G)) + j <= i && t >= 0 && t < T - 1) out[t] = (uint32_t)mine;
                    }
                }
            }
        }
    } else if (MODE == MODE_FORWARD) {
        double mx = neg_inf();
        double cand[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            cand[q] = (q < N) ? grp(cur, q) : neg_inf();
            mx = fmax(mx, cand[q]);
        }
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < G; ++q) sum += fast_exp_neg(cand[q] - mx, exp_tab);
        if (pair_ok && j == 0 && a.score_out) a.score_out[p] = (mx == neg_inf()) ? mx : mx + fast_log(sum, log_tab);
    }
}

// ------------------------------------------------------------------------------------------------
// general kernel: one CTA per (utterance, model)
constexpr int TG_THREADS = 256;

// termination over the last column prev[0..N): max / first argmax (hmm.py:169) or log-sum-exp (hmm.py:274)
template <int MODE>
__device__ __forceinline__ void trellis_terminate(const double* prev, int N, const TrellisArgs& a, int64_t p, double* red_v, int* red_i) {
    if (MODE == MODE_BACKWARD) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE == MODE_VITERBI) {
        double best = neg_inf();
        int arg = 0x7fffffff;
        for (int j = threadIdx.x; j < N; j += TG_THREADS) {
            const double c = prev[j];
            if (c > best) { best = c; arg = j; }      // ascending j per thread -> first max per thread
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = shfl_d(best, lane ^ o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) { red_v[warp] = best; red_i[warp] = arg; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < TG_THREADS / 32; ++w)
                if (red_v[w] > best || (red_v[w] == best && red_i[w] < arg)) { best = red_v[w]; arg = red_i[w]; }
            if (arg == 0x7fffffff) arg = 0;           // all -inf: np.argmax gives 0
            if (a.score_out) a.score_out[p] = best;
            if (a.last_state) a.last_state[p] = arg;
        }
    } else {
        double mx = neg_inf();
        for (int j = threadIdx.x; j < N; j += TG_THREADS) mx = fmax(mx, prev[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_d(mx, lane ^ o));
        if (lane == 0) red_v[warp] = mx;
        __syncthreads();
        mx = red_v[0];
        for (int w = 1; w < TG_THREADS / 32; ++w) mx = fmax(mx, red_v[w]);
        __syncthreads();
        double sum = 0.0;
        if (mx != neg_inf())
            for (int j = threadIdx.x; j < N; j += TG_THREADS) sum += (prev[j] == neg_inf()) ? 0.0 : exp(prev[j] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += shfl_d(sum, lane ^ o);
        if (lane == 0) red_v[warp] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < TG_THREADS / 32; ++w) sum += red_v[w];
            if (a.score_out) a.score_out[p] = (mx == neg_inf()) ? mx : mx + log(sum);
        }
    }
}

// Sparse-transition kernel for 32 < N <= 256*SPT with at most 4 finite predecessors (successors) per state: the
// usual case for large acoustic models (left-to-right: self, next, skip).  One CTA per (utterance, model); every
// thread OWNS states tid + 256 i for the whole utterance and keeps their transition lists in REGISTERS, so a time
// step touches HBM only for the emission row (read, prefetched one step ahead) and the backpointer row (write):
// 8N bytes per frame, the roofline of SURVEY.md 8(d).  The trellis column is double buffered in shared memory.
constexpr int TR_DEG = 4;

__device__ __forceinline__ void tr_cp_async(void* smem_dst, const void* gmem_src, int bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

// PF emission rows are kept in flight with cp.async (global -> shared, no registers): at ~1 us of HBM latency and
// ~0.3 us per time step a one-step register prefetch leaves the kernel latency-bound (measured: 11 % of HBM peak).
template <int SPT, typename BT, int MODE, int PF>
__global__ void __launch_bounds__(TG_THREADS, 2)
trellis_reg_kernel(TrellisArgs a) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    const int N = a.N;
    double* prev = (double*)tg_smem;
    double* next = prev + N;
    const int n_pad = (N + 3) & ~3;
    BT* erow = reinterpret_cast<BT*>(next + N);            // [PF][n_pad]
    __shared__ double red_v[TG_THREADS / 32];
    __shared__ int red_i[TG_THREADS / 32];
    const int64_t p = blockIdx.x;
    const int u = (int)(p / a.n_models), k = (int)(p % a.n_models);
    const int64_t f0 = __ldg(a.frame_off + u);
    const int T = (int)(__ldg(a.frame_off + u + 1) - f0);
    if (T == 0) return;
    const int64_t ncol = (int64_t)a.n_models * N;
    const int64_t cbase = (int64_t)k * N;
    const BT* b = (const BT*)a.b;
    const int deg = (MODE == MODE_BACKWARD) ? a.max_out : a.max_in;
    const int* eidx = ((MODE == MODE_BACKWARD) ? a.succ_idx : a.pred_idx) + (int64_t)k * deg * N;
    const double* eval = ((MODE == MODE_BACKWARD) ? a.succ_val : a.pred_val) + (int64_t)k * deg * N;
    // emission row used at step s
    auto erow_of = [&](int s) -> int64_t { return (MODE == MODE_BACKWARD) ? (f0 + T - 1 - s) : (f0 + s); };
    const bool vec16 = (((size_t)a.ld_b * sizeof(BT)) % 16 == 0) && (((size_t)cbase * sizeof(BT)) % 16 == 0) &&
                       (((size_t)N * sizeof(BT)) % 16 == 0) && (((uintptr_t)b) % 16 == 0);
    auto issue_row = [&](int s) {
        if (s < T) {
            const BT* src = b + erow_of(s) * a.ld_b + cbase;
            BT* dst = erow + (size_t)(s % PF) * n_pad;
            if (vec16) {
                constexpr int EPV = 16 / (int)sizeof(BT);
                for (int i = threadIdx.x * EPV; i < N; i += TG_THREADS * EPV) tr_cp_async(dst + i, src + i, 16);
            } else {
                for (int i = threadIdx.x; i < N; i += TG_THREADS) tr_cp_async(dst + i, src + i, (int)sizeof(BT));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < PF - 1; ++s) issue_row(s);
    int idx[SPT][TR_DEG];
    double val[SPT][TR_DEG];
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        const int j = threadIdx.x + TG_THREADS * i;
#pragma unroll
        for (int d = 0; d < TR_DEG; ++d) {
            const bool ok = j < N && d < deg;
            idx[i][d] = ok ? __ldg(eidx + (int64_t)d * N + j) : 0;
            val[i][d] = ok ? __ldg(eval + (int64_t)d * N + j) : neg_inf();
        }
    }
    // Viterbi decode without the API-parity outputs (v, uint32 bt) is the hot case (config 4): running pointers instead of
    // per-state 64-bit index arithmetic, no output tests inside the state loop
    const bool lean = (MODE == MODE_VITERBI) && !a.full_out && !a.bt_out && a.bt_ws;
    uint16_t* btp = lean ? a.bt_ws + f0 * ncol + cbase + threadIdx.x : nullptr;
    for (int s = 0; s < T; ++s) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");     // row s has landed (my copies) ...
        __syncthreads();                                                       // ... and everyone's; step s-1's column is complete
        issue_row(s + PF - 1);                                                 // its buffer was last read in step s-1
        const BT* er = erow + (size_t)(s % PF) * n_pad;
        const int64_t row = erow_of(s);
        if (lean && s > 0) {
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                const int j = threadIdx.x + TG_THREADS * i;
                if (j < N) {
                    double best = prev[idx[i][0]] + val[i][0];
                    int arg = idx[i][0];
#pragma unroll
                    for (int d = 1; d < TR_DEG; ++d) {
                        const double c = prev[idx[i][d]] + val[i][d];
                        if (c > best) { best = c; arg = idx[i][d]; }            // ascending predecessor index: first max
                    }
                    if (best == neg_inf()) arg = 0;                             // no finite predecessor: np.argmax gives 0
                    next[j] = best + (double)er[j];
                    btp[TG_THREADS * i] = (uint16_t)arg;
                }
            }
            btp += ncol;
            double* t = prev; prev = next; next = t;
            continue;
        }
        if (lean) btp += ncol;                                                 // s == 0 goes through the general code below
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const int j = threadIdx.x + TG_THREADS * i;
            if (j < N) {
                const double e = (double)er[j];
                if (s == 0) {
                    double v;
                    if (MODE == MODE_BACKWARD) {
                        if (a.full_out) a.full_out[row * ncol + cbase + j] = 0.0;     // beta[T-1] = 0
                        v = e;                                                        // buffer holds beta[t] + b[t]
                    } else {
                        v = __ldg(a.logpi + cbase + j) + e;
                        if (a.full_out) a.full_out[row * ncol + cbase + j] = v;
                        if (MODE == MODE_VITERBI) {
                            if (a.bt_ws) a.bt_ws[row * ncol + cbase + j] = 0;
                            if (a.bt_out) a.bt_out[row * ncol + cbase + j] = 0;
                        }
                    }
                    next[j] = v;
                    continue;
                }
                double c[TR_DEG];
#pragma unroll
                for (int d = 0; d < TR_DEG; ++d) c[d] = prev[idx[i][d]] + val[i][d];
                if (MODE == MODE_VITERBI) {
                    double best = neg_inf();
                    int arg = 0;
#pragma unroll
                    for (int d = 0; d < TR_DEG; ++d)
                        if (c[d] > best) { best = c[d]; arg = idx[i][d]; }        // ascending predecessor index: first max
                    const double nv = best + e;
                    if (a.bt_ws) a.bt_ws[row * ncol + cbase + j] = (uint16_t)arg;
                    if (a.bt_out) a.bt_out[row * ncol + cbase + j] = (uint32_t)arg;
                    if (a.full_out) a.full_out[row * ncol + cbase + j] = nv;
                    next[j] = nv;
                } else {
                    double mx = fmax(fmax(c[0], c[1]), fmax(c[2], c[3]));
                    double l = mx;
                    if (mx != neg_inf()) {
                        double sum = 0.0;
#pragma unroll
                        for (int d = 0; d < TR_DEG; ++d) sum += exp(c[d] - mx);     // exp(-inf) = 0 for the padding
                        l = mx + log(sum);
                    }
                    if (MODE == MODE_FORWARD) {
                        const double nv = l + e;
                        if (a.full_out) a.full_out[row * ncol + cbase + j] = nv;
                        next[j] = nv;
                    } else {
                        if (a.full_out) a.full_out[row * ncol + cbase + j] = l;    // beta[t, j]
                        next[j] = l + e;                                            // beta[t] + b[t] for the next step
                    }
                }
            }
        }
        double* t = prev; prev = next; next = t;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    trellis_terminate<MODE>(prev, N, a, p, red_v, red_i);
}

// Banded Viterbi for left-to-right models (every finite transition i -> j has 0 <= j - i <= 2: self loop, next, skip --
// the usual topology of large acoustic models, bank->band != 0).  A thread owns SPT CONSECUTIVE states and keeps their
// trellis values in registers across time steps; the only values it needs from elsewhere are the two states below its
// range, which the neighbouring thread publishes through a double-buffered shared-memory halo.  Per step and thread that
// is one 16-byte shared load + store instead of 4 x SPT scattered 8-byte loads, one 16-byte backpointer store instead of
// SPT 2-byte ones, and one barrier.  Ties resolve as in hmm.py:160-164 (np.argmax: lowest predecessor index).
template <int SPT, typename BT, int PF>
__global__ void __launch_bounds__(TG_THREADS, 2)
trellis_band_kernel(TrellisArgs a) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    const int N = a.N;
    const int n_pad = (N + 3) & ~3;
    double* col = (double*)tg_smem;                               // [N] the last column, for the termination
    double2* halo = reinterpret_cast<double2*>(col + n_pad);       // [2][TG_THREADS]: (v[j0 + SPT - 2], v[j0 + SPT - 1]) of every thread
    BT* erow = reinterpret_cast<BT*>(halo + 2 * TG_THREADS);       // [PF][n_pad]
    __shared__ double red_v[TG_THREADS / 32];
    __shared__ int red_i[TG_THREADS / 32];
    const int64_t p = blockIdx.x;
    const int u = (int)(p / a.n_models), k = (int)(p % a.n_models);
    const int64_t f0 = __ldg(a.frame_off + u);
    const int T = (int)(__ldg(a.frame_off + u + 1) - f0);
    if (T == 0) return;
    const int64_t ncol = (int64_t)a.n_models * N;
    const int64_t cbase = (int64_t)k * N;
    const BT* b = (const BT*)a.b;
    const int tid = threadIdx.x, j0 = tid * SPT;
    const bool vec16 = (((size_t)a.ld_b * sizeof(BT)) % 16 == 0) && (((size_t)cbase * sizeof(BT)) % 16 == 0) &&
                       (((size_t)N * sizeof(BT)) % 16 == 0) && (((uintptr_t)b) % 16 == 0);
    auto issue_row = [&](int s) {
        if (s < T) {
            const BT* src = b + (f0 + s) * a.ld_b + cbase;
            BT* dst = erow + (size_t)(s % PF) * n_pad;
            if (vec16) {
                constexpr int EPV = 16 / (int)sizeof(BT);
                for (int i = tid * EPV; i < N; i += TG_THREADS * EPV) tr_cp_async(dst + i, src + i, 16);
            } else {
                for (int i = tid; i < N; i += TG_THREADS) tr_cp_async(dst + i, src + i, (int)sizeof(BT));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < PF - 1; ++s) issue_row(s);
    // transitions into my states by distance: w[i][d] = log A[j - d, j], from the ELL predecessor lists
    double w[SPT][3];
    {
        const int deg = a.max_in;
        const int* eidx = a.pred_idx + (int64_t)k * deg * N;
        const double* eval = a.pred_val + (int64_t)k * deg * N;
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            w[i][0] = w[i][1] = w[i][2] = neg_inf();
            const int j = j0 + i;
            if (j < N)
                for (int d = 0; d < deg; ++d) {
                    const double val = __ldg(eval + (int64_t)d * N + j);
                    const int dist = j - __ldg(eidx + (int64_t)d * N + j);
                    if (val != neg_inf()) {
                        if (dist == 0) w[i][0] = val;
                        else if (dist == 1) w[i][1] = val;
                        else w[i][2] = val;
                    }
                }
        }
    }
    double v[SPT];
    // backpointer row segment of this thread: SPT consecutive uint16 -- one 16-byte store when SPT = 8 and aligned
    uint16_t* btp = a.bt_ws ? a.bt_ws + f0 * ncol + cbase + j0 : nullptr;
    const bool bt_vec = (SPT == 8) && a.bt_ws && ((ncol * sizeof(uint16_t)) % 16 == 0) && ((cbase * sizeof(uint16_t)) % 16 == 0) &&
                        (((uintptr_t)a.bt_ws) % 16 == 0) && (j0 + SPT <= N);
    for (int s = 0; s < T; ++s) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");     // row s has landed (my copies) ...
        __syncthreads();                                                       // ... and everyone's; the halo of step s-1 is complete
        issue_row(s + PF - 1);                                                 // its buffer was last read in step s-1
        const BT* er = erow + (size_t)(s % PF) * n_pad + j0;
        double e[SPT];
        if (j0 + SPT <= n_pad) {
#pragma unroll
            for (int i = 0; i < SPT; ++i) e[i] = (double)er[i];
        } else {
#pragma unroll
            for (int i = 0; i < SPT; ++i) e[i] = (j0 + i < N) ? (double)er[i] : 0.0;
        }
        uint32_t arg[SPT];
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                v[i] = (j0 + i < N) ? __ldg(a.logpi + cbase + j0 + i) + e[i] : neg_inf();
                arg[i] = 0;
            }
        } else {
            const double2 h = (tid > 0) ? halo[((s - 1) & 1) * TG_THREADS + tid - 1] : make_double2(neg_inf(), neg_inf());
            // descending i: v[i] is overwritten after the states above it have read it
#pragma unroll
            for (int i = SPT - 1; i >= 0; --i) {
                const double p2 = (i >= 2) ? v[i - 2] : (i == 1 ? h.y : h.x);
                const double p1 = (i >= 1) ? v[i - 1] : h.y;
                double best = p2 + w[i][2];
                int dist = 2;
                const double c1 = p1 + w[i][1];
                if (c1 > best) { best = c1; dist = 1; }
                const double c0 = v[i] + w[i][0];
                if (c0 > best) { best = c0; dist = 0; }
                arg[i] = (best == neg_inf()) ? 0u : (uint32_t)(j0 + i - dist);  // no finite predecessor: np.argmax gives 0
                v[i] = best + e[i];
            }
        }
        halo[(s & 1) * TG_THREADS + tid] = make_double2(v[SPT - 2], v[SPT - 1]);
        if (btp) {
            if (bt_vec) {
                uint4 pk;
                pk.x = arg[0] | (arg[1] << 16); pk.y = arg[2] | (arg[3] << 16);
                pk.z = arg[4] | (arg[5] << 16); pk.w = arg[6] | (arg[7] << 16);
                *reinterpret_cast<uint4*>(btp) = pk;
            } else {
#pragma unroll
                for (int i = 0; i < SPT; ++i)
                    if (j0 + i < N) btp[i] = (uint16_t)arg[i];
            }
            btp += ncol;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < SPT; ++i)
        if (j0 + i < N) col[j0 + i] = v[i];
    __syncthreads();
    trellis_terminate<MODE_VITERBI>(col, N, a, p, red_v, red_i);
}

// ------------------------------------------------------------------------------------------------
// Banded Viterbi, second form: the decode kernel of large left-to-right banks (config 4).  Same ownership as
// trellis_band_kernel (a thread keeps SPT consecutive states, their trellis values and their three transition weights in
// registers), with everything that made a time step slow taken out of it:
//   * NO block barrier in the time loop.  A thread needs the two values below its range: inside a warp they come by
//     __shfl_up, across warps through a small shared-memory ring guarded by mbarrier pairs (full / empty per slot), so warp
//     w runs one step behind warp w-1 in a skewed pipeline and only ever waits for its neighbour (bar.sync made all eight
//     warps wait for the slowest one every frame).
//   * emission rows are PRIVATE: every thread cp.asyncs exactly the SPT values it will consume (PF rows ahead), so
//     cp.async.wait_group alone makes them visible -- no barrier, and no second pass over shared memory.
//   * backpointers are 2-BIT distance codes (0 / 1 / 2 = j - i of the chosen predecessor): one uint16 per thread and frame
//     (SPT = 8) instead of eight -- 0.5 kB per frame for 2000 states instead of 4 kB.
//   * the backtrace (hmm.py:168-171) is FUSED: when the forward pass is done the CTA pulls its own backpointer rows back
//     (L2-resident, 64 rows at a time) into the shared memory the emission ring occupied, one thread chases the pointers
//     there (a shared-memory load per frame instead of a dependent global load), and the path leaves as coalesced rows.
// Needs: band topology, N % SPT == 0, 16-byte aligned emission rows (else trellis_band_kernel + backtrace_kernel run).
constexpr int TB2_RING = 8;      // halo ring depth: how far a warp may run ahead of its successor
constexpr int TB2_CH = 64;       // backpointer rows per backtrace chunk

// mbarrier hand-off of the halo between neighbouring warps: arrive has release, try_wait acquire semantics for the shared
// memory written before / read after, and the wait is a hardware-suspended SYNCS.PHASECHK instead of a polling loop
__device__ __forceinline__ void tb2_bar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tb2_bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tb2_bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

template <int SPT, typename BT, int PF>
__global__ void __launch_bounds__(TG_THREADS, 2)
trellis_band2_kernel(TrellisArgs a, int bt_pitch) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    constexpr int EPC = 16 / (int)sizeof(BT);                 // emission values per 16-byte chunk
    constexpr int CPT = SPT / EPC;                             // chunks per thread and row
    constexpr int NW = TG_THREADS / 32;
    typedef typename std::conditional<SPT == 8, uint16_t, uint8_t>::type CodeT;      // SPT 2-bit codes
    static_assert(SPT % EPC == 0 && CPT >= 1, "a thread's emission values must be whole 16-byte chunks");
    static_assert((PF & (PF - 1)) == 0 && (TB2_RING & (TB2_RING - 1)) == 0, "ring sizes must be powers of two");
    // shared memory: emission ring [PF][CPT][TG_THREADS] 16-byte chunks (re-used by the termination column and the backtrace)
    uint4* ering = reinterpret_cast<uint4*>(tg_smem);
    __shared__ double2 halo[TB2_RING][NW];
    __shared__ __align__(8) uint64_t h_full[TB2_RING][NW], h_empty[TB2_RING][NW];    // [slot][producer warp]
    __shared__ double red_v[NW];
    __shared__ int red_i[NW];
    __shared__ int s_last;
    __shared__ double s_best;
    const int N = a.N;
    const int64_t p = blockIdx.x;
    const int u = (int)(p / a.n_models), k = (int)(p % a.n_models);
    const int64_t f0 = __ldg(a.frame_off + u);
    const int T = (int)(__ldg(a.frame_off + u + 1) - f0);
    if (T == 0) return;
    const int64_t cbase = (int64_t)k * N;
    const BT* b = (const BT*)a.b;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, j0 = tid * SPT;
    const bool own = j0 < N;                                   // N % SPT == 0: an owning thread owns SPT states
    const int n_wp = ((N / SPT) + 31) >> 5;                    // warps that own states
    if (tid < TB2_RING * NW) {
        tb2_bar_init(&h_full[0][0] + tid);
        tb2_bar_init(&h_empty[0][0] + tid);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // backpointer rows: row (f0 + s) * n_models + k, bt_pitch bytes, thread t's SPT codes at t * sizeof(CodeT)
    uint8_t* bt_rows = a.bt_ws ? reinterpret_cast<uint8_t*>(a.bt_ws) + ((int64_t)f0 * a.n_models + k) * bt_pitch : nullptr;
    const int64_t bt_step = (int64_t)a.n_models * bt_pitch;
    double v[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) v[i] = neg_inf();
    if (wp < n_wp) {
        // running pointers: global emission row about to be fetched, its ring slot; the slot being consumed
        const BT* gsrc = b + f0 * a.ld_b + cbase + j0;
        const int64_t gstep = a.ld_b;
        int s_issue = 0;
        auto issue_row = [&]() {
            if (own && s_issue < T) {
                uint4* dst = ering + ((s_issue & (PF - 1)) * CPT) * TG_THREADS + tid;
#pragma unroll
                for (int c = 0; c < CPT; ++c) tr_cp_async(dst + c * TG_THREADS, gsrc + c * EPC, 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            gsrc += gstep;
            ++s_issue;
        };
#pragma unroll
        for (int i = 0; i < PF - 1; ++i) issue_row();
        // transitions into my states by distance: w[i][d] = log A[j - d, j], from the ELL predecessor lists
        double w[SPT][3];
        {
            const int deg = a.max_in;
            const int* eidx = a.pred_idx + (int64_t)k * deg * N;
            const double* eval = a.pred_val + (int64_t)k * deg * N;
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                w[i][0] = w[i][1] = w[i][2] = neg_inf();
                const int j = j0 + i;
                if (own)
                    for (int d = 0; d < deg; ++d) {
                        const double val = __ldg(eval + (int64_t)d * N + j);
                        const int dist = j - __ldg(eidx + (int64_t)d * N + j);
                        if (val != neg_inf()) {
                            if (dist == 0) w[i][0] = val;
                            else if (dist == 1) w[i][1] = val;
                            else w[i][2] = val;
                        }
                    }
            }
        }
        CodeT* btp = (bt_rows && own) ? reinterpret_cast<CodeT*>(bt_rows) + tid : nullptr;
        const bool publish = (lane == 31) && (wp + 1 < n_wp);
        const bool receive = (lane == 0) && (wp > 0);
        auto load_e = [&](int s, double* e) {
            const uint4* er = ering + ((s & (PF - 1)) * CPT) * TG_THREADS + tid;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const uint4 q = er[c * TG_THREADS];       // (threads past the last state read an unwritten slot: their values go nowhere)
                if (sizeof(BT) == 4) {
                    e[c * EPC + 0] = (double)__uint_as_float(q.x); e[c * EPC + 1] = (double)__uint_as_float(q.y);
                    e[c * EPC + 2] = (double)__uint_as_float(q.z); e[c * EPC + 3] = (double)__uint_as_float(q.w);
                } else {
                    e[c * EPC + 0] = __hiloint2double((int)q.y, (int)q.x);
                    e[c * EPC + (EPC > 1 ? 1 : 0)] = __hiloint2double((int)q.w, (int)q.z);
                }
            }
        };
        auto publish_halo = [&](int s) {                               // lane 31 of a warp with a successor
            const int slot = s & (TB2_RING - 1);
            tb2_bar_wait(&h_empty[slot][wp], (((uint32_t)s / TB2_RING) & 1u) ^ 1u);   // the slot's previous content has been read
            halo[slot][wp] = make_double2(v[SPT - 2], v[SPT - 1]);
            tb2_bar_arrive(&h_full[slot][wp]);
        };
        // ---- frame 0: v[0] = pi + b[:, 0]   (hmm.py:159)
        {
            asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");
            issue_row();
            double e[SPT];
            load_e(0, e);
#pragma unroll
            for (int i = 0; i < SPT; ++i) v[i] = own ? __ldg(a.logpi + cbase + j0 + i) + e[i] : neg_inf();
            if (publish) publish_halo(0);
            if (btp) { *btp = (CodeT)0; btp = reinterpret_cast<CodeT*>(reinterpret_cast<uint8_t*>(btp) + bt_step); }
        }
        for (int s = 1; s < T; ++s) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PF - 2) : "memory");     // my chunks of row s have landed
            issue_row();                                                           // into the slot I read in step s - 1
            double e[SPT];
            load_e(s, e);
            // the two values below my range at step s - 1: my left neighbour's top two states
            double2 h;
            h.x = __shfl_up_sync(0xffffffffu, v[SPT - 2], 1);
            h.y = __shfl_up_sync(0xffffffffu, v[SPT - 1], 1);
            if (lane == 0) h = make_double2(neg_inf(), neg_inf());
            if (receive) {
                const int slot = (s - 1) & (TB2_RING - 1);
                tb2_bar_wait(&h_full[slot][wp - 1], ((uint32_t)(s - 1) / TB2_RING) & 1u);     // warp wp-1 has published step s - 1
                h = halo[slot][wp - 1];
                tb2_bar_arrive(&h_empty[slot][wp - 1]);
            }
            __syncwarp();
            // descending i: v[i] is overwritten after the states above it have read it.  Codes: distance 0 / 1 / 2 to the chosen
            // predecessor, FIRST maximum in ascending predecessor order (np.argmax, hmm.py:166).  A column with no finite
            // candidate keeps code 2: such a state is never on a path that ends in a finite score, and an utterance whose
            // every final score is -inf is handled in the backtrace.
            uint32_t code = 0;
#pragma unroll
            for (int i = SPT - 1; i >= 0; --i) {
                const double p2 = (i >= 2) ? v[i - 2] : (i == 1 ? h.y : h.x);
                const double p1 = (i >= 1) ? v[i - 1] : h.y;
                double best = p2 + w[i][2];
                uint32_t dist = 2u << (2 * i);
                const double c1 = p1 + w[i][1];
                if (c1 > best) { best = c1; dist = 1u << (2 * i); }
                const double c0 = v[i] + w[i][0];
                if (c0 > best) { best = c0; dist = 0u; }
                code |= dist;
                v[i] = best + e[i];
            }
            if (publish) publish_halo(s);
            if (btp) { *btp = (CodeT)code; btp = reinterpret_cast<CodeT*>(reinterpret_cast<uint8_t*>(btp) + bt_step); }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                           // the ring is free; every backpointer store has been issued
    double* col = reinterpret_cast<double*>(tg_smem);
    if (own) {
#pragma unroll
        for (int i = 0; i < SPT; ++i) col[j0 + i] = v[i];
    }
    __syncthreads();
    {   // termination (hmm.py:169): max and FIRST argmax over the last column
        double best = neg_inf();
        int arg = 0x7fffffff;
        for (int j = tid; j < N; j += TG_THREADS) {
            const double c = col[j];
            if (c > best) { best = c; arg = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = shfl_d(best, lane ^ o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) { red_v[wp] = best; red_i[wp] = arg; }
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < NW; ++q)
                if (red_v[q] > best || (red_v[q] == best && red_i[q] < arg)) { best = red_v[q]; arg = red_i[q]; }
            if (arg == 0x7fffffff) arg = 0;                    // all -inf: np.argmax gives 0
            if (a.score_out) a.score_out[p] = best;
            if (a.last_state) a.last_state[p] = arg;
            s_last = arg;
            s_best = best;
        }
        __syncthreads();
    }
    if (!a.path_out || !bt_rows) return;
    // ---- fused backtrace: path[T-1] = argmax v[T-1]; path[t] = bt[t+1, path[t+1]]                    (hmm.py:168-171)
    uint32_t* out = a.path_out + (int64_t)k * a.total_frames + f0;
    if (!(s_best > neg_inf())) {
        // no state has a finite score at the last frame (or NaN emissions): path[T-1] = 0 and, in a band, state 0 has itself as
        // its only predecessor, so np.argmax returns 0 at every earlier frame whether column 0 is finite (its one candidate) or
        // not (all -inf -> 0): the reference's path is all zeros
        for (int i = tid; i < T; i += TG_THREADS) out[i] = 0u;
        return;
    }
    uint8_t* stage = tg_smem;                                  // [TB2_CH][bt_pitch]
    uint32_t* pst = reinterpret_cast<uint32_t*>(tg_smem + (size_t)TB2_CH * bt_pitch);   // [TB2_CH] states of frames lo-1 .. hi-1
    int st = s_last;                                           // state at frame hi (thread 0 carries it across chunks)
    if (tid == 0) out[T - 1] = (uint32_t)st;
    const int vec_per_row = bt_pitch / 16;
    for (int hi = T - 1; hi >= 1; hi -= TB2_CH) {
        const int lo = max(1, hi - TB2_CH + 1), nrows = hi - lo + 1;
        __syncthreads();                                       // the previous chunk's staging has been consumed
        for (int i = tid; i < nrows * vec_per_row; i += TG_THREADS) {
            const int r = i / vec_per_row, c = i - r * vec_per_row;
            reinterpret_cast<uint4*>(stage)[i] = __ldcg(reinterpret_cast<const uint4*>(bt_rows + (int64_t)(lo + r) * bt_step) + c);
        }
        __syncthreads();
        if (tid == 0) {
            for (int r = hi; r >= lo; --r) {
                const CodeT cw = reinterpret_cast<const CodeT*>(stage + (size_t)(r - lo) * bt_pitch)[st / SPT];
                st -= ((int)cw >> (2 * (st % SPT))) & 3;
                pst[r - lo] = (uint32_t)st;                    // the state of frame r - 1
            }
        }
        __syncthreads();
        for (int i = tid; i < nrows; i += TG_THREADS) out[lo - 1 + i] = pst[i];
    }
}

template <typename BT, int MODE>
__global__ void __launch_bounds__(TG_THREADS)
trellis_general_kernel(TrellisArgs a) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    const int N = a.N;
    double* buf0 = (double*)tg_smem;
    double* buf1 = buf0 + N;
    __shared__ double red_v[TG_THREADS / 32];
    __shared__ int red_i[TG_THREADS / 32];
    const int64_t p = blockIdx.x;
    const int u = (int)(p / a.n_models), k = (int)(p % a.n_models);
    const int64_t f0 = __ldg(a.frame_off + u);
    const int T = (int)(__ldg(a.frame_off + u + 1) - f0);
    if (T == 0) return;
    const int64_t ncol = (int64_t)a.n_models * N;
    const int64_t cbase = (int64_t)k * N;
    const BT* b = (const BT*)a.b;
    const int deg = (MODE == MODE_BACKWARD) ? a.max_out : a.max_in;
    const int* eidx = ((MODE == MODE_BACKWARD) ? a.succ_idx : a.pred_idx) + (int64_t)k * deg * N;
    const double* eval = ((MODE == MODE_BACKWARD) ? a.succ_val : a.pred_val) + (int64_t)k * deg * N;
    double* prev = buf0;
    double* next = buf1;
    // step 0
    for (int j = threadIdx.x; j < N; j += TG_THREADS) {
        double v;
        if (MODE == MODE_BACKWARD) {
            v = 0.0;
            if (a.full_out) a.full_out[(f0 + T - 1) * ncol + cbase + j] = 0.0;
            v += (double)__ldg(b + (f0 + T - 1) * a.ld_b + cbase + j);    // buffer holds beta[t]+b[t]
        } else {
            v = __ldg(a.logpi + cbase + j) + (double)__ldg(b + f0 * a.ld_b + cbase + j);
            if (a.full_out) a.full_out[f0 * ncol + cbase + j] = v;
            if (MODE == MODE_VITERBI) {
                if (a.bt_ws) a.bt_ws[f0 * ncol + cbase + j] = 0;
                if (a.bt_out) a.bt_out[f0 * ncol + cbase + j] = 0;
            }
        }
        prev[j] = v;
    }
    __syncthreads();
    for (int s = 1; s < T; ++s) {
        const int64_t row = (MODE == MODE_BACKWARD) ? (f0 + T - 1 - s) : (f0 + s);
        for (int j = threadIdx.x; j < N; j += TG_THREADS) {
            const double e = (double)__ldg(b + row * a.ld_b + cbase + j);
            double nv;
            if (MODE == MODE_VITERBI) {
                double best = neg_inf();
                int arg = 0;
                for (int d = 0; d < deg; ++d) {
                    const int i = __ldg(eidx + (int64_t)d * N + j);
                    const double c = prev[i] + __ldg(eval + (int64_t)d * N + j);
                    if (c > best) { best = c; arg = i; }
                }
                nv = best + e;
                if (a.bt_ws) a.bt_ws[row * ncol + cbase + j] = (uint16_t)arg;
                if (a.bt_out) a.bt_out[row * ncol + cbase + j] = (uint32_t)arg;
                if (a.full_out) a.full_out[row * ncol + cbase + j] = nv;
                next[j] = nv;
            } else {
                double mx = neg_inf();
                for (int d = 0; d < deg; ++d) {
                    const int i = __ldg(eidx + (int64_t)d * N + j);
                    mx = fmax(mx, prev[i] + __ldg(eval + (int64_t)d * N + j));
                }
                double sum = 0.0;
                if (mx != neg_inf()) {
                    for (int d = 0; d < deg; ++d) {
                        const int i = __ldg(eidx + (int64_t)d * N + j);
                        const double c = prev[i] + __ldg(eval + (int64_t)d * N + j);
                        sum += (c == neg_inf()) ? 0.0 : exp(c - mx);
                    }
                }
                const double l = (mx == neg_inf()) ? mx : mx + log(sum);
                if (MODE == MODE_FORWARD) {
                    nv = l + e;
                    if (a.full_out) a.full_out[row * ncol + cbase + j] = nv;
                    next[j] = nv;
                } else {
                    if (a.full_out) a.full_out[row * ncol + cbase + j] = l;   // beta[t, j]
                    next[j] = l + e;                                          // beta[t] + b[t] for the next step
                }
            }
        }
        __syncthreads();
        double* t = prev; prev = next; next = t;
    }
    trellis_terminate<MODE>(prev, N, a, p, red_v, red_i);
}

// path[T-1] = argmax v[T-1]; path[t] = bt[t+1, path[t+1]]                                  (hmm.py:168-171)
__global__ void backtrace_kernel(const uint16_t* __restrict__ bt, const int32_t* __restrict__ last_state,
                                 const int64_t* __restrict__ frame_off, int n_utt, int n_models, int N,
                                 int64_t total_frames, uint32_t* __restrict__ path) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)n_utt * n_models) return;
    const int u = (int)(p / n_models), k = (int)(p % n_models);
    const int64_t f0 = frame_off[u];
    const int T = (int)(frame_off[u + 1] - f0);
    if (T == 0) return;
    const int64_t ncol = (int64_t)n_models * N;
    uint32_t* out = path + (int64_t)k * total_frames + f0;
    int s = last_state[p];
    out[T - 1] = (uint32_t)s;
    for (int t = T - 2; t >= 0; --t) {
        s = bt[(f0 + t + 1) * ncol + (int64_t)k * N + s];
        out[t] = (uint32_t)s;
    }
}

// word = argmax_k score[u, k], first maximum (np.argmax); all -inf / NaN-free input assumed
__global__ void argmax_rows_kernel(const double* __restrict__ score, int n_rows, int n_cols, int32_t* __restrict__ arg,
                                   double* __restrict__ best_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double best = neg_inf();
    int a = 0;
    for (int k = 0; k < n_cols; ++k) {
        const double v = score[(int64_t)r * n_cols + k];
        if (v > best) { best = v; a = k; }
    }
    arg[r] = a;
    if (best_out) best_out[r] = best;
}

// ------------------------------------------------------------------------------------------------
template <typename BT, int MODE>
static int launch_small(const TrellisArgs& a, cudaStream_t stream) {
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const int G = a.N <= 8 ? a.N : (a.N <= 16 ? 16 : 32);
    const int ppw = 32 / G;
    const unsigned grid = (unsigned)ceil_div(n_pairs, (int64_t)TS_WARPS * ppw);
#define LNASR_TS(GV, PFV) trellis_small_kernel<GV, BT, MODE, PFV><<<grid, TS_WARPS * 32, 0, stream>>>(a)
    switch (G) {
        case 1:  LNASR_TS(1, 12); break;
        case 2:  LNASR_TS(2, 12); break;
        case 3:  LNASR_TS(3, 12); break;
        case 4:  LNASR_TS(4, 12); break;
        case 5:  LNASR_TS(5, 12); break;
        case 6:  LNASR_TS(6, 12); break;
        case 7:  LNASR_TS(7, 12); break;
        case 8:  LNASR_TS(8, 12); break;
        case 16: LNASR_TS(16, 4); break;
        default: LNASR_TS(32, 4); break;
    }
#undef LNASR_TS
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <typename BT, int MODE>
static int launch_general(const TrellisArgs& a, cudaStream_t stream) {
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const size_t smem = 2 * (size_t)a.N * sizeof(double);
    if (smem > 200 * 1024) { set_error("trellis: N=%d does not fit the shared-memory column buffers", a.N); return LNASR_EINVAL; }
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(trellis_general_kernel<BT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trellis_general_kernel<BT, MODE><<<(unsigned)n_pairs, TG_THREADS, smem, stream>>>(a);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <int SPT, typename BT, int MODE>
static int launch_reg(const TrellisArgs& a, cudaStream_t stream) {
    constexpr int PF = (sizeof(BT) == 4) ? 8 : 4;            // emission rows in flight
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const size_t n_pad = ((size_t)a.N + 3) & ~(size_t)3;
    const size_t smem = 2 * (size_t)a.N * sizeof(double) + (size_t)PF * n_pad * sizeof(BT);
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(trellis_reg_kernel<SPT, BT, MODE, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trellis_reg_kernel<SPT, BT, MODE, PF><<<(unsigned)n_pairs, TG_THREADS, smem, stream>>>(a);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <int SPT, typename BT>
static int launch_band(const TrellisArgs& a, cudaStream_t stream) {
    constexpr int PF = (sizeof(BT) == 4) ? 8 : 4;            // emission rows in flight
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const size_t n_pad = ((size_t)a.N + 3) & ~(size_t)3;
    const size_t smem = n_pad * sizeof(double) + 2 * TG_THREADS * sizeof(double2) + (size_t)PF * n_pad * sizeof(BT);
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(trellis_band_kernel<SPT, BT, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trellis_band_kernel<SPT, BT, PF><<<(unsigned)n_pairs, TG_THREADS, smem, stream>>>(a);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

// backpointer row pitch of trellis_band2_kernel: one code word per owning thread, rounded up to whole 16-byte vectors
static inline int band2_pitch(int N, int spt) { return ((N / spt) * (spt / 4) + 15) / 16 * 16; }
// lnasr_debug_set_band_kernel / environment LNASR_BAND_KERNEL: 1 = trellis_band_kernel + backtrace_kernel
static int g_band_version = []() { const char* e = getenv("LNASR_BAND_KERNEL"); return (e && e[0] == '1') ? 1 : 2; }();

// does the decode of this call run on trellis_band2_kernel (which also backtraces)?
static bool band2_eligible(const void* b, int b_dtype, int64_t ld_b, const lnasr_hmm_bank* bank, bool plain_decode) {
    if (g_band_version != 2 || !plain_decode) return false;
    const int N = bank->N;
    if (!(bank->band >= 1 && bank->band <= 3) || N <= 32 || N > 8 * TG_THREADS) return false;
    if (!bank->pred_idx || !bank->pred_val || bank->max_in < 1) return false;
    const int spt = (N <= 4 * TG_THREADS) ? 4 : 8;
    const size_t esz = (b_dtype == LNASR_DT_F32) ? 4 : 8;
    if (N % spt) return false;
    return ((size_t)ld_b * esz) % 16 == 0 && ((size_t)N * esz) % 16 == 0 && ((uintptr_t)b) % 16 == 0;
}

template <int SPT, typename BT>
static int launch_band2(const TrellisArgs& a, cudaStream_t stream) {
    constexpr int PF = (sizeof(BT) == 4) ? 8 : 4;            // emission rows in flight
    const int64_t n_pairs = (int64_t)a.n_utt * a.n_models;
    const int pitch = band2_pitch(a.N, SPT);
    size_t smem = (size_t)PF * (SPT * sizeof(BT)) * TG_THREADS;                    // the private emission ring
    smem = std::max(smem, (size_t)a.N * sizeof(double));                             // termination column
    smem = std::max(smem, (size_t)TB2_CH * pitch + TB2_CH * sizeof(uint32_t));       // backtrace staging
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(trellis_band2_kernel<SPT, BT, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trellis_band2_kernel<SPT, BT, PF><<<(unsigned)n_pairs, TG_THREADS, smem, stream>>>(a, pitch);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <int MODE>
static int run_trellis(TrellisArgs a, int b_dtype, const lnasr_hmm_bank* bank, cudaStream_t stream) {
    a.n_models = bank->n_models; a.N = bank->N; a.logA = bank->logA; a.logpi = bank->logpi;
    a.max_in = bank->max_in; a.max_out = bank->max_out;
    a.pred_idx = bank->pred_idx; a.pred_val = bank->pred_val; a.succ_idx = bank->succ_idx; a.succ_val = bank->succ_val;
    const bool f32 = (b_dtype == LNASR_DT_F32);
    if (a.N <= 32) return f32 ? launch_small<float, MODE>(a, stream) : launch_small<double, MODE>(a, stream);
    const bool have = (MODE == MODE_BACKWARD) ? (a.succ_idx && a.succ_val && a.max_out > 0) : (a.pred_idx && a.pred_val && a.max_in > 0);
    if (!have) { set_error("trellis: N=%d > 32 needs the packed transition lists (lnasr_trans_pack)", a.N); return LNASR_EINVAL; }
    const int deg = (MODE == MODE_BACKWARD) ? a.max_out : a.max_in;
    // left-to-right band (bank->band = 1 + the largest j - i of a finite transition i -> j, all within 0..2): the decode
    // path proper -- Viterbi without the API-parity outputs -- keeps the trellis column in registers
    if (MODE == MODE_VITERBI && band2_eligible(a.b, b_dtype, a.ld_b, bank, !a.full_out && !a.bt_out)) {
        if (a.N <= 4 * TG_THREADS) return f32 ? launch_band2<4, float>(a, stream) : launch_band2<4, double>(a, stream);
        return f32 ? launch_band2<8, float>(a, stream) : launch_band2<8, double>(a, stream);
    }
    if (MODE == MODE_VITERBI && bank->band >= 1 && bank->band <= 3 && !a.full_out && !a.bt_out && a.N <= 8 * TG_THREADS) {
        if (a.N <= 4 * TG_THREADS) return f32 ? launch_band<4, float>(a, stream) : launch_band<4, double>(a, stream);
        return f32 ? launch_band<8, float>(a, stream) : launch_band<8, double>(a, stream);
    }
    if (deg <= TR_DEG && a.N <= 4 * TG_THREADS)
        return f32 ? launch_reg<4, float, MODE>(a, stream) : launch_reg<4, double, MODE>(a, stream);
    if (deg <= TR_DEG && a.N <= 8 * TG_THREADS)
        return f32 ? launch_reg<8, float, MODE>(a, stream) : launch_reg<8, double, MODE>(a, stream);
    return f32 ? launch_general<float, MODE>(a, stream) : launch_general<double, MODE>(a, stream);
}

static int check_common(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                        const lnasr_hmm_bank* bank, const char* who) {
    if (!b || !frame_offsets || !bank || !bank->logA || !bank->logpi) { set_error("%s: null pointer", who); return LNASR_EINVAL; }
    if (b_dtype != LNASR_DT_F32 && b_dtype != LNASR_DT_F64) { set_error("%s: b dtype must be float32/float64", who); return LNASR_EINVAL; }
    if (bank->n_models < 1 || bank->N < 1 || bank->N > 65535) { set_error("%s: need 1 <= N <= 65535, n_models >= 1", who); return LNASR_EINVAL; }
    if (ld_b < (int64_t)bank->n_models * bank->N) { set_error("%s: ld_b too small", who); return LNASR_EINVAL; }
    if (n_utt < 0) { set_error("%s: n_utt < 0", who); return LNASR_EINVAL; }
    return LNASR_OK;
}

}  // namespace lnasr

using namespace lnasr;

extern "C" int lnasr_debug_set_band_kernel(int version) { g_band_version = (version == 1) ? 1 : 2; return LNASR_OK; }

extern "C" int lnasr_trans_degree(const double* logA, int32_t n_models, int32_t N, int32_t* deg_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(logA && deg_out && n_models >= 1 && N >= 1, "trans_degree: bad arguments");
    LNASR_CUDA(cudaMemsetAsync(deg_out, 0, 4 * sizeof(int32_t), stream));
    const int64_t n = (int64_t)n_models * N;
    trans_degree_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(logA, n_models, N, deg_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_trans_pack(const double* logA, int32_t n_models, int32_t N, int32_t max_in, int32_t max_out,
                                int32_t* pred_idx, double* pred_val, int32_t* succ_idx, double* succ_val, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(logA && pred_idx && pred_val && succ_idx && succ_val, "trans_pack: null pointer");
    LNASR_CHECK_ARG(n_models >= 1 && N >= 1 && max_in >= 1 && max_out >= 1, "trans_pack: bad shape");
    const int64_t n = (int64_t)n_models * N;
    trans_pack_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(logA, n_models, N, max_in, max_out, pred_idx, pred_val, succ_idx, succ_val);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" size_t lnasr_viterbi_workspace_bytes(int64_t total_frames, int32_t n_utt, int32_t n_models, int32_t N) {
    // uint16 backpointers [total_frames, n_models*N] + int32 final state per (utterance, model)
    size_t bt = (size_t)total_frames * n_models * N * sizeof(uint16_t);
    bt = (bt + 255) & ~(size_t)255;
    return bt + (size_t)n_utt * n_models * sizeof(int32_t);
}

extern "C" int lnasr_viterbi_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                                   int64_t total_frames, const lnasr_hmm_bank* bank, uint32_t* path_out,
                                   double* score_out, double* v_out, uint32_t* bt_out, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_common(b, b_dtype, ld_b, frame_offsets, n_utt, bank, "viterbi_batch");
    if (rc) return rc;
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    TrellisArgs a{};
    a.b = b; a.ld_b = ld_b; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.full_out = v_out; a.bt_out = bt_out; a.score_out = score_out;
    const int64_t n_pairs = (int64_t)n_utt * bank->n_models;
    if (path_out) {
        size_t bt_bytes = ((size_t)total_frames * bank->n_models * bank->N * sizeof(uint16_t) + 255) & ~(size_t)255;
        size_t need = bt_bytes + (size_t)n_pairs * sizeof(int32_t);
        if (!workspace || workspace_bytes < need) {
            set_error("viterbi_batch: workspace %zu < %zu bytes", workspace_bytes, need);
            return LNASR_EWORKSPACE;
        }
        a.bt_ws = (uint16_t*)workspace;
        a.last_state = (int32_t*)((char*)workspace + bt_bytes);
        a.path_out = path_out;
        a.total_frames = total_frames;
    }
    rc = run_trellis<MODE_VITERBI>(a, b_dtype, bank, stream);
    if (rc) return rc;
    // the small kernel (N <= 32) and the banded decode kernel backtrace in place
    if (path_out && bank->N > 32 && !band2_eligible(b, b_dtype, ld_b, bank, !v_out && !bt_out)) {
        backtrace_kernel<<<(unsigned)ceil_div(n_pairs, 64), 64, 0, stream>>>(a.bt_ws, a.last_state, frame_offsets, n_utt,
                                                                          bank->n_models, bank->N, total_frames, path_out);
        LNASR_LAUNCH_CHECK();
    }
    return LNASR_OK;
}

extern "C" int lnasr_forward_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                                   int64_t total_frames, const lnasr_hmm_bank* bank, double* alpha_out,
                                   double* logprob_out, void* stream_) {
    int rc = check_common(b, b_dtype, ld_b, frame_offsets, n_utt, bank, "forward_batch");
    if (rc) return rc;
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    TrellisArgs a{};
    a.b = b; a.ld_b = ld_b; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.full_out = alpha_out; a.score_out = logprob_out;
    return run_trellis<MODE_FORWARD>(a, b_dtype, bank, (cudaStream_t)stream_);
}

extern "C" int lnasr_backward_batch(const void* b, int b_dtype, int64_t ld_b, const int64_t* frame_offsets, int32_t n_utt,
                                    int64_t total_frames, const lnasr_hmm_bank* bank, double* beta_out, void* stream_) {
    int rc = check_common(b, b_dtype, ld_b, frame_offsets, n_utt, bank, "backward_batch");
    if (rc) return rc;
    LNASR_CHECK_ARG(beta_out, "backward_batch: beta_out is null");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    TrellisArgs a{};
    a.b = b; a.ld_b = ld_b; a.frame_off = frame_offsets; a.n_utt = n_utt;
    a.full_out = beta_out;
    return run_trellis<MODE_BACKWARD>(a, b_dtype, bank, (cudaStream_t)stream_);
}

extern "C" int lnasr_argmax_rows(const double* score, int32_t n_rows, int32_t n_cols, int32_t* arg_out, double* best_out,
                                 void* stream) {
    LNASR_CHECK_ARG(score && arg_out && n_rows >= 0 && n_cols >= 1, "argmax_rows: bad arguments");
    if (n_rows == 0) return LNASR_OK;
    argmax_rows_kernel<<<(unsigned)ceil_div(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(score, n_rows, n_cols, arg_out, best_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}
