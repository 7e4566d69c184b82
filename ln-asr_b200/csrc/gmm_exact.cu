// gmm_exact.cu -- float64 SIMT evaluation of diagonal-covariance GMM emission log-likelihoods.
//
// Replaces GMMHMM._map_b (/root/reference/lnasr/gmmhmm.py:46-67) + gaussian_multi_distribution_log
// (/root/reference/lnasr/gmm.py:59-68) for diagonal Sigma, for the SMALL models of configs 1-3 where the
// whole parameter set is a few hundred Gaussians: the quadratic form is evaluated as (x-mu)^2 * (-1/(2 var))
// in float64 (no cancellation, matches the reference to ~1e-13), the mixture log-sum-exp in float64.
// The tensor-core path for large models is gmm_tc.cu.
//
// Also: lnasr_emit_gather for discrete HMMs (HMM._map_b, /root/reference/lnasr/hmm.py:47-61).
#include <algorithm>
#include "common.cuh"

namespace lnasr {

constexpr int GX_THREADS = 128;                   // threads of one state group (and of the full-covariance kernel)
constexpr int GX_FT = 2;                          // frames per thread (register tile: 2 frames x 4 Gaussians)
constexpr int GX_TILE = GX_THREADS * GX_FT;       // frames per block
constexpr int GX_SGROUPS = 4;                     // state groups per block: 512 threads share one x tile
constexpr int GX_MAX_DIM = 64;

// Packed model: Gaussians in chunks of 4 mixtures of one state (n_mix padded with log w = -inf).  Per chunk
// (dim+1) rows of 8 doubles:  row d < dim: [a_0..a_3 | b_0..b_3],  a = -1/(2 var_d), b = mu_d/var_d;
// row dim: [c_0..c_3 | logw_0..logw_3],  c = -1/2 (D ln 2pi + sum ln var + sum mu^2/var).
// log N(x) = sum_d (a_d x_d + b_d) x_d + c   -- two DFMA per (Gaussian, dim); float64 makes the expansion harmless.
__host__ __device__ inline int gx_chunks(int n_mix) { return (n_mix + 3) / 4; }

// one WARP per (state, chunk, mixture): the model is re-packed every EM iteration, so the pack must not be a serial loop
__global__ void gmm_pack_exact_kernel(const double* __restrict__ logw, const double* __restrict__ mu,
                                      const double* __restrict__ var, int n_states, int n_mix, int dim,
                                      double* __restrict__ packed) {
    const int nch = gx_chunks(n_mix);
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // (state, chunk, g)
    const int lane = threadIdx.x & 31;
    if (idx >= n_states * nch * 4) return;
    const int g = idx & 3, ch = (idx >> 2) % nch, j = (idx >> 2) / nch;
    const int m = ch * 4 + g;
    double* o = packed + (int64_t)(j * nch + ch) * (dim + 1) * 8;
    if (m >= n_mix) {
        for (int d = lane; d < dim; d += 32) { o[d * 8 + g] = 0.0; o[d * 8 + 4 + g] = 0.0; }
        if (lane == 0) { o[dim * 8 + g] = 0.0; o[dim * 8 + 4 + g] = neg_inf(); }
        return;
    }
    const int64_t gi = (int64_t)j * n_mix + m;
    double c = 0.0;
    for (int d = lane; d < dim; d += 32) {
        const double v = var[gi * dim + d], mm = mu[gi * dim + d];
        o[d * 8 + g] = -0.5 / v;
        o[d * 8 + 4 + g] = mm / v;
        c += log(v) + mm * mm / v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if (lane == 0) {
        o[dim * 8 + g] = -0.5 * (c + (double)dim * 1.8378770664093453);   /* ln(2 pi) */
        o[dim * 8 + 4 + g] = logw[gi];
    }
}

__device__ __forceinline__ void lse_push(double v, double& mx, double& sum) {
    if (v > mx) {
        sum = sum * exp(mx - v) + 1.0;
        mx = v;
    } else if (v != neg_inf()) {
        sum += exp(v - mx);
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Block = 4 state groups x 128 threads sharing one x tile of 256 frames; thread = 2 frames x 4 Gaussians.
// The parameters of the 4 states being scored (one chunk of 4 mixtures each) are staged in shared memory with
// cp.async, double buffered, so the inner loop reads them as conflict-free broadcasts instead of waiting on L2.
template <typename FeatT, typename OutT>
__global__ void __launch_bounds__(GX_THREADS * GX_SGROUPS, 2)
gmm_loglik_exact_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames,
                        const double* __restrict__ packed, int n_states, int n_mix, int dim,
                        OutT* __restrict__ b_out, int64_t ld_b, double* __restrict__ bm_out, int states_per_block) {
    // x tile transposed: xs[d][frame] -> conflict-free column reads
    extern __shared__ __align__(16) unsigned char gx_smem[];
    double (*xs)[GX_TILE] = reinterpret_cast<double (*)[GX_TILE]>(gx_smem);
    const int chunk_doubles = (dim + 1) * 8;
    double* pbuf = reinterpret_cast<double*>(gx_smem) + (size_t)dim * GX_TILE;      // [2][GX_SGROUPS][chunk_doubles]
    const int tid = threadIdx.x & (GX_THREADS - 1);            // frame lane inside the state group
    const int sgroup = threadIdx.x / GX_THREADS;              // scores state j_begin + 4*jg + sgroup
    const int64_t t0 = (int64_t)blockIdx.x * GX_TILE;
    const int rows = (int)min((int64_t)GX_TILE, n_frames - t0);
    const int nch = gx_chunks(n_mix);
    const int j_begin = blockIdx.y * states_per_block;
    const int j_end = min(n_states, j_begin + states_per_block);
    const int n_units = ((j_end - j_begin + GX_SGROUPS - 1) / GX_SGROUPS) * nch;
    auto issue = [&](int u) {
        const int jg = u / nch, ch = u - jg * nch;
        double* dst = pbuf + (size_t)(u & 1) * GX_SGROUPS * chunk_doubles;
        const int pieces = chunk_doubles / 2;                  // 16-byte pieces per chunk
        for (int pidx = threadIdx.x; pidx < GX_SGROUPS * pieces; pidx += GX_THREADS * GX_SGROUPS) {
            const int sg = pidx / pieces, off = pidx - sg * pieces;
            const int j = j_begin + jg * GX_SGROUPS + sg;
            if (j < j_end) cp_async16(dst + (size_t)sg * chunk_doubles + off * 2, packed + (int64_t)(j * nch + ch) * chunk_doubles + off * 2);
        }
        cp_async_commit();
    };
    if (n_units > 0) issue(0);
    for (int i = threadIdx.x; i < GX_TILE * dim; i += GX_THREADS * GX_SGROUPS) {
        const int r = i / dim, d = i - r * dim;
        xs[d][r] = (r < rows) ? (double)feats[(t0 + r) * ld_feats + d] : 0.0;
    }
    const int64_t tA = t0 + tid, tB = t0 + tid + GX_THREADS;
    const bool okA = tA < n_frames, okB = tB < n_frames;
    const int64_t n_gauss = (int64_t)n_states * n_mix;
    double mxA = neg_inf(), sumA = 0.0, mxB = neg_inf(), sumB = 0.0;
    for (int u = 0; u < n_units; ++u) {
        if (u + 1 < n_units) { issue(u + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                                       // unit u's parameters (and, first time, the x tile) are visible
        const int jg = u / nch, ch = u - jg * nch;
        const int j = j_begin + jg * GX_SGROUPS + sgroup;
        if (j < j_end) {
            const double2* p = reinterpret_cast<const double2*>(pbuf + ((size_t)(u & 1) * GX_SGROUPS + sgroup) * chunk_doubles);
            if (ch == 0) { mxA = neg_inf(); sumA = 0.0; mxB = neg_inf(); sumB = 0.0; }
            double qA[4] = {0.0, 0.0, 0.0, 0.0}, qB[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 3
            for (int d = 0; d < dim; ++d) {
                const double xA = xs[d][tid], xB = xs[d][tid + GX_THREADS];
                const double2 a01 = p[d * 4], a23 = p[d * 4 + 1];
                const double2 b01 = p[d * 4 + 2], b23 = p[d * 4 + 3];
                qA[0] = fma(fma(a01.x, xA, b01.x), xA, qA[0]);
                qA[1] = fma(fma(a01.y, xA, b01.y), xA, qA[1]);
                qA[2] = fma(fma(a23.x, xA, b23.x), xA, qA[2]);
                qA[3] = fma(fma(a23.y, xA, b23.y), xA, qA[3]);
                qB[0] = fma(fma(a01.x, xB, b01.x), xB, qB[0]);
                qB[1] = fma(fma(a01.y, xB, b01.y), xB, qB[1]);
                qB[2] = fma(fma(a23.x, xB, b23.x), xB, qB[2]);
                qB[3] = fma(fma(a23.y, xB, b23.y), xB, qB[3]);
            }
            const double2 c01 = p[dim * 4], c23 = p[dim * 4 + 1];
            const double2 w01 = p[dim * 4 + 2], w23 = p[dim * 4 + 3];
            const double c[4] = {c01.x, c01.y, c23.x, c23.y}, lw[4] = {w01.x, w01.y, w23.x, w23.y};
            double vA[4], vB[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int m = ch * 4 + g;
                const double gA = qA[g] + c[g], gB = qB[g] + c[g];     // pure Gaussian log-density, gmm.py:59-68
                if (bm_out && m < n_mix) {                              // gmmhmm.py:66
                    if (okA) bm_out[tA * n_gauss + (int64_t)j * n_mix + m] = gA;
                    if (okB) bm_out[tB * n_gauss + (int64_t)j * n_mix + m] = gB;
                }
                vA[g] = gA + lw[g];                                     // + log w[j,m], gmmhmm.py:67
                vB[g] = gB + lw[g];
            }
            // branch-free streaming log-sum-exp over the chunk (exp(-inf) = 0 covers empty mixtures)
            const double nA = fmax(fmax(fmax(vA[0], vA[1]), fmax(vA[2], vA[3])), mxA);
            const double nB = fmax(fmax(fmax(vB[0], vB[1]), fmax(vB[2], vB[3])), mxB);
            if (nA != neg_inf()) {
                if (ch > 0) sumA *= exp(mxA - nA);
                sumA += (exp(vA[0] - nA) + exp(vA[1] - nA)) + (exp(vA[2] - nA) + exp(vA[3] - nA));
                mxA = nA;
            }
            if (nB != neg_inf()) {
                if (ch > 0) sumB *= exp(mxB - nB);
                sumB += (exp(vB[0] - nB) + exp(vB[1] - nB)) + (exp(vB[2] - nB) + exp(vB[3] - nB));
                mxB = nB;
            }
            if (ch == nch - 1) {
                if (okA) b_out[tA * ld_b + j] = (OutT)((mxA == neg_inf()) ? mxA : mxA + log(sumA));
                if (okB) b_out[tB * ld_b + j] = (OutT)((mxB == neg_inf()) ? mxB : mxB + log(sumB));
            }
        }
        __syncthreads();                                       // buffer u&1 is refilled by issue(u+2)
    }
}


// ------------------------------------------------------------------------------------------------
// FP64 tensor-core variant: the contraction [x^2 | x | 1] . [-1/(2 var) | mu/var | c]^T as
// mma.sync.m8n8k4.f64 (DMMA).  On B200 DMMA and DFMA both sustain 64 FMA/clk/SM (measured, tools/ubench/fp64_pipes.cu)
// but one DMMA carries 256 FMA per warp instruction and its operands arrive as one conflict-free LDS.64 per fragment
// that is reused MT x NT times, so the kernel leaves the shared-memory / issue limits of the SIMT kernel above and runs at
// the FP64 pipe.  Everything stays float64, so the expansion costs ~1e-13 absolute (K = 2D+1 <= 80 terms of magnitude
// <= 1e3) -- same bar as the SIMT kernel.
//
// Block = 8 warps x 16 frames (two blocks per SM: 16 warps hide the latency of the serial exp / log chains of the epilogue).  x tile in shared memory as doubles [128][LDX] with x[D] = 1 (the constant column); the A
// fragment of k-step j is x or x*x of dims 2j, 2j+1, formed in registers.  Gaussians are processed in groups of 40
// (5 n-blocks of 8); the group's parameter rows ([40][84] doubles: k < 80 weights, k = 80 log w) are streamed through a
// double buffer with cp.async.  The mixture log-sum-exp (gmmhmm.py:67) runs on the accumulator fragments: a thread holds
// 2 adjacent Gaussians of one frame, so the n_mix (padded to MP = 1, 2, 4 or 8) mixtures of a state are reached with
// 0, 0, 1 or 2 quad shuffles.
constexpr int GM_WARPS = 8;
constexpr int GM_MT = 2;                            // m-blocks (8 frames each) per warp
constexpr int GM_TILE = GM_WARPS * GM_MT * 8;       // 128 frames per block
constexpr int GM_NT = 5;                            // n-blocks (8 Gaussians each) per group
constexpr int GM_GROUP = GM_NT * 8;                 // 40 Gaussians per group
constexpr int GM_LDB = 84;                          // parameter row: 80 weights | log w | pad  (84 = 4 mod 16: conflict-free B fragments)
constexpr int GM_MAX_DIM = 39;                      // K = 2 dim + 1 <= 80

__host__ __device__ inline int gm_mpad(int n_mix) { return n_mix <= 1 ? 1 : (n_mix <= 2 ? 2 : (n_mix <= 4 ? 4 : 8)); }
__host__ __device__ inline int gm_ksteps(int dim) { return (2 * dim + 1 + 3) / 4; }
__host__ __device__ inline int gm_ldx(int dim) {   // x row stride: >= 2*ksteps slots, = 2 mod 4 (conflict-free A fragments)
    int l = 2 * gm_ksteps(dim);
    while ((l & 3) != 2) ++l;
    return l;
}
__host__ __device__ inline int64_t gm_groups(int n_states, int n_mix) {
    return ((int64_t)n_states * gm_mpad(n_mix) + GM_GROUP - 1) / GM_GROUP;
}
__host__ __device__ inline bool gm_supported(int n_mix, int dim) { return n_mix <= 8 && dim <= GM_MAX_DIM; }

// image: [n_groups*40][84] doubles, Gaussian (state j, mixture m) at row j*MP + m.  One warp per row.
__global__ void gmm_pack_mma_kernel(const double* __restrict__ logw, const double* __restrict__ mu,
                                    const double* __restrict__ var, int n_states, int n_mix, int dim,
                                    double* __restrict__ img) {
    const int MP = gm_mpad(n_mix);
    const int64_t rows = gm_groups(n_states, n_mix) * GM_GROUP;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double* o = img + row * GM_LDB;
    const int64_t j = row / MP;
    const int m = (int)(row % MP);
    const bool real = j < n_states && m < n_mix;
    const int64_t gi = j * n_mix + m;
    double c = 0.0;
    for (int k = lane; k < GM_LDB; k += 32) {
        double val = 0.0;
        const int d = k >> 1;
        if (real && k < 2 * dim) {
            const double v = var[gi * dim + d], mm = mu[gi * dim + d];
            if (k & 1) val = mm / v;
            else { val = -0.5 / v; c += log(v) + mm * mm / v; }
        }
        if (k != 2 * dim && k != 80) o[k] = val;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if (lane == 0) {
        o[2 * dim] = real ? -0.5 * (c + (double)dim * 1.8378770664093453) : 0.0;   /* ln(2 pi) */
        if (2 * dim != 80) o[80] = real ? logw[gi] : neg_inf();
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


// epilogue of both DMMA kernels: + log w, log-sum-exp over the MP (padded) mixtures of a state, stores.  `acc` are the
// accumulator fragments of GM_MT m-blocks x GM_NT n-blocks; tbase = first frame of m-block 0; W = this group's rows.
template <int MP, typename OutT>
__device__ __forceinline__ void gm_epilogue(double (&acc)[GM_MT][GM_NT][2], const double* __restrict__ W, int g, int64_t tbase,
                                            int r, int c, int64_t n_frames, int n_states, int n_mix, int64_t n_gauss,
                                            OutT* __restrict__ b_out, int64_t ld_b, double* __restrict__ bm_out,
                                            const double2* __restrict__ log_tab, const double* __restrict__ exp_tab) {
    // ---- epilogue: + log w, log-sum-exp over the MP (padded) mixtures of a state, stores
#pragma unroll
    for (int nb = 0; nb < GM_NT; ++nb) {
        const int gl = nb * 8 + 2 * c;                                   // my first Gaussian inside the group
        const int64_t gp = (int64_t)g * GM_GROUP + gl;                   // padded global Gaussian index (row of the image)
        const double lw0 = W[(size_t)gl * GM_LDB + 80], lw1 = W[(size_t)(gl + 1) * GM_LDB + 80];
        const int64_t st0 = gp / MP, st1 = (gp + 1) / MP;
        const int mx0 = (int)(gp % MP), mx1 = (int)((gp + 1) % MP);
        double mxv[GM_MT], sum[GM_MT];
#pragma unroll
        for (int mb = 0; mb < GM_MT; ++mb) {
            const int64_t t = tbase + mb * 8 + r;
            const bool ok = t < n_frames;
            const double g0 = acc[mb][nb][0], g1 = acc[mb][nb][1];       // pure Gaussian log-densities (gmm.py:59-68)
            if (bm_out && ok) {                                          // gmmhmm.py:66
                if (st0 < n_states && mx0 < n_mix) bm_out[t * n_gauss + st0 * n_mix + mx0] = g0;
                if (st1 < n_states && mx1 < n_mix) bm_out[t * n_gauss + st1 * n_mix + mx1] = g1;
            }
            const double v0 = g0 + lw0, v1 = g1 + lw1;                   // + log w[j,m] (gmmhmm.py:67)
            if (MP == 1) {
                if (ok && st0 < n_states) b_out[t * ld_b + st0] = (OutT)v0;
                if (ok && st1 < n_states) b_out[t * ld_b + st1] = (OutT)v1;
            } else {
                double mx = fmax(v0, v1);
                if (MP >= 4) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                if (MP >= 8) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                double sm = fast_exp_neg(v0 - mx, exp_tab) + fast_exp_neg(v1 - mx, exp_tab);    // mx = -inf: NaN -> 0
                if (MP >= 4) sm += __shfl_xor_sync(0xffffffffu, sm, 1);
                if (MP >= 8) sm += __shfl_xor_sync(0xffffffffu, sm, 2);
                mxv[mb] = mx;
                sum[mb] = sm;
            }
        }
        if (MP == 2) {
#pragma unroll
            for (int mb = 0; mb < GM_MT; ++mb) {
                const int64_t t = tbase + mb * 8 + r;
                if (t < n_frames && st0 < n_states)
                    b_out[t * ld_b + st0] = (OutT)((mxv[mb] == neg_inf()) ? mxv[mb] : mxv[mb] + fast_log(sum[mb], log_tab));
            }
        } else if (MP >= 4) {
            // the 2 (MP=4) or 4 (MP=8) lanes that share a state hold identical (max, sum) for every m-block: each of
            // them finishes a different m-block, so the log is taken once per result instead of once per lane
            static_assert(GM_MT == 2, "the m-block hand-out below assumes two m-blocks per warp");
            const int mb = c & 1;
            const double mx = mb ? mxv[1] : mxv[0], sm = mb ? sum[1] : sum[0];
            const int64_t t = tbase + mb * 8 + r;
            const bool writer = (MP == 4) || (c < 2);
            if (writer && t < n_frames && st0 < n_states)
                b_out[t * ld_b + st0] = (OutT)((mx == neg_inf()) ? mx : mx + fast_log(sm, log_tab));
        }
    }
}

template <int MP, typename FeatT, typename OutT>
__global__ void __launch_bounds__(GM_WARPS * 32, 2)
gmm_loglik_mma_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames, const double* __restrict__ img,
                      int n_states, int n_mix, int dim, int n_groups, OutT* __restrict__ b_out, int64_t ld_b,
                      double* __restrict__ bm_out) {
    extern __shared__ __align__(16) unsigned char gx_smem[];
    const int ldx = gm_ldx(dim), ks = gm_ksteps(dim);
    double* xs = reinterpret_cast<double*>(gx_smem);                         // [GM_TILE][ldx]
    double* wb = xs + (size_t)GM_TILE * ldx;                                 // [2][GM_GROUP][GM_LDB]
    double2* log_tab = reinterpret_cast<double2*>(wb + 2 * GM_GROUP * GM_LDB);
    double* exp_tab = reinterpret_cast<double*>(log_tab + 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
    const int64_t t0 = (int64_t)blockIdx.x * GM_TILE;
    auto issue = [&](int g) {
        double* dst = wb + (size_t)(g & 1) * GM_GROUP * GM_LDB;
        const double* src = img + (int64_t)g * GM_GROUP * GM_LDB;
        for (int i = tid; i < GM_GROUP * GM_LDB / 2; i += GM_WARPS * 32) cp_async16(dst + 2 * i, src + 2 * i);
        cp_async_commit();
    };
    issue(0);
    for (int i = tid; i < GM_TILE * ldx; i += GM_WARPS * 32) {
        const int row = i / ldx, col = i - row * ldx;
        double v = (col == dim) ? 1.0 : 0.0;
        if (col < dim && t0 + row < n_frames) v = (double)feats[(t0 + row) * ld_feats + col];
        xs[i] = v;
    }
    fill_log_exp_tables(log_tab, exp_tab, tid, GM_WARPS * 32);
    const double* xrow = xs + (size_t)(warp * GM_MT * 8 + r) * ldx + (c >> 1);   // + mb*8*ldx + 2j
    const bool sq = (c & 1) == 0;                                               // even k: x^2, odd k: x
    const int64_t n_gauss = (int64_t)n_states * n_mix;
    for (int g = 0; g < n_groups; ++g) {
        if (g + 1 < n_groups) { issue(g + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                                       // group g's rows (and, first time, the x tile + tables) are visible
        const double* W = wb + (size_t)(g & 1) * GM_GROUP * GM_LDB;
        const double* wrow = W + (size_t)r * GM_LDB + c;        // + nb*8*GM_LDB + 4j
        double acc[GM_MT][GM_NT][2];
#pragma unroll
        for (int mb = 0; mb < GM_MT; ++mb)
#pragma unroll
            for (int nb = 0; nb < GM_NT; ++nb) { acc[mb][nb][0] = 0.0; acc[mb][nb][1] = 0.0; }
#pragma unroll 2
        for (int j = 0; j < ks; ++j) {
            double a[GM_MT], bf[GM_NT];
#pragma unroll
            for (int mb = 0; mb < GM_MT; ++mb) {
                const double xv = xrow[(size_t)mb * 8 * ldx + 2 * j];
                a[mb] = sq ? xv * xv : xv;
            }
#pragma unroll
            for (int nb = 0; nb < GM_NT; ++nb) bf[nb] = wrow[(size_t)nb * 8 * GM_LDB + 4 * j];
#pragma unroll
            for (int mb = 0; mb < GM_MT; ++mb)
#pragma unroll
                for (int nb = 0; nb < GM_NT; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], bf[nb]);
        }
        gm_epilogue<MP, OutT>(acc, W, g, t0 + (int64_t)warp * GM_MT * 8, r, c, n_frames, n_states, n_mix, n_gauss, b_out, ld_b,
                              bm_out, log_tab, exp_tab);
        __syncthreads();                                       // buffer g&1 is refilled by issue(g+2)
    }
}

// Persistent variant for models whose whole DMMA image fits in shared memory (config 2: 200 Gaussians = 134 KB; configs 1
// and 3 a single group): one block per SM loads the image ONCE, then every WARP is an independent worker that walks
// 16-frame slabs -- its own x slab arrives by cp.async one slab ahead, and there is no block barrier after the prologue.
// Warps therefore drift out of phase and the DMMA phase of one warp overlaps the exp / log epilogue of another (the
// streaming kernel above keeps all warps of a block in the same phase with its per-group barriers: tensor pipe 51 %).
constexpr int GR_MAX_WARPS = 16;

template <typename FeatT>
__device__ __forceinline__ void cp_async_elem(FeatT* dst, const FeatT* src) {
    if (sizeof(FeatT) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int MP, typename FeatT, typename OutT>
__global__ void __launch_bounds__(GR_MAX_WARPS * 32, 1)
gmm_loglik_mma_resident_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames, const double* __restrict__ img,
                               int n_states, int n_mix, int dim, int n_groups, OutT* __restrict__ b_out, int64_t ld_b,
                               double* __restrict__ bm_out) {
    extern __shared__ __align__(16) unsigned char gx_smem[];
    const int ldx = gm_ldx(dim), ks = gm_ksteps(dim);
    const int n_warps = blockDim.x >> 5;
    double* W = reinterpret_cast<double*>(gx_smem);                          // [n_groups*40][84]
    double2* log_tab = reinterpret_cast<double2*>(W + (size_t)n_groups * GM_GROUP * GM_LDB);
    double* exp_tab = reinterpret_cast<double*>(log_tab + 128);
    FeatT* xall = reinterpret_cast<FeatT*>(exp_tab + 64);                    // [n_warps][2][16][ldx]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
    for (int64_t i = tid; i < (int64_t)n_groups * GM_GROUP * GM_LDB / 2; i += blockDim.x) cp_async16(W + 2 * i, img + 2 * i);
    cp_async_commit();
    fill_log_exp_tables(log_tab, exp_tab, tid, blockDim.x);
    FeatT* xw = xall + (size_t)warp * 2 * 16 * ldx;
    for (int i = lane; i < 2 * 16 * ldx; i += 32)          // constant column x[dim] = 1 and zero padding; columns < dim arrive by cp.async
        if ((i % ldx) >= dim) xw[i] = ((i % ldx) == dim) ? FeatT(1) : FeatT(0);
    const int64_t n_slabs = (n_frames + 15) >> 4;
    const int64_t wstride = (int64_t)gridDim.x * n_warps;
    int64_t slab = (int64_t)blockIdx.x * n_warps + warp;
    auto fetch = [&](int64_t sl, int buf) {                // x rows of slab sl -> buffer buf (columns < dim only)
        if (sl < n_slabs) {
            FeatT* dst = xw + (size_t)buf * 16 * ldx;
            const int64_t t0 = sl << 4;
            const int rows = (int)min((int64_t)16, n_frames - t0);
            for (int i = lane; i < rows * dim; i += 32) {
                const int row = i / dim, col = i - row * dim;
                cp_async_elem(dst + row * ldx + col, feats + (t0 + row) * ld_feats + col);
            }
        }
        cp_async_commit();
    };
    __syncwarp();
    fetch(slab, 0);
    cp_async_wait<0>();
    __syncthreads();                                       // the image, the tables and (per warp) the first slab are in place
    const bool sq = (c & 1) == 0;                          // even k: x^2, odd k: x
    const int64_t n_gauss = (int64_t)n_states * n_mix;
    for (int it = 0; slab < n_slabs; slab += wstride, ++it) {
        const int buf = it & 1;
        fetch(slab + wstride, buf ^ 1);                    // next slab flies during this one
        const FeatT* xrow = xw + (size_t)buf * 16 * ldx + (size_t)r * ldx + (c >> 1);
        const int64_t t0 = slab << 4;
        for (int g = 0; g < n_groups; ++g) {
            const double* Wg = W + (size_t)g * GM_GROUP * GM_LDB;
            const double* wrow = Wg + (size_t)r * GM_LDB + c;
            double acc[GM_MT][GM_NT][2];
#pragma unroll
            for (int mb = 0; mb < GM_MT; ++mb)
#pragma unroll
                for (int nb = 0; nb < GM_NT; ++nb) { acc[mb][nb][0] = 0.0; acc[mb][nb][1] = 0.0; }
            // fragments of k-step j+1 are loaded before the 10 DMMAs of k-step j are issued
            FeatT xa[GM_MT];
            double bf[GM_NT];
#pragma unroll
            for (int mb = 0; mb < GM_MT; ++mb) xa[mb] = xrow[(size_t)mb * 8 * ldx];
#pragma unroll
            for (int nb = 0; nb < GM_NT; ++nb) bf[nb] = wrow[(size_t)nb * 8 * GM_LDB];
#pragma unroll 2
            for (int j = 0; j < ks; ++j) {
                double a[GM_MT], bc[GM_NT];
#pragma unroll
                for (int mb = 0; mb < GM_MT; ++mb) {
                    const double xv = (double)xa[mb];
                    a[mb] = sq ? xv * xv : xv;
                }
#pragma unroll
                for (int nb = 0; nb < GM_NT; ++nb) bc[nb] = bf[nb];
                const int jn = (j + 1 < ks) ? j + 1 : j;
#pragma unroll
                for (int mb = 0; mb < GM_MT; ++mb) xa[mb] = xrow[(size_t)mb * 8 * ldx + 2 * jn];
#pragma unroll
                for (int nb = 0; nb < GM_NT; ++nb) bf[nb] = wrow[(size_t)nb * 8 * GM_LDB + 4 * jn];
#pragma unroll
                for (int mb = 0; mb < GM_MT; ++mb)
#pragma unroll
                    for (int nb = 0; nb < GM_NT; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], bc[nb]);
            }
            gm_epilogue<MP, OutT>(acc, Wg, g, t0, r, c, n_frames, n_states, n_mix, n_gauss, b_out, ld_b, bm_out, log_tab, exp_tab);
        }
        cp_async_wait<0>();
        __syncwarp();                                      // the next slab has landed and every lane is done with this buffer
    }
}

// Full-covariance variant (the reference's gaussian_multi_distribution_log takes a full Sigma, gmm.py:59-68):
// Sigma = L L^T is factorised on the host at model-pack time; here quad = | Linv x - nu |^2 with nu = Linv mu,
// c = -1/2 (D ln 2pi + 2 sum ln L_ii).  One thread per frame, parameters are warp-uniform loads.
template <typename FeatT, typename OutT>
__global__ void __launch_bounds__(GX_THREADS)
gmm_loglik_full_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames, const double* __restrict__ linv,
                       const double* __restrict__ nu, const double* __restrict__ cst, const double* __restrict__ logw,
                       int n_states, int n_mix, int dim, OutT* __restrict__ b_out, int64_t ld_b,
                       double* __restrict__ bm_out) {
    extern __shared__ __align__(16) unsigned char gx_smem[];
    double (*xs)[GX_THREADS + 1] = reinterpret_cast<double (*)[GX_THREADS + 1]>(gx_smem);
    const int64_t t0 = (int64_t)blockIdx.x * GX_THREADS;
    const int64_t t = t0 + threadIdx.x;
    const int rows = (int)min((int64_t)GX_THREADS, n_frames - t0);
    for (int i = threadIdx.x; i < rows * dim; i += GX_THREADS) {
        const int r = i / dim, d = i - r * dim;
        xs[d][r] = (double)feats[(t0 + r) * ld_feats + d];
    }
    __syncthreads();
    if (t >= n_frames) return;
    for (int j = blockIdx.y; j < n_states; j += gridDim.y) {
        double mx = neg_inf(), sum = 0.0;
        for (int m = 0; m < n_mix; ++m) {
            const int g = j * n_mix + m;
            const double* Lg = linv + (int64_t)g * dim * dim;
            double q = 0.0;
            for (int r = 0; r < dim; ++r) {
                double y = -__ldg(nu + (int64_t)g * dim + r);
                for (int e = 0; e <= r; ++e) y = fma(__ldg(Lg + r * dim + e), xs[e][threadIdx.x], y);
                q = fma(y, y, q);
            }
            const double gl = __ldg(cst + g) - 0.5 * q;
            if (bm_out) bm_out[t * ((int64_t)n_states * n_mix) + g] = gl;
            const double bm = gl + __ldg(logw + g);
            if (bm > mx) { sum = sum * exp(mx - bm) + 1.0; mx = bm; }
            else if (bm != neg_inf()) sum += exp(bm - mx);
        }
        b_out[t * ld_b + j] = (OutT)((mx == neg_inf()) ? mx : mx + log(sum));
    }
}

__global__ void emit_gather_kernel(const double* __restrict__ logB, const int32_t* __restrict__ obs, int64_t n_frames,
                                   int n_states, int n_symbols, double* __restrict__ b_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames * n_states) return;
    const int64_t t = i / n_states;
    const int j = (int)(i - t * n_states);
    const int o = obs[t];
    b_out[i] = (o >= 0 && o < n_symbols) ? logB[(int64_t)j * n_symbols + o] : neg_inf();
}

}  // namespace lnasr

using namespace lnasr;

// packed buffer = SIMT image (chunks of 4 mixtures) followed by the DMMA image ([groups*40][84]) when the shape allows it
static inline int64_t gx_simt_doubles(int n_states, int n_mix, int dim) { return (int64_t)n_states * gx_chunks(n_mix) * (dim + 1) * 8; }

extern "C" int64_t lnasr_gmm_exact_packed_doubles(int32_t n_states, int32_t n_mix, int32_t dim) {
    int64_t n = gx_simt_doubles(n_states, n_mix, dim);
    if (gm_supported(n_mix, dim)) n += gm_groups(n_states, n_mix) * GM_GROUP * GM_LDB;
    return n;
}

extern "C" int lnasr_gmm_pack_exact(const double* logw, const double* mu, const double* var, int32_t n_states,
                                    int32_t n_mix, int32_t dim, double* packed, void* stream) {
    LNASR_CHECK_ARG(logw && mu && var && packed, "gmm_pack_exact: null pointer");
    LNASR_CHECK_ARG(n_states >= 1 && n_mix >= 1 && dim >= 1, "gmm_pack_exact: bad shape");
    const int n = n_states * gx_chunks(n_mix) * 4;
    gmm_pack_exact_kernel<<<(n * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logw, mu, var, n_states, n_mix, dim, packed);
    LNASR_LAUNCH_CHECK();
    if (gm_supported(n_mix, dim)) {
        const int64_t rows = gm_groups(n_states, n_mix) * GM_GROUP;
        gmm_pack_mma_kernel<<<(unsigned)ceil_div(rows * 32, 128), 128, 0, (cudaStream_t)stream>>>(
            logw, mu, var, n_states, n_mix, dim, packed + gx_simt_doubles(n_states, n_mix, dim));
        LNASR_LAUNCH_CHECK();
    }
    return LNASR_OK;
}

template <typename FeatT, typename OutT>
static int launch_exact(const void* feats, int ld_feats, int64_t n_frames, const double* packed, int n_states,
                        int n_mix, int dim, void* b_out, int64_t ld_b, double* bm_out, cudaStream_t stream) {
    const int64_t tiles = ceil_div(n_frames, GX_TILE);
    // split the states over blockIdx.y until the grid fills the machine a few times over
    int y = 1;
    const int64_t want = 4LL * sm_count();
    while (tiles * y < want && y * 2 <= n_states) y *= 2;
    const int spb = (int)ceil_div(n_states, y);
    y = (int)ceil_div(n_states, spb);
    dim3 grid((unsigned)tiles, (unsigned)y);
    const size_t smem = ((size_t)dim * GX_TILE + (size_t)2 * GX_SGROUPS * (dim + 1) * 8) * sizeof(double);
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(gmm_loglik_exact_kernel<FeatT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_loglik_exact_kernel<FeatT, OutT><<<grid, GX_THREADS * GX_SGROUPS, smem, stream>>>(
        (const FeatT*)feats, ld_feats, n_frames, packed, n_states, n_mix, dim, (OutT*)b_out, ld_b, bm_out, spb);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}


static bool g_force_simt = false;     // bring-up / test switches: lnasr_debug_set_gmm_simt(1) keeps the SIMT kernel,
static bool g_force_stream = false;   // (2) keeps the streaming DMMA kernel where the persistent one applies

template <typename FeatT, typename OutT>
static int launch_mma(const void* feats, int ld_feats, int64_t n_frames, const double* packed, int n_states,
                      int n_mix, int dim, void* b_out, int64_t ld_b, double* bm_out, cudaStream_t stream) {
    const double* img = packed + gx_simt_doubles(n_states, n_mix, dim);
    const int n_groups = (int)gm_groups(n_states, n_mix);
    // persistent kernel with the whole image resident when it leaves room for >= 8 warps of double-buffered x slabs
    {
        const size_t fixed = ((size_t)n_groups * GM_GROUP * GM_LDB + 256 + 64) * sizeof(double);
        const size_t per_warp = (size_t)2 * 16 * gm_ldx(dim) * sizeof(FeatT);
        const size_t cap = 226 * 1024;
        int nw = fixed < cap ? (int)std::min<size_t>(GR_MAX_WARPS, (cap - fixed) / per_warp) : 0;
        const int64_t n_slabs = ceil_div(n_frames, 16);
        if (nw >= 8 && !g_force_stream) {
            const size_t smem_r = fixed + (size_t)nw * per_warp;
            const unsigned grid_r = (unsigned)std::min<int64_t>(sm_count(), ceil_div(n_slabs, nw));
#define LNASR_GR_LAUNCH(MPV)                                                                                              \
    do {                                                                                                                  \
        LNASR_CUDA(cudaFuncSetAttribute(gmm_loglik_mma_resident_kernel<MPV, FeatT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)); \
        gmm_loglik_mma_resident_kernel<MPV, FeatT, OutT><<<grid_r, nw * 32, smem_r, stream>>>(                            \
            (const FeatT*)feats, ld_feats, n_frames, img, n_states, n_mix, dim, n_groups, (OutT*)b_out, ld_b, bm_out);    \
    } while (0)
            switch (gm_mpad(n_mix)) {
                case 1: LNASR_GR_LAUNCH(1); break;
                case 2: LNASR_GR_LAUNCH(2); break;
                case 4: LNASR_GR_LAUNCH(4); break;
                default: LNASR_GR_LAUNCH(8); break;
            }
#undef LNASR_GR_LAUNCH
            LNASR_LAUNCH_CHECK();
            return LNASR_OK;
        }
    }
    const size_t smem = ((size_t)GM_TILE * gm_ldx(dim) + 2 * GM_GROUP * GM_LDB + 256 + 64) * sizeof(double);
    const unsigned grid = (unsigned)ceil_div(n_frames, GM_TILE);
#define LNASR_GM_LAUNCH(MPV)                                                                                              \
    do {                                                                                                                  \
        LNASR_CUDA(cudaFuncSetAttribute(gmm_loglik_mma_kernel<MPV, FeatT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gmm_loglik_mma_kernel<MPV, FeatT, OutT><<<grid, GM_WARPS * 32, smem, stream>>>(                                  \
            (const FeatT*)feats, ld_feats, n_frames, img, n_states, n_mix, dim, n_groups, (OutT*)b_out, ld_b, bm_out);    \
    } while (0)
    switch (gm_mpad(n_mix)) {
        case 1: LNASR_GM_LAUNCH(1); break;
        case 2: LNASR_GM_LAUNCH(2); break;
        case 4: LNASR_GM_LAUNCH(4); break;
        default: LNASR_GM_LAUNCH(8); break;
    }
#undef LNASR_GM_LAUNCH
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}


extern "C" int lnasr_debug_set_gmm_simt(int on) { g_force_simt = (on == 1); g_force_stream = (on == 2); return LNASR_OK; }

// A frame with an inf or NaN feature (the log energy of an all-zero frame is -inf, mfcc.py:158-160, and the deltas next to it
// are NaN): the reference evaluates (x - mu) Sigma^-1 (x - mu)^T with the FULL matrix product (gmm.py:59-68), where the zero
// off-diagonal entries of a diagonal Sigma^-1 meet the infinite difference as 0 * inf = NaN -- every Gaussian of such a frame is
// NaN there.  The expansion the kernels above evaluate would give -inf, +inf or NaN depending on the signs of mu, so those
// rows are overwritten here.  One block per 128 frames; it reads the features once and returns unless a frame is flagged.
template <typename FeatT, typename OutT>
__global__ void __launch_bounds__(256) gmm_nonfinite_rows_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames, int dim,
                                                                 int n_states, int n_mix, OutT* __restrict__ b_out, int64_t ld_b,
                                                                 double* __restrict__ bm_out) {
    __shared__ int s_flag[128];
    __shared__ int s_any;
    const int64_t f0 = (int64_t)blockIdx.x * 128;
    const int rows = (int)min((int64_t)128, n_frames - f0);
    if (threadIdx.x < 128) s_flag[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * dim; i += 256) {
        const int r = i / dim, d = i - r * dim;
        const double x = (double)feats[(f0 + r) * ld_feats + d];
        if (!(fabs(x) <= 1.7976931348623157e308)) { s_flag[r] = 1; s_any = 1; }
    }
    __syncthreads();
    if (!s_any) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int r = 0; r < rows; ++r) {
        if (!s_flag[r]) continue;
        for (int j = threadIdx.x; j < n_states; j += 256) b_out[(f0 + r) * ld_b + j] = (OutT)qnan;
        if (bm_out)
            for (int64_t j = threadIdx.x; j < (int64_t)n_states * n_mix; j += 256) bm_out[(f0 + r) * (int64_t)n_states * n_mix + j] = qnan;
    }
}

template <typename FeatT, typename OutT>
static int launch_nonfinite(const void* feats, int ld_feats, int64_t n_frames, int dim, int n_states, int n_mix, void* b_out, int64_t ld_b,
                            double* bm_out, cudaStream_t stream) {
    gmm_nonfinite_rows_kernel<FeatT, OutT><<<(unsigned)ceil_div(n_frames, 128), 256, 0, stream>>>(
        (const FeatT*)feats, ld_feats, n_frames, dim, n_states, n_mix, (OutT*)b_out, ld_b, bm_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_gmm_loglik_exact(const void* feats, int feats_dtype, int32_t ld_feats, int64_t n_frames,
                                      const double* packed, int32_t n_states, int32_t n_mix, int32_t dim,
                                      void* b_out, int b_dtype, int64_t ld_b, double* bm_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(feats && packed && b_out, "gmm_loglik_exact: null pointer");
    LNASR_CHECK_ARG(dim >= 1 && dim <= GX_MAX_DIM, "gmm_loglik_exact: dim must be in [1,%d]", GX_MAX_DIM);
    LNASR_CHECK_ARG(ld_feats >= dim && ld_b >= n_states && n_states >= 1 && n_mix >= 1, "gmm_loglik_exact: bad shape");
    LNASR_CHECK_ARG(n_frames < (1LL << 31) * GX_TILE, "gmm_loglik_exact: too many frames");
    if (n_frames == 0) return LNASR_OK;
    const bool f32 = feats_dtype == LNASR_DT_F32, f64 = feats_dtype == LNASR_DT_F64;
    const bool o32 = b_dtype == LNASR_DT_F32, o64 = b_dtype == LNASR_DT_F64;
    LNASR_CHECK_ARG((f32 || f64) && (o32 || o64), "gmm_loglik_exact: dtypes must be float32 or float64");
    int rc;
    if (gm_supported(n_mix, dim) && !g_force_simt) {
        if (f32 && o32) rc = launch_mma<float, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else if (f32 && o64) rc = launch_mma<float, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else if (f64 && o32) rc = launch_mma<double, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else rc = launch_mma<double, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    } else {
        if (f32 && o32) rc = launch_exact<float, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else if (f32 && o64) rc = launch_exact<float, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else if (f64 && o32) rc = launch_exact<double, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
        else rc = launch_exact<double, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    }
    if (rc) return rc;
    if (f32 && o32) return launch_nonfinite<float, float>(feats, ld_feats, n_frames, dim, n_states, n_mix, b_out, ld_b, bm_out, stream);
    if (f32 && o64) return launch_nonfinite<float, double>(feats, ld_feats, n_frames, dim, n_states, n_mix, b_out, ld_b, bm_out, stream);
    if (f64 && o32) return launch_nonfinite<double, float>(feats, ld_feats, n_frames, dim, n_states, n_mix, b_out, ld_b, bm_out, stream);
    return launch_nonfinite<double, double>(feats, ld_feats, n_frames, dim, n_states, n_mix, b_out, ld_b, bm_out, stream);
}

template <typename FeatT, typename OutT>
static int launch_full(const void* feats, int ld_feats, int64_t n_frames, const double* linv, const double* nu,
                       const double* cst, const double* logw, int n_states, int n_mix, int dim, void* b_out, int64_t ld_b,
                       double* bm_out, cudaStream_t stream) {
    const int64_t tiles = ceil_div(n_frames, GX_THREADS);
    int y = 1;
    while (tiles * y < 2LL * sm_count() * 4 && y * 2 <= n_states) y *= 2;
    dim3 grid((unsigned)tiles, (unsigned)y);
    const size_t smem = (size_t)dim * (GX_THREADS + 1) * sizeof(double);
    if (smem > 32 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(gmm_loglik_full_kernel<FeatT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_loglik_full_kernel<FeatT, OutT><<<grid, GX_THREADS, smem, stream>>>(
        (const FeatT*)feats, ld_feats, n_frames, linv, nu, cst, logw, n_states, n_mix, dim, (OutT*)b_out, ld_b, bm_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_gmm_loglik_full(const void* feats, int feats_dtype, int32_t ld_feats, int64_t n_frames,
                                     const double* linv, const double* nu, const double* cst, const double* logw,
                                     int32_t n_states, int32_t n_mix, int32_t dim, void* b_out, int b_dtype, int64_t ld_b,
                                     double* bm_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LNASR_CHECK_ARG(feats && linv && nu && cst && logw && b_out, "gmm_loglik_full: null pointer");
    LNASR_CHECK_ARG(dim >= 1 && dim <= GX_MAX_DIM, "gmm_loglik_full: dim must be in [1,%d]", GX_MAX_DIM);
    LNASR_CHECK_ARG(ld_feats >= dim && ld_b >= n_states && n_states >= 1 && n_mix >= 1, "gmm_loglik_full: bad shape");
    if (n_frames == 0) return LNASR_OK;
    const bool f32 = feats_dtype == LNASR_DT_F32, f64 = feats_dtype == LNASR_DT_F64;
    const bool o32 = b_dtype == LNASR_DT_F32, o64 = b_dtype == LNASR_DT_F64;
    LNASR_CHECK_ARG((f32 || f64) && (o32 || o64), "gmm_loglik_full: dtypes must be float32 or float64");
    if (f32 && o32) return launch_full<float, float>(feats, ld_feats, n_frames, linv, nu, cst, logw, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    if (f32 && o64) return launch_full<float, double>(feats, ld_feats, n_frames, linv, nu, cst, logw, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    if (f64 && o32) return launch_full<double, float>(feats, ld_feats, n_frames, linv, nu, cst, logw, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    return launch_full<double, double>(feats, ld_feats, n_frames, linv, nu, cst, logw, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
}

extern "C" int lnasr_emit_gather(const double* logB, const int32_t* obs, int64_t n_frames, int32_t n_states,
                                 int32_t n_symbols, double* b_out, void* stream) {
    LNASR_CHECK_ARG(logB && obs && b_out && n_states >= 1 && n_symbols >= 1, "emit_gather: bad arguments");
    if (n_frames == 0) return LNASR_OK;
    const int64_t n = n_frames * n_states;
    emit_gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(logB, obs, n_frames, n_states, n_symbols, b_out);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}
