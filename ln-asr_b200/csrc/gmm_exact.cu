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

constexpr int GX_THREADS = 128;   // one frame per thread
constexpr int GX_MAX_DIM = 64;

// packed[g] = [ a_d = -1/(2 var_d) (D) | mu_d (D) | c = -1/2 (D ln 2pi + sum ln var) | logw ]   (2D+2 doubles)
__global__ void gmm_pack_exact_kernel(const double* __restrict__ logw, const double* __restrict__ mu,
                                      const double* __restrict__ var, int n_gauss, int dim,
                                      double* __restrict__ packed) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gauss) return;
    double* o = packed + (int64_t)g * (2 * dim + 2);
    double sumlog = 0.0;
    for (int d = 0; d < dim; ++d) {
        const double v = var[(int64_t)g * dim + d];
        o[d] = -0.5 / v;
        o[dim + d] = mu[(int64_t)g * dim + d];
        sumlog += log(v);
    }
    o[2 * dim] = -0.5 * ((double)dim * 1.8378770664093453 /* ln(2 pi) */ + sumlog);
    o[2 * dim + 1] = logw[g];
}

template <typename FeatT, typename OutT>
__global__ void __launch_bounds__(GX_THREADS)
gmm_loglik_exact_kernel(const FeatT* __restrict__ feats, int ld_feats, int64_t n_frames,
                        const double* __restrict__ packed, int n_states, int n_mix, int dim,
                        OutT* __restrict__ b_out, int64_t ld_b, double* __restrict__ bm_out, int states_per_block) {
    // x tile transposed: xs[d][thread] -> conflict-free column reads
    extern __shared__ __align__(16) unsigned char gx_smem[];
    double (*xs)[GX_THREADS + 1] = reinterpret_cast<double (*)[GX_THREADS + 1]>(gx_smem);
    const int64_t t = (int64_t)blockIdx.x * GX_THREADS + threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * GX_THREADS;
    // coalesced load of the [128, dim] tile
    const int rows = (int)min((int64_t)GX_THREADS, n_frames - t0);
    for (int i = threadIdx.x; i < rows * dim; i += GX_THREADS) {
        const int r = i / dim, d = i - r * dim;
        xs[d][r] = (double)feats[(t0 + r) * ld_feats + d];
    }
    __syncthreads();
    if (t >= n_frames) return;
    const int j_begin = blockIdx.y * states_per_block;
    const int j_end = min(n_states, j_begin + states_per_block);
    const int stride = 2 * dim + 2;
    for (int j = j_begin; j < j_end; ++j) {
        // online log-sum-exp over the mixtures
        double mx = neg_inf(), sum = 0.0;
        for (int m = 0; m < n_mix; ++m) {
            const double* p = packed + (int64_t)(j * n_mix + m) * stride;
            double q = 0.0;
#pragma unroll 4
            for (int d = 0; d < dim; ++d) {
                const double diff = xs[d][threadIdx.x] - __ldg(p + dim + d);
                q = fma(diff * diff, __ldg(p + d), q);
            }
            const double g = q + __ldg(p + 2 * dim);         // pure Gaussian log-density, gmm.py:59-68
            if (bm_out) bm_out[t * ((int64_t)n_states * n_mix) + (int64_t)j * n_mix + m] = g;   // gmmhmm.py:66
            const double bm = g + __ldg(p + 2 * dim + 1);    // + log w[j,m], gmmhmm.py:67
            if (bm > mx) {
                sum = sum * exp(mx - bm) + 1.0;
                mx = bm;
            } else if (bm != neg_inf()) {
                sum += exp(bm - mx);
            }
        }
        const double r = (mx == neg_inf()) ? mx : mx + log(sum);
        b_out[t * ld_b + j] = (OutT)r;
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

extern "C" int lnasr_gmm_pack_exact(const double* logw, const double* mu, const double* var, int32_t n_states,
                                    int32_t n_mix, int32_t dim, double* packed, void* stream) {
    LNASR_CHECK_ARG(logw && mu && var && packed, "gmm_pack_exact: null pointer");
    LNASR_CHECK_ARG(n_states >= 1 && n_mix >= 1 && dim >= 1, "gmm_pack_exact: bad shape");
    const int n = n_states * n_mix;
    gmm_pack_exact_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logw, mu, var, n, dim, packed);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

template <typename FeatT, typename OutT>
static int launch_exact(const void* feats, int ld_feats, int64_t n_frames, const double* packed, int n_states,
                        int n_mix, int dim, void* b_out, int64_t ld_b, double* bm_out, cudaStream_t stream) {
    const int64_t tiles = ceil_div(n_frames, GX_THREADS);
    // split the states over blockIdx.y until the grid fills the machine twice
    int y = 1;
    const int64_t want = 2LL * sm_count() * 4;
    while (tiles * y < want && y * 2 <= n_states) y *= 2;
    const int spb = (int)ceil_div(n_states, y);
    y = (int)ceil_div(n_states, spb);
    dim3 grid((unsigned)tiles, (unsigned)y);
    const size_t smem = (size_t)dim * (GX_THREADS + 1) * sizeof(double);
    if (smem > 48 * 1024)
        LNASR_CUDA(cudaFuncSetAttribute(gmm_loglik_exact_kernel<FeatT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_loglik_exact_kernel<FeatT, OutT><<<grid, GX_THREADS, smem, stream>>>(
        (const FeatT*)feats, ld_feats, n_frames, packed, n_states, n_mix, dim, (OutT*)b_out, ld_b, bm_out, spb);
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
    LNASR_CHECK_ARG(n_frames < (1LL << 31) * GX_THREADS, "gmm_loglik_exact: too many frames");
    if (n_frames == 0) return LNASR_OK;
    const bool f32 = feats_dtype == LNASR_DT_F32, f64 = feats_dtype == LNASR_DT_F64;
    const bool o32 = b_dtype == LNASR_DT_F32, o64 = b_dtype == LNASR_DT_F64;
    LNASR_CHECK_ARG((f32 || f64) && (o32 || o64), "gmm_loglik_exact: dtypes must be float32 or float64");
    if (f32 && o32) return launch_exact<float, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    if (f32 && o64) return launch_exact<float, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    if (f64 && o32) return launch_exact<double, float>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
    return launch_exact<double, double>(feats, ld_feats, n_frames, packed, n_states, n_mix, dim, b_out, ld_b, bm_out, stream);
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
    if (smem > 48 * 1024)
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
