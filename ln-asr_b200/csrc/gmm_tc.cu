// gmm_tc.cu -- GMM emission log-likelihoods on the 5th-generation tensor cores (kernel (2) of the north star).
//
// Replaces GMMHMM._map_b (/root/reference/lnasr/gmmhmm.py:46-67) + gaussian_multi_distribution_log
// (/root/reference/lnasr/gmm.py:59-68) for diagonal covariances and LARGE models (config 4: 2000 states x 16
// mixtures x 39 dims = 32 000 Gaussians per frame):
//
//     log N(x; mu, diag var) + log w  =  [x^2 | x | 1] . [ -1/(2 var) | mu/var | c ]^T          
//
// is a dense [frames, K] x [K, states*mixtures] contraction with K = 2*dim+1 (79 -> 80).  It runs as
// tcgen05.mma kind::tf32 with 3xTF32 error compensation: every fp32 operand is split v = hi + lo with hi
// rounded to TF32, and   D = Ah.Bh + Ah.Bl + Al.Bh   is accumulated in fp32 in TMEM (the dropped Al.Bl term is
// 2^-22 relative).  The log-sum-exp over the mixtures of a state is fused into the TMEM epilogue: the packed
// weights put the M mixtures of a state in M adjacent accumulator columns, so one tcgen05.ld gives a thread
// (= one frame) all mixtures of a state and the reduction needs no shuffles.  Weights are pre-scaled by
// log2(e) so the epilogue is max / ex2 / lg2.
//
// CTA = 17 warps, one CTA per SM, each owning one W tile (NT <= 224 accumulator columns, resident in shared
// memory for the whole pass) and streaming 128-frame tiles:
//   warps 8-11  lo producers : thread = frame row; reads its 39 (centred) features from the tile-transposed staging
//   warps 12-15 hi producers   copy (coalesced), forms [x^2|x|1], splits hi/lo and writes the UMMA canonical K-major
//                              (no-swizzle) layout straight into a 4-slot ring (slot = half of K of the lo or hi
//                              image), fence.proxy.async, mbarrier arrive;
//   warp 16     MMA          : one thread issues 30 tcgen05.mma (M=128, N=NT, K=8) per tile, tcgen05.commit frees
//                              ring slots and publishes the accumulator stage (2 TMEM stages of NT columns);
//   warps 0-7   epilogue     : two warps per TMEM lane quadrant (half of the columns each): tcgen05.ld 32x32b.x16,
//                              max / ex2 / lg2 over the mixtures, + per-state constant, b[frame, state] fp32 stores.
// No tensor maps are needed: A never exists in global memory and W is a linear copy of a pre-packed image.
#include <algorithm>
#include "common.cuh"
#include "tc_common.cuh"

namespace lnasr {

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------ bring-up probe
// One tcgen05.mma (shape from idesc) on caller-provided shared-memory images; D [128 lanes, 128 columns] is returned.
__global__ void __launch_bounds__(128) umma_probe_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img,
                                                         int b_bytes, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                                                         uint32_t layout_type, uint32_t idesc, float* __restrict__ d_out,
                                                         uint32_t second_off) {
    extern __shared__ __align__(1024) uint8_t probe_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* sa = (float*)probe_smem;
    float* sb = (float*)(probe_smem + ((a_bytes + 1023) & ~1023));
    for (int i = threadIdx.x; i < a_bytes / 4; i += 128) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < b_bytes / 4; i += 128) sb[i] = b_img[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint64_t ad = make_smem_desc(smem_u32(sa), a_lbo, a_sbo, layout_type);
        const uint64_t bd = make_smem_desc(smem_u32(sb), b_lbo, b_sbo, layout_type);
        umma_tf32(tmem_base, ad, bd, idesc, 0u);
        if (second_off) {
            // a second, ACCUMULATING MMA on the K chunks that start `second_off` bytes further into both images: shows how the
            // fp32 accumulator in TMEM takes in a small addend (rounding mode, guard bits) -- tools/probe_umma_rounding.py
            const uint64_t ad2 = make_smem_desc(smem_u32(sa) + second_off, a_lbo, a_sbo, layout_type);
            const uint64_t bd2 = make_smem_desc(smem_u32(sb) + second_off, b_lbo, b_sbo, layout_type);
            umma_tf32(tmem_base, ad2, bd2, idesc, 1u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * 128 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ packed image
constexpr uint32_t TC_MAGIC = 0x4c4e5443u;   // "LNTC"
constexpr int TC_BM = 128;                    // frames per tile (UMMA M)
constexpr int TC_THREADS = 544;               // 8 epilogue + 4 lo-producer + 4 hi-producer + 1 MMA warp
constexpr int TC_MAX_DIM = 39;                // the K axis is laid out for 39 dims; smaller dims are zero-padded
constexpr int TC_KP = 80;                     // 2*39 + 1 -> 80
constexpr int TC_KH = TC_KP / 2;              // K values per ring slot
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr float TC_LN2 = 0.6931471805599453f;
constexpr float TC_LN2_LO = (float)(0.6931471805599453094 - (double)0.6931471805599453f);

struct TcHeader {       // 512 bytes at the start of the packed image
    uint32_t magic;
    int32_t n_states, n_mix, Mp, dim, Kp, SPT, NT, n_wtiles;
    int32_t pad[7];
    float shift[64];
    uint8_t reserved[512 - 16 * 4 - 64 * 4];
};
static_assert(sizeof(TcHeader) == 512, "TcHeader must be 512 bytes");

struct TcGeom { int Mp, Kp, SPT, NT, n_wtiles; size_t tile_bytes, tail_off, total_bytes, smem_bytes; };

static bool tc_geometry(int n_states, int n_mix, int dim, TcGeom* g) {
    if (n_states < 1 || n_mix < 1 || n_mix > 16 || dim < 1 || dim > TC_MAX_DIM) return false;
    int Mp = 1;
    while (Mp < n_mix) Mp <<= 1;
    const int Kp = TC_KP;                                      // fixed K axis [x0^2, x0, ... x38^2, x38 | 1 | 0]; dims beyond `dim` are zero
    const int gran = 16 / Mp;                                  // states per 16 accumulator columns
    // shared memory: W hi+lo = 2*NT*Kp*4, per-state constants, A ring = 4 * 128 * (Kp/2) * 4, barriers 1 KB
    const int ring = 4 * TC_BM * (Kp / 2) * 4;
    int nt_max = (TC_SMEM_LIMIT - ring - 2048) / (2 * Kp * 4);
    nt_max = std::min(nt_max, 256) / 16 * 16;
    if (nt_max < 16) return false;
    const int spt_max = nt_max / Mp;
    const int n_wtiles = (n_states + spt_max - 1) / spt_max;
    int spt = (n_states + n_wtiles - 1) / n_wtiles;
    spt = (spt + gran - 1) / gran * gran;
    g->Mp = Mp; g->Kp = Kp; g->SPT = spt; g->NT = spt * Mp;
    g->n_wtiles = (n_states + spt - 1) / spt;
    g->tile_bytes = (size_t)2 * g->NT * Kp * 4 + (size_t)((g->SPT * 4 + 15) / 16 * 16);   // W hi | W lo | per-state constant
    g->tail_off = (sizeof(TcHeader) + g->tile_bytes * g->n_wtiles + 15) / 16 * 16;       // exact tail (tc_common.cuh)
    g->total_bytes = g->tail_off + tc_tail_bytes(n_states, n_mix, dim);
    g->smem_bytes = g->tile_bytes + ring + 1024;
    return true;
}

// ------------------------------------------------------------------------------------------------ exact tail + fix-up
__global__ void tc_pack_tail_kernel(const double* __restrict__ logw, const double* __restrict__ mu, const double* __restrict__ var,
                                    int64_t n_gauss, int dim, double* __restrict__ tail) {
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_gauss) return;
    double c = (double)dim * 1.8378770664093453;
    for (int d = 0; d < dim; ++d) {
        const double v = var[gi * dim + d];
        c += log(v);
        tail[n_gauss + gi * dim + d] = mu[gi * dim + d];
        tail[n_gauss + n_gauss * dim + gi * dim + d] = -0.5 / v;
    }
    tail[gi] = logw[gi] - 0.5 * c;
}

int tc_pack_tail(const double* logw, const double* mu, const double* var, int n_states, int n_mix, int dim, double* tail, cudaStream_t stream) {
    const int64_t n_gauss = (int64_t)n_states * n_mix;
    tc_pack_tail_kernel<<<(unsigned)ceil_div(n_gauss, 128), 128, 0, stream>>>(logw, mu, var, n_gauss, dim, tail);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

// one block per 128-frame tile; returns at once unless the staging kernel flagged a frame of the tile
__global__ void __launch_bounds__(256) tc_fixup_kernel(const float* __restrict__ feats, int ld_feats, int64_t n_frames,
                                                       const uint8_t* __restrict__ flags, const double* __restrict__ tail, int n_states,
                                                       int n_mix, int dim, float* __restrict__ b_out, int64_t ld_b) {
    __shared__ int s_list[128];
    __shared__ int s_n;
    __shared__ double s_x[64];
    const int64_t f0 = (int64_t)blockIdx.x * 128;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (threadIdx.x < 128 && f0 + threadIdx.x < n_frames && flags[f0 + threadIdx.x]) s_list[atomicAdd(&s_n, 1)] = threadIdx.x;
    __syncthreads();
    const int n_flagged = s_n;
    if (n_flagged == 0) return;
    const int64_t n_gauss = (int64_t)n_states * n_mix;
    const double* t_mu = tail + n_gauss;
    const double* t_nh = tail + n_gauss + n_gauss * dim;
    for (int q = 0; q < n_flagged; ++q) {
        const int64_t f = f0 + s_list[q];
        __syncthreads();
        if (threadIdx.x < dim) s_x[threadIdx.x] = (double)feats[f * ld_feats + threadIdx.x];
        __syncthreads();
        bool finite = true;
        for (int d = 0; d < dim; ++d) finite = finite && (fabs(s_x[d]) <= 1.7976931348623157e308);
        if (!finite) {
            // inf or NaN feature: every Gaussian of the frame is NaN in the reference (0 * inf in its full-matrix quadratic form,
            // gmm.py:59-68; see gmm_nonfinite_rows_kernel in gmm_exact.cu)
            for (int j = threadIdx.x; j < n_states; j += 256) b_out[f * ld_b + j] = __int_as_float(0x7fc00000);
            continue;
        }
        for (int j = threadIdx.x; j < n_states; j += 256) {
            // log sum_m exp(cst + quad) with the reference's conventions: NaN in -> NaN out, every mixture -inf -> -inf
            double mx = neg_inf();
            bool is_nan = false;
            for (int m = 0; m < n_mix; ++m) {
                const int64_t gi = (int64_t)j * n_mix + m;
                double v = tail[gi];
                for (int d = 0; d < dim; ++d) {
                    const double dx = s_x[d] - t_mu[gi * dim + d];
                    v = fma(t_nh[gi * dim + d], dx * dx, v);
                }
                is_nan |= (v != v);
                mx = fmax(mx, v);
            }
            double r;
            if (is_nan) r = __longlong_as_double(0x7ff8000000000000LL);
            else if (mx == neg_inf()) r = mx;
            else {
                double sum = 0.0;
                for (int m = 0; m < n_mix; ++m) {
                    const int64_t gi = (int64_t)j * n_mix + m;
                    double v = tail[gi];
                    for (int d = 0; d < dim; ++d) {
                        const double dx = s_x[d] - t_mu[gi * dim + d];
                        v = fma(t_nh[gi * dim + d], dx * dx, v);
                    }
                    sum += exp(v - mx);
                }
                r = mx + log(sum);
            }
            b_out[f * ld_b + j] = (float)r;
        }
    }
}

int tc_fixup(const float* feats, int ld_feats, int64_t n_frames, const uint8_t* flags, const double* tail, int n_states, int n_mix,
             int dim, float* b_out, int64_t ld_b, cudaStream_t stream) {
    tc_fixup_kernel<<<(unsigned)ceil_div(n_frames, 128), 256, 0, stream>>>(feats, ld_feats, n_frames, flags, tail, n_states, n_mix, dim,
                                                                           b_out, ld_b);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

// byte offset of element (row n, k) inside one K-major no-swizzle image of `rows` x `K` fp32:
// core matrix = 8 rows x 16 bytes; [n/8][k/4][n%8][k%4]  ->  LBO (next K chunk) = 128 B, SBO (next 8 rows) = (K/4)*128 B
__host__ __device__ inline uint32_t kmajor_off(int n, int k, int K) {
    return (uint32_t)(((n >> 3) * (K >> 2) + (k >> 2)) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

__global__ void gmm_pack_tc_kernel(const double* __restrict__ logw, const double* __restrict__ mu, const double* __restrict__ var,
                                   int n_states, int n_mix, int dim, const double* __restrict__ shift, TcGeom g,
                                   uint8_t* __restrict__ packed) {
    const int64_t cols_total = (int64_t)g.n_wtiles * g.NT;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        TcHeader* h = reinterpret_cast<TcHeader*>(packed);
        h->magic = TC_MAGIC; h->n_states = n_states; h->n_mix = n_mix; h->Mp = g.Mp; h->dim = dim; h->Kp = g.Kp;
        h->SPT = g.SPT; h->NT = g.NT; h->n_wtiles = g.n_wtiles;
        for (int d = 0; d < 64; ++d) h->shift[d] = (shift && d < dim) ? (float)shift[d] : 0.f;
    }
    if (idx >= cols_total * g.Kp) return;
    const int64_t col = idx / g.Kp;
    const int k = (int)(idx - col * g.Kp);
    const int tile = (int)(col / g.NT), n = (int)(col - (int64_t)tile * g.NT);
    const int j = tile * g.SPT + n / g.Mp, m = n % g.Mp;
    double v = 0.0;
    if (j < n_states && m < n_mix) {
        const int64_t gi = (int64_t)j * n_mix + m;
        if (k < 2 * TC_MAX_DIM) {
            // K layout [x0^2, x0, x1^2, x1, ...| 1 | 0]: the two terms of a dimension adjacent (see H_KP in gmm_tc16.cu)
            const int d = k >> 1;
            if (d < dim) {
                const double sh = shift ? shift[d] : 0.0;
                v = (k & 1) ? (mu[gi * dim + d] - sh) / var[gi * dim + d] : -0.5 / var[gi * dim + d];
            }
        }
        v *= 1.4426950408889634;                          // log2(e): the epilogue works in base 2
    }
    if (k == 2 * TC_MAX_DIM) {
        // K column 78 multiplies the constant 1 of the A rows.  It carries the constant of the Gaussian (+ log w)
        // RELATIVE to the largest constant of its state; that per-state maximum is added in the epilogue in fp32 with
        // round-to-nearest.  Keeping the big common part (~ -55 in base 2) out of the tensor-core accumulator keeps
        // the accumulated magnitudes -- and with them the truncation of the fp32 accumulation -- small, at the
        // price of one add per state.  Padded mixtures / states get -1e30 (exp2 -> 0).
        auto gconst = [&](int mm_) -> double {
            const int64_t gi = (int64_t)j * n_mix + mm_;
            double c = (double)dim * 1.8378770664093453;
            for (int d = 0; d < dim; ++d) {
                const double sh = shift ? shift[d] : 0.0;
                const double dm = mu[gi * dim + d] - sh;
                c += log(var[gi * dim + d]) + dm * dm / var[gi * dim + d];
            }
            c = (logw[gi] - 0.5 * c) * 1.4426950408889634;
            return (c > -1e30) ? c : -1e30;               // log w = -inf
        };
        double cmax = -1e30;
        if (j < n_states)
            for (int mm_ = 0; mm_ < n_mix; ++mm_) cmax = fmax(cmax, gconst(mm_));
        v = -1e30;
        const double cmax_f = (cmax > -1e30) ? (double)(float)cmax : cmax;      // what the epilogue adds: the float32 rounding of cmax
        if (j < n_states && m < n_mix) {
            const double c = gconst(m);
            v = (c > -1e30) ? c - cmax_f : -1e30;
        }
        if (m == 0) {
            float* ctab = reinterpret_cast<float*>(packed + sizeof(TcHeader) + (size_t)tile * g.tile_bytes + (size_t)2 * g.NT * g.Kp * 4);
            ctab[n / g.Mp] = (float)cmax;
        }
    }
    const float hi = tf32_round((float)v);
    const float lo = (float)(v - (double)hi);
    uint8_t* base = packed + sizeof(TcHeader) + (size_t)tile * g.tile_bytes;
    const uint32_t off = kmajor_off(n, k, g.Kp);
    *reinterpret_cast<float*>(base + off) = hi;
    *reinterpret_cast<float*>(base + (size_t)g.NT * g.Kp * 4 + off) = lo;
}

// ------------------------------------------------------------------------------------------------ feature staging
// xt[tile][d < 40][row < 128] = feats[tile*128 + row][d] - shift[d]  (0 beyond n_frames / dim).  The W-resident main
// kernel makes every CTA stream ALL frames; reading row-major 39-float rows with thread = row costs 8 sectors per
// useful one out of L2.  One cheap pass to this tile-transposed copy (L2-resident: 160 B/frame) makes the
// producers' loads perfectly coalesced.
constexpr int TC_XT_D = 40;
constexpr float TC_X_LIMIT = 4096.f;      // |x - shift| beyond this (or inf / NaN): the frame goes to the exact fix-up path
__global__ void __launch_bounds__(256) gmm_tc_stage_kernel(const float* __restrict__ feats, int ld_feats, int64_t n_frames, int dim,
                                                           const uint8_t* __restrict__ packed, float* __restrict__ xt,
                                                           uint8_t* __restrict__ flags) {
    __shared__ float tile[TC_BM][TC_XT_D + 1];
    __shared__ int s_flag[TC_BM];
    if (threadIdx.x < TC_BM) s_flag[threadIdx.x] = 0;
    const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
    const int64_t f0 = (int64_t)blockIdx.x * TC_BM;
    const int rows = (int)min((int64_t)TC_BM, n_frames - f0);
    for (int i = threadIdx.x; i < TC_BM * TC_XT_D; i += 256) {
        const int r = i / TC_XT_D, d = i - r * TC_XT_D;
        tile[r][d] = 0.f;
    }
    __syncthreads();
    if (ld_feats == dim) {                                   // rows are contiguous: one flat coalesced copy
        const float* src = feats + f0 * ld_feats;
        for (int i = threadIdx.x; i < rows * dim; i += 256) {
            const int r = i / dim, d = i - r * dim;
            tile[r][d] = __ldg(src + i) - hdr->shift[d];
        }
    } else {
        for (int i = threadIdx.x; i < rows * dim; i += 256) {
            const int r = i / dim, d = i - r * dim;
            tile[r][d] = __ldg(feats + (f0 + r) * ld_feats + d) - hdr->shift[d];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * dim; i += 256) {
        const int r = i / dim, d = i - r * dim;
        if (!(fabsf(tile[r][d]) <= TC_X_LIMIT)) s_flag[r] = 1;              // also inf and NaN
    }
    __syncthreads();
    if (threadIdx.x < rows) flags[f0 + threadIdx.x] = (uint8_t)s_flag[threadIdx.x];
    float* dst = xt + (int64_t)blockIdx.x * TC_XT_D * TC_BM;
    for (int i = threadIdx.x; i < TC_XT_D * TC_BM; i += 256) {
        const int d = i / TC_BM, r = i - d * TC_BM;
        dst[i] = s_flag[r] ? 0.f : tile[r][d];
    }
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TcArgs {
    const float* xt; int64_t n_frames;
    const uint8_t* packed; float* b_out; int64_t ld_b;
    int n_states, dim, Mp, Kp, SPT, NT, n_wtiles, n_fsplit;
    int64_t frames_per_split;
    uint32_t swap_lbo_sbo;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MP>
__device__ __forceinline__ void lse_store(const float* v, const float* c_state, float* out_row, int st_local0, int st_global0, int st_end) {
    // 16 accumulator columns = 16/MP states, MP mixtures each, base-2 domain
#pragma unroll
    for (int g = 0; g < 16 / MP; ++g) {
        float mx = v[g * MP];
#pragma unroll
        for (int i = 1; i + 1 < MP; i += 2) mx = fmax3(mx, v[g * MP + i], v[g * MP + i + 1]);
        if ((MP & 1) == 0 && MP > 1) mx = fmaxf(mx, v[g * MP + MP - 1]);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MP; ++i) s += ex2_approx(v[g * MP + i] - mx);
        const float bits = (mx + lg2_approx(s)) + c_state[st_local0 + g];
        const float r = fmaf(bits, TC_LN2, bits * TC_LN2_LO);
        if (st_global0 + g < st_end) out_row[st_global0 + g] = r;
    }
}

constexpr int TC_EPI_WARPS = 8, TC_WARP_LO = 8, TC_WARP_HI = 12, TC_WARP_MMA = 16;

__global__ void __launch_bounds__(TC_THREADS, 1) gmm_tc_kernel(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int Kp = TC_KP, KH = TC_KH;
    const int NT = a.NT;
    const uint32_t w_part = (uint32_t)NT * Kp * 4;
    const uint32_t slot_bytes = TC_BM * KH * 4;
    const uint32_t tile_bytes = 2 * w_part + (uint32_t)((a.SPT * 4 + 15) / 16 * 16);
    uint8_t* s_w = tc_smem;                       // [W hi | W lo | c_state], the packed tile image verbatim
    const float* s_c = reinterpret_cast<const float*>(tc_smem + 2 * w_part);
    uint8_t* s_a = tc_smem + tile_bytes;          // 4 ring slots: lo-half0, lo-half1, hi-half0, hi-half1
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + 4 * slot_bytes);
    uint64_t* a_full = bars;                      // [4] producers -> MMA
    uint64_t* a_empty = bars + 4;                 // [4] MMA (commit) -> producers
    uint64_t* t_full = bars + 8;                  // [2] MMA (commit) -> epilogue
    uint64_t* t_empty = bars + 10;                // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, TC_EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == TC_WARP_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t idesc = make_idesc_tf32(TC_BM, NT);
    // descriptor strides of the K-major no-swizzle images (kmajor_off): next K chunk 128 B, next 8 rows (K/4)*128 B
    uint32_t a_lbo = 128, a_sbo = (uint32_t)(KH / 4) * 128, b_lbo = 128, b_sbo = (uint32_t)(Kp / 4) * 128;
    if (a.swap_lbo_sbo) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }

    uint32_t tile_cnt = 0;      // tiles processed so far by this role (every role advances it identically)
    int cur_wtile = -1;
    const int n_items = a.n_wtiles * a.n_fsplit;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int wt = item % a.n_wtiles, fsp = item / a.n_wtiles;
        const int64_t fs = (int64_t)fsp * a.frames_per_split;
        const int64_t fe = min(a.n_frames, fs + a.frames_per_split);
        if (fs >= fe) continue;
        if (wt != cur_wtile) {
            // (re)load the resident W tile: plain 16-byte copies, then make them visible to the tensor-core proxy
            __syncthreads();
            const uint4* src = reinterpret_cast<const uint4*>(a.packed + sizeof(TcHeader) + (size_t)wt * tile_bytes);
            uint4* dst = reinterpret_cast<uint4*>(s_w);
            for (uint32_t i = tid; i < tile_bytes / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
            fence_proxy_async();
            __syncthreads();
            cur_wtile = wt;
        }
        const int n_tiles = (int)((fe - fs + TC_BM - 1) / TC_BM);

        if (warp >= TC_WARP_LO && warp < TC_WARP_MMA) {
            // ===================== producers: thread = frame row; group 0 writes the lo slots (0,1), group 1 the hi slots (2,3)
            const bool hi_group = warp >= TC_WARP_HI;
            const int r = tid - (hi_group ? TC_WARP_HI : TC_WARP_LO) * 32;
            const uint32_t row_off = (uint32_t)((r >> 3) * (KH >> 2) * 128 + (r & 7) * 16);
            for (int t = 0; t < n_tiles; ++t) {
                const int64_t f = fs + (int64_t)t * TC_BM + r;
                // coalesced: lane = row inside the tile-transposed staging copy (already centred, zero padded)
                float x[TC_MAX_DIM];
                const float* xr = a.xt + (f / TC_BM) * (int64_t)(TC_XT_D * TC_BM) + r;
#pragma unroll
                for (int d = 0; d < TC_MAX_DIM; ++d) x[d] = __ldg(xr + d * TC_BM);
                const uint32_t par = ((tile_cnt + t) & 1) ^ 1;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int s = (hi_group ? 2 : 0) + half;
                    mbar_wait(a_empty + s, par);
                    uint8_t* slot = s_a + s * slot_bytes + row_off;
#pragma unroll
                    for (int cc = 0; cc < KH / 4; ++cc) {
                        float q[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = half * KH + cc * 4 + e;          // compile-time after unrolling
                            float v;                                        // K axis: [x0^2, x0, x1^2, x1, ... | 1 | 0]
                            if (k < 2 * TC_MAX_DIM) v = (k & 1) ? x[k >> 1] : x[k >> 1] * x[k >> 1];
                            else v = (k == 2 * TC_MAX_DIM) ? 1.f : 0.f;
                            const float hi = tf32_round(v);
                            q[e] = hi_group ? hi : (v - hi);
                        }
                        *reinterpret_cast<float4*>(slot + cc * 128) = make_float4(q[0], q[1], q[2], q[3]);
                    }
                    fence_proxy_async();
                    mbar_arrive(a_full + s);
                }
            }
        } else if (warp == TC_WARP_MMA) {
            // ===================== MMA issuer: one thread
            if (lane == 0) {
                const uint32_t wh = smem_u32(s_w), wl = wh + w_part;
                const uint32_t kadv_a = 2 * (a.swap_lbo_sbo ? a_sbo : a_lbo), kadv_b = 2 * (a.swap_lbo_sbo ? b_sbo : b_lbo);
                for (int t = 0; t < n_tiles; ++t) {
                    const uint32_t tc = tile_cnt + t;
                    const uint32_t stage = tc & 1;
                    mbar_wait(t_empty + stage, ((tc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + stage * (uint32_t)NT;
                    // Order: all correction products first (Al.Bh from slots 0,1; Ah.Bl from slots 2,3), the main
                    // Ah.Bh products last.  The tensor core truncates when it aligns fp32 addends, so the fewer
                    // accumulations happen at full magnitude the smaller the (systematic) error: 10 instead of 30.
                    uint32_t acc = 0;
                    auto issue = [&](int s, uint32_t wbase) {
                        const int half = s & 1;
                        const uint32_t sa = smem_u32(s_a + s * slot_bytes);
                        for (int j = 0; j < KH / 8; ++j) {
                            const int gk = half * (KH / 8) + j;
                            umma_tf32(d_tmem, make_smem_desc(sa + j * kadv_a, a_lbo, a_sbo, 0),
                                      make_smem_desc(wbase + gk * kadv_b, b_lbo, b_sbo, 0), idesc, acc);
                            acc = 1;
                        }
                    };
                    for (int s = 0; s < 2; ++s) {                       // Al . Bh
                        mbar_wait(a_full + s, tc & 1);
                        tc_fence_after();
                        issue(s, wh);
                        umma_commit(a_empty + s);                       // lo slots are free again
                    }
                    for (int s = 2; s < 4; ++s) {                       // Ah . Bl
                        mbar_wait(a_full + s, tc & 1);
                        tc_fence_after();
                        issue(s, wl);
                    }
                    for (int s = 2; s < 4; ++s) {                       // Ah . Bh
                        issue(s, wh);
                        umma_commit(a_empty + s);
                    }
                    umma_commit(t_full + stage);        // accumulator stage complete
                }
            }
            __syncwarp();
        } else {
            // ===================== epilogue: warps w and w+4 read TMEM lanes 32(w&3).. (= frame rows), half of the columns each
            const int quad = warp & 3, chalf = warp >> 2;
            const int row = quad * 32 + lane;
            const int st_tile0 = wt * a.SPT;
            const int st_end = min(a.n_states, st_tile0 + a.SPT);
            const int n16 = NT / 16;
            const int c16_begin = chalf ? (n16 + 1) / 2 : 0, c16_end = chalf ? n16 : (n16 + 1) / 2;
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t tc = tile_cnt + t;
                const uint32_t stage = tc & 1;
                mbar_wait(t_full + stage, (tc >> 1) & 1);
                tc_fence_after();
                const int64_t f = fs + (int64_t)t * TC_BM + row;
                float* out_row = a.b_out + f * a.ld_b;
                const int se = (f < fe) ? st_end : 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + stage * (uint32_t)NT;
                for (int c16 = c16_begin; c16 < c16_end; ++c16) {
                    float v[16];
                    tmem_ld16(taddr + c16 * 16, v);
                    const int sl0 = (c16 * 16) / a.Mp;
                    switch (a.Mp) {
                        case 16: lse_store<16>(v, s_c, out_row, sl0, st_tile0 + sl0, se); break;
                        case 8: lse_store<8>(v, s_c, out_row, sl0, st_tile0 + sl0, se); break;
                        case 4: lse_store<4>(v, s_c, out_row, sl0, st_tile0 + sl0, se); break;
                        case 2: lse_store<2>(v, s_c, out_row, sl0, st_tile0 + sl0, se); break;
                        default: lse_store<1>(v, s_c, out_row, sl0, st_tile0 + sl0, se); break;
                    }
                }
                tc_fence_before();
                mbar_arrive(t_empty + stage);
            }
        }
        tile_cnt += n_tiles;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) tmem_dealloc(tmem_base, 512);
}

static uint32_t g_umma_swap = 0;
static uint32_t g_probe_second_off = 0;

}  // namespace lnasr

using namespace lnasr;

extern "C" int lnasr_debug_set_umma_swap(int swap) { g_umma_swap = swap ? 1u : 0u; return LNASR_OK; }
extern "C" int lnasr_debug_set_probe_second_mma(uint32_t byte_offset) { g_probe_second_off = byte_offset; return LNASR_OK; }

extern "C" int lnasr_debug_umma_probe(const float* a_image, int32_t a_bytes, const float* b_image, int32_t b_bytes, uint32_t a_lbo,
                                      uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t layout_type, uint32_t idesc,
                                      float* d_out, void* stream) {
    LNASR_CHECK_ARG(a_image && b_image && d_out, "umma_probe: null pointer");
    LNASR_CHECK_ARG(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes <= 128 * 1024,
                    "umma_probe: image sizes must be multiples of 16 and fit 128 KB");
    const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 1024;
    LNASR_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a_image, a_bytes, b_image, b_bytes, a_lbo, a_sbo, b_lbo, b_sbo,
                                                             layout_type, idesc, d_out, g_probe_second_off);
    LNASR_LAUNCH_CHECK();
    return LNASR_OK;
}

extern "C" int lnasr_gmm_tc_packed_bytes(int32_t n_states, int32_t n_mix, int32_t dim, size_t* bytes) {
    TcGeom g;
    LNASR_CHECK_ARG(bytes, "gmm_tc_packed_bytes: null pointer");
    LNASR_CHECK_ARG(tc_geometry(n_states, n_mix, dim, &g), "gmm_tc: need n_mix <= 16 and dim <= %d (got n_mix=%d dim=%d)", TC_MAX_DIM, n_mix, dim);
    *bytes = g.total_bytes;
    return LNASR_OK;
}

extern "C" int lnasr_gmm_pack_tc(const double* logw, const double* mu, const double* var, int32_t n_states, int32_t n_mix,
                                 int32_t dim, const double* shift, void* packed, void* stream) {
    TcGeom g;
    LNASR_CHECK_ARG(logw && mu && var && packed, "gmm_pack_tc: null pointer");
    LNASR_CHECK_ARG(tc_geometry(n_states, n_mix, dim, &g), "gmm_tc: need n_mix <= 16 and dim <= %d (got n_mix=%d dim=%d)", TC_MAX_DIM, n_mix, dim);
    const int64_t n = (int64_t)g.n_wtiles * g.NT * g.Kp;
    gmm_pack_tc_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(logw, mu, var, n_states, n_mix, dim, shift, g,
                                                                                     (uint8_t*)packed);
    LNASR_LAUNCH_CHECK();
    return tc_pack_tail(logw, mu, var, n_states, n_mix, dim, reinterpret_cast<double*>((uint8_t*)packed + g.tail_off), (cudaStream_t)stream);
}

extern "C" int lnasr_gmm_tc_workspace_bytes(int64_t n_frames, size_t* bytes) {
    LNASR_CHECK_ARG(bytes && n_frames >= 0, "gmm_tc_workspace_bytes: bad arguments");
    // per 128-frame tile: the TF32 kernel stages 40 x 128 floats (20 480 B), the FP16-slice kernel its two operand images
    // (2 x 128 x 80 halves = 40 960 B); one size serves both
    // + one flag byte per frame (frames handed to the exact fix-up path)
    *bytes = (size_t)ceil_div(n_frames, TC_BM) * TC_BM * TC_XT_D * sizeof(float) * 2 + 256 + tc_flag_bytes(n_frames);
    return LNASR_OK;
}

extern "C" int lnasr_gmm_loglik_tc(const float* feats, int32_t ld_feats, int64_t n_frames, const void* packed, int32_t n_states,
                                   int32_t n_mix, int32_t dim, float* b_out, int64_t ld_b, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TcGeom g;
    LNASR_CHECK_ARG(feats && packed && b_out, "gmm_loglik_tc: null pointer");
    LNASR_CHECK_ARG(tc_geometry(n_states, n_mix, dim, &g), "gmm_tc: need n_mix <= 16 and dim <= %d (got n_mix=%d dim=%d)", TC_MAX_DIM, n_mix, dim);
    LNASR_CHECK_ARG(ld_feats >= dim && ld_b >= n_states, "gmm_loglik_tc: bad leading dimensions");
    if (n_frames == 0) return LNASR_OK;
    size_t need = 0;
    lnasr_gmm_tc_workspace_bytes(n_frames, &need);
    if (!workspace || workspace_bytes < need) {
        set_error("gmm_loglik_tc: workspace %zu < %zu bytes", workspace_bytes, need);
        return LNASR_EWORKSPACE;
    }
    float* xt = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int64_t frame_tiles = ceil_div(n_frames, TC_BM);
    uint8_t* flags = reinterpret_cast<uint8_t*>(xt) + (size_t)frame_tiles * TC_BM * TC_XT_D * sizeof(float) * 2;
    gmm_tc_stage_kernel<<<(unsigned)frame_tiles, 256, 0, stream>>>(feats, ld_feats, n_frames, dim, (const uint8_t*)packed, xt, flags);
    LNASR_LAUNCH_CHECK();
    TcArgs a{};
    a.xt = xt; a.n_frames = n_frames; a.packed = (const uint8_t*)packed; a.b_out = b_out; a.ld_b = ld_b;
    a.n_states = n_states; a.dim = dim; a.Mp = g.Mp; a.Kp = g.Kp; a.SPT = g.SPT; a.NT = g.NT; a.n_wtiles = g.n_wtiles;
    const int sms = sm_count();
    int n_fsplit = (int)std::max<int64_t>(1, std::min<int64_t>(frame_tiles, sms / std::max(1, g.n_wtiles)));
    const int64_t tiles_per_split = ceil_div(frame_tiles, n_fsplit);
    n_fsplit = (int)ceil_div(frame_tiles, tiles_per_split);
    a.n_fsplit = n_fsplit;
    a.frames_per_split = tiles_per_split * TC_BM;
    a.swap_lbo_sbo = g_umma_swap;
    const int grid = std::min(sms, g.n_wtiles * n_fsplit);
    LNASR_CUDA(cudaFuncSetAttribute(gmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    gmm_tc_kernel<<<grid, TC_THREADS, g.smem_bytes, stream>>>(a);
    LNASR_LAUNCH_CHECK();
    return tc_fixup(feats, ld_feats, n_frames, flags, reinterpret_cast<const double*>((const uint8_t*)packed + g.tail_off), n_states, n_mix,
                    dim, b_out, ld_b, stream);
}
