// HBM-bound fused kernels of the hdiff_b200 hot path (vectorised 16-byte accesses, warp/shared
// reductions, no re-reads beyond the algorithmic minimum):
//   K4  GroupNorm statistics / GroupNorm-apply + Swish (+ dropout) forward and backward
//       (reference: nn.GroupNorm + Swish + nn.Dropout, DiffusionFreeGuidence/ModelCondition.py:128-129,141-143,95,249-250)
//   K6  q_sample and noise-MSE forward/backward (DiffusionCondition.py:43-45)
//   K7  CFG mix + posterior mean + noise step (+ NaN flag, final clip) (DiffusionCondition.py:74-80,91-98)
//   colsum (bias / embedding-add gradients), fused global-norm clip + AdamW on the flat buffers
//       (TrainCondition.py:61-63).
#include "hd_common.cuh"

// second-generation bf16 GroupNorm kernels (hd_gn.cu); the generic templates below serve the fp32 check mode, shapes outside
// the new kernels' mapping and HDIFF_GN_V1=1
int hd_gn2_supported(int C0, int C1, int G, int64_t HW, int N);
int hd_gn2_stats(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, double* sums, cudaStream_t st);
int hd_gn2_apply(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                 const float* beta, float eps, int act, float p_drop, uint64_t seed, void* out, cudaStream_t st);
int hd_gn2_bwd_reduce(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                      const float* beta, float eps, int act, float p_drop, uint64_t seed, const void* dy, double* gsums,
                      float* dgamma, float* dbeta, cudaStream_t st);
int hd_gn2_bwd_apply(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                     const float* beta, float eps, int act, float p_drop, uint64_t seed, const void* dy, const double* gsums,
                     const void* add, const void* acc0, const void* acc1, void* dx0, void* dx1, float* cs_total, float* cs_per_n,
                     int64_t cs_ld, int cs_n, cudaStream_t st);
extern "C" int hd_gn_v2(int dtype, int C0, int C1, int G, int64_t HW, int N) { return dtype == HD_BF16 && hd_gn2_supported(C0, C1, G, HW, N); }

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int N = 4; using raw = float4; static constexpr bool fast = false; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; using raw = uint4; static constexpr bool fast = true; };

template <typename T> __device__ __forceinline__ void vec_load(const T* p, float* v);
template <> __device__ __forceinline__ void vec_load<float>(const float* p, float* v) {
    float4 r = __ldg(reinterpret_cast<const float4*>(p)); v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> __device__ __forceinline__ void vec_load<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));      // read-only path: loads may be hoisted above earlier stores
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// raw (still packed) vector loads: issued in batches BEFORE any of them is unpacked, so that every thread keeps several
// 16-byte requests in flight (these kernels are latency-bound otherwise: ~16 KB in flight per SM against the ~45 KB
// that 6.5 TB/s needs)
template <typename T> __device__ __forceinline__ typename Vec<T>::raw raw_load(const T* p) {
    return __ldg(reinterpret_cast<const typename Vec<T>::raw*>(p));
}
__device__ __forceinline__ void unpack(const float4& r, float* v) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
__device__ __forceinline__ void unpack(const uint4& r, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void vec_store(T* p, const float* v);
template <> __device__ __forceinline__ void vec_store<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vec_store<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}

// Common description of a (possibly two-source) NHWC tensor [N][HW][C0 + C1].
template <typename T>
struct Src2 {
    const T* p0; const T* p1; int C0, C1;
    __device__ __forceinline__ const T* at(int n, int64_t pix, int64_t HW, int c) const {
        return c < C0 ? p0 + ((int64_t)n * HW + pix) * C0 + c : p1 + ((int64_t)n * HW + pix) * C1 + (c - C0);
    }
};

struct GnParams {
    int N; int64_t HW; int C, G;
    const double* sums;       // [N][G][2] (sum, sum of squares)
    const float* gamma; const float* beta; float eps;
    int act; float p_drop; uint64_t seed;
};

__device__ __forceinline__ void gn_load_stats(const GnParams& g, int n, float* s_mean, float* s_rstd) {
    const double cnt = (double)(g.C / g.G) * (double)g.HW;
    for (int i = threadIdx.x; i < g.G; i += blockDim.x) {
        double s = g.sums[((int64_t)n * g.G + i) * 2], ss = g.sums[((int64_t)n * g.G + i) * 2 + 1];
        double m = s / cnt, var = ss / cnt - m * m;
        if (var < 0) var = 0;
        s_mean[i] = (float)m; s_rstd[i] = (float)(1.0 / sqrt(var + (double)g.eps));
    }
}

// ------------------------------- statistics -------------------------------------------------
// grid (chunks, N).  thread <-> fixed channel vector, strides over the pixels of the chunk.
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(Src2<T> x, int N, int64_t HW, int C, int G, int64_t pix_per_block, double* sums) {
    constexpr int V = Vec<T>::N;
    __shared__ float sg[64][2];
    const int lanes = C / V, cpg = C / G;
    const int ppi = blockDim.x / lanes;           // pixels per iteration
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < G; i += blockDim.x) { sg[i][0] = 0.f; sg[i][1] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x % lanes, sub = threadIdx.x / lanes;
    float s[V], ss[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s[i] = 0.f; ss[i] = 0.f; }
    if (sub < ppi) {
        const int64_t p0 = blockIdx.x * pix_per_block;
        const int64_t p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
        constexpr int U = 4;
        for (int64_t p = p0 + sub; p < p1; p += (int64_t)U * ppi) {
            typename Vec<T>::raw xr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (p + (int64_t)u * ppi < p1) xr[u] = raw_load(x.at(n, p + (int64_t)u * ppi, HW, lane * V));
#pragma unroll
            for (int u = 0; u < U; ++u) if (p + (int64_t)u * ppi < p1) {
                float v[V]; unpack(xr[u], v);
#pragma unroll
                for (int i = 0; i < V; ++i) { s[i] += v[i]; ss[i] += v[i] * v[i]; }
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) { int g = (lane * V + i) / cpg; atomicAdd(&sg[g][0], s[i]); atomicAdd(&sg[g][1], ss[i]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        atomicAdd(sums + ((int64_t)n * G + i) * 2, (double)sg[i][0]);
        atomicAdd(sums + ((int64_t)n * G + i) * 2 + 1, (double)sg[i][1]);
    }
}

// blocks_per_sm x waves: the grid is sized to whole waves of resident blocks (a kernel that keeps 3 blocks per SM resident and
// is launched with 8 blocks per SM runs 2.67 waves: the last one is two-thirds empty)
static inline int64_t pick_chunk(int N, int64_t HW, int ppi, int* chunks_out, int blocks_per_sm = 8) {
    int64_t target_blocks = (int64_t)hd_num_sms() * blocks_per_sm;
    int64_t chunks = (target_blocks + N - 1) / N;
    int64_t max_chunks = (HW + ppi - 1) / ppi;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t ppb = (HW + chunks - 1) / chunks;
    ppb = (ppb + ppi - 1) / ppi * ppi;
    chunks = (HW + ppb - 1) / ppb;
    *chunks_out = (int)chunks;
    return ppb;
}

template <typename T>
static int gn_check(int C0, int C1, int G) {
    constexpr int V = Vec<T>::N;
    int C = C0 + C1;
    if (C % G != 0 || G > 64 || C % V != 0 || C0 % V != 0 || C / V > 256) { hd_set_error("groupnorm: unsupported channel configuration"); return HD_ERR_UNSUPPORTED; }
    return HD_OK;
}

template <typename T>
static int gn_stats_t(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, double* sums, cudaStream_t st) {
    int rc = gn_check<T>(C0, C1, G); if (rc) return rc;
    int C = C0 + C1;
    if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * G, st) != cudaSuccess) return HD_ERR_CUDA;
    int ppi = 256 / (C / Vec<T>::N), chunks;
    int64_t ppb = pick_chunk(N, HW, ppi, &chunks);
    gn_stats_kernel<T><<<dim3(chunks, N), 256, 0, st>>>(Src2<T>{(const T*)in0, (const T*)in1, C0, C1}, N, HW, C, G, ppb, sums);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_gn_stats(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, double* sums, cudaStream_t stream) {
    HD_REQUIRE(in0 && sums && N > 0 && HW > 0 && C0 > 0 && (C1 == 0 || in1));
    if (hd_gn_v2(dtype, C0, C1, G, HW, N)) return hd_gn2_stats(in0, C0, in1, C1, N, HW, G, sums, stream);
    if (dtype == HD_F32) return gn_stats_t<float>(in0, C0, in1, C1, N, HW, G, sums, stream);
    if (dtype == HD_BF16) return gn_stats_t<__nv_bfloat16>(in0, C0, in1, C1, N, HW, G, sums, stream);
    return HD_ERR_ARG;
}

// GroupNorm statistics from per-channel sums that the producing convolutions left behind (hd_conv_tc `chan_sums`):
// sums[n][g] = sum over the channels of group g of cs[n][c]; the (possibly two-source) channel axis is C0 | C1.
__global__ void gn_group_sums_kernel(const double* cs0, int C0, const double* cs1, int C1, int N, int G, double* sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * G) return;
    const int n = i / G, g = i - n * G, cpg = (C0 + C1) / G;
    double s = 0.0, ss = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const double* p = c < C0 ? cs0 + ((long long)n * C0 + c) * 2 : cs1 + ((long long)n * C1 + (c - C0)) * 2;
        s += p[0]; ss += p[1];
    }
    sums[(long long)i * 2] = s; sums[(long long)i * 2 + 1] = ss;
}
extern "C" int hd_gn_group_sums(const double* cs0, int C0, const double* cs1, int C1, int N, int G, double* sums, cudaStream_t stream) {
    HD_REQUIRE(cs0 && sums && N > 0 && G > 0 && C0 > 0 && (C1 == 0 || cs1) && (C0 + C1) % G == 0);
    gn_group_sums_kernel<<<(N * G + 127) / 128, 128, 0, stream>>>(cs0, C0, cs1, C1, N, G, sums);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ------------------------------- apply (forward) --------------------------------------------
// out[n][pix][c] = drop( act( (x - mean) * rstd * gamma + beta ) ); grid (chunks, N).
// thread <-> fixed channel vector (scale / shift hoisted out of the pixel loop: one FMA per element before the activation)
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(Src2<T> x, GnParams g, int64_t pix_per_block, T* out) {
    constexpr int V = Vec<T>::N;
    __shared__ float s_mean[64], s_rstd[64];
    const int n = blockIdx.y;
    gn_load_stats(g, n, s_mean, s_rstd);
    __syncthreads();
    const int lanes = g.C / V, cpg = g.C / g.G;
    const int ppi = blockDim.x / lanes;
    const int lane = threadIdx.x % lanes, sub = threadIdx.x / lanes;
    if (sub >= ppi) return;
    const int c0 = lane * V;
    float a[V], b[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = c0 + k, gi = c / cpg;
        a[k] = s_rstd[gi] * g.gamma[c];
        b[k] = g.beta[c] - s_mean[gi] * a[k];
    }
    const int64_t p0 = blockIdx.x * pix_per_block;
    const int64_t p1 = p0 + pix_per_block < g.HW ? p0 + pix_per_block : g.HW;
    const bool drop = g.p_drop > 0.f;
    constexpr int U = 4;
    for (int64_t pb = p0 + sub; pb < p1; pb += (int64_t)U * ppi) {
        typename Vec<T>::raw xr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (pb + (int64_t)u * ppi < p1) xr[u] = raw_load(x.at(n, pb + (int64_t)u * ppi, g.HW, c0));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t pix = pb + (int64_t)u * ppi;
            if (pix < p1) {
                float v[V]; unpack(xr[u], v);
                const int64_t obase = ((int64_t)n * g.HW + pix) * g.C + c0;
                float ds[V];
                if (drop) hd_dropout_vec<V>(g.seed, (uint64_t)obase, g.p_drop, ds);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    float z = fmaf(v[k], a[k], b[k]);
                    if (g.act) z = hd_swish(z);      // forward keeps the EX2 + RCP sigmoid: its error is absolute, not relative
                    if (drop) z *= ds[k];
                    v[k] = z;
                }
                vec_store(out + obase, v);
            }
        }
    }
}
template <typename T>
static int gn_apply_t(const void* in0, int C0, const void* in1, int C1, GnParams g, void* out, cudaStream_t st) {
    int rc = gn_check<T>(C0, C1, g.G); if (rc) return rc;
    int ppi = 256 / (g.C / Vec<T>::N), chunks;
    static int per_sm = 0;
    if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_apply_kernel<T>, 256, 0) != cudaSuccess || per_sm < 1)) per_sm = 2;
    int64_t ppb = pick_chunk(g.N, g.HW, ppi, &chunks, 3 * per_sm);
    gn_apply_kernel<T><<<dim3((unsigned)chunks, g.N), 256, 0, st>>>(Src2<T>{(const T*)in0, (const T*)in1, C0, C1}, g, ppb, (T*)out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_gn_apply(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums,
                           const float* gamma, const float* beta, float eps, int act, float p_drop, uint64_t seed, void* out,
                           cudaStream_t stream) {
    HD_REQUIRE(in0 && sums && gamma && beta && out && N > 0 && HW > 0 && (C1 == 0 || in1));
    if (hd_gn_v2(dtype, C0, C1, G, HW, N)) return hd_gn2_apply(in0, C0, in1, C1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, out, stream);
    GnParams g{N, HW, C0 + C1, G, sums, gamma, beta, eps, act, p_drop, hd_seed_mix(seed)};
    if (dtype == HD_F32) return gn_apply_t<float>(in0, C0, in1, C1, g, out, stream);
    if (dtype == HD_BF16) return gn_apply_t<__nv_bfloat16>(in0, C0, in1, C1, g, out, stream);
    return HD_ERR_ARG;
}

// ------------------------------- backward, reduction pass -----------------------------------
// dy' = dy * drop * act'(z).  Per channel: dgamma += sum dy' xhat, dbeta += sum dy'.
// Per (n, group): gsums = (sum gamma dy', sum gamma dy' xhat).
template <typename T>
__device__ __forceinline__ void gn_bwd_reduce_body(const Src2<T>& x, const GnParams& g, const T* dy, int64_t pix_per_block,
                                                   double* gsums, float* dgamma, float* dbeta, int n, int blk,
                                                   float* s_mean, float* s_rstd, float (*sg)[2], float* s_ch, T* dy_act = nullptr) {
    constexpr int V = Vec<T>::N;
    gn_load_stats(g, n, s_mean, s_rstd);
    for (int i = threadIdx.x; i < g.G; i += blockDim.x) { sg[i][0] = 0.f; sg[i][1] = 0.f; }
    for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) s_ch[i] = 0.f;
    __syncthreads();
    const int lanes = g.C / V, cpg = g.C / g.G;
    const int ppi = blockDim.x / lanes;
    const int lane = threadIdx.x % lanes, sub = threadIdx.x / lanes;
    if (sub < ppi) {
        const int c0 = lane * V;
        // z = v*A1 + B1 is the pre-activation; the sums kept per thread are sum(dy') and sum(dy' * v): xhat = v*rstd - mean*rstd
        // is applied once at the end (fewer live registers -> more loads in flight)
        float A1[V], B1[V], s1[V], s2[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int gi = (c0 + k) / cpg;
            A1[k] = s_rstd[gi] * g.gamma[c0 + k];
            B1[k] = g.beta[c0 + k] - s_mean[gi] * A1[k];
            s1[k] = 0.f; s2[k] = 0.f;
        }
        const bool drop = g.p_drop > 0.f;
        const int64_t p0 = blk * pix_per_block;
        const int64_t p1 = p0 + pix_per_block < g.HW ? p0 + pix_per_block : g.HW;
        constexpr int U = 4;
        for (int64_t pb = p0 + sub; pb < p1; pb += (int64_t)U * ppi) {
          typename Vec<T>::raw xr[U], dr[U];
#pragma unroll
          for (int u = 0; u < U; ++u) if (pb + (int64_t)u * ppi < p1) {
              xr[u] = raw_load(x.at(n, pb + (int64_t)u * ppi, g.HW, c0));
              dr[u] = raw_load(dy + ((int64_t)n * g.HW + pb + (int64_t)u * ppi) * g.C + c0);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) if (pb + (int64_t)u * ppi < p1) {
            const int64_t p = pb + (int64_t)u * ppi;
            float v[V], d[V];
            unpack(xr[u], v);
            const int64_t obase = ((int64_t)n * g.HW + p) * g.C + c0;
            unpack(dr[u], d);
            float ds[V];
            if (drop) hd_dropout_vec<V>(g.seed, (uint64_t)obase, g.p_drop, ds);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float dd = d[k];
                if (drop) dd *= ds[k];
                if (g.act) dd *= hd_swish_grad_t<Vec<T>::fast>(fmaf(v[k], A1[k], B1[k]));
                s1[k] += dd; s2[k] = fmaf(dd, v[k], s2[k]);
                d[k] = dd;
            }
            // dy' = dy * dropout mask * act'(z) handed to the apply pass (usually in place over dy), which then needs neither
            // the sigmoid nor the dropout hash: both passes are bound by their instruction count, not by HBM
            if (dy_act) vec_store(dy_act + obase, d);
          }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = c0 + k, gi = c / cpg;
            const float gam = g.gamma[c];
            const float sxh = s_rstd[gi] * (s2[k] - s_mean[gi] * s1[k]);       // sum(dy' * xhat)
            atomicAdd(&s_ch[c], sxh); atomicAdd(&s_ch[g.C + c], s1[k]);        // block-level first: one global atomic per channel per block
            atomicAdd(&sg[gi][0], gam * s1[k]); atomicAdd(&sg[gi][1], gam * sxh);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) { atomicAdd(dgamma + i, s_ch[i]); atomicAdd(dbeta + i, s_ch[g.C + i]); }
    for (int i = threadIdx.x; i < g.G; i += blockDim.x) {
        atomicAdd(gsums + ((int64_t)n * g.G + i) * 2, (double)sg[i][0]);
        atomicAdd(gsums + ((int64_t)n * g.G + i) * 2 + 1, (double)sg[i][1]);
    }
    __syncthreads();                              // the shared partials may be reused by the caller's next work item
}
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(Src2<T> x, GnParams g, const T* dy, int64_t pix_per_block,
                                                            double* gsums, float* dgamma, float* dbeta, T* dy_act) {
    __shared__ float s_mean[64], s_rstd[64], sg[64][2];
    extern __shared__ float s_ch[];               // [2][C]: per-channel (dgamma, dbeta) partials of this block
    gn_bwd_reduce_body<T>(x, g, dy, pix_per_block, gsums, dgamma, dbeta, blockIdx.y, blockIdx.x, s_mean, s_rstd, sg, s_ch, dy_act);
}
template <typename T>
static int gn_bwd_reduce_t(const void* in0, int C0, const void* in1, int C1, GnParams g, const void* dy, double* gsums,
                           float* dgamma, float* dbeta, void* dy_act, cudaStream_t st) {
    int rc = gn_check<T>(C0, C1, g.G); if (rc) return rc;
    if (cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * (size_t)g.N * g.G, st) != cudaSuccess) return HD_ERR_CUDA;
    int ppi = 256 / (g.C / Vec<T>::N), chunks;
    int64_t ppb = pick_chunk(g.N, g.HW, ppi, &chunks);
    gn_bwd_reduce_kernel<T><<<dim3(chunks, g.N), 256, 2 * g.C * sizeof(float), st>>>(Src2<T>{(const T*)in0, (const T*)in1, C0, C1}, g, (const T*)dy, ppb, gsums, dgamma, dbeta, (T*)dy_act);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_gn_bwd_reduce(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                                const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                                uint64_t seed, const void* dy, double* gsums, float* dgamma, float* dbeta, void* dy_act,
                                cudaStream_t stream) {
    HD_REQUIRE(in0 && sums && gamma && beta && dy && gsums && dgamma && dbeta && (C1 == 0 || in1));
    if (hd_gn_v2(dtype, C0, C1, G, HW, N)) {
        if (dy_act) { hd_set_error("hd_gn_bwd_reduce: the second-generation kernels do not hand dy' over (ask hd_gn_v2 first)"); return HD_ERR_ARG; }
        return hd_gn2_bwd_reduce(in0, C0, in1, C1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, dy, gsums, dgamma, dbeta, stream);
    }
    GnParams g{N, HW, C0 + C1, G, sums, gamma, beta, eps, act, p_drop, hd_seed_mix(seed)};
    if (dtype == HD_F32) return gn_bwd_reduce_t<float>(in0, C0, in1, C1, g, dy, gsums, dgamma, dbeta, dy_act, stream);
    if (dtype == HD_BF16) return gn_bwd_reduce_t<__nv_bfloat16>(in0, C0, in1, C1, g, dy, gsums, dgamma, dbeta, dy_act, stream);
    return HD_ERR_ARG;
}

// ------------------------------- backward, apply pass ---------------------------------------
// dx = rstd * (gamma dy' - A/m - xhat B/m) + add + acc ; written to two destination tensors.
// thread <-> fixed channel vector, per-channel constants hoisted out of the pixel loop.
template <typename T>
__device__ __forceinline__ void gn_bwd_apply_body(const Src2<T>& x, const GnParams& g, const T* dy, const double* gsums, const T* add,
                                                  const T* acc0, const T* acc1, T* dx0, T* dx1, int64_t pix_per_block, int n, int blk,
                                                  float* s_mean, float* s_rstd, float* s_a, float* s_b,
                                                  float* cs_total, float* cs_per_n, int64_t cs_ld, int cs_n, float* s_cs,
                                                  bool dy_is_act = false) {
    constexpr int V = Vec<T>::N;
    const bool want_cs = cs_total || cs_per_n;     // column sums of dx: the bias / embedding-add gradients of the producing conv
    if (want_cs) for (int i = threadIdx.x; i < g.C; i += blockDim.x) s_cs[i] = 0.f;
    gn_load_stats(g, n, s_mean, s_rstd);
    const double m = (double)(g.C / g.G) * (double)g.HW;
    for (int i = threadIdx.x; i < g.G; i += blockDim.x) {
        s_a[i] = (float)(__ldcg(gsums + ((int64_t)n * g.G + i) * 2) / m);        // L2: written by atomics (maybe of this launch)
        s_b[i] = (float)(__ldcg(gsums + ((int64_t)n * g.G + i) * 2 + 1) / m);
    }
    __syncthreads();
    const int lanes = g.C / V, cpg = g.C / g.G;
    const int ppi = blockDim.x / lanes;
    const int lane = threadIdx.x % lanes, sub = threadIdx.x / lanes;
    if (sub < ppi) {
    const int c0 = lane * V;
    // z = v*A1 + B1 (pre-activation);  dx = A1*dy' - C1 - v*D1   with xhat = v*rstd + mr folded into C1 / D1
    float A1[V], B1[V], C1[V], D1[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = c0 + k, gi = c / cpg;
        const float rs = s_rstd[gi], mr = -s_mean[gi] * rs;
        A1[k] = rs * g.gamma[c];
        B1[k] = fmaf(mr, g.gamma[c], g.beta[c]);
        C1[k] = rs * (s_a[gi] + mr * s_b[gi]);
        D1[k] = rs * rs * s_b[gi];
    }
    const bool first = c0 < x.C0;
    const int Cd = first ? x.C0 : x.C1, cd = first ? c0 : c0 - x.C0;
    const T* acc = first ? acc0 : acc1;
    T* dx = first ? dx0 : dx1;
    const int64_t p0 = blk * pix_per_block;
    const int64_t p1 = p0 + pix_per_block < g.HW ? p0 + pix_per_block : g.HW;
    const bool drop = g.p_drop > 0.f && !dy_is_act;            // dy_is_act: dy already carries the dropout mask and act'(z)
    const bool act = g.act && !dy_is_act;
    float cs[V];
#pragma unroll
    for (int k = 0; k < V; ++k) cs[k] = 0.f;
    // software pipeline: the loads of pixel i+1 are in flight while pixel i is processed (two named register buffers).
    // Measured: -14 % on this kernel; the same restructuring made the reduce / stats / forward-apply kernels slower
    // (register pressure), so those keep the batched-load form.
    using Raw = typename Vec<T>::raw;
    auto load = [&](int64_t pix, Raw& xr, Raw& dr, Raw& ar, Raw& cr) {
        xr = raw_load(x.at(n, pix, g.HW, c0));
        dr = raw_load(dy + ((int64_t)n * g.HW + pix) * g.C + c0);
        if (add) ar = raw_load(add + ((int64_t)n * g.HW + pix) * g.C + c0);
        if (acc) cr = raw_load(acc + ((int64_t)n * g.HW + pix) * Cd + cd);
    };
    auto process = [&](int64_t pix, const Raw& xr, const Raw& dr, const Raw& ar, const Raw& cr) {
        float v[V], d[V], r[V];
        unpack(xr, v);
        const int64_t obase = ((int64_t)n * g.HW + pix) * g.C + c0;
        unpack(dr, d);
        float ds[V];
        if (drop) hd_dropout_vec<V>(g.seed, (uint64_t)obase, g.p_drop, ds);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            float dd = d[k];
            if (drop) dd *= ds[k];
            if (act) dd *= hd_swish_grad_t<Vec<T>::fast>(fmaf(v[k], A1[k], B1[k]));
            r[k] = fmaf(A1[k], dd, -fmaf(v[k], D1[k], C1[k]));
        }
        if (add) { float t[V]; unpack(ar, t);
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] += t[k]; }
        const int64_t o = ((int64_t)n * g.HW + pix) * Cd + cd;
        if (acc) { float t[V]; unpack(cr, t);
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] += t[k]; }
        vec_store(dx + o, r);
#pragma unroll
        for (int k = 0; k < V; ++k) cs[k] += r[k];
    };
    {
        Raw xa, da, aa, ca, xb, db, ab, cb;
        int64_t pix = p0 + sub;
        if (pix < p1) load(pix, xa, da, aa, ca);
        while (pix < p1) {
            const int64_t pix1 = pix + ppi, pix2 = pix + 2 * (int64_t)ppi;
            if (pix1 < p1) load(pix1, xb, db, ab, cb);
            process(pix, xa, da, aa, ca);
            if (pix1 >= p1) break;
            if (pix2 < p1) load(pix2, xa, da, aa, ca);
            process(pix1, xb, db, ab, cb);
            pix = pix2;
        }
    }
    if (want_cs) {
#pragma unroll
        for (int k = 0; k < V; ++k) atomicAdd(&s_cs[c0 + k], cs[k]);
    }
    }
    __syncthreads();                              // the shared statistics may be reused by the caller's next work item
    if (want_cs) {
        for (int i = threadIdx.x; i < cs_n; i += blockDim.x) {      // only the leading cs_n channels (the first source tensor)
            if (cs_per_n) atomicAdd(cs_per_n + (int64_t)n * cs_ld + i, s_cs[i]);
            if (cs_total) atomicAdd(cs_total + i, s_cs[i]);
        }
        __syncthreads();
    }
}
template <typename T>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(Src2<T> x, GnParams g, const T* dy, const double* gsums, const T* add,
                                                           const T* acc0, const T* acc1, T* dx0, T* dx1, int64_t pix_per_block,
                                                           float* cs_total, float* cs_per_n, int64_t cs_ld, int cs_n, int dy_is_act) {
    __shared__ float s_mean[64], s_rstd[64], s_a[64], s_b[64];
    extern __shared__ float s_cs[];               // [C] column sums of dx of this block
    gn_bwd_apply_body<T>(x, g, dy, gsums, add, acc0, acc1, dx0, dx1, pix_per_block, blockIdx.y, blockIdx.x, s_mean, s_rstd, s_a, s_b,
                         cs_total, cs_per_n, cs_ld, cs_n, s_cs, dy_is_act != 0);
}

// ------------------------------- backward, both passes in one launch ------------------------
// The reduction and the apply pass both read x and dy.  Walking the batch in groups of images whose x + dy fit the L2
// (all CTAs reduce a group, meet at a grid barrier, then apply to the same group) turns the second read into L2 hits:
// HBM traffic drops from 5 tensor passes to 3.  Cooperative launch (all CTAs co-resident), one barrier per group.
__device__ __forceinline__ void hd_grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    __syncthreads();
}
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_fused_kernel(Src2<T> x, GnParams g, const T* dy, double* gsums, float* dgamma, float* dbeta,
                                                           const T* add, const T* acc0, const T* acc1, T* dx0, T* dx1,
                                                           int64_t pix_per_block, int chunks, int group, unsigned* counter) {
    __shared__ float s_mean[64], s_rstd[64], sg[64][2], s_a[64], s_b[64];
    extern __shared__ float s_ch[];
    unsigned round = 0;
    for (int n0 = 0; n0 < g.N; n0 += group) {
        const int nimg = g.N - n0 < group ? g.N - n0 : group;
        for (int vb = blockIdx.x; vb < nimg * chunks; vb += gridDim.x)
            gn_bwd_reduce_body<T>(x, g, dy, pix_per_block, gsums, dgamma, dbeta, n0 + vb / chunks, vb % chunks, s_mean, s_rstd, sg, s_ch);
        hd_grid_barrier(counter, ++round * gridDim.x);
        for (int vb = blockIdx.x; vb < nimg * chunks; vb += gridDim.x)
            gn_bwd_apply_body<T>(x, g, dy, gsums, add, acc0, acc1, dx0, dx1, pix_per_block, n0 + vb / chunks, vb % chunks, s_mean, s_rstd, s_a, s_b,
                                 nullptr, nullptr, 0, 0, s_ch);
    }
}
template <typename T>
static int gn_bwd_fused_t(const void* in0, int C0, const void* in1, int C1, GnParams g, const void* dy, double* gsums, float* dgamma,
                          float* dbeta, const void* add, const void* acc0, const void* acc1, void* dx0, void* dx1, unsigned* counter,
                          cudaStream_t st) {
    int rc = gn_check<T>(C0, C1, g.G); if (rc) return rc;
    if (cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * (size_t)g.N * g.G, st) != cudaSuccess) return HD_ERR_CUDA;
    if (cudaMemsetAsync(counter, 0, sizeof(unsigned), st) != cudaSuccess) return HD_ERR_CUDA;
    const size_t dsm = 2 * g.C * sizeof(float);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_bwd_fused_kernel<T>, 256, dsm) != cudaSuccess || per_sm < 1) {
        hd_set_error("gn_bwd_fused: occupancy query failed"); return HD_ERR_CUDA;
    }
    const int grid = per_sm * hd_num_sms();
    // images per group: x + dy of the group within ~48 MB of the 126 MB L2
    const int64_t per_img = 2 * g.HW * g.C * (int64_t)sizeof(T);
    int group = (int)((48ll << 20) / per_img); if (group < 1) group = 1; if (group > g.N) group = g.N;
    const int ppi = 256 / (g.C / Vec<T>::N);
    int chunks = grid / group; if (chunks < 1) chunks = 1;
    int64_t ppb = (g.HW + chunks - 1) / chunks; ppb = (ppb + ppi - 1) / ppi * ppi;
    chunks = (int)((g.HW + ppb - 1) / ppb);
    Src2<T> x{(const T*)in0, (const T*)in1, C0, C1};
    const T* dyp = (const T*)dy; const T* addp = (const T*)add; const T* a0 = (const T*)acc0; const T* a1 = (const T*)acc1;
    T* d0 = (T*)dx0; T* d1 = (T*)dx1;
    void* args[] = {&x, &g, &dyp, &gsums, &dgamma, &dbeta, &addp, &a0, &a1, &d0, &d1, &ppb, &chunks, &group, &counter};
    if (cudaLaunchCooperativeKernel((const void*)gn_bwd_fused_kernel<T>, dim3(grid), dim3(256), args, dsm, st) != cudaSuccess) {
        hd_set_error(cudaGetErrorString(cudaGetLastError())); return HD_ERR_CUDA;
    }
    return HD_OK;
}
// gsums: [N][G][2] fp64 scratch, counter: one zero-initialisable unsigned of scratch (grid barrier)
extern "C" int hd_gn_bwd_fused(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                               const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                               uint64_t seed, const void* dy, double* gsums, float* dgamma, float* dbeta, const void* add,
                               const void* acc0, const void* acc1, void* dx0, void* dx1, unsigned* counter, cudaStream_t stream) {
    HD_REQUIRE(in0 && sums && gamma && beta && dy && gsums && dgamma && dbeta && dx0 && counter && (C1 == 0 || (in1 && dx1)));
    GnParams g{N, HW, C0 + C1, G, sums, gamma, beta, eps, act, p_drop, hd_seed_mix(seed)};
    if (dtype == HD_F32) return gn_bwd_fused_t<float>(in0, C0, in1, C1, g, dy, gsums, dgamma, dbeta, add, acc0, acc1, dx0, dx1, counter, stream);
    if (dtype == HD_BF16) return gn_bwd_fused_t<__nv_bfloat16>(in0, C0, in1, C1, g, dy, gsums, dgamma, dbeta, add, acc0, acc1, dx0, dx1, counter, stream);
    return HD_ERR_ARG;
}
template <typename T>
static int gn_bwd_apply_t(const void* in0, int C0, const void* in1, int C1, GnParams g, const void* dy, const double* gsums,
                          const void* add, const void* acc0, const void* acc1, void* dx0, void* dx1, float* cs_total, float* cs_per_n,
                          int64_t cs_ld, int cs_n, int dy_is_act, cudaStream_t st) {
    int rc = gn_check<T>(C0, C1, g.G); if (rc) return rc;
    int ppi = 256 / (g.C / Vec<T>::N), chunks;
    int64_t ppb = pick_chunk(g.N, g.HW, ppi, &chunks);
    gn_bwd_apply_kernel<T><<<dim3((unsigned)chunks, g.N), 256, g.C * sizeof(float), st>>>(Src2<T>{(const T*)in0, (const T*)in1, C0, C1}, g, (const T*)dy, gsums,
                                                                  (const T*)add, (const T*)acc0, (const T*)acc1, (T*)dx0, (T*)dx1, ppb,
                                                                  cs_total, cs_per_n, cs_ld, cs_n, dy_is_act);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_gn_bwd_apply(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                               const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                               uint64_t seed, const void* dy, const double* gsums, const void* add, const void* acc0,
                               const void* acc1, void* dx0, void* dx1, float* cs_total, float* cs_per_n, int64_t cs_ld,
                               int cs_n, int dy_is_act, cudaStream_t stream) {
    HD_REQUIRE(in0 && sums && gamma && beta && dy && gsums && dx0 && (C1 == 0 || (in1 && dx1)));
    HD_REQUIRE(cs_n >= 0 && cs_n <= C0 + C1);
    if (hd_gn_v2(dtype, C0, C1, G, HW, N)) {
        if (dy_is_act) { hd_set_error("hd_gn_bwd_apply: the second-generation kernels take the raw dy (ask hd_gn_v2 first)"); return HD_ERR_ARG; }
        return hd_gn2_bwd_apply(in0, C0, in1, C1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, dy, gsums, add, acc0, acc1, dx0, dx1,
                                cs_total, cs_per_n, cs_ld, cs_n, stream);
    }
    GnParams g{N, HW, C0 + C1, G, sums, gamma, beta, eps, act, p_drop, hd_seed_mix(seed)};
    if (dtype == HD_F32) return gn_bwd_apply_t<float>(in0, C0, in1, C1, g, dy, gsums, add, acc0, acc1, dx0, dx1, cs_total, cs_per_n, cs_ld, cs_n, dy_is_act, stream);
    if (dtype == HD_BF16) return gn_bwd_apply_t<__nv_bfloat16>(in0, C0, in1, C1, g, dy, gsums, add, acc0, acc1, dx0, dx1, cs_total, cs_per_n, cs_ld, cs_n, dy_is_act, stream);
    return HD_ERR_ARG;
}

// ------------------------------- column sums -------------------------------------------------
// per_n[n][c] = sum_pix t[n][pix][c] (optional);  total[c] += sum_n per_n[n][c] (optional)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* t, int N, int64_t HW, int C, int64_t pix_per_block, float* per_n, int64_t ld, float* total) {
    constexpr int V = Vec<T>::N;
    const int lanes = C / V, ppi = blockDim.x / lanes;
    const int lane = threadIdx.x % lanes, sub = threadIdx.x / lanes, n = blockIdx.y;
    extern __shared__ float s_col[];              // [C] partial of this block: one global atomic per channel per block
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_col[i] = 0.f;
    __syncthreads();
    if (sub < ppi) {
        float s[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] = 0.f;
        const int64_t p0 = blockIdx.x * pix_per_block;
        const int64_t p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
            constexpr int U = 4;
        for (int64_t p = p0 + sub; p < p1; p += (int64_t)U * ppi) {
            typename Vec<T>::raw xr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (p + (int64_t)u * ppi < p1) xr[u] = raw_load(t + ((int64_t)n * HW + p + (int64_t)u * ppi) * C + lane * V);
#pragma unroll
            for (int u = 0; u < U; ++u) if (p + (int64_t)u * ppi < p1) {
                float v[V]; unpack(xr[u], v);
#pragma unroll
                for (int k = 0; k < V; ++k) s[k] += v[k];
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) atomicAdd(&s_col[lane * V + k], s[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (per_n) atomicAdd(per_n + (int64_t)n * ld + i, s_col[i]);
        if (total) atomicAdd(total + i, s_col[i]);
    }
}
// any channel count (the 4 / 8-channel maps of the image condition encoder): thread <-> element, shared partials per block
template <typename T>
__global__ void __launch_bounds__(256) colsum_generic_kernel(const T* t, int64_t HW, int C, int64_t pix_per_block, float* per_n, int64_t ld, float* total) {
    extern __shared__ float s_col[];
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_col[i] = 0.f;
    __syncthreads();
    const int64_t p0 = blockIdx.x * pix_per_block, p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
    const T* base = t + (int64_t)n * HW * C;
    for (int64_t e = p0 * C + threadIdx.x; e < p1 * C; e += blockDim.x) atomicAdd(&s_col[e % C], hd_ld(base + e));
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (per_n) atomicAdd(per_n + (int64_t)n * ld + i, s_col[i]);
        if (total) atomicAdd(total + i, s_col[i]);
    }
}
__global__ void colsum_nchw_kernel(const float* t, int C, int64_t HW, float* per_n, int64_t ld, float* total) {
    // block per (n, c)
    __shared__ float red[8];
    const int c = blockIdx.x, n = blockIdx.y;
    const float* p = t + ((int64_t)n * C + c) * HW;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) s += p[i];
    s = hd_warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
        if (per_n) atomicAdd(per_n + (int64_t)n * ld + c, tot);
        if (total) atomicAdd(total + c, tot);
    }
}
// Both outputs ACCUMULATE (atomics): the caller zeroes them (they live in the flat gradient buffers).
extern "C" int hd_colsum(int dtype, const void* t, int nchw_f32, int N, int64_t HW, int C, float* per_n, int64_t ld_per_n, float* total, cudaStream_t stream) {
    HD_REQUIRE(t && (per_n || total) && N > 0 && HW > 0 && C > 0);
    if (nchw_f32) {
        colsum_nchw_kernel<<<dim3(C, N), 256, 0, stream>>>((const float*)t, C, HW, per_n, ld_per_n, total);
        HD_CHECK_LAUNCH();
        return HD_OK;
    }
    int V = dtype == HD_F32 ? 4 : 8;
    if (C % V != 0 && C <= 4096 && (dtype == HD_F32 || dtype == HD_BF16)) {
        int chunks; const int64_t ppb = pick_chunk(N, HW, 64, &chunks);
        if (dtype == HD_F32) colsum_generic_kernel<float><<<dim3(chunks, N), 256, C * sizeof(float), stream>>>((const float*)t, HW, C, ppb, per_n, ld_per_n, total);
        else colsum_generic_kernel<__nv_bfloat16><<<dim3(chunks, N), 256, C * sizeof(float), stream>>>((const __nv_bfloat16*)t, HW, C, ppb, per_n, ld_per_n, total);
        HD_CHECK_LAUNCH();
        return HD_OK;
    }
    if (C % V != 0 || C / V > 256) { hd_set_error("colsum: unsupported channel count"); return HD_ERR_UNSUPPORTED; }
    int ppi = 256 / (C / V), chunks;
    int64_t ppb = pick_chunk(N, HW, ppi, &chunks);
    if (dtype == HD_F32) colsum_kernel<float><<<dim3(chunks, N), 256, C * sizeof(float), stream>>>((const float*)t, N, HW, C, ppb, per_n, ld_per_n, total);
    else if (dtype == HD_BF16) colsum_kernel<__nv_bfloat16><<<dim3(chunks, N), 256, C * sizeof(float), stream>>>((const __nv_bfloat16*)t, N, HW, C, ppb, per_n, ld_per_n, total);
    else return HD_ERR_ARG;
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ------------------------------- NCHW fp32 -> NHWC bf16, channel-padded ----------------------
// The 3-channel network input (and the gradient of the 3-channel output) enter the tcgen05 convolution as a
// 64-channel NHWC tensor whose channels >= Cin are zero (reference: head / tail nn.Conv2d, ModelCondition.py:219,251).
// thread <-> pixel: coalesced plane reads, 128 contiguous bytes written per thread.
__global__ void pad_nchw_kernel(const float* in, int Cin, __nv_bfloat16* out, int64_t HW, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / HW, pix = i - n * HW;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < Cin ? __ldg(in + (n * Cin + c) * HW + pix) : 0.f;
        uint4 first; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&first);
#pragma unroll
        for (int c = 0; c < 4; ++c) h[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
        uint4* dst = reinterpret_cast<uint4*>(out + i * 64);
        dst[0] = first;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 1; c < 8; ++c) dst[c] = z;
    }
}
extern "C" int hd_pad_nchw(const float* in, int Cin, void* out, int N, int64_t HW, cudaStream_t stream) {
    HD_REQUIRE(in && out && Cin > 0 && Cin <= 8 && N > 0 && HW > 0);
    const int64_t total = (int64_t)N * HW;
    int64_t blocks = (total + 255) / 256; if (blocks > (int64_t)hd_num_sms() * 16) blocks = (int64_t)hd_num_sms() * 16;
    pad_nchw_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, Cin, (__nv_bfloat16*)out, HW, total);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ------------------------------- K6: q_sample / MSE ------------------------------------------
// x_t = sqrt_ab[t[n]] * x0 + sqrt_1m_ab[t[n]] * noise     (fp32 NCHW, chw elements per sample)
__global__ void q_sample_kernel(const float4* x0, const float4* noise, const int64_t* t, const float* sab, const float* s1ab,
                                float4* xt, int64_t chw4, int T) {
    const int n = blockIdx.y;
    const int64_t tn = t[n];
    if (tn < 0 || tn >= T) { if (threadIdx.x == 0 && blockIdx.x == 0) printf("hd_q_sample: t = %lld out of range [0, %d)\n", (long long)tn, T); __trap(); }
    const float a = sab[tn], b = s1ab[tn];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chw4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 u = x0[n * chw4 + i], v = noise[n * chw4 + i];
        xt[n * chw4 + i] = make_float4(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z, a * u.w + b * v.w);
    }
}
extern "C" int hd_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sab, const float* s1ab, float* xt,
                           int N, int64_t chw, int T, cudaStream_t stream) {
    HD_REQUIRE(x0 && noise && t && sab && s1ab && xt && N > 0 && chw > 0 && chw % 4 == 0 && T > 0);
    int64_t chw4 = chw / 4;
    int64_t bx = (chw4 + 255) / 256; if (bx > 1024) bx = 1024;
    q_sample_kernel<<<dim3((unsigned)bx, N), 256, 0, stream>>>((const float4*)x0, (const float4*)noise, t, sab, s1ab, (float4*)xt, chw4, T);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
// loss = (pred - noise)^2 ; dpred = 2 (pred - noise) * g
__global__ void mse_fwd_kernel(const float4* pred, const float4* noise, float4* loss, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pred[i], q = noise[i];
        float4 d = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
        loss[i] = make_float4(d.x * d.x, d.y * d.y, d.z * d.z, d.w * d.w);
    }
}
__global__ void mse_bwd_kernel(const float4* pred, const float4* noise, const float4* g, float4* dpred, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pred[i], q = noise[i], w = g[i];
        dpred[i] = make_float4(2.f * (p.x - q.x) * w.x, 2.f * (p.y - q.y) * w.y, 2.f * (p.z - q.z) * w.z, 2.f * (p.w - q.w) * w.w);
    }
}
extern "C" int hd_mse_fwd(const float* pred, const float* noise, float* loss, int64_t n, cudaStream_t stream) {
    HD_REQUIRE(pred && noise && loss && n > 0 && n % 4 == 0);
    int64_t b = (n / 4 + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    mse_fwd_kernel<<<(unsigned)b, 256, 0, stream>>>((const float4*)pred, (const float4*)noise, (float4*)loss, n / 4);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_mse_bwd(const float* pred, const float* noise, const float* g, float* dpred, int64_t n, cudaStream_t stream) {
    HD_REQUIRE(pred && noise && g && dpred && n > 0 && n % 4 == 0);
    int64_t b = (n / 4 + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    mse_bwd_kernel<<<(unsigned)b, 256, 0, stream>>>((const float4*)pred, (const float4*)noise, (const float4*)g, (float4*)dpred, n / 4);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ------------------------------- K7: sampler step --------------------------------------------
// eps = w1 eps_c - w eps_u (w1 = fp32(1+w), rounded on the host like torch does) ; mean = c1 x - c2 eps ; x' = mean + sqrt_var z ; NaN flag ; final clip.
// coef points at a device table [T][3] = (coeff1, coeff2, sqrt(var)); the step index is read
// from device memory so that one captured CUDA graph serves every time step.
__global__ void sampler_step_kernel(float4* x, const float4* eps_c, const float4* eps_u, const float4* z, float w1, float w,
                                    const float* coef, const int* step_ptr, int last_step_clip, int* nan_flag, int64_t n4) {
    const int step = *step_ptr;
    const float c1 = coef[step * 3], c2 = coef[step * 3 + 1], sv = coef[step * 3 + 2];
    const bool last = (step == 0);
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 xv = x[i], e = eps_c[i];
        if (eps_u) { float4 u = eps_u[i]; e = make_float4(w1 * e.x - w * u.x, w1 * e.y - w * u.y, w1 * e.z - w * u.z, w1 * e.w - w * u.w); }
        float4 r = make_float4(c1 * xv.x - c2 * e.x, c1 * xv.y - c2 * e.y, c1 * xv.z - c2 * e.z, c1 * xv.w - c2 * e.w);
        if (!last) { float4 zz = z[i]; r.x += sv * zz.x; r.y += sv * zz.y; r.z += sv * zz.z; r.w += sv * zz.w; }
        bad |= isnan(r.x) | isnan(r.y) | isnan(r.z) | isnan(r.w);
        if (last && last_step_clip) {
            r.x = fminf(fmaxf(r.x, -1.f), 1.f); r.y = fminf(fmaxf(r.y, -1.f), 1.f);
            r.z = fminf(fmaxf(r.z, -1.f), 1.f); r.w = fminf(fmaxf(r.w, -1.f), 1.f);
        }
        x[i] = r;
    }
    if (bad) atomicOr(nan_flag, 1);
}
extern "C" int hd_sampler_step(float* x, const float* eps_c, const float* eps_u, const float* z, float w1, float w, const float* coef,
                               const int* step_ptr, int last_step_clip, int* nan_flag, int64_t n, cudaStream_t stream) {
    HD_REQUIRE(x && eps_c && z && coef && step_ptr && nan_flag && n > 0 && n % 4 == 0);
    int64_t b = (n / 4 + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    sampler_step_kernel<<<(unsigned)b, 256, 0, stream>>>((float4*)x, (const float4*)eps_c, (const float4*)eps_u, (const float4*)z, w1, w, coef,
                                                         step_ptr, last_step_clip, nan_flag, n / 4);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
__global__ void add_int_kernel(int* p, int delta) { *p += delta; }
extern "C" int hd_add_int(int* p, int delta, cudaStream_t stream) {
    HD_REQUIRE(p);
    add_int_kernel<<<1, 1, 0, stream>>>(p, delta);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ------------------------------- fused clip + AdamW on flat buffers --------------------------
__global__ void sqnorm_kernel(const float4* g, int64_t n4, double* out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = g[i]; s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = hd_warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; atomicAdd(out, (double)t); }
}
extern "C" int hd_sqnorm(const float* g, int64_t n, double* out, int accumulate, cudaStream_t stream) {
    HD_REQUIRE(g && out && n > 0 && n % 4 == 0);
    if (!accumulate && cudaMemsetAsync(out, 0, sizeof(double), stream) != cudaSuccess) return HD_ERR_CUDA;
    int64_t b = (n / 4 + 255) / 256; if (b > 148 * 8) b = 148 * 8;
    sqnorm_kernel<<<(unsigned)b, 256, 0, stream>>>((const float4*)g, n / 4, out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.AdamW (decoupled weight decay):
//   clip = min(1, max_norm / (||g|| + 1e-6)); g *= clip
//   p *= 1 - lr wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adamw_kernel(float4* p, float4* g, float4* m, float4* v, int64_t n4, const double* sqnorm, float max_norm,
                             float lr, float b1, float b2, float eps, float decay, float step_size, float bc2_sqrt) {
    float clip = 1.f;
    if (max_norm > 0.f) { float nrm = (float)sqrt(*sqnorm); clip = fminf(1.f, max_norm / (nrm + 1e-6f)); }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
        float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gk = gg[k] * clip;
            gg[k] = gk;
            pp[k] *= decay;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vq[k] = b2 * vq[k] + (1.f - b2) * gk * gk;
            float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
            pp[k] -= step_size * (mm[k] / denom);
        }
        p[i] = pv; g[i] = gv; m[i] = mv; v[i] = vv;
    }
}
extern "C" int hd_adamw_flat(float* p, float* g, float* m, float* v, int64_t n, const double* sqnorm, float max_norm, double lr,
                             double b1, double b2, double eps, double wd, int step, cudaStream_t stream) {
    HD_REQUIRE(p && g && m && v && n > 0 && n % 4 == 0 && step >= 1 && (max_norm <= 0.f || sqnorm));
    // bias corrections in double from the caller's double hyper-parameters, as torch.optim.AdamW forms them (in fp32,
    // 1 - 0.999^step cancels to ~1e-5 relative at small step counts)
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    int64_t b = (n / 4 + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    adamw_kernel<<<(unsigned)b, 256, 0, stream>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n / 4, sqnorm, max_norm, (float)lr, (float)b1, (float)b2,
                                                  (float)eps, (float)(1.0 - lr * wd), (float)(lr / bc1), (float)sqrt(bc2));
    HD_CHECK_LAUNCH();
    return HD_OK;
}
