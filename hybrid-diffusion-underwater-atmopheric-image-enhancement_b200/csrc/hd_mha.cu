// Multi-head self-attention of the reference's MHA ResBlock: nn.MultiheadAttention(out_ch, num_heads=8) applied to the
// flattened feature map, q = k = v (DiffusionFreeGuidence/ModelCondition.py:189,203-208 == diffusion/Model.py:290,304-309;
// DynamicUNet's four middle blocks, diffusion/Model.py:425-431).  The packed in / out projections run as 1x1 convolutions on
// the tcgen05 kernel (in_proj_weight [3C][C] IS the packed GEMM layout); this file is the attention core
//     o[n, i, h] = softmax_j(q[n, i, h] . k[n, j, h] / sqrt(hd)) v[n, j, h],   hd = C / heads
// for head dims 4..64.  Head dims of 8-16 (C = 64-128 with 8 heads) are too narrow for a tcgen05 K step (16 bf16 per
// instruction row, 128-lane accumulators), so this is a flash-style CUDA-core kernel: fp32 math, one query (or key) row per
// thread (two threads for hd = 64), the other operand streamed through shared memory and read by broadcast, online softmax
// in chunks of 8 keys, no [S, S] tensor.  Backward is two deterministic passes (query-stationary dQ, key-stationary dK / dV).
// qkv is [N][S][3C] (q | k | v along the channel axis, head h = channels [h hd, (h+1) hd) of each), out / dout [N][S][C],
// lse / delta [N][heads][S] fp32.
#include "hd_common.cuh"
#include <math.h>

namespace {

constexpr int kRows = 128;      // rows (queries or keys) per CTA = threads / TPR
constexpr int kTile = 64;       // streamed rows per shared-memory tile
constexpr float kLog2e = 1.4426950408889634f;

template <int TPR> __device__ __forceinline__ float row_sum(float v) {
    if (TPR == 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// stage `rows` x DHT floats of channel block [c0, c0 + HD) of tensor rows [r0, r0 + kTile) into shared memory [kTile][HD]
template <typename T, int HD>
__device__ __forceinline__ void stage(float* dst, const T* src, int64_t row_stride, int r0, int S, int c0, int nthreads) {
    for (int e = threadIdx.x; e < kTile * HD; e += nthreads) {
        const int r = e / HD, d = e - r * HD;
        dst[e] = (r0 + r < S) ? hd_ld(src + (int64_t)(r0 + r) * row_stride + c0 + d) : 0.f;
    }
}

// ------------------------------- forward ----------------------------------------------------
template <typename T, int HD, int TPR>
__global__ void __launch_bounds__(kRows * TPR) mha_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse,
                                                               int S, int C, int heads, float scale) {
    constexpr int DH = HD / TPR;
    __shared__ float Ks[kTile * HD], Vs[kTile * HD];
    const int n = blockIdx.z, h = blockIdx.y;
    const int row = blockIdx.x * kRows + threadIdx.x / TPR, part = threadIdx.x % TPR;
    const bool valid = row < S;
    const T* base = qkv + (int64_t)n * S * 3 * C;
    const int c0 = h * HD;
    float q[DH], o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
        q[d] = valid ? hd_ld(base + (int64_t)row * 3 * C + c0 + part * DH + d) * (scale * kLog2e) : 0.f;
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < S; k0 += kTile) {
        __syncthreads();
        stage<T, HD>(Ks, base + C, 3 * C, k0, S, c0, kRows * TPR);
        stage<T, HD>(Vs, base + 2 * C, 3 * C, k0, S, c0, kRows * TPR);
        __syncthreads();
        const int nk = min(kTile, S - k0);
        for (int j0 = 0; j0 < nk; j0 += 8) {
            float s[8];
            float mx = m;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* kr = Ks + (j0 + u) * HD + part * DH;
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) a = fmaf(q[d], kr[d], a);
                a = row_sum<TPR>(a);
                s[u] = (j0 + u < nk) ? a : -INFINITY;
                mx = fmaxf(mx, s[u]);
            }
            const float corr = exp2f(m - mx);
            l *= corr;
#pragma unroll
            for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float p = exp2f(s[u] - mx);
                l += p;
                const float* vr = Vs + (j0 + u) * HD + part * DH;
#pragma unroll
                for (int d = 0; d < DH; ++d) o[d] = fmaf(p, vr[d], o[d]);
            }
            m = mx;
        }
    }
    if (valid) {
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < DH; ++d) hd_st(out + ((int64_t)n * S + row) * C + c0 + part * DH + d, o[d] * inv);
        if (part == 0) lse[((int64_t)n * heads + h) * S + row] = (m + log2f(l)) * (1.f / kLog2e);
    }
}

// ------------------------------- backward, query-stationary: delta and dQ ---------------------
template <typename T, int HD, int TPR>
__global__ void __launch_bounds__(kRows * TPR) mha_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ o, const T* __restrict__ dout,
                                                                  const float* __restrict__ lse, float* __restrict__ delta, T* __restrict__ dqkv,
                                                                  int S, int C, int heads, float scale) {
    constexpr int DH = HD / TPR;
    __shared__ float Ks[kTile * HD], Vs[kTile * HD];
    const int n = blockIdx.z, h = blockIdx.y;
    const int row = blockIdx.x * kRows + threadIdx.x / TPR, part = threadIdx.x % TPR;
    const bool valid = row < S;
    const T* base = qkv + (int64_t)n * S * 3 * C;
    const int c0 = h * HD, cp = c0 + part * DH;
    float q[DH], g[DH], dq[DH];
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
        q[d] = valid ? hd_ld(base + (int64_t)row * 3 * C + cp + d) * (scale * kLog2e) : 0.f;
        g[d] = valid ? hd_ld(dout + ((int64_t)n * S + row) * C + cp + d) : 0.f;
        dl = fmaf(g[d], valid ? hd_ld(o + ((int64_t)n * S + row) * C + cp + d) : 0.f, dl);
        dq[d] = 0.f;
    }
    dl = row_sum<TPR>(dl);
    const float l2 = valid ? lse[((int64_t)n * heads + h) * S + row] * kLog2e : 0.f;
    if (valid && part == 0) delta[((int64_t)n * heads + h) * S + row] = dl;
    for (int k0 = 0; k0 < S; k0 += kTile) {
        __syncthreads();
        stage<T, HD>(Ks, base + C, 3 * C, k0, S, c0, kRows * TPR);
        stage<T, HD>(Vs, base + 2 * C, 3 * C, k0, S, c0, kRows * TPR);
        __syncthreads();
        const int nk = min(kTile, S - k0);
        for (int j = 0; j < nk; ++j) {
            const float* kr = Ks + j * HD + part * DH;
            const float* vr = Vs + j * HD + part * DH;
            float a = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) { a = fmaf(q[d], kr[d], a); dp = fmaf(g[d], vr[d], dp); }
            a = row_sum<TPR>(a); dp = row_sum<TPR>(dp);
            const float ds = exp2f(a - l2) * (dp - dl) * scale;
#pragma unroll
            for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, kr[d], dq[d]);
        }
    }
    if (valid) {
#pragma unroll
        for (int d = 0; d < DH; ++d) hd_st(dqkv + ((int64_t)n * S + row) * 3 * C + cp + d, dq[d]);
    }
}

// ------------------------------- backward, key-stationary: dK and dV --------------------------
template <typename T, int HD, int TPR>
__global__ void __launch_bounds__(kRows * TPR) mha_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ lse,
                                                                   const float* __restrict__ delta, T* __restrict__ dqkv,
                                                                   int S, int C, int heads, float scale) {
    constexpr int DH = HD / TPR;
    __shared__ float Qs[kTile * HD], Gs[kTile * HD], Ls[kTile], Ds[kTile];
    const int n = blockIdx.z, h = blockIdx.y;
    const int row = blockIdx.x * kRows + threadIdx.x / TPR, part = threadIdx.x % TPR;
    const bool valid = row < S;
    const T* base = qkv + (int64_t)n * S * 3 * C;
    const int c0 = h * HD, cp = c0 + part * DH;
    float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
        k[d] = valid ? hd_ld(base + (int64_t)row * 3 * C + C + cp + d) * (scale * kLog2e) : 0.f;
        v[d] = valid ? hd_ld(base + (int64_t)row * 3 * C + 2 * C + cp + d) : 0.f;
        dk[d] = 0.f; dv[d] = 0.f;
    }
    const float* lse_h = lse + ((int64_t)n * heads + h) * S;
    const float* dl_h = delta + ((int64_t)n * heads + h) * S;
    for (int i0 = 0; i0 < S; i0 += kTile) {
        __syncthreads();
        stage<T, HD>(Qs, base, 3 * C, i0, S, c0, kRows * TPR);
        stage<T, HD>(Gs, dout + (int64_t)n * S * C, C, i0, S, c0, kRows * TPR);
        if (threadIdx.x < kTile) {
            const bool in = i0 + threadIdx.x < S;
            Ls[threadIdx.x] = in ? lse_h[i0 + threadIdx.x] * kLog2e : INFINITY;      // exp2(s - inf) = 0: rows past S contribute nothing
            Ds[threadIdx.x] = in ? dl_h[i0 + threadIdx.x] : 0.f;
        }
        __syncthreads();
        const int nq = min(kTile, S - i0);
        for (int i = 0; i < nq; ++i) {
            const float* qr = Qs + i * HD + part * DH;
            const float* gr = Gs + i * HD + part * DH;
            float a = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) { a = fmaf(qr[d], k[d], a); dp = fmaf(gr[d], v[d], dp); }
            a = row_sum<TPR>(a); dp = row_sum<TPR>(dp);
            const float p = exp2f(a - Ls[i]);
            const float ds = p * (dp - Ds[i]) * scale;
#pragma unroll
            for (int d = 0; d < DH; ++d) { dv[d] = fmaf(p, gr[d], dv[d]); dk[d] = fmaf(ds, qr[d], dk[d]); }
        }
    }
    if (valid) {
#pragma unroll
        for (int d = 0; d < DH; ++d) {
            hd_st(dqkv + ((int64_t)n * S + row) * 3 * C + C + cp + d, dk[d]);
            hd_st(dqkv + ((int64_t)n * S + row) * 3 * C + 2 * C + cp + d, dv[d]);
        }
    }
}

template <typename T, int HD, int TPR>
int mha_fwd_t(const void* qkv, void* out, float* lse, int N, int S, int C, int heads, cudaStream_t st) {
    const dim3 grid((S + kRows - 1) / kRows, heads, N);
    mha_fwd_kernel<T, HD, TPR><<<grid, kRows * TPR, 0, st>>>((const T*)qkv, (T*)out, lse, S, C, heads, 1.f / sqrtf((float)HD));
    HD_CHECK_LAUNCH();
    return HD_OK;
}
template <typename T, int HD, int TPR>
int mha_bwd_t(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int N, int S, int C,
              int heads, cudaStream_t st) {
    const dim3 grid((S + kRows - 1) / kRows, heads, N);
    const float scale = 1.f / sqrtf((float)HD);
    mha_bwd_dq_kernel<T, HD, TPR><<<grid, kRows * TPR, 0, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, delta, (T*)dqkv, S, C, heads, scale);
    mha_bwd_dkv_kernel<T, HD, TPR><<<grid, kRows * TPR, 0, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, S, C, heads, scale);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

#define MHA_DISPATCH(FN, ...)                                                               \
    switch (hd) {                                                                           \
        case 4: return FN<T, 4, 1>(__VA_ARGS__);                                            \
        case 8: return FN<T, 8, 1>(__VA_ARGS__);                                            \
        case 16: return FN<T, 16, 1>(__VA_ARGS__);                                          \
        case 32: return FN<T, 32, 1>(__VA_ARGS__);                                          \
        case 64: return FN<T, 64, 2>(__VA_ARGS__);                                          \
        default: hd_set_error("mha: head dim must be 4, 8, 16, 32 or 64"); return HD_ERR_UNSUPPORTED; \
    }
template <typename T> int mha_fwd_d(int hd, const void* qkv, void* out, float* lse, int N, int S, int C, int heads, cudaStream_t st) {
    MHA_DISPATCH(mha_fwd_t, qkv, out, lse, N, S, C, heads, st)
}
template <typename T> int mha_bwd_d(int hd, const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                                    int N, int S, int C, int heads, cudaStream_t st) {
    MHA_DISPATCH(mha_bwd_t, qkv, out, dout, lse, delta, dqkv, N, S, C, heads, st)
}

// ---- head packing for the tensor-core route ---------------------------------------------------------------------------------
// The tcgen05 attention kernels (hd_attn_tc.cu) are built for head dim 128.  For head dims 8-64 (multiples of 8) the heads are
// zero-padded to 128 channels and run there as N * heads independent sequences: 2-16x the useful MMA work, still an order of
// magnitude faster than the CUDA-core kernel above (hd = 32, S = 1024, 128 sequences: 0.11 vs 0.65 ms forward).  Zero channels
// change neither the scores nor the outputs, and their gradients come out as zeros.
//   pack   : src [N][S][parts * C] -> dst [N * heads][S][parts * 128], dst(n h, s, p, d) = d < hd ? src(n, s, p, h hd + d) * (p == 0 ? scale0 : 1) : 0
//   unpack : the inverse for d < hd (same scale on part 0)
// One thread per 16-byte chunk of a destination row.
__global__ void mha_pack_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int S, int C, int heads, int hd,
                                int parts, float scale0, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % 16);                         // chunk of 8 channels inside the padded head
        int64_t r = i / 16;
        const int p = (int)(r % parts); r /= parts;
        const int s = (int)(r % S); r /= S;
        const int h = (int)(r % heads);
        const int64_t n = r / heads;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (j * 8 < hd) {
            v = __ldg(reinterpret_cast<const uint4*>(src + ((n * S + s) * (int64_t)parts + p) * C + h * hd + j * 8));
            if (p == 0 && scale0 != 1.f) {
                __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(q[k]); q[k] = __floats2bfloat162_rn(f.x * scale0, f.y * scale0); }
            }
        }
        reinterpret_cast<uint4*>(dst)[i] = v;
    }
}
__global__ void mha_unpack_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int S, int C, int heads, int hd,
                                  int parts, float scale0, int64_t total) {
    const int cph = hd / 8;                                  // chunks per head
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % cph);
        int64_t r = i / cph;
        const int h = (int)(r % heads); r /= heads;
        const int p = (int)(r % parts); r /= parts;
        const int s = (int)(r % S);
        const int64_t n = r / S;
        uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (((n * heads + h) * S + s) * (int64_t)parts + p) * 128 + j * 8));
        if (p == 0 && scale0 != 1.f) {
            __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(q[k]); q[k] = __floats2bfloat162_rn(f.x * scale0, f.y * scale0); }
        }
        *reinterpret_cast<uint4*>(dst + ((n * S + s) * (int64_t)parts + p) * C + h * hd + j * 8) = v;
    }
}

}  // namespace

// bf16 only.  parts = 3 (q | k | v) or 1 (o, dO).  scale0 multiplies part 0 (1 in the product path: the softmax scale hd^-1/2 goes to
// hd_attn_*_tc_scaled directly, so packing is exact).
extern "C" int hd_mha_pack_heads(const void* src, void* dst, int N, int S, int C, int heads, int parts, float scale0, cudaStream_t stream) {
    HD_REQUIRE(src && dst && N > 0 && S > 0 && heads > 0 && C % heads == 0 && (parts == 1 || parts == 3));
    const int hd = C / heads;
    HD_REQUIRE(hd % 8 == 0 && hd <= 128);
    const int64_t total = (int64_t)N * heads * S * parts * 16;
    int64_t blocks = (total + 255) / 256; if (blocks > (int64_t)hd_num_sms() * 16) blocks = (int64_t)hd_num_sms() * 16;
    mha_pack_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, S, C, heads, hd, parts, scale0, total);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_mha_unpack_heads(const void* src, void* dst, int N, int S, int C, int heads, int parts, float scale0, cudaStream_t stream) {
    HD_REQUIRE(src && dst && N > 0 && S > 0 && heads > 0 && C % heads == 0 && (parts == 1 || parts == 3));
    const int hd = C / heads;
    HD_REQUIRE(hd % 8 == 0 && hd <= 128);
    const int64_t total = (int64_t)N * S * parts * heads * (hd / 8);
    int64_t blocks = (total + 255) / 256; if (blocks > (int64_t)hd_num_sms() * 16) blocks = (int64_t)hd_num_sms() * 16;
    mha_unpack_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, S, C, heads, hd, parts, scale0, total);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

extern "C" int hd_mha_supported(int C, int heads) {
    if (heads <= 0 || C % heads != 0) return 0;
    const int hd = C / heads;
    return hd == 4 || hd == 8 || hd == 16 || hd == 32 || hd == 64;
}
extern "C" int hd_mha_fwd(int dtype, const void* qkv, void* out, float* lse, int N, int S, int C, int heads, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && lse && N > 0 && S > 0 && C > 0 && heads > 0 && C % heads == 0);
    if (dtype == HD_F32) return mha_fwd_d<float>(C / heads, qkv, out, lse, N, S, C, heads, stream);
    if (dtype == HD_BF16) return mha_fwd_d<__nv_bfloat16>(C / heads, qkv, out, lse, N, S, C, heads, stream);
    return HD_ERR_ARG;
}
extern "C" int hd_mha_bwd(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                          int N, int S, int C, int heads, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && dout && lse && delta && dqkv && N > 0 && S > 0 && C > 0 && heads > 0 && C % heads == 0);
    if (dtype == HD_F32) return mha_bwd_d<float>(C / heads, qkv, out, dout, lse, delta, dqkv, N, S, C, heads, stream);
    if (dtype == HD_BF16) return mha_bwd_d<__nv_bfloat16>(C / heads, qkv, out, dout, lse, delta, dqkv, N, S, C, heads, stream);
    return HD_ERR_ARG;
}
