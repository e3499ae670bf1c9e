// CUDA-core kernels of the hdiff_b200 hot path.
//
//  * generic (any shape) convolution forward / weight-gradient on logical views, templated on the
//    activation type: the fp32 instantiation is the "fp32 check mode" of the path (BASELINE.json
//    north_star: <= 1e-5 against the reference), the bf16 instantiation covers the shapes the
//    tcgen05 kernels do not take (3-channel head / tail, channel counts that are not multiples
//    of 64);
//  * reference-grade attention (scores of one query row in shared memory) for the same purpose;
//  * the tiny fp32 kernels of the embedding path (Linear / Embedding forward + backward);
//  * gather-pack / scatter-unpack between the reference's OIHW fp32 parameters and the packed
//    GEMM layouts.
//
// Reference call sites these replace are cited next to each C-ABI entry in include/hdiff_b200.h.
#include "hd_common.cuh"
#include <string.h>
#include <mutex>

static thread_local char g_err[512] = {0};
void hd_set_error(const char* msg) { strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1); }
extern "C" const char* hd_last_error() { return g_err; }
extern "C" int hd_abi_version() { return 1; }

// ---------------------------------------------------------------------------------------------
// Generic convolution forward: out(n,y,x,co) = bias[co] + emb[n][co] + res + sum_{tap,ci} w[co][tap][ci] * in(n,y+ty-pad,x+tx-pad,ci)
// one thread per (pixel, co); co fastest so that weight rows of neighbouring threads share x.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct ConvArgs {
    const T* in0; const T* in1; const float* in_nchw;  // in_nchw != null: fp32 NCHW input (head)
    int C0, C1, P_in;
    const T* w; const float* bias; const float* emb; int64_t emb_stride;
    const T* res; T* out; float* out_nchw;             // out_nchw != null: fp32 NCHW output (tail)
    int Cout, P_out;
    int N, H, W, k;
    int nchw_c;                                        // channels stored through out_nchw (<= Cout)
};

template <typename T>
__device__ __forceinline__ float conv_in_load(const ConvArgs<T>& a, int n, int y, int x, int j) {
    if (a.in_nchw) {
        if (a.P_in == 1) return a.in_nchw[(((int64_t)n * a.C0 + j) * a.H + y) * a.W + x];
        // 2x2 space-to-depth view of an NCHW image (stride-2 convolutions of the image condition encoder, diffusion/Model.py:119-121)
        const int q = j / a.C0, c = j - q * a.C0;
        return a.in_nchw[(((int64_t)n * a.C0 + c) * (2 * a.H) + (2 * y + (q >> 1))) * (2 * a.W) + (2 * x + (q & 1))];
    }
    if (a.P_in == 1) {
        if (j < a.C0) return hd_ld(a.in0 + ((((int64_t)n * a.H + y) * a.W + x) * a.C0 + j));
        return hd_ld(a.in1 + ((((int64_t)n * a.H + y) * a.W + x) * a.C1 + (j - a.C0)));
    }
    return hd_ld(a.in0 + hd_view_off(a.C0, 2, a.H, a.W, n, y, x, j));
}

template <typename T>
__global__ void conv_simt_kernel(ConvArgs<T> a) {
    const int CinL = (a.C0 + a.C1) * a.P_in * a.P_in;
    const int CoutL = a.Cout * a.P_out * a.P_out;
    const int64_t total = (int64_t)a.N * a.H * a.W * CoutL;
    const int pad = a.k / 2, kk = a.k * a.k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int co = (int)(i % CoutL);
        int64_t pix = i / CoutL;
        int x = (int)(pix % a.W); int64_t r = pix / a.W;
        int y = (int)(r % a.H); int n = (int)(r / a.H);
        float acc = 0.f;
        for (int ty = 0; ty < a.k; ++ty) {
            int yy = y + ty - pad; if (yy < 0 || yy >= a.H) continue;
            for (int tx = 0; tx < a.k; ++tx) {
                int xx = x + tx - pad; if (xx < 0 || xx >= a.W) continue;
                const T* wr = a.w + ((int64_t)co * kk + ty * a.k + tx) * CinL;
                for (int j = 0; j < CinL; ++j) acc += hd_ld(wr + j) * conv_in_load(a, n, yy, xx, j);
            }
        }
        if (a.bias) acc += a.bias[co];
        if (a.emb) acc += a.emb[(int64_t)n * a.emb_stride + co];
        if (a.out_nchw) {
            if (co < a.nchw_c) a.out_nchw[(((int64_t)n * a.nchw_c + co) * a.H + y) * a.W + x] = acc;
        } else {
            int64_t off = hd_view_off(a.Cout, a.P_out, a.H, a.W, n, y, x, co);
            if (a.res) acc += hd_ld(a.res + off);
            hd_st(a.out + off, acc);
        }
    }
}

extern "C" int hd_conv_simt(int dtype, const void* in0, int C0, const void* in1, int C1, int P_in, int in_nchw_f32,
                            const void* w, const float* bias, const float* emb, int64_t emb_stride,
                            const void* res, void* out, int Cout, int P_out, int out_nchw_f32,
                            int N, int H, int W, int ksize, cudaStream_t stream) {
    HD_REQUIRE(in0 && w && out);
    HD_REQUIRE(ksize == 1 || ksize == 3);
    HD_REQUIRE(P_in == 1 || (P_in == 2 && C1 == 0));
    HD_REQUIRE(P_out == 1 || (P_out == 2 && !out_nchw_f32 && !emb));
    HD_REQUIRE(N > 0 && H > 0 && W > 0 && C0 > 0 && Cout > 0 && out_nchw_f32 >= 0 && out_nchw_f32 <= Cout);
    int64_t total = (int64_t)N * H * W * Cout * P_out * P_out;
    int block = 128;
    int grid = (int)((total + block - 1) / block < (int64_t)hd_num_sms() * 32 ? (total + block - 1) / block : (int64_t)hd_num_sms() * 32);
    if (dtype == HD_F32) {
        ConvArgs<float> a{in_nchw_f32 ? nullptr : (const float*)in0, (const float*)in1, in_nchw_f32 ? (const float*)in0 : nullptr, C0, C1, P_in,
                          (const float*)w, bias, emb, emb_stride, (const float*)res, out_nchw_f32 ? nullptr : (float*)out,
                          out_nchw_f32 ? (float*)out : nullptr, Cout, P_out, N, H, W, ksize, out_nchw_f32};
        conv_simt_kernel<float><<<grid, block, 0, stream>>>(a);
    } else if (dtype == HD_BF16) {
        using B = __nv_bfloat16;
        ConvArgs<B> a{in_nchw_f32 ? nullptr : (const B*)in0, (const B*)in1, in_nchw_f32 ? (const float*)in0 : nullptr, C0, C1, P_in,
                      (const B*)w, bias, emb, emb_stride, (const B*)res, out_nchw_f32 ? nullptr : (B*)out,
                      out_nchw_f32 ? (float*)out : nullptr, Cout, P_out, N, H, W, ksize, out_nchw_f32};
        conv_simt_kernel<B><<<grid, block, 0, stream>>>(a);
    } else { HD_REQUIRE(!"dtype"); }
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ---------------------------------------------------------------------------------------------
// Generic weight gradient: dw[co][tap][ci] = sum_{n,y,x} dy(n,y,x,co) * in(n,y+ty-pad,x+tx-pad,ci)
// grid = (pixel chunks, CoutL); threads stride over (tap, ci); fp32 atomics across chunks.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct WgradArgs {
    ConvArgs<T> c;            // input side (in0/in1/in_nchw, C0, C1, P_in, N, H, W, k)
    const T* dy; const float* dy_nchw; int Cdy, P_dy;
    float* dw;
    int pix_per_block;
};

template <typename T>
__global__ void wgrad_simt_kernel(WgradArgs<T> a) {
    const ConvArgs<T>& c = a.c;
    const int CinL = (c.C0 + c.C1) * c.P_in * c.P_in;
    const int kk = c.k * c.k, pad = c.k / 2;
    const int co = blockIdx.y;
    const int64_t npix = (int64_t)c.N * c.H * c.W;
    const int64_t p0 = (int64_t)blockIdx.x * a.pix_per_block;
    const int64_t p1 = p0 + a.pix_per_block < npix ? p0 + a.pix_per_block : npix;
    for (int e = threadIdx.x; e < kk * CinL; e += blockDim.x) {
        int tap = e / CinL, j = e - tap * CinL;
        int ty = tap / c.k, tx = tap - ty * c.k;
        float acc = 0.f;
        for (int64_t pix = p0; pix < p1; ++pix) {
            int x = (int)(pix % c.W); int64_t r = pix / c.W;
            int y = (int)(r % c.H); int n = (int)(r / c.H);
            int yy = y + ty - pad, xx = x + tx - pad;
            if (yy < 0 || yy >= c.H || xx < 0 || xx >= c.W) continue;
            float g = a.dy_nchw ? a.dy_nchw[(((int64_t)n * a.Cdy + co) * c.H + y) * c.W + x]
                                : hd_ld(a.dy + hd_view_off(a.Cdy, a.P_dy, c.H, c.W, n, y, x, co));
            acc += g * conv_in_load(c, n, yy, xx, j);
        }
        atomicAdd(a.dw + ((int64_t)co * kk + tap) * CinL + j, acc);
    }
}

extern "C" int hd_wgrad_simt(int dtype, const void* in0, int C0, const void* in1, int C1, int P_in, int in_nchw_f32,
                             const void* dy, int Cdy, int P_dy, int dy_nchw_f32, float* dw,
                             int N, int H, int W, int ksize, cudaStream_t stream) {
    HD_REQUIRE(in0 && dy && dw);
    HD_REQUIRE(ksize == 1 || ksize == 3);
    HD_REQUIRE(P_in == 1 || (P_in == 2 && C1 == 0));
    HD_REQUIRE(P_dy == 1 || (P_dy == 2 && !dy_nchw_f32));
    const int CoutL = Cdy * P_dy * P_dy, CinL = (C0 + C1) * P_in * P_in;
    const int64_t npix = (int64_t)N * H * W;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)CoutL * ksize * ksize * CinL, stream) != cudaSuccess) return HD_ERR_CUDA;
    int chunks = (int)(npix < 4096 ? (npix + 63) / 64 : 256);
    if (chunks < 1) chunks = 1;
    int ppb = (int)((npix + chunks - 1) / chunks);
    chunks = (int)((npix + ppb - 1) / ppb);
    dim3 grid(chunks, CoutL);
    int block = 128;
    if (dtype == HD_F32) {
        WgradArgs<float> a{{in_nchw_f32 ? nullptr : (const float*)in0, (const float*)in1, in_nchw_f32 ? (const float*)in0 : nullptr, C0, C1, P_in,
                            nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 1, N, H, W, ksize, 0},
                           dy_nchw_f32 ? nullptr : (const float*)dy, dy_nchw_f32 ? (const float*)dy : nullptr, Cdy, P_dy, dw, ppb};
        wgrad_simt_kernel<float><<<grid, block, 0, stream>>>(a);
    } else if (dtype == HD_BF16) {
        using B = __nv_bfloat16;
        WgradArgs<B> a{{in_nchw_f32 ? nullptr : (const B*)in0, (const B*)in1, in_nchw_f32 ? (const float*)in0 : nullptr, C0, C1, P_in,
                        nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 1, N, H, W, ksize, 0},
                       dy_nchw_f32 ? nullptr : (const B*)dy, dy_nchw_f32 ? (const float*)dy : nullptr, Cdy, P_dy, dw, ppb};
        wgrad_simt_kernel<B><<<grid, block, 0, stream>>>(a);
    } else { HD_REQUIRE(!"dtype"); }
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ---------------------------------------------------------------------------------------------
// Attention, reference-grade: one block per (n, row).  Scores of the row live in shared memory.
//   fwd : p = softmax(q_i . K^T * scale), o_i = p V, lse_i = log sum exp
//   bwd : delta_i = dO_i . O_i
//         dQ_i = sum_j dS_ij K_j,  dS_ij = p_ij (dO_i . V_j - delta_i) * scale       (row pass)
//         dK_j = sum_i dS_ij Q_i,  dV_j = sum_i p_ij dO_i                            (column pass)
// qkv is [N][S][3C] (q | k | v along the channel axis), everything else [N][S][C].
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void attn_fwd_simt_kernel(const T* qkv, T* out, float* lse, int S, int C, float scale) {
    extern __shared__ float sm[];
    float* sc = sm;            // [S]
    float* q = sm + S;         // [C]
    __shared__ float red[32];
    const int n = blockIdx.y, i = blockIdx.x;
    const T* base = qkv + (int64_t)n * S * 3 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) q[c] = hd_ld(base + (int64_t)i * 3 * C + c);
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const T* kr = base + (int64_t)j * 3 * C + C;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += q[c] * hd_ld(kr + c);
        s *= scale; sc[j] = s; mx = fmaxf(mx, s);
    }
    mx = hd_warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = -INFINITY;
    for (int w = 0; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < S; j += blockDim.x) { float e = __expf(sc[j] - mx); sc[j] = e; sum += e; }
    sum = hd_warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
    float inv = 1.f / sum;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float o = 0.f;
        for (int j = 0; j < S; ++j) o += sc[j] * hd_ld(base + (int64_t)j * 3 * C + 2 * C + c);
        hd_st(out + ((int64_t)n * S + i) * C + c, o * inv);
    }
    if (threadIdx.x == 0) lse[(int64_t)n * S + i] = mx + __logf(sum);
}

template <typename T>
__global__ void attn_delta_kernel(const T* o, const T* dout, float* delta, int64_t rows, int C) {
    int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float s = 0.f;
    for (int c = threadIdx.x & 31; c < C; c += 32) s += hd_ld(o + r * C + c) * hd_ld(dout + r * C + c);
    s = hd_warp_sum(s);
    if ((threadIdx.x & 31) == 0) delta[r] = s;
}

// mode 0: row pass (block = query i): dQ_i.   mode 1: column pass (block = key j): dK_j and dV_j.
template <typename T>
__global__ void attn_bwd_simt_kernel(const T* qkv, const T* dout, const float* lse, const float* delta, T* dqkv,
                                     int S, int C, float scale, int mode) {
    extern __shared__ float sm[];
    float* ds = sm;            // [S]  dS (row) or dS^T (column)
    float* pp = sm + S;        // [S]  p  (column pass only)
    float* a = sm + 2 * S;     // [C]  q_i or k_j
    float* b = sm + 2 * S + C; // [C]  dO_i or v_j
    const int n = blockIdx.y, r = blockIdx.x;
    const T* base = qkv + (int64_t)n * S * 3 * C;
    const T* dob = dout + (int64_t)n * S * C;
    const float* lse_n = lse + (int64_t)n * S;
    const float* dl_n = delta + (int64_t)n * S;
    if (mode == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { a[c] = hd_ld(base + (int64_t)r * 3 * C + c); b[c] = hd_ld(dob + (int64_t)r * C + c); }
        __syncthreads();
        float l = lse_n[r], d = dl_n[r];
        for (int j = threadIdx.x; j < S; j += blockDim.x) {
            const T* kr = base + (int64_t)j * 3 * C + C; const T* vr = kr + C;
            float s = 0.f, dp = 0.f;
            for (int c = 0; c < C; ++c) { s += a[c] * hd_ld(kr + c); dp += b[c] * hd_ld(vr + c); }
            float p = __expf(s * scale - l);
            ds[j] = p * (dp - d) * scale;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float g = 0.f;
            for (int j = 0; j < S; ++j) g += ds[j] * hd_ld(base + (int64_t)j * 3 * C + C + c);
            hd_st(dqkv + ((int64_t)n * S + r) * 3 * C + c, g);
        }
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { a[c] = hd_ld(base + (int64_t)r * 3 * C + C + c); b[c] = hd_ld(base + (int64_t)r * 3 * C + 2 * C + c); }
        __syncthreads();
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const T* qr = base + (int64_t)i * 3 * C; const T* dor = dob + (int64_t)i * C;
            float s = 0.f, dp = 0.f;
            for (int c = 0; c < C; ++c) { s += a[c] * hd_ld(qr + c); dp += b[c] * hd_ld(dor + c); }
            float p = __expf(s * scale - lse_n[i]);
            pp[i] = p; ds[i] = p * (dp - dl_n[i]) * scale;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float gk = 0.f, gv = 0.f;
            for (int i = 0; i < S; ++i) { gk += ds[i] * hd_ld(base + (int64_t)i * 3 * C + c); gv += pp[i] * hd_ld(dob + (int64_t)i * C + c); }
            hd_st(dqkv + ((int64_t)n * S + r) * 3 * C + C + c, gk);
            hd_st(dqkv + ((int64_t)n * S + r) * 3 * C + 2 * C + c, gv);
        }
    }
}

template <typename T>
static int attn_fwd_simt_t(const void* qkv, void* out, float* lse, int N, int S, int C, cudaStream_t st) {
    size_t smem = sizeof(float) * ((size_t)S + C);
    if (smem > 200 * 1024) { hd_set_error("attn_simt: S too large for the check-mode kernel"); return HD_ERR_UNSUPPORTED; }
    cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_fwd_simt_kernel<T><<<dim3(S, N), 128, smem, st>>>((const T*)qkv, (T*)out, lse, S, C, rsqrtf((float)C));
    HD_CHECK_LAUNCH();
    return HD_OK;
}
template <typename T>
static int attn_bwd_simt_t(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                           int N, int S, int C, cudaStream_t st) {
    size_t smem = sizeof(float) * (2 * (size_t)S + 2 * C);
    if (smem > 200 * 1024) { hd_set_error("attn_simt: S too large for the check-mode kernel"); return HD_ERR_UNSUPPORTED; }
    int64_t rows = (int64_t)N * S;
    attn_delta_kernel<T><<<(unsigned)((rows + 3) / 4), 128, 0, st>>>((const T*)out, (const T*)dout, delta, rows, C);
    cudaFuncSetAttribute(attn_bwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float scale = rsqrtf((float)C);
    attn_bwd_simt_kernel<T><<<dim3(S, N), 128, smem, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, S, C, scale, 0);
    attn_bwd_simt_kernel<T><<<dim3(S, N), 128, smem, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, S, C, scale, 1);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

extern "C" int hd_attn_fwd_simt(int dtype, const void* qkv, void* out, float* lse, int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && lse && N > 0 && S > 0 && C > 0);
    if (dtype == HD_F32) return attn_fwd_simt_t<float>(qkv, out, lse, N, S, C, stream);
    if (dtype == HD_BF16) return attn_fwd_simt_t<__nv_bfloat16>(qkv, out, lse, N, S, C, stream);
    HD_REQUIRE(!"dtype");
}
extern "C" int hd_attn_bwd_simt(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                                void* dqkv, int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && dout && lse && delta && dqkv && N > 0 && S > 0 && C > 0);
    if (dtype == HD_F32) return attn_bwd_simt_t<float>(qkv, out, dout, lse, delta, dqkv, N, S, C, stream);
    if (dtype == HD_BF16) return attn_bwd_simt_t<__nv_bfloat16>(qkv, out, dout, lse, delta, dqkv, N, S, C, stream);
    HD_REQUIRE(!"dtype");
}

// ---------------------------------------------------------------------------------------------
// Embedding path (fp32, tiny): y[M][Nout] (+)= act(x)[M][K] . w[Nout][K]^T + b
// ---------------------------------------------------------------------------------------------
__global__ void linear_fwd_kernel(const float* x, int M, int K, int64_t ldx, const float* w, const float* b,
                                  float* y, int Nout, int64_t ldy, int in_swish, int accumulate) {
    // one warp per output element
    int64_t o = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= (int64_t)M * Nout) return;
    int m = (int)(o / Nout), n = (int)(o % Nout);
    float s = 0.f;
    for (int k = threadIdx.x & 31; k < K; k += 32) {
        float v = x[m * ldx + k];
        if (in_swish) v = hd_swish(v);
        s += v * w[(int64_t)n * K + k];
    }
    s = hd_warp_sum(s);
    if ((threadIdx.x & 31) == 0) {
        if (b) s += b[n];
        if (accumulate) y[m * ldy + n] += s; else y[m * ldy + n] = s;
    }
}
extern "C" int hd_linear_fwd(const float* x, int M, int K, int64_t ldx, const float* w, const float* b, float* y, int Nout,
                             int64_t ldy, int in_swish, int accumulate, cudaStream_t stream) {
    HD_REQUIRE(x && w && y && M > 0 && K > 0 && Nout > 0);
    int64_t outs = (int64_t)M * Nout;
    linear_fwd_kernel<<<(unsigned)((outs + 3) / 4), 128, 0, stream>>>(x, M, K, ldx, w, b, y, Nout, ldy, in_swish, accumulate);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// dx[m][k] (+)= (sum_n dy[m][n] w[n][k]) * (x_pre ? swish'(x_pre[m][k]) : 1)
__global__ void linear_bwd_x_kernel(const float* dy, int M, int Nout, int64_t lddy, const float* w, int K,
                                    const float* x_pre, int64_t ldx, float* dx, int64_t lddx, int accumulate) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= (int64_t)M * K) return;
    int m = (int)(o / K), k = (int)(o % K);
    float s = 0.f;
    for (int n = 0; n < Nout; ++n) s += dy[m * lddy + n] * w[(int64_t)n * K + k];
    if (x_pre) s *= hd_swish_grad(x_pre[m * ldx + k]);
    if (accumulate) dx[m * lddx + k] += s; else dx[m * lddx + k] = s;
}
extern "C" int hd_linear_bwd_x(const float* dy, int M, int Nout, int64_t lddy, const float* w, int K, const float* x_pre,
                               int64_t ldx, float* dx, int64_t lddx, int accumulate, cudaStream_t stream) {
    HD_REQUIRE(dy && w && dx && M > 0 && K > 0 && Nout > 0);
    int64_t outs = (int64_t)M * K;
    linear_bwd_x_kernel<<<(unsigned)((outs + 127) / 128), 128, 0, stream>>>(dy, M, Nout, lddy, w, K, x_pre, ldx, dx, lddx, accumulate);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// dw[n][k] += sum_m dy[m][n] act(x[m][k]);  db[n] += sum_m dy[m][n]
__global__ void linear_bwd_w_kernel(const float* dy, int M, int Nout, int64_t lddy, const float* x, int K, int64_t ldx,
                                    int in_swish, float* dw, float* db) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= (int64_t)Nout * K) return;
    int n = (int)(o / K), k = (int)(o % K);
    float s = 0.f, sb = 0.f;
    for (int m = 0; m < M; ++m) {
        float v = x[m * ldx + k];
        if (in_swish) v = hd_swish(v);
        float g = dy[m * lddy + n];
        s += g * v; sb += g;
    }
    dw[o] += s;
    if (db && k == 0) db[n] += sb;
}
extern "C" int hd_linear_bwd_w(const float* dy, int M, int Nout, int64_t lddy, const float* x, int K, int64_t ldx, int in_swish,
                               float* dw, float* db, cudaStream_t stream) {
    HD_REQUIRE(dy && x && dw && M > 0 && K > 0 && Nout > 0);
    int64_t outs = (int64_t)Nout * K;
    linear_bwd_w_kernel<<<(unsigned)((outs + 127) / 128), 128, 0, stream>>>(dy, M, Nout, lddy, x, K, ldx, in_swish, dw, db);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// An index outside [0, rows) is a caller bug (t >= T, label > num_labels): like torch's device-side assert in nn.Embedding it
// aborts the launch (sticky error on the context) instead of reading out of bounds.
__global__ void embedding_fwd_kernel(const float* table, int rows, int dim, const int64_t* idx, int M, float* out) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= (int64_t)M * dim) return;
    int m = (int)(o / dim), d = (int)(o % dim);
    const int64_t r = idx[m];
    if (r < 0 || r >= rows) { if (d == 0) printf("hd_embedding_fwd: index %lld out of range [0, %d)\n", (long long)r, rows); __trap(); }
    out[o] = table[r * dim + d];
}
__global__ void embedding_bwd_kernel(const float* dout, int dim, const int64_t* idx, int M, float* dtable, int64_t padding_idx) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= (int64_t)M * dim) return;
    int m = (int)(o / dim), d = (int)(o % dim);
    int64_t r = idx[m];
    if (r == padding_idx) return;
    atomicAdd(dtable + r * dim + d, dout[o]);
}
extern "C" int hd_embedding_fwd(const float* table, int rows, int dim, const int64_t* idx, int M, float* out, cudaStream_t stream) {
    HD_REQUIRE(table && idx && out && rows > 0 && dim > 0 && M > 0);
    int64_t n = (int64_t)M * dim;
    embedding_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(table, rows, dim, idx, M, out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int hd_embedding_bwd(const float* dout, int dim, const int64_t* idx, int M, float* dtable, int64_t padding_idx, cudaStream_t stream) {
    HD_REQUIRE(dout && idx && dtable && dim > 0 && M > 0);
    int64_t n = (int64_t)M * dim;
    embedding_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(dout, dim, idx, M, dtable, padding_idx);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// ---------------------------------------------------------------------------------------------
// Parameter (re)packing.  The reference's parameters stay fp32 in their own layouts (OIHW conv
// weights, IOHW transposed-conv weights) inside one flat buffer; the GEMM kernels want
// [CoutL][tap][CinL] in the compute dtype.  idx tables are built once on the host.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_pack_kernel(const float* src, const int32_t* ia, const int32_t* ib, int64_t n, T* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t a = ia[i];
        float v = a >= 0 ? src[a] : 0.f;
        if (ib) { int32_t b = ib[i]; if (b >= 0) v += src[b]; }
        hd_st(out + i, v);
    }
}
extern "C" int hd_gather_pack(int out_dtype, const float* src, const int32_t* ia, const int32_t* ib, int64_t n, void* out, cudaStream_t stream) {
    HD_REQUIRE(src && ia && out && n > 0);
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (out_dtype == HD_F32) gather_pack_kernel<float><<<grid, 256, 0, stream>>>(src, ia, ib, n, (float*)out);
    else if (out_dtype == HD_BF16) gather_pack_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, ia, ib, n, (__nv_bfloat16*)out);
    else HD_REQUIRE(!"dtype");
    HD_CHECK_LAUNCH();
    return HD_OK;
}
// dst[j] += packed[inv[j]] for every j with inv[j] >= 0
__global__ void scatter_unpack_kernel(const float* packed, const int32_t* inv, int64_t n, float* dst) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int32_t s = inv[j];
        if (s >= 0) dst[j] += packed[s];
    }
}
extern "C" int hd_scatter_unpack(const float* packed, const int32_t* inv, int64_t n, float* dst, cudaStream_t stream) {
    HD_REQUIRE(packed && inv && dst && n > 0);
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    scatter_unpack_kernel<<<grid, 256, 0, stream>>>(packed, inv, n, dst);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
