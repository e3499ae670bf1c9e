// Small HBM-bound kernels around the hybrid (image-conditioned) pipeline of the reference:
//   * nearest-neighbour up-sampling of a skip tensor and its backward (DynamicUNet's skip fix-up,
//     diffusion/Model.py:506-510: F.interpolate(skip_h, size=h.shape[2:], mode="nearest")), NHWC, integer factor;
//   * uint8 / float image -> fp32 with an affine map (the hybrid trainer's (x.float() / 255) * 2 - 1 and the sampler's
//     x.float() / 255, diffusion/Diffusion.py:56-57,221).
#include "hd_common.cuh"

namespace {

// one thread per 16-byte channel chunk of an OUTPUT pixel
template <typename T>
__global__ void upsample_nearest_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C, int fy, int fx, int64_t total) {
    constexpr int V = 16 / sizeof(T);
    const int cv = C / V, OW = W * fx, OH = H * fy;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t pix = i / cv;
        const int ox = (int)(pix % OW); pix /= OW;
        const int oy = (int)(pix % OH);
        const int64_t n = pix / OH;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((n * H + oy / fy) * W + ox / fx) * C)) + c);
        reinterpret_cast<uint4*>(out + ((n * OH + oy) * OW + ox) * (int64_t)C)[c] = v;
    }
}

// one thread per 16-byte channel chunk of an INPUT-resolution pixel: sums the f x f block of the output gradient in fp32
template <typename T> struct Chunk;
template <> struct Chunk<float> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void add(float* a, const uint4& r) { const float* p = reinterpret_cast<const float*>(&r); for (int k = 0; k < 4; ++k) a[k] += p[k]; }
    static __device__ __forceinline__ uint4 pack(const float* a) { uint4 r; float* p = reinterpret_cast<float*>(&r); for (int k = 0; k < 4; ++k) p[k] = a[k]; return r; }
};
template <> struct Chunk<__nv_bfloat16> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void add(float* a, const uint4& r) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
        for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); a[2 * k] += f.x; a[2 * k + 1] += f.y; }
    }
    static __device__ __forceinline__ uint4 pack(const float* a) {
        uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(a[2 * k], a[2 * k + 1]);
        return r;
    }
};
template <typename T>
__global__ void upsample_nearest_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, int H, int W, int C, int fy, int fx, int64_t total) {
    constexpr int V = Chunk<T>::V;
    const int cv = C / V, OW = W * fx, OH = H * fy;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t pix = i / cv;
        const int x = (int)(pix % W); pix /= W;
        const int y = (int)(pix % H);
        const int64_t n = pix / H;
        float a[V];
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] = 0.f;
        for (int dy = 0; dy < fy; ++dy)
            for (int dx = 0; dx < fx; ++dx)
                Chunk<T>::add(a, __ldg(reinterpret_cast<const uint4*>(dout + ((n * OH + (y * fy + dy)) * OW + (x * fx + dx)) * (int64_t)C) + c));
        reinterpret_cast<uint4*>(din + ((n * H + y) * W + x) * (int64_t)C)[c] = Chunk<T>::pack(a);
    }
}

template <typename S>
__global__ void image_affine_kernel(const S* __restrict__ in, float* __restrict__ out, float scale, float shift, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmaf((float)in[i], scale, shift);
}

int grid_for(int64_t total) {
    int64_t b = (total + 255) / 256, cap = (int64_t)hd_num_sms() * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int hd_upsample_nearest(int dtype, const void* in, void* out, int N, int H, int W, int C, int fy, int fx, cudaStream_t stream) {
    HD_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && fy >= 1 && fx >= 1);
    if (dtype == HD_F32) {
        HD_REQUIRE(C % 4 == 0);
        const int64_t total = (int64_t)N * H * fy * W * fx * (C / 4);
        upsample_nearest_kernel<float><<<grid_for(total), 256, 0, stream>>>((const float*)in, (float*)out, H, W, C, fy, fx, total);
    } else if (dtype == HD_BF16) {
        HD_REQUIRE(C % 8 == 0);
        const int64_t total = (int64_t)N * H * fy * W * fx * (C / 8);
        upsample_nearest_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, H, W, C, fy, fx, total);
    } else return HD_ERR_ARG;
    HD_CHECK_LAUNCH();
    return HD_OK;
}

extern "C" int hd_upsample_nearest_bwd(int dtype, const void* dout, void* din, int N, int H, int W, int C, int fy, int fx, cudaStream_t stream) {
    HD_REQUIRE(dout && din && N > 0 && H > 0 && W > 0 && C > 0 && fy >= 1 && fx >= 1);
    if (dtype == HD_F32) {
        HD_REQUIRE(C % 4 == 0);
        const int64_t total = (int64_t)N * H * W * (C / 4);
        upsample_nearest_bwd_kernel<float><<<grid_for(total), 256, 0, stream>>>((const float*)dout, (float*)din, H, W, C, fy, fx, total);
    } else if (dtype == HD_BF16) {
        HD_REQUIRE(C % 8 == 0);
        const int64_t total = (int64_t)N * H * W * (C / 8);
        upsample_nearest_bwd_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)din, H, W, C, fy, fx, total);
    } else return HD_ERR_ARG;
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// out[i] = in[i] * scale + shift;  src_u8 != 0: `in` is uint8 (dataset images), else fp32
extern "C" int hd_image_affine(const void* in, int src_u8, float* out, float scale, float shift, int64_t n, cudaStream_t stream) {
    HD_REQUIRE(in && out && n > 0);
    if (src_u8) image_affine_kernel<uint8_t><<<grid_for(n), 256, 0, stream>>>((const uint8_t*)in, out, scale, shift, n);
    else image_affine_kernel<float><<<grid_for(n), 256, 0, stream>>>((const float*)in, out, scale, shift, n);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
