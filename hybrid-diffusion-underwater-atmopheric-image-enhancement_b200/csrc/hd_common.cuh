// Shared device/host helpers for the hdiff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define HD_OK 0
#define HD_ERR_ARG (-1)       // bad shape / alignment / null pointer
#define HD_ERR_UNSUPPORTED (-2)  // shape outside what this kernel family covers
#define HD_ERR_CUDA (-3)      // launch failed (cudaGetLastError)
#define HD_ERR_DRIVER (-4)    // driver entry point (tensor map encode) unavailable

#define HD_F32 0
#define HD_BF16 1

#define HD_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { hd_set_error(cudaGetErrorString(e_)); return HD_ERR_CUDA; } } while (0)
#define HD_REQUIRE(cond) do { if (!(cond)) { hd_set_error("requirement failed: " #cond); return HD_ERR_ARG; } } while (0)

void hd_set_error(const char* msg);

// ---------------------------------------------------------------------------------------------
// Logical tensor views.  Every activation is physically NHWC [N][PH][PW][C].  A view with
// P == 2 exposes the 2x2 space-to-depth rearrangement of that tensor as a logical
// [N][H][W][4C] tensor (H = PH/2, W = PW/2, logical channel j = (py*2+px)*C + c).  Strided
// convolutions and the transposed convolution become stride-1 convolutions on such views
// ("views").
// ---------------------------------------------------------------------------------------------
struct HdView {
    const void* p;   // base pointer (T)
    int C;           // physical channels
    int P;           // 1 or 2
};

__host__ __device__ __forceinline__ int64_t hd_view_off(int C, int P, int H, int W, int n, int y, int x, int j) {
    // element offset of logical (n, y, x, j) inside the physical tensor
    if (P == 1) return (((int64_t)n * H + y) * W + x) * C + j;
    int q = j / C, c = j - q * C;
    int py = q >> 1, px = q & 1;
    return (((int64_t)n * (2 * H) + (2 * y + py)) * (2 * W) + (2 * x + px)) * C + c;
}

template <typename T> __device__ __forceinline__ float hd_ld(const T* p);
template <> __device__ __forceinline__ float hd_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float hd_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void hd_st(T* p, float v);
template <> __device__ __forceinline__ void hd_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void hd_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// MUFU.EX2 + MUFU.RCP (a few ulp): an IEEE division here costs more issue slots than the rest of a GroupNorm element
__device__ __forceinline__ float hd_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// single-MUFU sigmoid (tanh.approx, |error| ~ 2.5e-4): used by the bf16 kernels only, where it is below the storage rounding;
// it halves the transcendental-pipe work of the GroupNorm kernels, which are issue-bound rather than bandwidth-bound
__device__ __forceinline__ float hd_sigmoid_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
template <bool kFast> __device__ __forceinline__ float hd_sigmoid_t(float x) { return kFast ? hd_sigmoid_fast(x) : hd_sigmoid(x); }
template <bool kFast> __device__ __forceinline__ float hd_swish_t(float x) { return x * hd_sigmoid_t<kFast>(x); }
template <bool kFast> __device__ __forceinline__ float hd_swish_grad_t(float x) { float s = hd_sigmoid_t<kFast>(x); return s * fmaf(x, 1.f - s, 1.f); }
__device__ __forceinline__ float hd_swish(float x) { return x * hd_sigmoid(x); }
__device__ __forceinline__ float hd_swish_grad(float x) { float s = hd_sigmoid(x); return s * (1.f + x * (1.f - s)); }

// Counter-based RNG for dropout: one 32-bit hash per PAIR of consecutive elements, 15 uniform bits per element
// (drop probability resolution 2^-15); the backward pass regenerates the mask from the same (seed, index) instead
// of storing it.  The caller's 64-bit seed is mixed ONCE on the host (splitmix64) into two 32-bit keys; the per-pair hash is
// two multiply / xor-shift rounds keyed before and between them (6 integer instructions per pair), checked for rate,
// uniformity and neighbour / cross-seed correlation (all at the 1e-3 sampling-noise level over 4 M pairs).
__host__ __device__ __forceinline__ uint64_t hd_seed_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// `mixed` = hd_seed_mix(seed); `pair` = (element index >> 1), truncated to 32 bits
__host__ __device__ __forceinline__ uint32_t hd_hash_pair(uint64_t mixed, uint32_t pair) {
    uint32_t h = pair * 0x9E3779B1u + (uint32_t)mixed;
    h ^= h >> 16;
    h = h * 0x85EBCA6Bu + (uint32_t)(mixed >> 32);
    h ^= h >> 13;
    return h;
}
// element 2j uses bits [0, 15) of the pair's hash, element 2j+1 bits [16, 31); dropped iff field < thr15
__host__ __device__ __forceinline__ uint32_t hd_dropout_thr15(float p_drop) {
    uint32_t t = (uint32_t)(p_drop * 32768.f + 0.5f);
    return t > 0x7FFFu ? 0x7FFFu : t;
}
// scale[k] = 0 (dropped) or 1/(1-p) for the V consecutive elements starting at the EVEN element index `base`
template <int V>
__host__ __device__ __forceinline__ void hd_dropout_vec(uint64_t mixed, uint64_t base, float p_drop, float* scale) {
    const uint32_t thr = hd_dropout_thr15(p_drop);
    const float keep = 1.f / (1.f - p_drop);
#pragma unroll
    for (int k = 0; k < V; k += 2) {
        const uint32_t h = hd_hash_pair(mixed, (uint32_t)(base >> 1) + (uint32_t)(k >> 1));
        scale[k] = (h & 0x7FFFu) < thr ? 0.f : keep;
        scale[k + 1] = ((h >> 16) & 0x7FFFu) < thr ? 0.f : keep;
    }
}
#ifdef __CUDACC__
// the same decision for a packed bf16 pair: 0xFFFF in every half that is KEPT.  thr2 = thr15 * 0x00010001.
// ((field | 0x8000) - thr15 keeps bit 15 iff field >= thr15; PRMT replicates the two sign bits over their halves)
__device__ __forceinline__ uint32_t hd_keep_mask2(uint32_t h, uint32_t thr2) {
    const uint32_t w = ((h & 0x7FFF7FFFu) | 0x80008000u) - thr2;
    uint32_t m;      // selector nibbles 9 / B = "replicate the sign of byte 1 / byte 3" (__byte_perm masks that mode bit away)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(w), "r"(0u), "r"(0xBB99u));
    return m;
}
#endif

__device__ __forceinline__ float hd_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float hd_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


// Kernel attributes (opt-in shared memory) are PER DEVICE: `seen` is one flag word per call site, one bit per device ordinal.
// Usage: if (!hd_seen_on_device(&flags)) { cudaFuncSetAttribute(...); hd_mark_on_device(&flags); }  (setting twice is harmless)
static inline unsigned long long hd_device_bit() { int dev = 0; cudaGetDevice(&dev); return 1ull << (dev & 63); }
static inline bool hd_seen_on_device(const unsigned long long* seen) { return (__atomic_load_n(seen, __ATOMIC_ACQUIRE) & hd_device_bit()) != 0; }
static inline void hd_mark_on_device(unsigned long long* seen) { __atomic_fetch_or(seen, hd_device_bit(), __ATOMIC_ACQ_REL); }

static inline int hd_num_sms() {
    static int n = 0;
    if (n == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}
