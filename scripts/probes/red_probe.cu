// Hardware probe (lab equipment, not product code): sustained throughput of fp32 reductions into global memory (L2 atomics)
// in the access pattern a one-pass attention backward would use for dQ: every CTA adds 32 KB tiles (128 x 64 fp32) into an
// 8 MB per-image region that several CTAs walk at different offsets.
//   mode 0: cp.reduce.async.bulk.global.shared::cta.add.f32 of a 32 KB shared-memory tile (one thread issues, 2 in flight)
//   mode 1: red.global.add.v4.f32 from registers, 128 threads x 16 vectors per tile
//   mode 2: plain st.global.v4.f32 of the same tile (the store-bandwidth reference)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/red_probe scripts/probes/red_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

constexpr int kTileBytes = 32768, kTilesPerImg = 256;

__global__ void __launch_bounds__(128) red_kernel(float* dst, int n_img, int iters, int mode) {
    extern __shared__ __align__(128) float tile[];
    for (int i = threadIdx.x; i < kTileBytes / 4; i += blockDim.x) tile[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int img = blockIdx.x % n_img;
    float* base = dst + (size_t)img * kTilesPerImg * (kTileBytes / 4);
    int t = (blockIdx.x / n_img * 37 + blockIdx.x * 11) % kTilesPerImg;
    if (mode == 0) {
        if (threadIdx.x == 0) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(tile);
            for (int i = 0; i < iters; ++i) {
                float* g = base + (size_t)t * (kTileBytes / 4);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(kTileBytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                t = (t + 1) % kTilesPerImg;
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
        for (int i = 0; i < iters; ++i) {
            float* g = base + (size_t)t * (kTileBytes / 4);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float* p = g + (j * 128 + threadIdx.x) * 4;                   // a warp covers 512 contiguous bytes per instruction
                if (mode == 1) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                else *reinterpret_cast<float4*>(p) = v;
            }
            t = (t + 1) % kTilesPerImg;
        }
    }
}

int main(int argc, char** argv) {
    const int n_img = 32;
    const size_t bytes = (size_t)n_img * kTilesPerImg * kTileBytes;
    float* dst;
    cudaMalloc(&dst, bytes);
    cudaMemset(dst, 0, bytes);
    cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode)
        for (int per_sm = 1; per_sm <= 4; per_sm *= 2) {
            const int grid = 148 * per_sm, iters = 2048 / per_sm;
            red_kernel<<<grid, 128, kTileBytes>>>(dst, n_img, 64, mode);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            red_kernel<<<grid, 128, kTileBytes>>>(dst, n_img, iters, mode);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double tb = (double)grid * iters * kTileBytes / 1e12;
            printf("mode %d (%s) CTAs/SM %d: %.3f ms, %.2f TB/s of fp32 adds (%s)\n", mode, mode == 0 ? "bulk reduce" : mode == 1 ? "red.v4.f32" : "st.v4",
                   per_sm, ms, tb / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
        }
    // spot check: the bulk-reduce result is a sum of ones
    float h[4]; cudaMemcpy(h, dst, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dst[0..3] = %.0f %.0f %.0f %.0f\n", h[0], h[1], h[2], h[3]);
    return 0;
}
