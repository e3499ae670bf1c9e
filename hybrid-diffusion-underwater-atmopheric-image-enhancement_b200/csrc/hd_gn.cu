// K4, second generation: GroupNorm (+ Swish, + dropout) forward and backward for bf16 NHWC tensors.
// (reference: nn.GroupNorm + Swish + nn.Dropout, DiffusionFreeGuidence/ModelCondition.py:128-129,141-143,95,249-250 and
// their autograd backward)
//
// Why a second generation.  The first one (hd_fused.cu; kept for the fp32 check mode and as the A/B partner, HDIFF_GN_V1=1)
// spent ~50 SASS instructions per element (64-bit address arithmetic per load, run-time feature branches, EX2 + RCP sigmoid,
// a 10-instruction hash per element pair) at 117-128 registers = 24 % occupancy; ncu: dram 31-38 %, issue-bound.  A leaner
// register-loaded rewrite (14 instructions per element) reached 0.55-0.62 of the HBM copy rate and stopped there: ptxas
// interleaves the "batched" loads with the arithmetic to stay inside 64 registers, so a warp has 2-3 loads in flight, not 8
// (profiles/r02_gn_register_loaded_variants.txt).  Here the memory pipeline is decoupled from the registers:
//   * a PRODUCER warp streams the CTA's contiguous pixel range through a shared-memory ring with cp.async.bulk (one bulk
//     copy per operand tensor and tile, completion on an mbarrier); 3 CTAs x ~48 KB are in flight per SM whatever the
//     consumers do.  The two sources of a fused torch.cat are two streams of the same ring;
//   * 8 CONSUMER warps read their fixed channel quad with conflict-free 8-byte shared loads, release the stage as soon as the
//     raw values are in registers, compute, and store straight to global memory;
//   * features are template parameters; per-channel constants live in 8-20 registers;
//   * Swish through ONE MUFU: with h = z/2 and t = tanh.approx(h): swish(z) = h + h t, 2 swish'(z) = (1 + t) + h (1 - t^2);
//     the factor 2 and the dropout keep-scale ride in the per-channel constants;
//   * dropout mask per bf16 PAIR by integer SWAR (hd_keep_mask2) applied to the packed bits;
//   * the backward reduce pass no longer rewrites dy (a saving only while the kernels were issue-bound).
#include "hd_tc_common.cuh"
#include <stdlib.h>

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kConsumers = 256;            // 8 consumer warps
constexpr int kThreads = kConsumers + 32;  // + the producer warp
constexpr int kMaxStreams = 6;
constexpr int kMaxStages = 8;
constexpr int kRingBudget = 64 * 1024;     // bytes of shared memory per CTA for the ring (3 CTAs per SM)

struct GnArgs {
    const bf16* x0; const bf16* x1; int C0, C1;       // two-source NHWC input [N][HW][C0 | C1]
    int N, HW, C, G;
    const double* sums;                                // [N][G][2] (sum, sum of squares) of the input
    const float* gamma; const float* beta; float eps;
    float p_drop; uint64_t mixed;                      // hd_seed_mix(seed)
};

// ---- the ring ---------------------------------------------------------------------------------------------------------
struct Ring {
    int nstreams, stages, TP;                          // TP = pixels per tile
    uint32_t stage_bytes;
    const uint8_t* base[kMaxStreams];                  // pixel 0 of image 0
    uint32_t bpp[kMaxStreams];                         // bytes per pixel
    uint32_t off[kMaxStreams];                         // offset of the stream's tile inside a stage
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct RingState {
    uint64_t* full; uint64_t* empty; uint8_t* buf;
    // dyn: [kMaxStages] full | [kMaxStages] empty (128 bytes) | stages x stage_bytes
    __device__ __forceinline__ void init(const Ring& r, uint8_t* dyn) {
        full = reinterpret_cast<uint64_t*>(dyn);
        empty = full + kMaxStages;
        buf = dyn + 128;
        if (threadIdx.x == 0) {
            for (int s = 0; s < r.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kConsumers / 32); }
            fence_barrier_init();
        }
    }
};
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

// Ring position, kept across the segments (images) a CTA walks.
struct Cursor { int s = 0; uint32_t ph = 0; bool wrapped = false; };
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // the eight consumer warps only

// A CTA's share of the batch is one contiguous range of the flattened pixel axis [0, N * HW): `seg(n, p0, p1)` is called for
// every image the range touches.  One wave of CTAs covers the batch exactly (a per-image split leaves floor(444 / N) * N of the
// 444 resident CTAs busy: 6 % idle at N = 32).
template <typename F>
__device__ __forceinline__ void for_each_segment(int per, int total, int HW, F seg) {       // N * HW < 2^31 (checked on the host)
    int g0 = (int)blockIdx.x * per;
    const int g1 = min(g0 + per, total);
    int n = g0 / HW, p0 = g0 - n * HW;
    while (g0 < g1) {
        const int p1 = min(HW, p0 + (g1 - g0));
        seg(n, p0, p1);
        g0 += p1 - p0;
        ++n; p0 = 0;
    }
}

// producer lane: tiles of image n, pixels [p0, p1).  No divisions: (stage, phase) advance incrementally.
__device__ __forceinline__ void produce(const Ring& r, const RingState& rs, Cursor& cu, uint32_t bpp_sum, int n, int HW, int p0, int p1) {
    for (int p = p0; p < p1; p += r.TP) {
        if (cu.wrapped && !mbar_try(rs.empty + cu.s, cu.ph ^ 1u)) mbar_wait(rs.empty + cu.s, cu.ph ^ 1u);    // consumers released the previous use
        const uint32_t npx = (uint32_t)min(r.TP, p1 - p);
        mbar_arrive_expect_tx(rs.full + cu.s, npx * bpp_sum);
        uint8_t* dst = rs.buf + (size_t)cu.s * r.stage_bytes;
        for (int k = 0; k < r.nstreams; ++k)
            bulk_g2s(dst + r.off[k], r.base[k] + ((size_t)n * HW + p) * r.bpp[k], npx * r.bpp[k], rs.full + cu.s);
        if (++cu.s == r.stages) { cu.s = 0; cu.ph ^= 1u; cu.wrapped = true; }
    }
}
__device__ __forceinline__ void produce_all(const Ring& r, const RingState& rs, int per, int total, int HW) {
    uint32_t bpp_sum = 0;
    for (int k = 0; k < r.nstreams; ++k) bpp_sum += r.bpp[k];
    Cursor cu;
    for_each_segment(per, total, HW, [&](int n, int p0, int p1) { produce(r, rs, cu, bpp_sum, n, HW, p0, p1); });
}

__device__ __forceinline__ uint2 lds8(const uint8_t* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void unpack4(const uint2& r, float* v) {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xFFFF0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float tanh_fast(float x) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }

// per-group mean / rstd of image n into shared memory (threads < G)
__device__ __forceinline__ void load_stats(const GnArgs& g, int n, float* s_mean, float* s_rstd) {
    if (threadIdx.x < g.G) {
        const double cnt = (double)(g.C / g.G) * (double)g.HW;
        const double s = g.sums[((size_t)n * g.G + threadIdx.x) * 2], ss = g.sums[((size_t)n * g.G + threadIdx.x) * 2 + 1];
        const double m = s / cnt;
        double var = ss / cnt - m * m;
        if (var < 0) var = 0;
        s_mean[threadIdx.x] = (float)m;
        s_rstd[threadIdx.x] = rsqrtf((float)var + g.eps);
    }
}

// consumer thread -> (channel quad, pixel phase).  lanes = C/4 channel quads per pixel, ppi = 256 / lanes pixels per step.
// A thread's k-th pixel slot of a tile is pixel sub + k ppi; its byte offset inside a stream's tile is o + k * step.
struct Map {
    int c, sub, ppi; bool active, first;
    int xs, cd;                        // channels and channel offset of the SOURCE tensor the quad lives in
    uint32_t xo, xstep;                // x (two-source) tile: offset of slot 0, bytes between slots
    __device__ __forceinline__ Map(const GnArgs& g, const Ring& r) {
        const int lanes = g.C >> 2;
        ppi = kConsumers / lanes;
        const int lane = threadIdx.x % lanes;
        sub = threadIdx.x / lanes;
        active = threadIdx.x < kConsumers && sub < ppi;
        c = lane * 4;
        first = c < g.C0;
        xs = first ? g.C0 : g.C1;
        cd = first ? c : c - g.C0;
        xo = (first ? r.off[0] : r.off[1]) + (uint32_t)(sub * xs + cd) * 2u;
        xstep = (uint32_t)(ppi * xs) * 2u;
    }
    // a C-wide stream (dy, add, out)
    __device__ __forceinline__ uint32_t wide_off(const GnArgs& g, uint32_t stream_off) const { return stream_off + (uint32_t)(sub * g.C + c) * 2u; }
    __device__ __forceinline__ uint32_t wide_step(const GnArgs& g) const { return (uint32_t)(ppi * g.C) * 2u; }
};

// first pair index (32 bits, wraps) of (image n, pixel p, channel quad c) in the logical [N][HW][C] tensor: ((n HW + p) C + c) / 2
__device__ __forceinline__ void mask4(const GnArgs& g, uint32_t thr2, uint32_t pix_global, int c, uint2& bits) {
    const uint32_t pi = pix_global * (uint32_t)(g.C >> 1) + (uint32_t)(c >> 1);
    bits.x &= hd_keep_mask2(hd_hash_pair(g.mixed, pi), thr2);
    bits.y &= hd_keep_mask2(hd_hash_pair(g.mixed, pi + 1), thr2);
}

// The consumer loop shared by all kernels.  Per tile: wait for the stage, load(stage base, k) for the thread's KPT pixel slots,
// release the stage (the raw values are in registers), math(pixel, k).  Full tiles run without bounds checks; the CTA's last
// (partial) tile checks each slot.  (stage, phase) advance incrementally: an integer division per tile costs more than the
// tile's arithmetic.
template <int KPT, typename Load, typename Math>
__device__ __forceinline__ void consume(const Ring& r, const RingState& rs, const Map& m, Cursor& cu, int p0, int p1, Load load, Math math) {
    const bool lane0 = (threadIdx.x & 31) == 0;
    int p = p0;
    for (; p + r.TP <= p1; p += r.TP) {
        if (!mbar_try(rs.full + cu.s, cu.ph)) mbar_wait(rs.full + cu.s, cu.ph);
        const uint8_t* sb = rs.buf + (size_t)cu.s * r.stage_bytes;
        if (m.active) {
#pragma unroll
            for (int k = 0; k < KPT; ++k) load(sb, k);
        }
        __syncwarp();
        if (lane0) mbar_arrive(rs.empty + cu.s);
        if (m.active) {
#pragma unroll
            for (int k = 0; k < KPT; ++k) math(p + m.sub + k * m.ppi, k);
        }
        if (++cu.s == r.stages) { cu.s = 0; cu.ph ^= 1u; }
    }
    if (p < p1) {
        const int npx = p1 - p;
        mbar_wait(rs.full + cu.s, cu.ph);
        const uint8_t* sb = rs.buf + (size_t)cu.s * r.stage_bytes;
        if (m.active) {
#pragma unroll
            for (int k = 0; k < KPT; ++k) if (m.sub + k * m.ppi < npx) load(sb, k);
        }
        __syncwarp();
        if (lane0) mbar_arrive(rs.empty + cu.s);
        if (m.active) {
#pragma unroll
            for (int k = 0; k < KPT; ++k) if (m.sub + k * m.ppi < npx) math(p + m.sub + k * m.ppi, k);
        }
        if (++cu.s == r.stages) { cu.s = 0; cu.ph ^= 1u; }
    }
}

// ------------------------------- statistics -------------------------------------------------
constexpr int kKPTStats = 8, kKPTFwd = 4, kKPTReduce = 4;
__global__ void __launch_bounds__(kThreads, 3) gn2_stats_kernel(GnArgs g, Ring r, int per, double* sums) {
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ float sg[64][2];
    RingState rs; rs.init(r, dyn);
    __syncthreads();
    const int total = g.N * g.HW;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) produce_all(r, rs, per, total, g.HW);
        return;
    }
    const Map m(g, r);
    const int cpg = g.C / g.G;
    Cursor cu;
    for_each_segment(per, total, g.HW, [&](int n, int p0, int p1) {
        if (threadIdx.x < 64) { sg[threadIdx.x][0] = 0.f; sg[threadIdx.x][1] = 0.f; }
        consumer_sync();
        float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
        uint2 xr[kKPTStats];
        consume<kKPTStats>(r, rs, m, cu, p0, p1,
                [&](const uint8_t* sb, int k) { xr[k] = lds8(sb + m.xo + k * m.xstep); },
                [&](int, int k) {
                    float v[4]; unpack4(xr[k], v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { s[i] += v[i]; ss[i] = fmaf(v[i], v[i], ss[i]); }
                });
        if (m.active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int gi = (m.c + k) / cpg; atomicAdd(&sg[gi][0], s[k]); atomicAdd(&sg[gi][1], ss[k]); }
        }
        consumer_sync();
        if (threadIdx.x < g.G) {
            atomicAdd(sums + ((size_t)n * g.G + threadIdx.x) * 2, (double)sg[threadIdx.x][0]);
            atomicAdd(sums + ((size_t)n * g.G + threadIdx.x) * 2 + 1, (double)sg[threadIdx.x][1]);
        }
        consumer_sync();
    });
}

// ------------------------------- apply (forward) --------------------------------------------
// out = drop(act(x * a + b));  ACT: swish through tanh.approx (|abs error| <= 5e-4 |z|/2, below the bf16 rounding of the
// tensor's O(1) values)
template <bool ACT, bool DROP>
__global__ void __launch_bounds__(kThreads, 3) gn2_apply_kernel(GnArgs g, Ring r, int per, bf16* out) {
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ float s_mean[64], s_rstd[64];
    RingState rs; rs.init(r, dyn);
    __syncthreads();
    const int total = g.N * g.HW;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) produce_all(r, rs, per, total, g.HW);
        return;
    }
    const Map m(g, r);
    const int cpg = g.C / g.G;
    const float keep = DROP ? 1.f / (1.f - g.p_drop) : 1.f;
    const uint32_t thr2 = hd_dropout_thr15(g.p_drop) * 0x00010001u;
    Cursor cu;
    for_each_segment(per, total, g.HW, [&](int n, int p0, int p1) {
        consumer_sync();                               // the previous image's statistics are no longer read
        load_stats(g, n, s_mean, s_rstd);
        consumer_sync();
        float A[4], B[4];
        if (m.active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int gi = (m.c + k) / cpg;
                const float a = s_rstd[gi] * g.gamma[m.c + k];
                const float sc = ACT ? 0.5f : keep;
                A[k] = sc * a;
                B[k] = sc * (g.beta[m.c + k] - s_mean[gi] * a);
            }
        }
        bf16* op = out + (size_t)n * g.HW * g.C + m.c;
        const uint32_t pix0 = (uint32_t)n * (uint32_t)g.HW;
        uint2 xr[kKPTFwd];
        consume<kKPTFwd>(r, rs, m, cu, p0, p1,
                [&](const uint8_t* sb, int k) { xr[k] = lds8(sb + m.xo + k * m.xstep); },
                [&](int p, int k) {
                    float v[4], y[4]; unpack4(xr[k], v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float h = fmaf(v[i], A[i], B[i]);
                        if (ACT) { const float hk = DROP ? h * keep : h; y[i] = fmaf(hk, tanh_fast(h), hk); }
                        else y[i] = h;
                    }
                    uint2 o = make_uint2(pack2(y[0], y[1]), pack2(y[2], y[3]));
                    if (DROP) mask4(g, thr2, pix0 + (uint32_t)p, m.c, o);
                    *reinterpret_cast<uint2*>(op + (size_t)p * g.C) = o;
                });
    });
}

// ------------------------------- backward: shared element math ------------------------------
// Returns dd = K * dy' with dy' = dy * mask * act'(z), K = 2 when ACT (the factor rides in the caller's constants) and the
// dropout keep-scale NOT applied (ditto).  d holds the (already masked) upstream gradient.
template <bool ACT>
__device__ __forceinline__ float dd_of(float v, float d, float A, float B) {
    if (!ACT) return d;
    const float h = fmaf(v, A, B);                     // z / 2
    const float t = tanh_fast(h);
    const float s2 = t + 1.f;                          // 2 sigmoid(z)
    const float w2 = fmaf(-t, t, 1.f);                 // 4 sigmoid (1 - sigmoid)
    return d * fmaf(h, w2, s2);                        // d * 2 swish'(z)
}

// ------------------------------- backward, reduction pass -----------------------------------
// Per channel: dgamma += sum dy' xhat, dbeta += sum dy'.  Per (n, group): gsums = (sum gamma dy', sum gamma dy' xhat).
// streams: x0 [, x1], dy (last)
template <bool ACT, bool DROP>
__global__ void __launch_bounds__(kThreads, 3) gn2_bwd_reduce_kernel(GnArgs g, Ring r, int per, double* gsums, float* dgamma, float* dbeta) {
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ float s_mean[64], s_rstd[64], sg[64][2];
    RingState rs; rs.init(r, dyn);
    float* s_ch = reinterpret_cast<float*>(dyn + 128 + (size_t)r.stages * r.stage_bytes);      // [2][C] behind the ring
    __syncthreads();
    const int total = g.N * g.HW;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) produce_all(r, rs, per, total, g.HW);
        return;
    }
    const Map m(g, r);
    const int cpg = g.C / g.G;
    const uint32_t thr2 = hd_dropout_thr15(g.p_drop) * 0x00010001u;
    const uint32_t dyo = m.wide_off(g, r.off[r.nstreams - 1]), dstep = m.wide_step(g);
    Cursor cu;
    for_each_segment(per, total, g.HW, [&](int n, int p0, int p1) {
        load_stats(g, n, s_mean, s_rstd);
        if (threadIdx.x < 64) { sg[threadIdx.x][0] = 0.f; sg[threadIdx.x][1] = 0.f; }
        for (int i = threadIdx.x; i < 2 * g.C; i += kConsumers) s_ch[i] = 0.f;
        consumer_sync();
        float A[4], B[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        if (m.active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int gi = (m.c + k) / cpg;
                const float a = s_rstd[gi] * g.gamma[m.c + k];
                A[k] = 0.5f * a;
                B[k] = 0.5f * (g.beta[m.c + k] - s_mean[gi] * a);
            }
        }
        const uint32_t pix0 = (uint32_t)n * (uint32_t)g.HW;
        uint2 xr[kKPTReduce], dr[kKPTReduce];
        consume<kKPTReduce>(r, rs, m, cu, p0, p1,
                [&](const uint8_t* sb, int k) { xr[k] = lds8(sb + m.xo + k * m.xstep); dr[k] = lds8(sb + dyo + k * dstep); },
                [&](int p, int k) {
                    if (DROP) mask4(g, thr2, pix0 + (uint32_t)p, m.c, dr[k]);
                    float v[4], d[4]; unpack4(xr[k], v); unpack4(dr[k], d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float dd = dd_of<ACT>(v[i], d[i], A[i], B[i]);
                        s1[i] += dd; s2[i] = fmaf(dd, v[i], s2[i]);
                    }
                });
        if (m.active) {
            const float K = (ACT ? 0.5f : 1.f) * (DROP ? 1.f / (1.f - g.p_drop) : 1.f);     // undo the factor 2, apply the keep-scale
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = m.c + k, gi = c / cpg;
                const float gam = g.gamma[c];
                const float t1 = K * s1[k];
                const float sxh = s_rstd[gi] * (K * s2[k] - s_mean[gi] * t1);     // sum(dy' * xhat)
                atomicAdd(&s_ch[c], sxh); atomicAdd(&s_ch[g.C + c], t1);
                atomicAdd(&sg[gi][0], gam * t1); atomicAdd(&sg[gi][1], gam * sxh);
            }
        }
        consumer_sync();
        for (int i = threadIdx.x; i < g.C; i += kConsumers) { atomicAdd(dgamma + i, s_ch[i]); atomicAdd(dbeta + i, s_ch[g.C + i]); }
        if (threadIdx.x < g.G) {
            atomicAdd(gsums + ((size_t)n * g.G + threadIdx.x) * 2, (double)sg[threadIdx.x][0]);
            atomicAdd(gsums + ((size_t)n * g.G + threadIdx.x) * 2 + 1, (double)sg[threadIdx.x][1]);
        }
        consumer_sync();
    });
}

// ------------------------------- backward, apply pass ---------------------------------------
// dx = rstd (gamma dy' - a - xhat b) + add + acc, a = gsums0 / m, b = gsums1 / m; written to the two source tensors' shapes.
// Column sums of dx (bias / embedding-add gradients of the convolution that produced x) for the leading cs_n channels.
// streams: x0 [, x1], dy [, add] [, acc0] [, acc1]
struct GnBwdOut {
    bf16* dx0; bf16* dx1;
    float* cs_total; float* cs_per_n; int64_t cs_ld; int cs_n;
    int s_dy, s_add, s_acc0, s_acc1;                   // stream index of each operand (-1: absent)
};
template <bool ACT, bool DROP, int KPT>
__global__ void __launch_bounds__(kThreads, 3) gn2_bwd_apply_kernel(GnArgs g, Ring r, int per, const double* gsums, GnBwdOut o) {
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ float s_mean[64], s_rstd[64], s_a[64], s_b[64];
    RingState rs; rs.init(r, dyn);
    float* s_cs = reinterpret_cast<float*>(dyn + 128 + (size_t)r.stages * r.stage_bytes);      // [C] behind the ring
    __syncthreads();
    const int total = g.N * g.HW;
    if (threadIdx.x >= kConsumers) {
        if (threadIdx.x == kConsumers) produce_all(r, rs, per, total, g.HW);
        return;
    }
    const Map m(g, r);
    const int cpg = g.C / g.G;
    const bool want_cs = o.cs_total || o.cs_per_n;
    const float keep = DROP ? 1.f / (1.f - g.p_drop) : 1.f;
    const uint32_t thr2 = hd_dropout_thr15(g.p_drop) * 0x00010001u;
    const uint32_t dstep = m.wide_step(g);
    const uint32_t dyo = m.wide_off(g, r.off[o.s_dy]);
    constexpr bool EXTRA = KPT == 2;                   // the four-slot instantiation is launched without addends only
    const bool has_add = EXTRA && o.s_add >= 0;
    const uint32_t ado = has_add ? m.wide_off(g, r.off[o.s_add]) : 0u;
    const int s_acc = m.first ? o.s_acc0 : o.s_acc1;
    const bool has_acc = EXTRA && s_acc >= 0;
    const uint32_t aco = has_acc ? r.off[s_acc] + (uint32_t)(m.sub * m.xs + m.cd) * 2u : 0u;
    Cursor cu;
    for_each_segment(per, total, g.HW, [&](int n, int p0, int p1) {
        if (want_cs) for (int i = threadIdx.x; i < g.C; i += kConsumers) s_cs[i] = 0.f;
        load_stats(g, n, s_mean, s_rstd);
        if (threadIdx.x < g.G) {
            const double cnt = (double)(g.C / g.G) * (double)g.HW;
            s_a[threadIdx.x] = (float)(__ldcg(gsums + ((size_t)n * g.G + threadIdx.x) * 2) / cnt);
            s_b[threadIdx.x] = (float)(__ldcg(gsums + ((size_t)n * g.G + threadIdx.x) * 2 + 1) / cnt);
        }
        consumer_sync();
        // dd = K' dy' (K' = 2 for ACT, keep-scale missing): dx = A (keep dd) - C1 - v D1 with A = rstd gamma / K'.  ACT: the same A
        // forms h = z / 2 (one array serves both, keep is applied to dd); no ACT: A carries the keep-scale itself.
        float A[4], B[4], C1[4], D1[4], cs[4] = {0.f, 0.f, 0.f, 0.f};
        if (m.active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = m.c + k, gi = c / cpg;
                const float rstd = s_rstd[gi], mr = -s_mean[gi] * rstd, a = rstd * g.gamma[c];
                A[k] = ACT ? 0.5f * a : keep * a;
                B[k] = 0.5f * fmaf(mr, g.gamma[c], g.beta[c]);
                C1[k] = rstd * (s_a[gi] + mr * s_b[gi]);
                D1[k] = rstd * rstd * s_b[gi];
            }
        }
        bf16* op = (m.first ? o.dx0 : o.dx1) + (size_t)n * g.HW * m.xs + m.cd;
        const uint32_t pix0 = (uint32_t)n * (uint32_t)g.HW;
        uint2 xr[KPT], dr[KPT], ar[EXTRA ? KPT : 1], cr[EXTRA ? KPT : 1];
        consume<KPT>(r, rs, m, cu, p0, p1,
                [&](const uint8_t* sb, int k) {
                    xr[k] = lds8(sb + m.xo + k * m.xstep); dr[k] = lds8(sb + dyo + k * dstep);
                    if (EXTRA && has_add) ar[k] = lds8(sb + ado + k * dstep);
                    if (EXTRA && has_acc) cr[k] = lds8(sb + aco + k * m.xstep);
                },
                [&](int p, int k) {
                    if (DROP) mask4(g, thr2, pix0 + (uint32_t)p, m.c, dr[k]);
                    float v[4], d[4], res[4]; unpack4(xr[k], v); unpack4(dr[k], d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float dd = dd_of<ACT>(v[i], d[i], A[i], B[i]);
                        if (ACT && DROP) dd *= keep;
                        res[i] = fmaf(A[i], dd, -fmaf(v[i], D1[i], C1[i]));
                    }
                    if (EXTRA && has_add) { float t[4]; unpack4(ar[k], t);
#pragma unroll
                        for (int i = 0; i < 4; ++i) res[i] += t[i]; }
                    if (EXTRA && has_acc) { float t[4]; unpack4(cr[k], t);
#pragma unroll
                        for (int i = 0; i < 4; ++i) res[i] += t[i]; }
                    *reinterpret_cast<uint2*>(op + (size_t)p * m.xs) = make_uint2(pack2(res[0], res[1]), pack2(res[2], res[3]));
#pragma unroll
                    for (int i = 0; i < 4; ++i) cs[i] += res[i];
                });
        if (want_cs) {
            if (m.active) {
#pragma unroll
                for (int k = 0; k < 4; ++k) atomicAdd(&s_cs[m.c + k], cs[k]);
            }
            consumer_sync();
            for (int i = threadIdx.x; i < o.cs_n; i += kConsumers) {
                if (o.cs_per_n) atomicAdd(o.cs_per_n + (int64_t)n * o.cs_ld + i, s_cs[i]);
                if (o.cs_total) atomicAdd(o.cs_total + i, s_cs[i]);
            }
        }
        consumer_sync();                               // shared statistics / partials are reused by the next image
    });
}

// ---- host side ------------------------------------------------------------------------------------------------------------
struct Plan { Ring r; int per; int grid; size_t smem; };

int add_stream(Ring& r, const void* base, int channels) {
    const int k = r.nstreams++;
    r.base[k] = (const uint8_t*)base;
    r.bpp[k] = (uint32_t)channels * 2u;
    return k;
}
// TP pixels per tile (kpt steps of ppi pixels), as many stages as fit the budget; ONE wave of CTAs (3 per SM), each taking a
// contiguous, tile-aligned share of the flattened pixel axis [0, N * HW)
Plan finish_plan(Ring r, int N, int HW, int C, size_t tail_bytes, int kpt) {
    Plan pl;
    const int ppi = kConsumers / (C / 4);
    r.TP = ppi * kpt;
    uint32_t off = 0;
    for (int k = 0; k < r.nstreams; ++k) { r.off[k] = off; off += (uint32_t)r.TP * r.bpp[k]; off = (off + 127u) & ~127u; }
    r.stage_bytes = off;
    int stages = (int)(kRingBudget / r.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    r.stages = stages;
    const int64_t total = (int64_t)N * HW;
    int64_t ctas = (int64_t)hd_num_sms() * 3;
    const int64_t tiles = (total + r.TP - 1) / r.TP;
    if (ctas > tiles) ctas = tiles;
    int64_t per = (total + ctas - 1) / ctas;
    per = (per + r.TP - 1) / r.TP * r.TP;
    pl.per = (int)per;
    pl.grid = (int)((total + per - 1) / per);
    pl.r = r;
    pl.smem = 128 + (size_t)r.stages * r.stage_bytes + tail_bytes;
    return pl;
}

template <typename K> int set_smem(K kernel, size_t smem) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        hd_set_error("hd_gn: cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); return HD_ERR_CUDA;
    }
    return HD_OK;
}

bool gn2_shape_ok(int C0, int C1, int G, int64_t HW, int N) {
    const int C = C0 + C1;
    // every stream's pixel row must be a multiple of 16 bytes (bulk copies): channel counts in multiples of 8
    return G > 0 && G <= 64 && C % G == 0 && C0 % 8 == 0 && C1 % 8 == 0 && C / 4 <= kConsumers && HW < (1ll << 31)
        && (int64_t)N * HW * C < (1ll << 40) && (int64_t)N * HW < (1ll << 31) - (1 << 20);
}

#define GN2_FLAGS(KERNEL, ...)                                                                                          \
    do {                                                                                                                \
        if (act && drop) { auto kf = KERNEL<true, true>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }          \
        else if (act) { auto kf = KERNEL<true, false>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }            \
        else if (drop) { auto kf = KERNEL<false, true>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }           \
        else { auto kf = KERNEL<false, false>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }                    \
    } while (0)

#define GN2_FLAGS3(KERNEL, KPT, ...)                                                                                    \
    do {                                                                                                                \
        if (act && drop) { auto kf = KERNEL<true, true, KPT>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }     \
        else if (act) { auto kf = KERNEL<true, false, KPT>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }       \
        else if (drop) { auto kf = KERNEL<false, true, KPT>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }      \
        else { auto kf = KERNEL<false, false, KPT>; rc = set_smem(kf, pl.smem); if (!rc) kf __VA_ARGS__; }               \
    } while (0)

}  // namespace

// second-generation entry points (bf16 only), called from the hd_gn_* dispatchers in hd_fused.cu
int hd_gn2_supported(int C0, int C1, int G, int64_t HW, int N) {
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("HDIFF_GN_V1"); v1 = (e && e[0] == '1') ? 1 : 0; }
    return !v1 && gn2_shape_ok(C0, C1, G, HW, N);
}

int hd_gn2_stats(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, double* sums, cudaStream_t st) {
    GnArgs g{(const bf16*)in0, (const bf16*)in1, C0, C1, N, (int)HW, C0 + C1, G, nullptr, nullptr, nullptr, 0.f, 0.f, 0};
    if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * G, st) != cudaSuccess) return HD_ERR_CUDA;
    Ring r{}; add_stream(r, in0, C0); if (C1) add_stream(r, in1, C1);
    const Plan pl = finish_plan(r, N, (int)HW, g.C, 0, kKPTStats);
    int rc = set_smem(gn2_stats_kernel, pl.smem); if (rc) return rc;
    gn2_stats_kernel<<<pl.grid, kThreads, pl.smem, st>>>(g, pl.r, pl.per, sums);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

int hd_gn2_apply(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                 const float* beta, float eps, int act, float p_drop, uint64_t seed, void* out, cudaStream_t st) {
    GnArgs g{(const bf16*)in0, (const bf16*)in1, C0, C1, N, (int)HW, C0 + C1, G, sums, gamma, beta, eps, p_drop, hd_seed_mix(seed)};
    Ring r{}; add_stream(r, in0, C0); if (C1) add_stream(r, in1, C1);
    const Plan pl = finish_plan(r, N, (int)HW, g.C, 0, kKPTFwd);
    const int grid = pl.grid;
    const bool drop = p_drop > 0.f;
    int rc = HD_OK;
    GN2_FLAGS(gn2_apply_kernel, <<<grid, kThreads, pl.smem, st>>>(g, pl.r, pl.per, (bf16*)out));
    if (rc) return rc;
    HD_CHECK_LAUNCH();
    return HD_OK;
}

int hd_gn2_bwd_reduce(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                      const float* beta, float eps, int act, float p_drop, uint64_t seed, const void* dy, double* gsums,
                      float* dgamma, float* dbeta, cudaStream_t st) {
    GnArgs g{(const bf16*)in0, (const bf16*)in1, C0, C1, N, (int)HW, C0 + C1, G, sums, gamma, beta, eps, p_drop, hd_seed_mix(seed)};
    if (cudaMemsetAsync(gsums, 0, sizeof(double) * 2 * (size_t)N * G, st) != cudaSuccess) return HD_ERR_CUDA;
    Ring r{}; add_stream(r, in0, C0); if (C1) add_stream(r, in1, C1);
    add_stream(r, dy, g.C);
    const Plan pl = finish_plan(r, N, (int)HW, g.C, 2 * g.C * sizeof(float), kKPTReduce);
    const int grid = pl.grid;
    const bool drop = p_drop > 0.f;
    int rc = HD_OK;
    GN2_FLAGS(gn2_bwd_reduce_kernel, <<<grid, kThreads, pl.smem, st>>>(g, pl.r, pl.per, gsums, dgamma, dbeta));
    if (rc) return rc;
    HD_CHECK_LAUNCH();
    return HD_OK;
}

int hd_gn2_bwd_apply(const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums, const float* gamma,
                     const float* beta, float eps, int act, float p_drop, uint64_t seed, const void* dy, const double* gsums,
                     const void* add, const void* acc0, const void* acc1, void* dx0, void* dx1, float* cs_total, float* cs_per_n,
                     int64_t cs_ld, int cs_n, cudaStream_t st) {
    GnArgs g{(const bf16*)in0, (const bf16*)in1, C0, C1, N, (int)HW, C0 + C1, G, sums, gamma, beta, eps, p_drop, hd_seed_mix(seed)};
    GnBwdOut o{(bf16*)dx0, (bf16*)dx1, cs_total, cs_per_n, cs_ld, cs_n, -1, -1, -1, -1};
    Ring r{}; add_stream(r, in0, C0); if (C1) add_stream(r, in1, C1);
    o.s_dy = add_stream(r, dy, g.C);
    if (add) o.s_add = add_stream(r, add, g.C);
    if (acc0) o.s_acc0 = add_stream(r, acc0, C0);
    if (acc1 && C1) o.s_acc1 = add_stream(r, acc1, C1);
    // two operand streams per pixel (x, dy): four pixel slots per thread and tile; with addends: two (16 raw registers either way)
    const bool lean = !add && !acc0 && !(acc1 && C1);
    const Plan pl = finish_plan(r, N, (int)HW, g.C, g.C * sizeof(float), lean ? 4 : 2);
    const int grid = pl.grid;
    const bool drop = p_drop > 0.f;
    int rc = HD_OK;
    if (lean) GN2_FLAGS3(gn2_bwd_apply_kernel, 4, <<<grid, kThreads, pl.smem, st>>>(g, pl.r, pl.per, gsums, o));
    else GN2_FLAGS3(gn2_bwd_apply_kernel, 2, <<<grid, kThreads, pl.smem, st>>>(g, pl.r, pl.per, gsums, o));
    if (rc) return rc;
    HD_CHECK_LAUNCH();
    return HD_OK;
}
