// K5w: AttnBlock's spatial self-attention on tcgen05 for WIDE heads (single head, head_dim = C a multiple of 128 up to 1024):
//   the C = 256 (64x64) and C = 512 (32x32) attention levels of BASELINE.json configs[4]; hd_attn_tc.cu keeps C = 128.
//   reference: DiffusionFreeGuidence/ModelCondition.py:101-120 (and AttnBlock_old, diffusion/Model.py:194-222).
//
// One design for every width: nothing is resident in shared memory.  The contraction over channels runs in chunks of 64
// that travel through one ring of uniform slots, and a CTA produces a 128-COLUMN CHUNK z of the output (blockIdx.z), so
// tensor memory holds the same 128-column accumulators whatever C is.  The scores are recomputed once per output chunk:
// (C/128) x the exponentials of an ideal kernel — these levels are a few per cent of a training step next to the
// convolutions, and correctness on the tensor cores (instead of the fp32 SIMT fallback, ~100x slower) is what matters here.
//
// Slot kinds (ring order = consumption order of the single MMA thread, so a slot is never skipped):
//   score slot : [128 rows x 64 ch] of the row-stationary operand (16 KB) + [64 rows x 64 ch] of the walking operand (8 KB)
//   tail  slot : the walking tile's [64 rows x 128 ch] chunk z, read MN-major by the accumulating MMAs (16 KB each)
// Sequence: scores(0), scores(1), tail(0), scores(2), tail(1), ...  — scores of tile j+1 are issued before the element-wise
// result of tile j is awaited, exactly as in the C = 128 kernels.
#include "hd_tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128;                 // stationary rows per CTA (TMEM lanes)
constexpr int BN = 64;                  // walking rows per tile
constexpr int ZC = 128;                 // output columns per CTA
constexpr int kXChunk = BM * 64 * 2;    // 16 KB
constexpr int kYChunk = BN * 64 * 2;    // 8 KB
constexpr int kScoreTx = kXChunk + kYChunk;
constexpr float kRescaleThreshold = 8.f;

struct Ring {
    int slot; uint32_t phase; int depth;
    __device__ __forceinline__ void next() { if (++slot == depth) { slot = 0; phase ^= 1; } }
};

int make_rows_map(CUtensorMap* m, const void* base, int N, int S, int Crow) {
    uint64_t dims[3] = {(uint64_t)Crow, (uint64_t)S, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)Crow, (uint64_t)S * Crow};
    uint32_t box[3] = {64, 64, 1};
    return hd_make_tmap_bf16(m, base, 3, dims, str, box);
}

// ---------------------------------------------------------------------------------------------
// Forward: the CTA owns 128 query rows and output columns [z*128, z*128+128).
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax (thread == query row).
// TMEM (256 columns, two CTAs per SM): S 64 | P 2 x 32 | O 128.
// ---------------------------------------------------------------------------------------------
constexpr int kWfThreads = 192;
constexpr int kWfStages = 4;
constexpr int kWfSlot = kScoreTx;       // 24 KB
constexpr uint32_t kColS = 0, kColP = 64, kColO = 128, kWfTmemCols = 256;

struct WideFwdParams {
    int N, S, C, tiles, nch;
    float scale_log2;
    __nv_bfloat16* out; float* lse;
};

__global__ void __launch_bounds__(kWfThreads, 2)
attn_fwd_wide_kernel(const __grid_constant__ CUtensorMap mapQKV, const WideFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWfStages * kWfSlot);
    uint64_t* full = bars;                    // [4]
    uint64_t* empty = bars + kWfStages;       // [4]
    uint64_t* s_full = bars + 2 * kWfStages;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BM, n = blockIdx.y, z = blockIdx.z;
    const int T = p.tiles, C = p.C, nch = p.nch;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        for (int s = 0; s < kWfStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kWfTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            Ring r{0, 0, kWfStages};
            auto load_scores = [&](int j) {
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(&empty[r.slot], r.phase ^ 1);
                    mbar_arrive_expect_tx(&full[r.slot], kScoreTx);
                    uint8_t* st = ring + r.slot * kWfSlot;
                    tma_load_3d(st, &mapQKV, &full[r.slot], c * 64, q0, n);
                    tma_load_3d(st + 8192, &mapQKV, &full[r.slot], c * 64, q0 + 64, n);
                    tma_load_3d(st + kXChunk, &mapQKV, &full[r.slot], C + c * 64, j * BN, n);
                    r.next();
                }
            };
            load_scores(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) load_scores(j + 1);
                mbar_wait(&empty[r.slot], r.phase ^ 1);
                mbar_arrive_expect_tx(&full[r.slot], 2 * kYChunk);
                uint8_t* st = ring + r.slot * kWfSlot;
                tma_load_3d(st, &mapQKV, &full[r.slot], 2 * C + z * ZC, j * BN, n);
                tma_load_3d(st + 8192, &mapQKV, &full[r.slot], 2 * C + z * ZC + 64, j * BN, n);
                r.next();
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
            const uint32_t idesc_pv = umma_idesc_bf16(BM, ZC, 0, 1);
            Ring r{0, 0, kWfStages};
            auto issue_s = [&]() {
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(&full[r.slot], r.phase);
                    tc_fence_after();
                    const uint32_t a = smem_u32(ring + r.slot * kWfSlot);
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub)
                        umma_bf16(tmem_base + kColS, umma_smem_desc(a + sub * 32, 16, 1024),
                                  umma_smem_desc(a + kXChunk + sub * 32, 16, 1024), idesc_s, (c | sub) != 0);
                    umma_commit(&empty[r.slot]);
                    r.next();
                }
                umma_commit(s_full);
            };
            issue_s();
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait(s_empty, j & 1);               // S_j is in registers: the columns are free
                    tc_fence_after();
                    issue_s();
                }
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                mbar_wait(&full[r.slot], r.phase);
                tc_fence_after();
                const uint32_t aV = smem_u32(ring + r.slot * kWfSlot);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)               // K = 16 keys per instruction = 2048 B of V rows
                    umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + (j & 1) * 32 + kk * 8, umma_smem_desc(aV + kk * 2048, 8192, 1024),
                                 idesc_pv, (j | kk) != 0);
                umma_commit(&empty[r.slot]);
                r.next();
                umma_commit(pv_done);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = p.scale_log2;
        float m_ref = 0.f, l = 0.f;
        for (int j = 0; j < T; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld32(lane_base + kColS, v);
            tmem_ld32(lane_base + kColS + 32, v + 32);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);
            float mx = __uint_as_float(v[0]);
#pragma unroll
            for (int i = 1; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            mx *= sl2;
            if (j == 0) {
                m_ref = mx;
            } else {
                const bool grow = mx > m_ref + kRescaleThreshold;
                if (__any_sync(0xffffffffu, grow)) {
                    float alpha = 1.f;
                    if (grow) { alpha = fast_exp2(m_ref - mx); m_ref = mx; l *= alpha; }
                    mbar_wait(pv_done, (j - 1) & 1);         // O must be at rest: P_{j-1} V_{j-1} has completed
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < ZC; c += 32) {
                        uint32_t o[32];
                        tmem_ld32(lane_base + kColO + c, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(lane_base + kColO + c, o);
                    }
                    tmem_wait_st();
                }
            }
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float a = fast_exp2(fmaf(__uint_as_float(v[2 * i]), sl2, -m_ref));
                const float b = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m_ref));
                l += a + b;
                pk[i] = pack_bf16x2(a, b);
            }
            // P is double-buffered (see hd_attn_tc.cu): S_{j+2}, whose completion gates the next write to this buffer, is
            // issued after P_j V_j
            tmem_st32(lane_base + kColP + (j & 1) * 32, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (T - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        __nv_bfloat16* orow = p.out + ((long long)n * p.S + q0 + row) * C + z * ZC;
#pragma unroll 1
        for (int c = 0; c < ZC; c += 32) {
            uint32_t o[32];
            tmem_ld32(lane_base + kColO + c, o);
            tmem_wait_ld();
            uint4* dst = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                dst[i] = w;
            }
        }
        if (z == 0) p.lse[(long long)n * p.S + q0 + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kWfTmemCols); }
}

// ---------------------------------------------------------------------------------------------
// Backward: two launches of one templated kernel (deterministic, no atomics), as in hd_attn_tc.cu:
//   kDQ = true : the CTA owns 128 QUERY rows (X0 = Q, X1 = dO) and walks key tiles (Y0 = K, Y1 = V):   dQ[:, z] += dS K[:, z]
//   kDQ = false: the CTA owns 128 KEY rows   (X0 = K, X1 = V ) and walks query tiles (Y0 = Q, Y1 = dO):
//                dK[:, z] += dS^T Q[:, z], dV[:, z] += P^T dO[:, z]
// St[b] = X0 Y0^T and dPt[b] = X1 Y1^T over all C channels (score slots), then the 8 element-wise warps write P / dS as bf16
// in place over their own score columns, then acc0 += dS Y0[:, z] (+ acc1 += P Y1[:, z]) from the tail slot.
// TMEM (512 columns): St 2 x 64 | dPt 2 x 64 | acc0 128 | acc1 128.
// ---------------------------------------------------------------------------------------------
constexpr int kWbThreads = 64 + 256;
constexpr int kWbStages = 6;
constexpr int kWbSlot = 32768;
constexpr uint32_t kColSt = 0, kColdPt = 128, kColAcc0 = 256, kColAcc1 = 384;

struct WideBwdParams {
    int N, S, C, tiles, nch;
    float scale_log2, scale;
    const float2* stats;
    __nv_bfloat16* dqkv;
};

template <bool kDQ>
__global__ void __launch_bounds__(kWbThreads, 1)
attn_bwd_wide_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO, const WideBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float2* sStat = reinterpret_cast<float2*>(smem + kWbStages * kWbSlot);      // [8 warps][2 buffers][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWbStages * kWbSlot + 8 * 2 * 32 * 8);
    uint64_t* full = bars;                     // [6]
    uint64_t* empty = bars + kWbStages;        // [6]
    uint64_t* s_full = bars + 2 * kWbStages;   // [2]
    uint64_t* p_full = s_full + 2;             // [2]
    uint64_t* acc_done = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * BM, n = blockIdx.y, z = blockIdx.z;
    const int T = p.tiles, C = p.C, nch = p.nch;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV); tma_prefetch_desc(&mapDO);
        for (int s = 0; s < kWbStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // (tensor map, first channel) of the four operands inside the qkv / dO rows
            const CUtensorMap* mX[2] = {&mapQKV, kDQ ? &mapDO : &mapQKV};
            const CUtensorMap* mY[2] = {&mapQKV, kDQ ? &mapQKV : &mapDO};
            const int cX[2] = {kDQ ? 0 : C, kDQ ? 0 : 2 * C};
            const int cY[2] = {kDQ ? C : 0, kDQ ? 2 * C : 0};
            Ring r{0, 0, kWbStages};
            auto load_scores = [&](int j) {
                for (int w = 0; w < 2; ++w)
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait(&empty[r.slot], r.phase ^ 1);
                        mbar_arrive_expect_tx(&full[r.slot], kScoreTx);
                        uint8_t* st = ring + r.slot * kWbSlot;
                        tma_load_3d(st, mX[w], &full[r.slot], cX[w] + c * 64, r0, n);
                        tma_load_3d(st + 8192, mX[w], &full[r.slot], cX[w] + c * 64, r0 + 64, n);
                        tma_load_3d(st + kXChunk, mY[w], &full[r.slot], cY[w] + c * 64, j * BN, n);
                        r.next();
                    }
            };
            load_scores(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) load_scores(j + 1);
                mbar_wait(&empty[r.slot], r.phase ^ 1);
                mbar_arrive_expect_tx(&full[r.slot], (kDQ ? 1 : 2) * 2 * kYChunk);
                uint8_t* st = ring + r.slot * kWbSlot;
                for (int w = 0; w < (kDQ ? 1 : 2); ++w)
                    for (int blk = 0; blk < 2; ++blk)
                        tma_load_3d(st + w * 16384 + blk * 8192, mY[w], &full[r.slot], cY[w] + z * ZC + blk * 64, j * BN, n);
                r.next();
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
            const uint32_t idesc_acc = umma_idesc_bf16(BM, ZC, 0, 1);
            Ring r{0, 0, kWbStages};
            auto issue_s = [&](int j) {
                const uint32_t b = (uint32_t)(j & 1) * 64;
                for (int w = 0; w < 2; ++w)
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait(&full[r.slot], r.phase);
                        tc_fence_after();
                        const uint32_t a = smem_u32(ring + r.slot * kWbSlot);
#pragma unroll
                        for (int sub = 0; sub < 4; ++sub)
                            umma_bf16(tmem_base + (w == 0 ? kColSt : kColdPt) + b, umma_smem_desc(a + sub * 32, 16, 1024),
                                      umma_smem_desc(a + kXChunk + sub * 32, 16, 1024), idesc_s, (c | sub) != 0);
                        umma_commit(&empty[r.slot]);
                        r.next();
                    }
                umma_commit(&s_full[j & 1]);
            };
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                // scores of tile j+1 go to the other buffer; its last readers (accumulating MMAs of tile j-1, issued earlier
                // on the in-order tensor pipe, and the element-wise loads of tile j-1) are done — see hd_attn_tc.cu
                if (j + 1 < T) issue_s(j + 1);
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                tc_fence_after();
                mbar_wait(&full[r.slot], r.phase);
                tc_fence_after();
                const uint32_t aY0 = smem_u32(ring + r.slot * kWbSlot), aY1 = aY0 + 16384;
                const uint32_t b = (uint32_t)(j & 1) * 64;
                // P / dS sit in the first 16 columns of each warp's 32-column range: K steps 0,1 -> +0,+8 ; 2,3 -> +32,+40
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ts(tmem_base + kColAcc0, tmem_base + kColdPt + b + (kk >> 1) * 32 + (kk & 1) * 8,
                                 umma_smem_desc(aY0 + kk * 2048, 8192, 1024), idesc_acc, (j | kk) != 0);
                if (!kDQ) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ts(tmem_base + kColAcc1, tmem_base + kColSt + b + (kk >> 1) * 32 + (kk & 1) * 8,
                                     umma_smem_desc(aY1 + kk * 2048, 8192, 1024), idesc_acc, (j | kk) != 0);
                }
                umma_commit(&empty[r.slot]);
                r.next();
            }
            umma_commit(acc_done);
        }
    } else {
        const int quarter = warp & 3;
        const int h = (warp - 2) >> 2;               // which 32 of the 64 score columns this warp owns
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = p.scale_log2;
        float2 my = make_float2(0.f, 0.f);
        if (kDQ) my = __ldg(p.stats + (long long)n * p.S + r0 + row);
        // key-row pass: the (lse*log2e, delta) pairs belong to the COLUMNS (queries); each warp stages the 32 pairs of its
        // columns in a private shared-memory buffer (double-buffered, fetched one tile ahead) and reads them as broadcasts
        float2* wstat = sStat + (warp - 2) * 64;
        const float2* gstat = p.stats + (long long)n * p.S + h * 32 + lane;
        float2 nxt = make_float2(0.f, 0.f);
        if (!kDQ) nxt = __ldg(gstat);
        for (int j = 0; j < T; ++j) {
            if (!kDQ) {
                wstat[(j & 1) * 32 + lane] = nxt;
                if (j + 1 < T) nxt = __ldg(gstat + (long long)(j + 1) * BN);
                __syncwarp();
            }
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t tS = lane_base + kColSt + (uint32_t)(j & 1) * 64 + h * 32;
            const uint32_t tD = lane_base + kColdPt + (uint32_t)(j & 1) * 64 + h * 32;
            uint32_t sv[32], dv[32];
            tmem_ld32(tS, sv);
            tmem_ld32(tD, dv);
            tmem_wait_ld();
            const float4* st4 = reinterpret_cast<const float4*>(wstat + (j & 1) * 32);
            uint32_t pp[16], ds[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float2 s0 = my, s1 = my;
                if (!kDQ) {
                    const float4 q = st4[i];         // same address in every lane: a shared-memory broadcast
                    s0 = make_float2(q.x, q.y); s1 = make_float2(q.z, q.w);
                }
                const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -s0.x));
                const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -s1.x));
                const float d0 = p0 * (__uint_as_float(dv[2 * i]) - s0.y);
                const float d1 = p1 * (__uint_as_float(dv[2 * i + 1]) - s1.y);
                pp[i] = pack_bf16x2(p0, p1);
                ds[i] = pack_bf16x2(d0, d1);
            }
            tmem_st16(tD, ds);
            if (!kDQ) tmem_st16(tS, pp);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[j & 1]);
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __nv_bfloat16* orow = p.dqkv + ((long long)n * p.S + r0 + row) * (3 * C) + (kDQ ? 0 : C) + z * ZC;
#pragma unroll 1
        for (int a = 0; a < (kDQ ? 1 : 2); ++a) {
            const float mul = a == 0 ? p.scale : 1.f;
#pragma unroll 1
            for (int c = h * 64; c < h * 64 + 64; c += 32) {     // the two warps of a lane quarter split the 128 columns
                uint32_t o[32];
                tmem_ld32(lane_base + (a == 0 ? kColAcc0 : kColAcc1) + c, o);
                tmem_wait_ld();
                uint4* dst = reinterpret_cast<uint4*>(orow + a * C + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(o[8 * i]) * mul, __uint_as_float(o[8 * i + 1]) * mul);
                    w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * mul, __uint_as_float(o[8 * i + 3]) * mul);
                    w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * mul, __uint_as_float(o[8 * i + 5]) * mul);
                    w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * mul, __uint_as_float(o[8 * i + 7]) * mul);
                    dst[i] = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// stats[row] = (lse * log2e, sum_c dO * O); one warp per row of C channels (C % 128 == 0)
__global__ void attn_stats_wide_kernel(const __nv_bfloat16* o, const __nv_bfloat16* dout, const float* lse, float2* stats,
                                       long long rows, int C) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(o + r * C) + i);
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(dout + r * C) + i);
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float2 x = __bfloat1622float2(ah[k]), y = __bfloat1622float2(bh[k]);
            acc += x.x * y.x + x.y * y.y;
        }
    }
    acc = hd_warp_sum(acc);
    if (lane == 0) stats[r] = make_float2(lse[r] * 1.4426950408889634f, acc);
}

}  // namespace

extern "C" int hd_attn_wide_tc_supported(int S, int C) {
    return (C >= 128 && C <= 1024 && C % ZC == 0 && S >= BM && S % BM == 0) ? 1 : 0;
}

extern "C" int hd_attn_fwd_wide_tc(const void* qkv, void* out, float* lse, int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && lse && N > 0);
    if (!hd_attn_wide_tc_supported(S, C)) { hd_set_error("hd_attn_fwd_wide_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    CUtensorMap m;
    int rc = make_rows_map(&m, qkv, N, S, 3 * C); if (rc) return rc;
    WideFwdParams p{};
    p.N = N; p.S = S; p.C = C; p.tiles = S / BN; p.nch = C / 64;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)C);
    p.out = (__nv_bfloat16*)out; p.lse = lse;
    const size_t smem = kWfStages * kWfSlot + 1024 + 16 * 8;
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(attn_fwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { hd_set_error("cudaFuncSetAttribute(attn_fwd_wide_kernel)"); return HD_ERR_CUDA; }
        hd_mark_on_device(&attr_set);
    }
    attn_fwd_wide_kernel<<<dim3(S / BM, N, C / ZC), kWfThreads, smem, stream>>>(m, p);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

extern "C" int hd_attn_bwd_wide_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                                   int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && dout && lse && stats && dqkv && N > 0);
    if (!hd_attn_wide_tc_supported(S, C)) { hd_set_error("hd_attn_bwd_wide_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    CUtensorMap mQKV, mDO;
    int rc = make_rows_map(&mQKV, qkv, N, S, 3 * C); if (rc) return rc;
    rc = make_rows_map(&mDO, dout, N, S, C); if (rc) return rc;
    const long long rows = (long long)N * S;
    attn_stats_wide_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse, (float2*)stats, rows, C);
    HD_CHECK_LAUNCH();
    WideBwdParams p{};
    p.N = N; p.S = S; p.C = C; p.tiles = S / BN; p.nch = C / 64;
    p.scale = 1.f / sqrtf((float)C);
    p.scale_log2 = 1.4426950408889634f * p.scale;
    p.stats = (const float2*)stats; p.dqkv = (__nv_bfloat16*)dqkv;
    const size_t smem = kWbStages * kWbSlot + 8 * 2 * 32 * 8 + 1024 + 24 * 8;
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(attn_bwd_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(attn_bwd_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            hd_set_error("cudaFuncSetAttribute(attn_bwd_wide_kernel)"); return HD_ERR_CUDA;
        }
        hd_mark_on_device(&attr_set);
    }
    const dim3 grid(S / BM, N, C / ZC);
    attn_bwd_wide_kernel<false><<<grid, kWbThreads, smem, stream>>>(mQKV, mDO, p);
    HD_CHECK_LAUNCH();
    attn_bwd_wide_kernel<true><<<grid, kWbThreads, smem, stream>>>(mQKV, mDO, p);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
