// K5: AttnBlock's spatial self-attention as a flash-style tcgen05 kernel (single head, head_dim = C = 128).
//   reference: DiffusionFreeGuidence/ModelCondition.py:101-120 (q,k,v = 1x1 convs; w = softmax(q^T k * C^-1/2) over all
//   H*W keys; h = w v) — the [S,S] score matrix never reaches HBM here.
// Input is the fused QKV projection [N][S][3C] bf16 (q | k | v per pixel), output [N][S][C] bf16 + logsumexp [N][S] fp32.
//
// Forward: one CTA owns 128 query rows (TMEM lanes) and walks the keys in tiles of 64.
//   warp 0    TMA producer: Q once, then (K_j, V_j) into a 2-stage ring
//   warp 1    MMA issuer : S_j = Q K_j^T (SS, fp32 in TMEM) ; O += P_j V_j (A = P_j from TMEM, B = V_j MN-major)
//   warps 2-5 softmax    : thread == query row; S_j -> registers, running max with lazy rescale of O, exp2, P_j -> TMEM (bf16,
//                          two buffers)
// S_{j+1} is issued as soon as S_j has been read into registers, so the tensor pipe works under the exponentials.
// Two CTAs fit one SM (256 TMEM columns and ~99 KB of shared memory each), which overlaps the rest.
#include "hd_tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 192;
constexpr int BM = 128;                 // query rows per CTA
constexpr int BN = 64;                  // keys per tile
constexpr int D = 128;                  // head dim (= channels)
constexpr int kQBytes = BM * D * 2;     // 32 KB: two [128 rows][64 ch] blocks
constexpr int kKBytes = BN * D * 2;     // 16 KB: two [64 keys][64 ch] blocks
constexpr int kStageBytes = 2 * kKBytes;
constexpr int kStages = 2;
constexpr uint32_t kColS = 0, kColP = 64, kColO = 128, kTmemCols = 256;
constexpr float kRescaleThreshold = 8.f;   // log2 units: P <= 2^8 relative to the stale reference max

struct AttnFwdParams {
    int N, S, tiles;
    float scale_log2;
    __nv_bfloat16* out; float* lse;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const AttnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + kQBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kQBytes + kStages * kStageBytes);
    uint64_t* q_full = bars;              // [1]
    uint64_t* kv_full = bars + 1;         // [2]
    uint64_t* kv_empty = bars + 3;        // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_empty = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* pv_done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BM, n = blockIdx.y;
    const int T = p.tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, kQBytes);
            for (int blk = 0; blk < 2; ++blk)
                for (int half = 0; half < 2; ++half)
                    tma_load_3d(sQ + blk * (kQBytes / 2) + half * 8192, &mapQKV, q_full, blk * 64, q0 + half * 64, n);
            for (int j = 0; j < T; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], kStageBytes);
                uint8_t* st = sKV + s * kStageBytes;
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + blk * 8192, &mapQKV, &kv_full[s], D + blk * 64, j * BN, n);
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + kKBytes + blk * 8192, &mapQKV, &kv_full[s], 2 * D + blk * 64, j * BN, n);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
            const uint32_t idesc_pv = umma_idesc_bf16(BM, D, 0, 1);
            const uint32_t aQ = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sKV + s * kStageBytes);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {            // K = 16 channels per instruction
                    const int blk = kk >> 2, sub = kk & 3;
                    umma_bf16(tmem_base + kColS, umma_smem_desc(aQ + blk * (kQBytes / 2) + sub * 32, 16, 1024),
                              umma_smem_desc(aK + blk * 8192 + sub * 32, 16, 1024), idesc_s, kk != 0);
                }
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait(s_empty, j & 1);               // S_j is in registers: the columns are free
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                const int s = j & 1;
                const uint32_t aV = smem_u32(sKV + s * kStageBytes + kKBytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)               // K = 16 keys per instruction = 2048 B of V rows
                    umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + (j & 1) * 32 + kk * 8, umma_smem_desc(aV + kk * 2048, 8192, 1024),
                                 idesc_pv, (j | kk) != 0);
                umma_commit(&kv_empty[s]);
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
                    for (int c = 0; c < D; c += 32) {
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
            // P is double-buffered: P_j V_j may still be running when P_{j+1} is written; S_{j+2} (whose completion
            // gates the next write to this buffer) is issued after P_j V_j, so no further wait is needed
            tmem_st32(lane_base + kColP + (j & 1) * 32, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (T - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        __nv_bfloat16* orow = p.out + ((long long)n * p.S + q0 + row) * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
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
        p.lse[(long long)n * p.S + q0 + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------------
// Forward, second generation (used when S % 256 == 0): one CTA per SM owns TWO 128-row query groups.
//   * Q lives in TENSOR memory (written once by the softmax threads), so the score MMAs read only the K tile from shared
//     memory: S = Q K^T with A from TMEM.  P is written in place over the first half of the score columns.
//   * every K/V tile in shared memory serves both groups: half the TMA traffic per query row.
//   Shared-memory traffic per (128 queries x 64 keys) drops from 96 KB (first generation, smem-bandwidth bound) to 48 KB.
//   * the two groups ping-pong: while the softmax warps of one group work, the tensor pipe runs the other group's MMAs.
// warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: softmax of group 0, warps 6-9: softmax of group 1.
// TMEM (512 columns): group g at g*256: S/P 64 | Q 64 | O 128.
// ---------------------------------------------------------------------------------------------
// (A second MMA issuer, one per query group, was tried and made this kernel SLOWER: 1.07 -> 1.52 ms at N = 8, S = 16384.  The single
// issuer is what staggers the two groups — S of one group is issued right behind P V of the other — and two free-running issuers
// let both groups reach their softmax phase together, leaving the tensor pipe idle.  Starting the second issuer half a period late
// gave the same 1.52 ms, so the cause may be elsewhere; not understood.)
constexpr int kF2Threads = 64 + 256;
constexpr int kF2Stages = 6;
constexpr uint32_t kF2S = 0, kF2Q = 64, kF2O = 128, kF2Group = 256;

__global__ void __launch_bounds__(kF2Threads, 1)
attn_fwd2_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __nv_bfloat16* __restrict__ qkv, const AttnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kF2Stages * kStageBytes);
    uint64_t* kv_full = bars;                     // [6]
    uint64_t* kv_empty = bars + kF2Stages;        // [6]
    uint64_t* q_ready = bars + 2 * kF2Stages;     // [2]
    uint64_t* s_full = q_ready + 2;               // [2]
    uint64_t* p_full = q_ready + 4;               // [2]
    uint64_t* pv_done = q_ready + 6;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 2 * BM, n = blockIdx.y;
    const int T = p.tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        for (int s = 0; s < kF2Stages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&q_ready[g], 4); mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < T; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&kv_full[stage], kStageBytes);
                uint8_t* st = sKV + stage * kStageBytes;
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + blk * 8192, &mapQKV, &kv_full[stage], D + blk * 64, j * BN, n);
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + kKBytes + blk * 8192, &mapQKV, &kv_full[stage], 2 * D + blk * 64, j * BN, n);
                if (++stage == kF2Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
            const uint32_t idesc_pv = umma_idesc_bf16(BM, D, 0, 1);
            // S_g = Q_g K^T: A = Q_g from tensor memory (8 columns per 16 channels), B = K tile (K-major) of `stage`
            auto issue_s = [&](int g, int stage) {
                const uint32_t aK = smem_u32(sKV + stage * kStageBytes);
                const uint32_t tg = tmem_base + g * kF2Group;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int blk = kk >> 2, sub = kk & 3;
                    umma_bf16_ts(tg + kF2S, tg + kF2Q + kk * 8, umma_smem_desc(aK + blk * 8192 + sub * 32, 16, 1024), idesc_s, kk != 0);
                }
                umma_commit(&s_full[g]);
            };
            mbar_wait(&q_ready[0], 0);
            mbar_wait(&q_ready[1], 0);
            tc_fence_after();
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            int stage = 0, nstage = 1; uint32_t nphase = 0;      // stage of tile j, stage / phase of tile j + 1
            for (int j = 0; j < T; ++j) {
                const uint32_t aV = smem_u32(sKV + stage * kStageBytes + kKBytes);
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {
                    const uint32_t tg = tmem_base + g * kF2Group;
                    mbar_wait(&p_full[g], j & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ts(tg + kF2O, tg + kF2S + kk * 8, umma_smem_desc(aV + kk * 2048, 8192, 1024), idesc_pv, (j | kk) != 0);
                    umma_commit(&pv_done[g]);
                    if (g == 1) umma_commit(&kv_empty[stage]);   // every MMA that reads tile j has been issued
                    if (j + 1 < T) {
                        if (g == 0) { mbar_wait(&kv_full[nstage], nphase); tc_fence_after(); }
                        issue_s(g, nstage);                      // its P (in place) was consumed by the P V just issued
                    }
                }
                stage = nstage;
                if (++nstage == kF2Stages) { nstage = 0; nphase ^= 1; }
            }
        }
    } else {
        const int g = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + g * kF2Group;
        const long long grow = (long long)n * p.S + q0 + g * BM + row;
        {   // this thread's query row -> tensor memory (column c = channels 2c, 2c+1)
            const uint4* src = reinterpret_cast<const uint4*>(qkv + grow * (3 * D));
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                uint32_t q[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 v = __ldg(src + hlf * 8 + i);
                    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
                }
                tmem_st32(lane_base + kF2Q + hlf * 32, q);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_ready[g]);
        }
        const float sl2 = p.scale_log2;
        float m_ref = 0.f, l = 0.f;
        for (int j = 0; j < T; ++j) {
            mbar_wait(&s_full[g], j & 1);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld32(lane_base + kF2S, v);
            tmem_ld32(lane_base + kF2S + 32, v + 32);
            tmem_wait_ld();
            // row maximum by four independent chains (one serial chain of 32 FMNMX3 was ~150 cycles on the critical path of a tile)
            auto row_max = [&]() {
                float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
                for (int i = 4; i < 64; i += 4) {
                    m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                    m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
                }
                return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
            };
            uint32_t pk[32];
            float lsum;
            auto exps = [&]() {                       // P = exp2(S * scale - m_ref), its row sum
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float a = fast_exp2(fmaf(__uint_as_float(v[2 * i]), sl2, -m_ref));
                    const float b = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m_ref));
                    s0 += a; s1 += b;
                    pk[i] = pack_bf16x2(a, b);
                }
                lsum = s0 + s1;
            };
            if (j == 0) {
                m_ref = row_max();
                exps();
            } else {
                // The reference maximum only moves when a tile exceeds it by more than the rescale threshold (rare after the first
                // tiles), so the exponentials are formed SPECULATIVELY against the current one while the maximum of this tile is
                // reduced beside them (MUFU and ALU pipes in parallel); a row that does grow recomputes them from its scores.
                exps();
                const float mx = row_max();
                const bool grow_m = mx > m_ref + kRescaleThreshold;
                if (__any_sync(0xffffffffu, grow_m)) {
                    float alpha = 1.f;
                    if (grow_m) { alpha = fast_exp2(m_ref - mx); m_ref = mx; l *= alpha; exps(); }
                    // O is at rest: S_j completed, and P_{j-1} V was issued before S_j on the in-order tensor pipe
#pragma unroll 1
                    for (int c = 0; c < D; c += 32) {
                        uint32_t o[32];
                        tmem_ld32(lane_base + kF2O + c, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(lane_base + kF2O + c, o);
                    }
                    tmem_wait_st();
                }
            }
            l += lsum;
            tmem_st32(lane_base + kF2S, pk);          // in place: all 64 score columns of this row are in registers
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
        }
        mbar_wait(&pv_done[g], (T - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        __nv_bfloat16* orow = p.out + grow * D;
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t o[32];
            tmem_ld32(lane_base + kF2O + c, o);
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
        p.lse[grow] = (m_ref + log2f(l)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int make_qkv_map(CUtensorMap* m, const void* qkv, int N, int S, int C3, int box_rows) {
    uint64_t dims[3] = {(uint64_t)C3, (uint64_t)S, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)C3, (uint64_t)S * C3};
    uint32_t box[3] = {64, (uint32_t)box_rows, 1};
    return hd_make_tmap_bf16(m, qkv, 3, dims, str, box);
}

// ---------------------------------------------------------------------------------------------
// Backward.  Two launches of one templated kernel, no atomics, deterministic:
//   kDQ = true : the CTA owns 128 QUERY rows (X0 = Q, X1 = dO) and walks key tiles (Y0 = K, Y1 = V):   dQ  += dS K
//   kDQ = false: the CTA owns 128 KEY rows   (X0 = K, X1 = V ) and walks query tiles (Y0 = Q, Y1 = dO): dK += dS^T Q, dV += P^T dO
// Per ring tile (64 rows of Y), score buffers b = tile & 1:
//   St[b]  = X0 Y0^T   (scores, or their transpose)       SS MMA, fp32 in TMEM
//   dPt[b] = X1 Y1^T   (dO V^T, or its transpose)         SS MMA
//   8 element-wise warps (lane = row of X; two warps per lane quarter, 32 columns each):
//       P = exp2(St * scale*log2e - lse*log2e), dS = P * (dPt - delta), written as bf16 IN PLACE over the first half of the
//       warp's own St / dPt columns (so P / dS cost no extra tensor memory and the score buffers can be double-buffered)
//   acc0 += dS Y0   [+ acc1 += P Y1]                      A from TMEM, B = the same Y tiles read MN-major
// The MMA thread issues St/dPt of tile j+1 BEFORE it waits for the element-wise result of tile j, so the tensor pipe always
// has queued work and the mbarrier / TMEM-load latencies of the element-wise warps stay off the critical path.
// `stats` = [N][S][2] fp32 (lse*log2e, delta = rowsum(dO * O)), written by attn_stats_kernel; in the key-row pass the 64
// (lse, delta) pairs of a query tile travel with the tile through the shared-memory ring (one bulk copy).
// ---------------------------------------------------------------------------------------------
constexpr int kBwdStages = 3;
constexpr int kXBytes = BM * D * 2;        // 32 KB per stationary operand
constexpr int kYBytes = BN * D * 2;        // 16 KB per ring operand
constexpr int kStatBytes = BN * 8;         // 64 x (lse*log2e, delta)
constexpr int kBwdStageBytes = 2 * kYBytes + kStatBytes;
constexpr uint32_t kColSt = 0 /* 2 x 64 */, kColdPt = 128 /* 2 x 64 */, kColAcc0 = 256, kColAcc1 = 384;

struct AttnBwdParams {
    int N, S, tiles;
    float scale_log2, scale;
    const float2* stats;
    __nv_bfloat16* dqkv;
    const __nv_bfloat16* qkv; const __nv_bfloat16* dout;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 8 element-wise warps: two per TMEM lane quarter, each owning 32 of the 64 score columns of a tile (a single warp per
// scheduler cannot hide the MUFU / TMEM-load latencies of 128 values per thread and would leave the tensor pipe idle)
// warp 10: second MMA issuer.  One thread issues a tcgen05.mma every ~58 cycles at best (scripts/probe_queue.py) and a tile
// needs 24 (20 in the dQ pass): 1 390 cycles of issue against 1 024 of tensor-pipe time.  The work is split by TMEM buffer
// family so that every write-after-read ordering stays inside one in-order issuer:
//   issuer 0: St = X0 Y0^T  and  acc1 += P Y1    (P lives in the St columns)
//   issuer 1: dPt = X1 Y1^T and  acc0 += dS Y0   (dS lives in the dPt columns)
// (16 element-wise warps with 16 columns each instead of 8 x 32 were tried: identical time, 2.82 ms at N = 8, S = 16384 — the
// element-wise chain is not the limit.  The pass is bound by its MMA mix: 16 SS MMAs of N = 64 per tile cost 48-57 cycles
// each (probe: 57 for one issuer, 48 for two, against 32 of math) + 8 of N = 128 at 64 = 1 300-1 400 of the ~1 570 cycles.)
constexpr int kBwdSecondMma = 10;
constexpr int kBwdThreads = 64 + 256 + 32;

template <bool kDQ>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO, const AttnBwdParams p) {
    // dQ pass: only 384 TMEM columns are needed for scores + accumulator, so the stationary operands (Q, dO) live in the
    // remaining 128 columns (written once by the element-wise threads) and the score MMAs read just the K / V tile from
    // shared memory (A from TMEM): they run at the full N = 64 rate instead of being shared-memory bound, and the 64 KB of
    // shared memory they occupied become three more ring stages.
    constexpr bool kXT = kDQ;
    constexpr int kNS = kXT ? 5 : kBwdStages;
    constexpr uint32_t kColX0 = 384, kColX1 = 448;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX0 = smem;
    uint8_t* sX1 = smem + kXBytes;
    uint8_t* sY = smem + (kXT ? 0 : 2 * kXBytes);            // stages of (Y0 16 KB | Y1 16 KB)
    uint8_t* sStat = sY + kNS * 2 * kYBytes;                 // stages of 512 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + kNS * kStatBytes);
    uint64_t* x_full = bars;                       // [1]
    uint64_t* y_full = bars + 1;                   // [3]
    uint64_t* y_empty = bars + 1 + kNS;     // [3]
    uint64_t* s_full = bars + 1 + 2 * kNS;  // [2]
    uint64_t* p_full = s_full + 2;                 // [2]
    uint64_t* acc_done = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * BM, n = blockIdx.y;
    const int T = p.tiles;
    // channel offsets inside the qkv rows
    constexpr int cX0 = kDQ ? 0 : D, cX1 = kDQ ? 0 : 2 * D, cY0 = kDQ ? D : 0, cY1 = kDQ ? 2 * D : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV); tma_prefetch_desc(&mapDO);
        mbar_init(x_full, kXT ? 8 : 1);               // kXT: one arrival per element-wise warp (operands stored to TMEM)
        for (int s = 0; s < kNS; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 2); }       // both issuers release a tile
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 2); mbar_init(&p_full[b], 8); }
        mbar_init(acc_done, 2);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const CUtensorMap* mX1 = kDQ ? &mapDO : &mapQKV;
            const CUtensorMap* mY1 = kDQ ? &mapQKV : &mapDO;
            if (!kXT) {
                mbar_arrive_expect_tx(x_full, 2 * kXBytes);
                for (int blk = 0; blk < 2; ++blk)
                    for (int half = 0; half < 2; ++half) {
                        tma_load_3d(sX0 + blk * (kXBytes / 2) + half * 8192, &mapQKV, x_full, cX0 + blk * 64, r0 + half * 64, n);
                        tma_load_3d(sX1 + blk * (kXBytes / 2) + half * 8192, mX1, x_full, cX1 + blk * 64, r0 + half * 64, n);
                    }
            }
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < T; ++j) {
                mbar_wait(&y_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&y_full[stage], kDQ ? 2 * kYBytes : kBwdStageBytes);
                uint8_t* st = sY + stage * 2 * kYBytes;
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + blk * 8192, &mapQKV, &y_full[stage], cY0 + blk * 64, j * BN, n);
                for (int blk = 0; blk < 2; ++blk) tma_load_3d(st + kYBytes + blk * 8192, mY1, &y_full[stage], cY1 + blk * 64, j * BN, n);
                if (!kDQ) bulk_load_1d(sStat + stage * kStatBytes, p.stats + (long long)n * p.S + j * BN, kStatBytes, &y_full[stage]);
                if (++stage == kNS) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == kBwdSecondMma) {
        const int me = warp == 1 ? 0 : 1;
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
            const uint32_t idesc_acc = umma_idesc_bf16(BM, D, 0, 1);
            // issuer 0: scores from (X0, Y0), accumulates acc1 from P (St columns) and Y1
            // issuer 1: scores from (X1, Y1), accumulates acc0 from dS (dPt columns) and Y0
            const uint32_t aX = smem_u32(me == 0 ? sX0 : sX1);
            const uint32_t colS = me == 0 ? kColSt : kColdPt;                  // this issuer's score buffers
            const uint32_t colX = me == 0 ? kColX0 : kColX1;                   // its stationary operand in TMEM (dQ pass)
            const uint32_t colAcc = me == 0 ? kColAcc1 : kColAcc0;
            const uint32_t yS = me == 0 ? 0 : kYBytes;                         // score operand inside a ring stage
            const uint32_t yA = me == 0 ? kYBytes : 0;                         // accumulation operand (the OTHER Y tile, MN-major)
            const bool has_acc = !(kDQ && me == 0);                            // the dQ pass has no acc1
            int ld_stage = 0; uint32_t ld_phase = 0;
            auto issue_s = [&](int j) {
                mbar_wait(&y_full[ld_stage], ld_phase);
                tc_fence_after();
                const uint32_t aY = smem_u32(sY + ld_stage * 2 * kYBytes) + yS;
                const uint32_t b = (uint32_t)(j & 1) * 64;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int blk = kk >> 2, sub = kk & 3;
                    const uint64_t bd = umma_smem_desc(aY + blk * 8192 + sub * 32, 16, 1024);
                    if (kXT) umma_bf16_ts(tmem_base + colS + b, tmem_base + colX + kk * 8, bd, idesc_s, kk != 0);
                    else umma_bf16(tmem_base + colS + b, umma_smem_desc(aX + blk * (kXBytes / 2) + sub * 32, 16, 1024), bd, idesc_s, kk != 0);
                }
                umma_commit(&s_full[j & 1]);
                if (++ld_stage == kNS) { ld_stage = 0; ld_phase ^= 1; }
            };
            mbar_wait(x_full, 0);
            tc_fence_after();
            issue_s(0);
            int stage = 0;
            for (int j = 0; j < T; ++j) {
                // scores of tile j+1 go to the other buffer.  Its last readers were this issuer's accumulation MMAs of tile j-1
                // (the tensor pipe executes one thread's MMAs in issue order) and the element-wise loads of tile j-1, which
                // completed before p_full of tile j-1 — already waited for.
                if (j + 1 < T) issue_s(j + 1);
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                tc_fence_after();
                if (has_acc) {
                    const uint32_t aY = smem_u32(sY + stage * 2 * kYBytes) + yA;
                    const uint32_t b = (uint32_t)(j & 1) * 64;
                    // P / dS sit in the first 16 columns of each warp's 32-column range: K steps 0,1 -> +0,+8 ; 2,3 -> +32,+40
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ts(tmem_base + colAcc, tmem_base + colS + b + (kk >> 1) * 32 + (kk & 1) * 8,
                                     umma_smem_desc(aY + kk * 2048, 8192, 1024), idesc_acc, (j | kk) != 0);
                }
                umma_commit(&y_empty[stage]);        // covers this issuer's score MMAs on the tile as well
                if (++stage == kNS) stage = 0;
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
        if (kXT) {   // this thread's row of Q (h == 0) or dO (h == 1) -> tensor memory, column c = channels 2c, 2c+1
            const long long grow = (long long)n * p.S + r0 + row;
            const uint4* src = h == 0 ? reinterpret_cast<const uint4*>(p.qkv + grow * (3 * D))
                                      : reinterpret_cast<const uint4*>(p.dout + grow * D);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t q[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 v = __ldg(src + hf * 8 + i);
                    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
                }
                tmem_st32(lane_base + (h == 0 ? kColX0 : kColX1) + hf * 32, q);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full);
        }
        int stage = 0; uint32_t y_phase = 0;
        for (int j = 0; j < T; ++j) {
            if (!kDQ) mbar_wait(&y_full[stage], y_phase);    // the tile's (lse, delta) pairs were bulk-copied with it
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t tS = lane_base + kColSt + (uint32_t)(j & 1) * 64 + h * 32;
            const uint32_t tD = lane_base + kColdPt + (uint32_t)(j & 1) * 64 + h * 32;
            uint32_t sv[32], dv[32];
            tmem_ld32(tS, sv);
            tmem_ld32(tD, dv);
            tmem_wait_ld();
            const float4* st4 = reinterpret_cast<const float4*>(sStat + stage * kStatBytes) + h * 16;   // s_full implies y_full
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
            if (++stage == kNS) { stage = 0; y_phase ^= 1; }
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __nv_bfloat16* orow = p.dqkv + ((long long)n * p.S + r0 + row) * (3 * D) + (kDQ ? 0 : D);
#pragma unroll 1
        for (int a = 0; a < (kDQ ? 1 : 2); ++a) {
            const float mul = a == 0 ? p.scale : 1.f;
#pragma unroll 1
            for (int c = h * 64; c < h * 64 + 64; c += 32) {     // the two warps of a lane quarter split the 128 columns
                uint32_t o[32];
                tmem_ld32(lane_base + (a == 0 ? kColAcc0 : kColAcc1) + c, o);
                tmem_wait_ld();
                uint4* dst = reinterpret_cast<uint4*>(orow + a * D + c);
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

// stats[row] = (lse * log2e, sum_c dO * O); one warp per row of C = 128 channels
__global__ void attn_stats_kernel(const __nv_bfloat16* o, const __nv_bfloat16* dout, const float* lse, float2* stats, long long rows) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(o + r * D) + lane);
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(dout + r * D) + lane);
    const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float2 x = __bfloat1622float2(ah[i]), y = __bfloat1622float2(bh[i]);
        acc += x.x * y.x + x.y * y.y;
    }
    acc = hd_warp_sum(acc);
    if (lane == 0) stats[r] = make_float2(lse[r] * 1.4426950408889634f, acc);
}

}  // namespace

// wide heads (C = 256, 384, ...): hd_attn_wide_tc.cu
extern "C" int hd_attn_wide_tc_supported(int S, int C);
extern "C" int hd_attn_fwd_wide_tc(const void* qkv, void* out, float* lse, int N, int S, int C, cudaStream_t stream);
extern "C" int hd_attn_bwd_wide_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                                   int N, int S, int C, cudaStream_t stream);

extern "C" int hd_attn_tc_supported(int S, int C) {
    if (C != D) return hd_attn_wide_tc_supported(S, C);
    return (S >= BM && S % BM == 0) ? 1 : 0;
}
extern "C" int hd_attn_bwd_tc_supported(int S, int C) { return hd_attn_tc_supported(S, C); }

extern "C" int hd_attn_fwd_tc_scaled(const void* qkv, void* out, float* lse, int N, int S, float scale, cudaStream_t stream);
extern "C" int hd_attn_bwd_tc_scaled(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                                     int N, int S, float scale, cudaStream_t stream);

extern "C" int hd_attn_fwd_tc(const void* qkv, void* out, float* lse, int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && lse && N > 0);
    if (!hd_attn_tc_supported(S, C)) { hd_set_error("hd_attn_fwd_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    if (C != D) return hd_attn_fwd_wide_tc(qkv, out, lse, N, S, C, stream);
    return hd_attn_fwd_tc_scaled(qkv, out, lse, N, S, 1.f / sqrtf((float)C), stream);
}

// The 128-channel kernel with the softmax scale given by the caller: softmax(q k^T * scale) v.  The multi-head route (hd_mha.cu)
// runs zero-padded heads of dim hd here with scale = hd^-1/2.
extern "C" int hd_attn_fwd_tc_scaled(const void* qkv, void* out, float* lse, int N, int S, float scale, cudaStream_t stream) {
    const int C = D;
    HD_REQUIRE(qkv && out && lse && N > 0 && N <= 65535 && scale > 0.f);
    if (!hd_attn_tc_supported(S, C)) { hd_set_error("hd_attn_fwd_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    CUtensorMap m;
    int rc = make_qkv_map(&m, qkv, N, S, 3 * C, 64); if (rc) return rc;
    AttnFwdParams p{};
    p.N = N; p.S = S; p.tiles = S / BN;
    p.scale_log2 = 1.4426950408889634f * scale;
    p.out = (__nv_bfloat16*)out; p.lse = lse;
    const size_t smem = kQBytes + kStages * kStageBytes + 1024 + 16 * 8;
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { hd_set_error("cudaFuncSetAttribute(attn_fwd_tc_kernel)"); return HD_ERR_CUDA; }
        hd_mark_on_device(&attr_set);
    }
    static const bool use_v2 = !(getenv("HDIFF_ATTN_FWD_V1"));
    if (use_v2 && S % (2 * BM) == 0) {
        const size_t smem2 = kF2Stages * kStageBytes + 1024 + 32 * 8;
        static unsigned long long attr2 = 0;
        if (!hd_seen_on_device(&attr2)) {
            if (cudaFuncSetAttribute(attn_fwd2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) { hd_set_error("cudaFuncSetAttribute(attn_fwd2_tc_kernel)"); return HD_ERR_CUDA; }
            hd_mark_on_device(&attr2);
        }
        attn_fwd2_tc_kernel<<<dim3(S / (2 * BM), N), kF2Threads, smem2, stream>>>(m, (const __nv_bfloat16*)qkv, p);
        HD_CHECK_LAUNCH();
        return HD_OK;
    }
    attn_fwd_tc_kernel<<<dim3(S / BM, N), kThreads, smem, stream>>>(m, p);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// stats: scratch of N*S*2 floats ((lse*log2e, delta) per row)
extern "C" int hd_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                              int N, int S, int C, cudaStream_t stream) {
    HD_REQUIRE(qkv && out && dout && lse && stats && dqkv && N > 0);
    if (!hd_attn_bwd_tc_supported(S, C)) { hd_set_error("hd_attn_bwd_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    if (C != D) return hd_attn_bwd_wide_tc(qkv, out, dout, lse, stats, dqkv, N, S, C, stream);
    return hd_attn_bwd_tc_scaled(qkv, out, dout, lse, stats, dqkv, N, S, 1.f / sqrtf((float)C), stream);
}

extern "C" int hd_attn_bwd_tc_scaled(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                                     int N, int S, float scale, cudaStream_t stream) {
    const int C = D;
    HD_REQUIRE(qkv && out && dout && lse && stats && dqkv && N > 0 && N <= 65535 && scale > 0.f);
    if (!hd_attn_bwd_tc_supported(S, C)) { hd_set_error("hd_attn_bwd_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    CUtensorMap mQKV, mDO;
    int rc = make_qkv_map(&mQKV, qkv, N, S, 3 * C, 64); if (rc) return rc;
    rc = make_qkv_map(&mDO, dout, N, S, C, 64); if (rc) return rc;
    const long long rows = (long long)N * S;
    attn_stats_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse, (float2*)stats, rows);
    HD_CHECK_LAUNCH();
    AttnBwdParams p{};
    p.N = N; p.S = S; p.tiles = S / BN;
    p.scale = scale;
    p.scale_log2 = 1.4426950408889634f * p.scale;
    p.stats = (const float2*)stats; p.dqkv = (__nv_bfloat16*)dqkv;
    p.qkv = (const __nv_bfloat16*)qkv; p.dout = (const __nv_bfloat16*)dout;
    const size_t smem = 168 * 1024;    // key-row pass: 2 x 32 KB + 3 stages; dQ pass: 5 stages (operands in TMEM); + barriers, alignment
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            hd_set_error("cudaFuncSetAttribute(attn_bwd_tc_kernel)"); return HD_ERR_CUDA;
        }
        hd_mark_on_device(&attr_set);
    }
    attn_bwd_tc_kernel<false><<<dim3(S / BM, N), kBwdThreads, smem, stream>>>(mQKV, mDO, p);
    HD_CHECK_LAUNCH();
    attn_bwd_tc_kernel<true><<<dim3(S / BM, N), kBwdThreads, smem, stream>>>(mQKV, mDO, p);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
