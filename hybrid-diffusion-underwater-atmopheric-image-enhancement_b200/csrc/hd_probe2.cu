// Hardware probe (test infrastructure for the next kernel generation, not on the product path):
// a CTA-PAIR GEMM with tcgen05.mma.cta_group::2, next to the same GEMM issued by each CTA on its own (cta_group::1).
//   C[256][N] (fp32) = A[256][K] * B[N][K]^T      bf16 in, N <= 256, K a multiple of 64 (all K blocks resident)
// Why: the one-CTA SS-mode MMAs of the convolution / wgrad kernels are bound by their shared-memory operand fetch
// (~(4096 + 32 N) / 64 cycles per 128xNx16 MMA).  In a pair each CTA fetches its own 128 rows of A but only
// HALF of B, so the fetch per MMA drops to (4096 + 16 N) bytes.  The probe checks the mechanics (pair TMEM allocation, TMA
// completion on the leader's mbarrier, one issuing thread for both SMs, multicast commit) and measures cycles per MMA in
// both modes by repeating the MMA sequence `reps` times.
//   mode 1: every CTA loads A (its 128 rows) and the whole B, issues M = 128 MMAs itself
//   mode 2: every CTA loads A (its 128 rows) and HALF of B (rows rank*N/2 ..), the leader issues M = 256 MMAs for the pair
#include "hd_tc_common.cuh"

namespace {

__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
probe_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out, int N, int K, int reps,
                  long long* cycles, int shift, int fill, int cper, int ring) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = K / 64;
    const int Nb = kMode == 2 ? N / 2 : N;                 // rows of B held by this CTA
    uint8_t* sA = smem;                                    // [nkb][128 rows][128 B]
    uint8_t* sB = smem + nkb * 16384;                      // [nkb][Nb rows][128 B]
    uint8_t* sF = sB + nkb * Nb * 128;                     // [4][16 KB] scratch refilled by TMA while the MMAs run (fill > 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(sF + 4 * 16384);
    uint64_t* done = full + 1;
    uint64_t* fbar = full + 2;                             // [4]
    uint64_t* dummy = full + 6;                            // target of the periodic commits (cper > 0)
    uint64_t* rfull = full + 7;                            // [8] ring emulation (ring > 0): the handshake of a real pipeline,
    uint64_t* rempty = full + 15;                          // [8] without any data movement
    uint32_t* slot = reinterpret_cast<uint32_t*>(full + 23);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); for (int i = 0; i < 4; ++i) mbar_init(&fbar[i], 1); mbar_init(dummy, 1); for (int i = 0; i < 8; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 1); } fence_barrier_init(); }
    if (warp == 1) { if (kMode == 2) tmem_alloc_pair(slot, 256); else tmem_alloc(slot, 256); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t bytes = (uint32_t)(nkb * (16384 + Nb * 128));
            if (kMode == 2) {
                const uint32_t lead_bar = mapa_shared(smem_u32(full), 0);
                if (rank == 0) mbar_arrive_expect_tx(full, 2 * bytes);
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d_pair(sA + kb * 16384, &mapA, lead_bar, kb * 64, (int)rank * 128);
                    tma_load_2d_pair(sB + kb * Nb * 128, &mapB, lead_bar, kb * 64, (int)rank * Nb);
                }
            } else {
                mbar_arrive_expect_tx(full, bytes);
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(sA + kb * 16384, &mapA, full, kb * 64, (int)rank * 128);
                    tma_load_2d(sB + kb * Nb * 128, &mapB, full, kb * 64, 0);
                }
            }
            if (ring > 0 && kMode == 1) {      // ring emulation, producer side: wait for the stage to be released, hand it back
                const int total = reps * nkb * 4 / 12;
                int st = 0; uint32_t ph = 0;
                for (int i = 0; i < total; ++i) {
                    mbar_wait(&rempty[st], ph ^ 1);
                    mbar_arrive(&rfull[st]);
                    if (++st == ring) { st = 0; ph ^= 1; }
                }
            }
            if (fill > 0) {        // background fill: 16 KB boxes (L2 hits), four in flight, while the MMA thread works
                mbar_wait(full, 0);
                const long long t0 = clock64();
                for (int i = 0; i < fill; ++i) {
                    if (i >= 4) mbar_wait(&fbar[i & 3], ((i >> 2) - 1) & 1);
                    mbar_arrive_expect_tx(&fbar[i & 3], 16384);
                    tma_load_2d(sF + (i & 3) * 16384, &mapA, &fbar[i & 3], 0, (int)rank * 128);
                }
                for (int i = fill > 4 ? fill - 4 : 0; i < fill; ++i) mbar_wait(&fbar[i & 3], (i >> 2) & 1);
                if (blockIdx.x < 2) cycles[2 + rank] = clock64() - t0;
            }
        }
    } else if (warp == 1) {
        if (elect_one() && (kMode == 1 || rank == 0)) {
            mbar_wait(full, 0);
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(kMode == 2 ? 256 : 128, N, 0, 0);
            int since = 0;
            int rst = 0, rcnt = 0; uint32_t rph = 0;
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r)
                for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (ring > 0 && kMode == 1 && rcnt == 0) { mbar_wait(&rfull[rst], rph); tc_fence_after(); }
                        const uint64_t ad = umma_smem_desc(smem_u32(sA + kb * 16384) + shift * 128 + k * 32, 16, 1024);   // shift > 0: timing only
                        const uint64_t bd = umma_smem_desc(smem_u32(sB + kb * Nb * 128) + k * 32, 16, 1024);
                        if (kMode == 2) umma_bf16_pair(tmem, ad, bd, idesc, (r | kb | k) != 0);
                        else umma_bf16(tmem, ad, bd, idesc, (r | kb | k) != 0);
                        // a real kernel commits once per ring stage (to release it): does that cost tensor-pipe time?
                        if (cper > 0 && ++since == cper) { since = 0; if (kMode == 2) umma_commit_pair(dummy, 1); else umma_commit(dummy); }
                        if (ring < 0 && ++rcnt == 12) {            // ring < 0: the issuing thread idles -ring cycles after every 12 MMAs
                            rcnt = 0;
                            const long long w0 = clock64();
                            while (clock64() - w0 < -ring) { }
                        }
                        if (ring > 0 && kMode == 1 && ++rcnt == 12) {
                            rcnt = 0;
                            umma_commit(&rempty[rst]);
                            if (++rst == ring) { rst = 0; rph ^= 1; }
                        }
                    }
            if (kMode == 2) umma_commit_pair(done, 3);
            else umma_commit(done);
            mbar_wait(done, 0);
            if (blockIdx.x < 2) cycles[rank] = clock64() - t0;
        }
    } else {
        mbar_wait(done, 0);
        tc_fence_after();
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        float* orow = out + ((long long)rank * 128 + row) * N;
        for (int c = 0; c < N && blockIdx.x < 2; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + c, v);
            tmem_wait_ld();
            for (int i = 0; i < 16; ++i) orow[c + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (kMode == 2) tmem_dealloc_pair(tmem, 256); else tmem_dealloc(tmem, 256);
    }
}

// ---- how far can the issuing thread run ahead of the tensor pipe? ----
// One CTA, A[128][192] and B[N][192] resident.  `groups` times: 12 fully unrolled MMAs (descriptor low words with constant
// offsets: the minimal instruction stream), then the issuing thread busy-waits `gap` cycles.  While gap < (queue depth) x
// (cycles per MMA) the time per group stays 12 x (cycles per MMA); beyond that the pipe runs dry and the time grows.
__global__ void __launch_bounds__(256, 1)
probe_queue_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N, int groups, int gap,
                   int issuers, int second_warp, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                         // [3][128 rows][128 B]
    uint8_t* sB = smem + 3 * 16384;             // [3][N rows][128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + 3 * N * 128);
    uint64_t* done = full + 1;                  // [2]
    uint32_t* slot = reinterpret_cast<uint32_t*>(full + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(&done[0], 1); mbar_init(&done[1], 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(full, (uint32_t)(3 * (16384 + N * 128)));
        for (int kb = 0; kb < 3; ++kb) {
            tma_load_2d(sA + kb * 16384, &mapA, full, kb * 64, 0);
            tma_load_2d(sB + kb * N * 128, &mapB, full, kb * 64, 0);
        }
    }
    // issuing threads: lane 0 of warp 0 and (issuers == 2) of warp `second_warp`, each with its own accumulator
    const int who = warp == 0 ? 0 : (issuers == 2 && warp == second_warp ? 1 : -1);
    if (who >= 0 && lane == 0) {
        mbar_wait(full, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB));
        const uint32_t bkb = (uint32_t)(N * 128) >> 4;
        const uint32_t d = tmem + who * 256;
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
#pragma unroll
            for (int i = 0; i < 12; ++i)
                umma_bf16_lo(d, a_lo + (i >> 2) * 1024 + (i & 3) * 2, b_lo + (i >> 2) * bkb + (i & 3) * 2, idesc, (g | i) != 0);
            if (gap > 0) { const long long w0 = clock64(); while (clock64() - w0 < gap) { } }
        }
        umma_commit(&done[who]);
        mbar_wait(&done[who], 0);
        cycles[who] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
}  // namespace

// a: [256][K] bf16, b: [N][K] bf16, out: [256][N] fp32 (the sum of `reps` identical products), cycles: [4] int64 (MMA span per CTA, fill span per CTA)
// shift: the A operand starts `shift` rows into its swizzled box (the convolution's shifted-operand mode); results are only
// meaningful for shift == 0, the cycle counts for any shift.  fill: number of 16 KB TMA loads into scratch shared memory issued
// concurrently with the MMAs (mode 1 only).  cper: a tcgen05.commit to a dummy mbarrier every cper MMAs (0 = none).
// nclusters: run that many identical pairs at once (only the first reports): does the per-SM MMA rate hold when all SMs issue?
// ring: emulate the full/empty handshake of a `ring`-stage pipeline with 12 MMAs per stage (mode 1; no data is moved);
// ring < 0: the issuing thread busy-waits -ring cycles after every 12 MMAs (how deep is the tensor pipe's queue?): how much does the ring fill of a real kernel slow the MMA operand fetch down?
extern "C" int hd_probe_pair(const void* a, const void* b, float* out, int N, int K, int mode, int reps, long long* cycles, int shift, int fill, int cper,
                             int nclusters, int ring, cudaStream_t stream) {
    HD_REQUIRE(a && b && out && cycles && (mode == 1 || mode == 2) && reps >= 1 && nclusters >= 1 && nclusters <= 74 && ring <= 8);
    HD_REQUIRE(ring == 0 || (reps * (K / 16)) % 12 == 0);
    HD_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0 && K >= 64 && K % 64 == 0);
    const int Nb = mode == 2 ? N / 2 : N;
    const size_t smem = (size_t)(K / 64) * (16384 + Nb * 128) + 4 * 16384 + 1024 + 256;
    HD_REQUIRE(smem <= 220 * 1024);
    CUtensorMap mA, mB;
    uint64_t da[2] = {(uint64_t)K, 256}, sa[1] = {(uint64_t)K}; uint32_t ba[2] = {64, 128};
    int rc = hd_make_tmap_bf16(&mA, a, 2, da, sa, ba); if (rc) return rc;
    uint64_t db[2] = {(uint64_t)K, (uint64_t)N}, sb[1] = {(uint64_t)K}; uint32_t bb[2] = {64, (uint32_t)Nb};
    rc = hd_make_tmap_bf16(&mB, b, 2, db, sb, bb); if (rc) return rc;
    if (mode == 2) {
        if (cudaFuncSetAttribute(probe_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HD_ERR_CUDA;
        probe_pair_kernel<2><<<2 * nclusters, 192, smem, stream>>>(mA, mB, out, N, K, reps, cycles, shift, fill, cper, ring);
    } else {
        if (cudaFuncSetAttribute(probe_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HD_ERR_CUDA;
        probe_pair_kernel<1><<<2 * nclusters, 192, smem, stream>>>(mA, mB, out, N, K, reps, cycles, shift, fill, cper, ring);
    }
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// a: [128][192] bf16, b: [N][192] bf16; cycles[i] = clock64 span of `groups` x (12 MMAs + `gap` idle cycles) of issuing thread i;
// issuers = 2: a second thread (lane 0 of warp `second_warp`) issues the same stream into its own accumulator at the same time
extern "C" int hd_probe_queue(const void* a, const void* b, int N, int groups, int gap, int issuers, int second_warp, long long* cycles,
                              cudaStream_t stream) {
    HD_REQUIRE(a && b && cycles && N >= 32 && N <= 256 && N % 32 == 0 && groups >= 1 && gap >= 0);
    HD_REQUIRE((issuers == 1 || issuers == 2) && second_warp >= 2 && second_warp <= 7);
    CUtensorMap mA, mB;
    uint64_t da[2] = {192, 128}, sa[1] = {192}; uint32_t ba[2] = {64, 128};
    int rc = hd_make_tmap_bf16(&mA, a, 2, da, sa, ba); if (rc) return rc;
    uint64_t db[2] = {192, (uint64_t)N}, sb[1] = {192}; uint32_t bb[2] = {64, (uint32_t)N};
    rc = hd_make_tmap_bf16(&mB, b, 2, db, sb, bb); if (rc) return rc;
    const size_t smem = 3 * (16384 + (size_t)N * 128) + 1024 + 64;
    if (cudaFuncSetAttribute(probe_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HD_ERR_CUDA;
    probe_queue_kernel<<<1, 256, smem, stream>>>(mA, mB, N, groups, gap, issuers, second_warp, cycles);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
