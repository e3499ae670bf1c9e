// K1/K2: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16 in,
// fp32 accumulate.  One kernel covers every convolution of the path except the 3-channel head/tail:
//   * 3x3 and 1x1 stride-1 convolutions (ResBlock conv1/conv2, shortcut, AttnBlock q/k/v/proj, UpSample.c)
//       reference: nn.Conv2d call sites DiffusionFreeGuidence/ModelCondition.py:82,96-99,130,144,147
//   * DownSample's conv3x3 s2 + conv5x5 s2 (ModelCondition.py:71-76) and UpSample's ConvTranspose2d 5x5 s2
//       (ModelCondition.py:83) as stride-1 3x3 convolutions on 2x2 space-to-depth views
//   * the skip concatenation torch.cat([h, skip]) (ModelCondition.py:271) as a K-split over two sources
//   * every data gradient (dgrad) = the same kernel on flipped / transposed packed weights
// GEMM view: M = output pixels (tile of 128 = TH x TW patch of one image), N = output channels, K = taps x Cin.
// A tiles are TMA boxes of the NHWC activation tensor shifted by the tap offset (out-of-bounds = zero
// padding); B tiles are TMA boxes of the packed K-major weights.  Accumulators live in TMEM (two buffers,
// so the epilogue of tile i overlaps the main loop of tile i+1).  Epilogue: + bias + per-sample embedding
// + residual, bf16 store.
#include "hd_tc_common.cuh"
#include <mutex>
#include <stdio.h>
#include <stdlib.h>

namespace {

// warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue, warps 6..: further TMA producers.  One thread issues a
// K block's two tensor copies in ~600 cycles (measured: the kernel ran at a constant ~640 cycles per K block whatever the
// MMA width), so the K blocks are dealt round-robin to kProducers issuing threads.
constexpr int kProducers = 4;
// warps 2-9: epilogue.  Two warps per TMEM lane quarter, each taking every other 16-column chunk of the tile: with one warp
// per scheduler the epilogue (address arithmetic, TMEM loads, packing, stores) took ~3300 of the ~3600 cycles per tile of a
// K = 576 convolution (ncu: epilogue warps 93 % busy, main loop idle) and was the bottleneck.
constexpr int kEpiWarps = 8;
// Second MMA issuer (warp 10, another scheduler than warp 1): the two warps take alternate tiles, each with its own TMEM
// accumulator.  Measured (scripts/probe_queue.py): one thread issues a tcgen05.mma every ~58 cycles at best and everything
// else it executes (barrier waits, descriptor arithmetic, commits) is added on top — for N <= 128 the tensor pipe never gets
// ahead of a single issuer, so its gaps are tensor-pipe idle time; a second issuer fills them (24 MMAs in 1154 cycles at
// N = 64 whatever the gaps, against 693 + gaps per 12 for one).
constexpr int kSecondMma = 2 + kEpiWarps;
constexpr int kFirstExtraProducer = kSecondMma + 1;
constexpr int kThreads = 32 * (3 + kEpiWarps + kProducers - 1);
constexpr int kABytes = 128 * 128;     // 128 pixels x 64 bf16
constexpr int kMaxStages = 8;

// Timing experiments (HDIFF_CONV_DBG) exist only in the lab build (-DHDIFF_LAB, libhdiff_b200_lab.so): the product library
// carries neither the switch nor its branches.  & 4: CTA 0 records (clock64, globaltimer) at its start and end.
#ifdef HDIFF_LAB
__device__ long long g_conv_dbg[4];
#define HD_CONV_DBG(p) ((p).dbg)
#else
#define HD_CONV_DBG(p) 0
#endif

// residual add on the packed output: out = bf16(bf16(acc + bias + emb) + res), one HADD2.BF16 per pair instead of two unpacks and two
// fp32 adds (the epilogue of a K = 576 tile is as long as its main loop, and the residual doubled its instruction count: 64->64
// + residual 0.27 ms against 0.19 ms without).  The intermediate rounding is what h + shortcut(x) does in any bf16 evaluation of
// the reference module (ModelCondition.py:161: conv output, then the add).
__device__ __forceinline__ void add_res_bf16x8(uint4& o, const uint4& r) {
    __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&o);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hadd2(a[i], b[i]);
}

// Division by a launch constant without the ~25-instruction integer-division sequence: q = umulhi(x, mul) >> shr for 0 <= x < 2^31
// (mul = ceil(2^(31 + ceil_log2 d) / d)).  ncu source view of a 64->64 launch: the three div/mod pairs of the tile decomposition were
// 82 of ~400 instructions of an epilogue warp per tile, and that warp's in-order instruction chain IS the tile period of the K = 576 layers.
struct FastDiv {
    uint32_t mul, shr, d;
    __device__ __forceinline__ int div(int x) const { return d == 1 ? x : (int)(__umulhi((uint32_t)x, mul) >> shr); }
    __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = div(x); r = x - q * (int)d; }
};
static FastDiv make_fastdiv(int d) {
    FastDiv f{0u, 0u, (uint32_t)d};
    if (d > 1) {
        int lg = 0; while ((1ll << lg) < d) ++lg;
        const int pw = 31 + lg;
        f.mul = (uint32_t)(((1ull << pw) + (uint64_t)d - 1) / (uint64_t)d);
        f.shr = (uint32_t)(pw - 32);
    }
    return f;
}

struct ConvTcParams {
    int N, H, W, TH, TW, tiles_x, tiles_y, m_tiles, n_tiles, NT;
    int k, pad, P_in, nchunk0, nchunk_c, kblocks, stages;
    int Cout, P_out;
    const float* bias; const float* emb; long long emb_stride;
    const __nv_bfloat16* res; __nv_bfloat16* out;
    float* out_nchw; int nchw_c;     // tail: store only the first nchw_c (<= 16) channels, fp32 NCHW
    int txm;                         // 3x3, TH == 1: one (TW+2)-pixel halo box per (tap row, chunk) serves the 3 tap columns
    int a_slot, a_tx;                // shared-memory bytes reserved for / transferred into the A part of a stage
    double* chan_sums;               // optional [N][Cout][2]: per-image, per-channel sum / sum of squares of the STORED output
                                     // (the GroupNorm statistics of the next layer, without another pass over the tensor)
    int stage_bufs;                  // staging buffers of that epilogue: 2, or 4 (tensor stores of up to three earlier blocks may still be
                                     // reading shared memory while a block is staged; never with a residual, whose prefetch protocol alternates two)
    int stage_out;                   // epilogue stages 64-channel blocks of the tile in shared memory (128-byte swizzle) and writes them
                                     // with TMA tensor stores: whole 128-byte lines instead of 32-byte pieces per thread
    int wres, kb_w;                  // wres: the whole packed weight matrix (n_tiles x kb_w blocks of [NT][64] bf16) is loaded ONCE per CTA
                                     // and stays in shared memory; the ring then carries activations only
    int mma2;                        // two MMA-issuing warps: 1 = alternate tiles (one accumulator each), 2 = both work on every stage,
                                     // each on half of its K slices into its own PARTIAL accumulator (column offset 128; NT <= 128);
                                     // the epilogue adds the two partials
    int contig;                      // each CTA walks one contiguous range of tiles instead of striding over the grid
    int dbg;                         // timing experiments only (HDIFF_CONV_DBG): 1 = epilogue does no work, 2 = producers load nothing
    int nprod;                       // issuing threads in use (<= stages: a producer must never be two ring laps ahead,
                                     // the parity wait on `empty` cannot tell 0 completed phases from 2)
    FastDiv fd_nt, fd_tx, fd_ty, fd_txy;   // n_tiles, tiles_x, tiles_y, tiles_x * tiles_y
    int pair;                        // CTA pair (cluster of 2, tcgen05 cta_group::2): the two CTAs take two adjacent pixel tiles of the same
                                     // output-channel tile; each loads its own A and HALF of the weight tile, the leader issues M = 256
                                     // MMAs for both.  Halves the weight traffic L2 -> shared memory (the ring fill that bounds the N = 128
                                     // layers) and shrinks a stage from 65 to 41 KB (deeper ring, room for the staged epilogue)
};

template <bool kStats, bool kRes, bool kPair>   // kPair: CTA-pair mode (a separate instantiation: a kernel that CONTAINS cta_group::2
                                    // instructions cannot be launched without a cluster — "cluster misconfiguration"); kStats: the epilogue also accumulates p.chan_sums; kRes: a residual tensor is added (its loads are
                                    // prefetched).  Separate instantiations: the extra registers must not slow the plain epilogue down
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
               const __grid_constant__ CUtensorMap mapRes, const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int b_rows = kPair ? p.NT / 2 : p.NT;          // rows of a weight tile held by this CTA
    const int b_bytes = b_rows * 128;
    const int stage_bytes = p.a_slot + (p.wres ? 0 : (p.txm ? 3 : 1) * b_bytes);
    const int stage_tx = p.a_tx + (p.wres ? 0 : (p.txm ? 3 : 1) * b_bytes);
    uint8_t* stage_buf = smem + (size_t)p.stages * stage_bytes;                  // [2][128 pixels][64 channels] bf16 when p.stage_out
    uint8_t* wres_buf = stage_buf + (p.stage_out ? p.stage_bufs * kABytes : 0);             // [n_tiles][kb_w][NT][64] bf16 when p.wres
    uint64_t* bars = reinterpret_cast<uint64_t*>(wres_buf + (p.wres ? (size_t)p.n_tiles * p.kb_w * b_bytes : 0));
    uint64_t* full = bars;                       // [stages]
    uint64_t* empty = bars + kMaxStages;         // [stages]
    uint64_t* tfull = bars + 2 * kMaxStages;     // [2]
    uint64_t* tempty = bars + 2 * kMaxStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
    uint64_t* wfull = bars + 2 * kMaxStages + 5;   // resident weights have landed
    uint64_t* res_full = bars + 2 * kMaxStages + 6; // [2] residual block has landed in staging buffer b (staged epilogue + residual)
    float* addend = reinterpret_cast<float*>(bars + 2 * kMaxStages + 8);     // [2][256]: bias + embedding row of a tile
    float* stat_s = addend + 2 * 256;                                        // [2][256][2]: per-tile channel sums (p.chan_sums)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef HDIFF_LAB
    if ((p.dbg & 4) && blockIdx.x == 0 && threadIdx.x == 0) {
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_conv_dbg[0] = clock64(); g_conv_dbg[1] = t;
    }
#endif
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB);
        if (p.stage_out) { tma_prefetch_desc(&mapOut); if (kRes) tma_prefetch_desc(&mapRes); }
        // two issuers: a stage is released by BOTH (its owner's tcgen05.commit + a plain arrive of the other, who only watched
        // it fill), so neither can fall a ring lap behind — a parity wait cannot tell phase L from phase L + 2
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], p.mma2 ? 2 : 1); }
        // pair: the leader's `tempty` collects the epilogue warps of BOTH CTAs
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], p.mma2 == 2 ? 2 : 1); mbar_init(&tempty[a], kPair ? 2 * kEpiWarps : kEpiWarps); }
        mbar_init(wfull, 1);
        mbar_init(&res_full[0], 1); mbar_init(&res_full[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (kPair) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();       // pair: the peer's barriers must be initialised before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // pair: the loops below walk PAIR tiles q = (pair of adjacent pixel tiles, output-channel tile); this CTA's tile of q:
    const int total_tiles = kPair ? (p.m_tiles / 2) * p.n_tiles : p.m_tiles * p.n_tiles;
    auto tile_of = [&](int q) {
        if (!kPair) return q;
        int mp, nt; p.fd_nt.divmod(q, mp, nt);
        return (mp * 2 + (int)rank) * p.n_tiles + nt;
    };
    // tile -> (output-channel tile, x tile, y tile, image)
    auto decompose = [&](int tile, int& n_tile, int& tx_i, int& ty_i, int& n) {
        int m_tile, r;
        p.fd_nt.divmod(tile, m_tile, n_tile);
        p.fd_tx.divmod(m_tile, r, tx_i);
        p.fd_ty.divmod(r, n, ty_i);
    };
    const uint32_t full_lead = kPair ? mapa_shared(smem_u32(full), 0) : 0u;       // the leader's barriers, shared::cluster addresses
    const uint32_t tempty_lead = kPair ? mapa_shared(smem_u32(tempty), 0) : 0u;
    // tiles of this CTA: strided over the grid, or (p.contig) one contiguous range — then a CTA stays inside one image for
    // ~100 tiles, which the staged statistics need (one flush of fp64 atomics per image change)
    int t_begin = blockIdx.x, t_end = total_tiles, t_step = gridDim.x;
    if (kPair) { t_begin = blockIdx.x >> 1; t_step = gridDim.x >> 1; }
    if (p.contig) {
        const int walkers = kPair ? (int)gridDim.x >> 1 : (int)gridDim.x, me = kPair ? (int)blockIdx.x >> 1 : (int)blockIdx.x;
        const int per = (total_tiles + walkers - 1) / walkers;
        t_begin = me * per; t_end = t_begin + per < total_tiles ? t_begin + per : total_tiles; t_step = 1;
    }

    if (warp == 0 || warp >= kFirstExtraProducer) {
        const int prod = warp == 0 ? 0 : warp - kFirstExtraProducer + 1;   // which share of the K blocks this warp's elected thread issues
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int turn = 0;                                // global K-block counter modulo kProducers
            if (p.wres && prod == 0) {
                mbar_arrive_expect_tx(wfull, (uint32_t)(p.n_tiles * p.kb_w * b_bytes));
                for (int nt = 0; nt < p.n_tiles; ++nt)
                    for (int kb = 0; kb < p.kb_w; ++kb)
                        tma_load_2d(wres_buf + (size_t)(nt * p.kb_w + kb) * b_bytes, &mapB, wfull, kb * 64, nt * p.NT);
            }
            for (int q = t_begin; q < t_end; q += t_step) {
                const int tile = tile_of(q);
                int n_tile, tx_i, ty_i, n;
                decompose(tile, n_tile, tx_i, ty_i, n);
                const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
                if (kPair) {
                    // Both CTAs' copies complete on the LEADER's full barrier: its producer announces the bytes of both, the peer only
                    // issues its copies.  A stage is released in both CTAs at once by the leader's multicast commit.
                    const int b_row0 = n_tile * p.NT + (int)rank * b_rows;
                    if (p.txm) {
                        const int tap_stride = p.P_in * p.nchunk_c;
                        for (int ty = 0; ty < 3; ++ty)
                            for (int py = 0; py < p.P_in; ++py)
                                for (int cc = 0; cc < p.nchunk_c; ++cc) {
                                    if (turn == prod) {
                                        mbar_wait(&empty[stage], phase ^ 1);
                                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (uint32_t)stage_tx);
                                        const uint32_t fb = full_lead + 8u * (uint32_t)stage;
                                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                                        if (cc < p.nchunk0) tma_load_5d_pair(sa, &mapA0, fb, cc * 64, x0 - 1, py, y0 + ty - 1, n);
                                        else tma_load_5d_pair(sa, &mapA1, fb, (cc - p.nchunk0) * 64, x0 - 1, py, y0 + ty - 1, n);
                                        const int kb0 = (ty * 3 * p.P_in + py) * p.nchunk_c + cc;
                                        for (int tx = 0; tx < 3; ++tx)
                                            tma_load_2d_pair(sa + p.a_slot + tx * b_bytes, &mapB, fb, (kb0 + tx * tap_stride) * 64, b_row0);
                                    }
                                    if (++turn == p.nprod) turn = 0;
                                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                                }
                    } else {
                        int kb = 0;
                        for (int ty = 0; ty < p.k; ++ty)
                            for (int tx = 0; tx < p.k; ++tx)
                                for (int py = 0; py < p.P_in; ++py)
                                    for (int cc = 0; cc < p.nchunk_c; ++cc, ++kb) {
                                        if (turn == prod) {
                                            mbar_wait(&empty[stage], phase ^ 1);
                                            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (uint32_t)stage_tx);
                                            const uint32_t fb = full_lead + 8u * (uint32_t)stage;
                                            uint8_t* sa = smem + (size_t)stage * stage_bytes;
                                            if (cc < p.nchunk0) tma_load_5d_pair(sa, &mapA0, fb, cc * 64, x0 + tx - p.pad, py, y0 + ty - p.pad, n);
                                            else tma_load_5d_pair(sa, &mapA1, fb, (cc - p.nchunk0) * 64, x0 + tx - p.pad, py, y0 + ty - p.pad, n);
                                            tma_load_2d_pair(sa + p.a_slot, &mapB, fb, kb * 64, b_row0);
                                        }
                                        if (++turn == p.nprod) turn = 0;
                                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                                    }
                    }
                    continue;
                }
                // K order = (tap row, tap column, parity row, 64-channel chunk); plain counters, no divisions: this
                // single thread's loop rate bounds how fast shared memory can be filled
                if (p.txm) {
                    // shifted-operand mode: the MMA reads the three tap columns out of ONE halo box at row offsets 0 / 1 / 2
                    // (a tcgen05 shared-memory operand may start at any 128-byte row of a swizzled box, scripts/probe_shift.py).
                    // A stage then carries 12 MMAs instead of 4 for 1.7x the bytes: the ring is bounded by shared-memory
                    // capacity x TMA latency, so this raises the MMA work in flight.
                    const int tap_stride = p.P_in * p.nchunk_c;          // K blocks between two tap columns
                    for (int ty = 0; ty < 3; ++ty)
                        for (int py = 0; py < p.P_in; ++py)
                            for (int cc = 0; cc < p.nchunk_c; ++cc) {
                                if (turn == prod) {
                                    mbar_wait(&empty[stage], phase ^ 1);
                                    if (HD_CONV_DBG(p) & 2) { mbar_arrive(&full[stage]); goto txm_next; }
                                    {
                                    mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_tx);
                                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                                    if (cc < p.nchunk0) tma_load_5d(sa, &mapA0, &full[stage], cc * 64, x0 - 1, py, y0 + ty - 1, n);
                                    else tma_load_5d(sa, &mapA1, &full[stage], (cc - p.nchunk0) * 64, x0 - 1, py, y0 + ty - 1, n);
                                    const int kb0 = (ty * 3 * p.P_in + py) * p.nchunk_c + cc;
                                    if (!p.wres)
                                        for (int tx = 0; tx < 3; ++tx)
                                            tma_load_2d(sa + p.a_slot + tx * b_bytes, &mapB, &full[stage], (kb0 + tx * tap_stride) * 64, n_tile * p.NT);
                                    }
                                }
                                txm_next:
                                if (++turn == p.nprod) turn = 0;
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            }
                    continue;
                }
                int kb = 0;
                for (int ty = 0; ty < p.k; ++ty)
                    for (int tx = 0; tx < p.k; ++tx)
                        for (int py = 0; py < p.P_in; ++py)
                            for (int cc = 0; cc < p.nchunk_c; ++cc, ++kb) {
                                if (turn == prod) {
                                    mbar_wait(&empty[stage], phase ^ 1);
                                    mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_tx);
                                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                                    if (cc < p.nchunk0) tma_load_5d(sa, &mapA0, &full[stage], cc * 64, x0 + tx - p.pad, py, y0 + ty - p.pad, n);
                                    else tma_load_5d(sa, &mapA1, &full[stage], (cc - p.nchunk0) * 64, x0 + tx - p.pad, py, y0 + ty - p.pad, n);
                                    if (!p.wres) tma_load_2d(sa + p.a_slot, &mapB, &full[stage], kb * 64, n_tile * p.NT);
                                }
                                if (++turn == p.nprod) turn = 0;
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            }
            }
        }
    } else if (warp == 1 || warp == kSecondMma) {
        const int me = warp == 1 ? 0 : 1;            // this issuer's tile parity = its accumulator
        if (kPair) {
            // the leader's first issuer drives both SMs: M = 256 (this CTA's 128 pixels + the peer's), B halves from both CTAs
            if (me == 0 && rank == 0 && elect_one()) {
                const uint32_t idesc = umma_idesc_bf16(256, p.NT, 0, 0);
                int stage = 0; uint32_t phase = 0; int it = 0;
                const uint32_t ring_lo = umma_desc_lo(smem_u32(smem));
                const uint32_t st16 = (uint32_t)stage_bytes >> 4, aslot16 = (uint32_t)p.a_slot >> 4, b16 = (uint32_t)b_bytes >> 4;
                uint32_t a_lo = ring_lo;
                for (int q = t_begin; q < t_end; q += t_step, ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 256;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_lo = a_lo + aslot16;
                        if (p.txm) {
#pragma unroll
                            for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_lo_pair(d_tmem, a_lo + tx * 8 + 2 * k, b_lo + tx * b16 + 2 * k, idesc, (kb | tx | k) != 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_lo_pair(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit_pair(&empty[stage], 3);
                        a_lo += st16;
                        if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = ring_lo; }
                    }
                    umma_commit_pair(&tfull[acc], 3);
                }
            }
        } else if ((me == 0 || p.mma2) && elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, p.NT, 0, 0);
            int stage = 0; uint32_t phase = 0; int it = 0;
            if (p.wres) { mbar_wait(wfull, 0); tc_fence_after(); }
            const int per_row = p.P_in * p.nchunk_c;     // txm: stages per tap row = K blocks between two tap columns
            // descriptor low words (see umma_desc_lo): everything below is 32-bit adds on them
            const uint32_t ring_lo = umma_desc_lo(smem_u32(smem));
            const uint32_t st16 = (uint32_t)stage_bytes >> 4, aslot16 = (uint32_t)p.a_slot >> 4, b16 = (uint32_t)b_bytes >> 4;
            const uint32_t wres_lo = umma_desc_lo(smem_u32(wres_buf));
            const uint32_t bstep = p.wres ? (uint32_t)per_row * b16 : b16;         // between the B tiles of two tap columns (txm)
            uint32_t a_lo = ring_lo;                     // A operand of the current stage
            for (int q = t_begin; q < t_end; q += t_step, ++it) {
                const int tile = tile_of(q);
                const int acc = it & 1;
                if (p.mma2 == 2) {
                    // both issuers, every stage: K slices k = me, me + 2 of each operand pair into partial accumulator `me`
                    mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_part = tmem_base + acc * 256 + me * 128;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_lo = a_lo + aslot16;
                        if (p.txm) {
#pragma unroll
                            for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                                for (int i = 0; i < 2; ++i)
                                    umma_bf16_lo(d_part, a_lo + tx * 8 + 4 * i + 2 * me, b_lo + tx * b16 + 4 * i + 2 * me, idesc, (kb | tx | i) != 0);
                        } else {
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                umma_bf16_lo(d_part, a_lo + 4 * i + 2 * me, b_lo + 4 * i + 2 * me, idesc, (kb | i) != 0);
                        }
                        umma_commit(&empty[stage]);
                        a_lo += st16;
                        if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = ring_lo; }
                    }
                    umma_commit(&tfull[acc]);
                    continue;
                }
                if (p.mma2 && acc != me) {               // the other issuer's tile: watch its stages fill, in order
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        mbar_arrive(&empty[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    a_lo = ring_lo + (uint32_t)stage * st16;
                    continue;
                }
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                const uint32_t wb_lo = wres_lo + (uint32_t)((tile - p.fd_nt.div(tile) * p.n_tiles) * p.kb_w) * b16;
                uint32_t wk_lo = wb_lo;                  // resident weights: B tile of this stage's first K block
                int in_row = 0;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_lo = p.wres ? wk_lo : a_lo + aslot16;
                    if (p.txm) {
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx)                    // halo box shifted by tx pixels = tx rows of 128 B
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_lo(d_tmem, a_lo + tx * 8 + 2 * k, b_lo + tx * bstep + 2 * k, idesc, (kb | tx | k) != 0);
                        wk_lo += b16;
                        if (++in_row == per_row) { in_row = 0; wk_lo += 2 * (uint32_t)per_row * b16; }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // 4 x (K = 16): advance 32 bytes inside the 128-byte swizzle atom
                            umma_bf16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0);
                        wk_lo += b16;
                    }
                    umma_commit(&empty[stage]);
                    a_lo += st16;
                    if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = ring_lo; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;            // which of the two interleaved sets of 16-column chunks
        const int row = quarter * 32 + lane;
        const int etid = (warp - 2) * 32 + lane;     // 0 .. 255 among the epilogue threads
        const int ty = row / p.TW, tx = row % p.TW;
        int it = 0;
        uint32_t sb = 0;                             // staged 64-channel blocks so far (p.stage_out): buffer = sb % p.stage_bufs
        // staged epilogue + residual: the residual block is brought into the staging buffer by TMA one block ahead (issued
        // by thread 0 right after the barrier that frees the buffer) and the output is formed in place over it
        auto res_load = [&](int tl, int blk, uint32_t buf) {
            int nt, txi, tyi, nn;
            decompose(tl, nt, txi, tyi, nn);
            const int j0 = nt * p.NT + blk * 64;
            const int cv_w = p.P_out * p.Cout;
            const int py = j0 / cv_w;
            mbar_arrive_expect_tx(&res_full[buf], kABytes);
            tma_load_5d(stage_buf + buf * kABytes, &mapRes, &res_full[buf], j0 - py * cv_w, txi * p.TW, py, tyi * p.TH, nn);
        };
        // The residual loads are issued by lane 0 of the SECOND epilogue warp: one thread needs ~300 cycles to issue a tensor copy, and
        // thread 0 already issues the tile's tensor store (ncu: with both on thread 0 the other seven warps sat 11 % of their samples
        // at the next tile's barrier; 64->64 + residual 291 us against 199 us without one)
        constexpr int kResIssuer = 32;
        const bool early_res = kRes && p.stage_out && p.NT == 64;
        if (kRes && p.stage_out && etid == kResIssuer && t_begin < t_end) res_load(tile_of(t_begin), 0, 0);
        // hand an accumulator back to the MMA issuer (pair: the leader's barrier collects both CTAs' warps)
        auto release_acc = [&](int acc) {
            if (kPair) mbar_arrive_cluster(tempty_lead + 8u * (uint32_t)acc); else mbar_arrive(&tempty[acc]);
        };
        // staged epilogue + statistics (Cout == 64: one block per tile): per-channel (sum, sum of squares) of the STORED bf16
        // values are read back out of the staged tile; every warp keeps the sums of its 16 pixel rows in registers (lane:
        // channels 2 lane, 2 lane + 1) across this CTA's consecutive tiles of one image and adds them to p.chan_sums (fp64
        // atomics) when the image changes — a CTA walks a contiguous range of tiles, so that is once or twice per launch
        const bool sstat = p.stage_out && p.chan_sums != nullptr;
        int n_prev = -1;
        float st_s0 = 0.f, st_q0 = 0.f, st_s1 = 0.f, st_q1 = 0.f;
        auto stat_flush = [&](int nn) {
            double* d = p.chan_sums + ((long long)nn * p.Cout + 2 * lane) * 2;
            atomicAdd(d, (double)st_s0); atomicAdd(d + 1, (double)st_q0); atomicAdd(d + 2, (double)st_s1); atomicAdd(d + 3, (double)st_q1);
            st_s0 = st_q0 = st_s1 = st_q1 = 0.f;
        };
        for (int q = t_begin; q < t_end; q += t_step, ++it) {
            const int tile = tile_of(q);
            const int acc = it & 1;
            int n_tile, tx_i, ty_i, n;
            decompose(tile, n_tile, tx_i, ty_i, n);
            const int y = ty_i * p.TH + ty, x = tx_i * p.TW + tx;
            const bool valid = y < p.H;
            // per-channel addend (bias + this image's embedding row) staged once per tile while the MMAs still run; the
            // element loop then reads it as shared-memory broadcasts instead of 32 dependent global loads per chunk
            float* add_t = addend + acc * 256;
            // NT <= 256 = number of epilogue threads: one addend element per thread.  The loads for the NEXT tile are
            // issued here and only stored to shared memory after this tile's chunks, so their latency (a global round trip
            // per tile, which used to sit in front of the barrier below) overlaps the chunk processing.
            auto addend_load = [&](int tl) {
                float a = 0.f;
                if (etid < p.NT) {
                    int mt, nt; p.fd_nt.divmod(tl, mt, nt);
                    const int nn = p.fd_txy.div(mt);
                    const int j = nt * p.NT + etid;
                    if (p.bias) a = __ldg(p.bias + j);
                    if (p.emb) a += __ldg(p.emb + (long long)nn * p.emb_stride + j);
                }
                return a;
            };
            if (it == 0 && etid < p.NT) add_t[etid] = addend_load(tile);
            const bool has_next = q + t_step < t_end;
            const int tile_next = has_next ? tile_of(q + t_step) : tile;
            float a_next = 0.f;
            if (has_next) a_next = addend_load(tile_next);
            if (kStats) for (int c = etid; c < 2 * p.NT; c += 32 * kEpiWarps) stat_s[acc * 512 + c] = 0.f;
            asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight epilogue warps only: this tile's addend is staged
            if (sstat && n != n_prev) {
                if (n_prev >= 0) stat_flush(n_prev);
                n_prev = n;
            }
            // element offset of this thread's pixel for output-channel 0 of each parity (P_out == 2 stores the four
            // parities of the transposed convolution through the 2x2 view); hoisted out of the chunk loop
            long long pix1 = 0, pix2 = 0;
            if (!p.stage_out) {          // only the direct-store epilogue addresses global memory itself
                pix1 = (((long long)n * p.H + y) * p.W + x) * p.Cout;
                pix2 = (((long long)n * (2 * p.H) + 2 * y) * (2 * p.W) + 2 * x) * p.Cout;
            }
            const long long q_dx = p.Cout, q_dy = 2ll * p.W * p.Cout;
            uint4 pre_a[2], pre_b[2];
            const bool have_pre = kRes && valid && !p.out_nchw && !p.stage_out && half * 16 < p.NT;
            if (have_pre) {
                const int c0 = half * 16;
                long long o;
                {
                    const int j = n_tile * p.NT + c0;
                    if (p.P_out == 1) o = pix1 + j;
                    else { const int q = j / p.Cout, cph = j - q * p.Cout; o = pix2 + (q >> 1) * q_dy + (q & 1) * q_dx + cph; }
                }
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + o);
                pre_a[0] = __ldg(rp); pre_a[1] = __ldg(rp + 1);
                if (c0 + 32 < p.NT) {
                    const int j = n_tile * p.NT + c0 + 32;
                    if (p.P_out == 1) o = pix1 + j;
                    else { const int q = j / p.Cout, cph = j - q * p.Cout; o = pix2 + (q >> 1) * q_dy + (q & 1) * q_dx + cph; }
                    const uint4* rq = reinterpret_cast<const uint4*>(p.res + o);
                    pre_b[0] = __ldg(rq); pre_b[1] = __ldg(rq + 1);
                }
            }
            if (early_res && has_next && etid == 0) {
                // one block per tile: the other staging buffer held tile i-1's output, whose store was issued a whole tile ago.
                // Requesting tile i+1's residual NOW (not after this tile's store) puts it a full epilogue earlier into the TMA
                // queue, which the four producer threads keep ~3 operand stages deep.
                bulk_wait_group_read0();
                res_load(tile_next, 0, (sb + 1) & 1);
            }
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            if (HD_CONV_DBG(p) & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                if (has_next && etid < p.NT) addend[(acc ^ 1) * 256 + etid] = a_next;
                continue;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
            if (p.out_nchw) {
                if (n_tile == 0 && half == 0) {
                    uint32_t v[16];
                    tmem_ld16(taddr, v);
                    tmem_wait_ld();
                    if (p.mma2 == 2) {
                        uint32_t v2[16];
                        tmem_ld16(taddr + 128, v2);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
                    }
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (i < p.nchw_c) {
                                const float f = __uint_as_float(v[i]) + add_t[i];
                                p.out_nchw[(((long long)n * p.nchw_c + i) * p.H + y) * p.W + x] = f;
                            }
                        }
                    }
                }
            } else if (p.stage_out) {
                // One loop iteration of the eight warps completes a [128 pixels][64 channels] block (this warp: the 16-column
                // chunks half*16 and half*16 + 32 of its 32 pixels).  Buffer protocol, one barrier per block: the issuing thread
                // waits until every earlier store has been READ out of shared memory before it joins the barrier of block sb, so
                // after that barrier the other buffer (store sb-1) is free for block sb+1, and buffer sb&1 is complete.
                auto stage_emit = [&](const uint32_t* v, int c, uint8_t* srow, int cb, const uint4* pre) {
                    const float4* a4 = reinterpret_cast<const float4*>(add_t + c);
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = a4[i];
                        f[4 * i] = __uint_as_float(v[4 * i]) + a.x; f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + a.y;
                        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + a.z; f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + a.w;
                    }
                    uint4 o0, o1;
                    o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
                    o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                    if (kRes) {          // the residual block sits (swizzled) where the output is about to be written
                        const uint4 r0 = *reinterpret_cast<const uint4*>(srow + ((cb ^ (row & 7)) << 4));
                        const uint4 r1 = *reinterpret_cast<const uint4*>(srow + (((cb + 1) ^ (row & 7)) << 4));
                        add_res_bf16x8(o0, r0); add_res_bf16x8(o1, r1);
                    }
                    *reinterpret_cast<uint4*>(srow + ((cb ^ (row & 7)) << 4)) = o0;             // 128-byte swizzle: chunk ^= row % 8
                    *reinterpret_cast<uint4*>(srow + (((cb + 1) ^ (row & 7)) << 4)) = o1;
                };
                for (int c = half * 16; c < p.NT; c += 64) {
                    uint32_t va[16], vb[16];
                    const bool first = c == half * 16;
                    tmem_ld16(taddr + c, va);
                    tmem_ld16(taddr + c + 32, vb);
                    tmem_wait_ld();
                    if (p.mma2 == 2) {                  // add the second issuer's partial accumulator
                        uint32_t wa[16], wb[16];
                        tmem_ld16(taddr + 128 + c, wa);
                        tmem_ld16(taddr + 128 + c + 32, wb);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            va[i] = __float_as_uint(__uint_as_float(va[i]) + __uint_as_float(wa[i]));
                            vb[i] = __float_as_uint(__uint_as_float(vb[i]) + __uint_as_float(wb[i]));
                        }
                    }
                    if (c + 64 >= p.NT) {               // last TMEM read of the tile: hand the accumulator back to the MMA issuers now,
                        tc_fence_before();              // not after the staging, the barrier and the statistics below
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                    }
                    uint8_t* sbuf = stage_buf + (sb & (uint32_t)(p.stage_bufs - 1)) * kABytes;
                    if (kRes) mbar_wait(&res_full[sb & 1], (sb >> 1) & 1);
                    stage_emit(va, c, sbuf + row * 128, half * 2, nullptr);
                    stage_emit(vb, c + 32, sbuf + row * 128, half * 2 + 4, nullptr);
                    fence_proxy_async_smem();
                    if (etid == 0) { if (p.stage_bufs == 4) bulk_wait_group_read2(); else bulk_wait_group_read0(); }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (etid == 0) {
                        // logical channel block j0 of the tile -> (channel, parity row) of the output view (P_out == 2: the four
                        // parities of a transposed convolution are channel ranges q * Cout, q = 2 py + px)
                        const int j0 = n_tile * p.NT + (c - half * 16);
                        const int cv_w = p.P_out * p.Cout;
                        const int py = j0 / cv_w;
                        tma_store_5d(&mapOut, sbuf, j0 - py * cv_w, tx_i * p.TW, py, ty_i * p.TH, n);
                        bulk_commit_group();
                    }
                    if (kRes && etid == kResIssuer) {      // the other buffer is free (thread 0 saw its store read out before it joined the
                                                           // barrier above): fetch the next residual block
                        if (c + 64 < p.NT) res_load(tile, (c - half * 16) / 64 + 1, (sb + 1) & 1);
                        else if (has_next && !early_res) res_load(tile_next, 0, (sb + 1) & 1);
                    }
                    if (sstat) {         // this warp: 16 of the 128 staged pixel rows; lane: channels 2 lane, 2 lane + 1 of the block
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f, s2 = 0.f, s3 = 0.f, q2 = 0.f, q3 = 0.f;
                        const int w8 = warp - 2;
                        int rows_ok = (p.H - ty_i * p.TH) * p.TW;        // rows past the image are not stored
                        if (rows_ok > 128) rows_ok = 128;
                        const uint8_t* col = sbuf + (lane & 3) * 4;
                        const uint32_t ch16 = (uint32_t)lane >> 2;
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {                // two independent chains
                            const int r = w8 * 16 + i;
                            uint32_t wa = 0, wb = 0;
                            if (r < rows_ok) wa = *reinterpret_cast<const uint32_t*>(col + r * 128 + ((ch16 ^ (r & 7)) << 4));
                            if (r + 1 < rows_ok) wb = *reinterpret_cast<const uint32_t*>(col + (r + 1) * 128 + ((ch16 ^ ((r + 1) & 7)) << 4));
                            const float a0 = __uint_as_float(wa << 16), a1 = __uint_as_float(wa & 0xFFFF0000u);
                            const float b0 = __uint_as_float(wb << 16), b1 = __uint_as_float(wb & 0xFFFF0000u);
                            s0 += a0; q0 = fmaf(a0, a0, q0); s1 += a1; q1 = fmaf(a1, a1, q1);
                            s2 += b0; q2 = fmaf(b0, b0, q2); s3 += b1; q3 = fmaf(b1, b1, q3);
                        }
                        st_s0 += s0 + s2; st_q0 += q0 + q2; st_s1 += s1 + s3; st_q1 += q1 + q3;
                    }
                    ++sb;
                }
            } else {
                // residual of the first chunk pair of this warp, requested BEFORE the accumulator is awaited (it does not depend
                // on the MMAs): for NT <= 64 that is every residual load of the tile
                auto emit = [&](const uint32_t* v, int c, float* f, const uint4* pre) {
                    const int j = n_tile * p.NT + c;            // logical output channel of v[0]
                    long long off;
                    if (p.P_out == 1) {
                        off = pix1 + j;
                    } else {
                        const int q = j / p.Cout, cph = j - q * p.Cout;
                        off = pix2 + (q >> 1) * q_dy + (q & 1) * q_dx + cph;
                    }
                    const float4* a4 = reinterpret_cast<const float4*>(add_t + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = a4[i];
                        f[4 * i] = __uint_as_float(v[4 * i]) + a.x; f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + a.y;
                        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + a.z; f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + a.w;
                    }
                    uint4 o0, o1;
                    o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
                    o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                    if (kRes) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.res + off);
                        uint4 r0, r1;
                        if (pre) { r0 = pre[0]; r1 = pre[1]; } else { r0 = __ldg(rp); r1 = __ldg(rp + 1); }
                        add_res_bf16x8(o0, r0); add_res_bf16x8(o1, r1);
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.out + off);
                    op[0] = o0; op[1] = o1;
                    // the values as stored (bf16), for the statistics
                    if (kStats) {
                        const uint32_t pk[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) { f[2 * i] = __uint_as_float(pk[i] << 16); f[2 * i + 1] = __uint_as_float(pk[i] & 0xFFFF0000u); }
                    }
                };
                // column sums over the 32 pixels of this warp by a transposing butterfly: 32 quantities per lane in, the total
                // of quantity l on lane l out, 31 shuffles (a plain butterfly would take 160)
                auto stats = [&](const float* f, bool on, int c) {
                    float r[32];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { r[i] = on ? f[i] : 0.f; r[16 + i] = on ? f[i] * f[i] : 0.f; }
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const bool upper = (lane & step) != 0;
#pragma unroll
                        for (int i = 0; i < step; ++i) {
                            const float send = upper ? r[i] : r[i + step];
                            const float keep = upper ? r[i + step] : r[i];
                            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                        }
                    }
                    atomicAdd(&stat_s[acc * 512 + (c + (lane & 15)) * 2 + (lane >> 4)], r[0]);
                };
                // this warp's chunks: half*16, half*16 + 32, ...; two TMEM loads in flight per wait
                for (int c = half * 16; c < p.NT; c += 64) {
                    uint32_t va[16], vb[16];
                    const bool two = c + 32 < p.NT;
                    const bool first = c == half * 16;
                    tmem_ld16(taddr + c, va);
                    if (two) tmem_ld16(taddr + c + 32, vb);
                    tmem_wait_ld();
                    if (p.mma2 == 2) {                  // add the second issuer's partial accumulator
                        uint32_t wa[16], wb[16];
                        tmem_ld16(taddr + 128 + c, wa);
                        if (two) tmem_ld16(taddr + 128 + c + 32, wb);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            va[i] = __float_as_uint(__uint_as_float(va[i]) + __uint_as_float(wa[i]));
                            if (two) vb[i] = __float_as_uint(__uint_as_float(vb[i]) + __uint_as_float(wb[i]));
                        }
                    }
                    float fa[16], fb[16];
                    if (valid) {
                        emit(va, c, fa, first && have_pre ? pre_a : nullptr);
                        if (two) emit(vb, c + 32, fb, first && have_pre ? pre_b : nullptr);
                    }
                    if (kStats) {
                        stats(fa, valid, c);
                        if (two) stats(fb, valid, c + 32);
                    }
                }
            }
            if (!p.stage_out) {                      // (the staged epilogue released the accumulator after its last TMEM load)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            if (has_next && etid < p.NT) addend[(acc ^ 1) * 256 + etid] = a_next;     // visible after the next iteration's barrier
            if (kStats) {
                asm volatile("bar.sync 1, 256;" ::: "memory");          // every warp's contribution to this tile is in shared memory
                for (int c = etid; c < 2 * p.NT; c += 32 * kEpiWarps)
                    atomicAdd(p.chan_sums + ((long long)n * p.Cout + n_tile * p.NT) * 2 + c, (double)stat_s[acc * 512 + c]);
            }
        }
        if (sstat && n_prev >= 0) stat_flush(n_prev);
        if (p.stage_out && etid == 0) bulk_wait_group0();           // shared memory must outlive the last tensor store
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();       // pair: the leader's MMAs read the peer's shared memory until the last tile
    if (warp == 1) { tc_fence_after(); if (kPair) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
#ifdef HDIFF_LAB
    if ((p.dbg & 4) && blockIdx.x == 0 && threadIdx.x == 0) {
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_conv_dbg[2] = clock64(); g_conv_dbg[3] = t;
    }
#endif
}

int pick_nt(int CoutL) {
    if (CoutL <= 256) return (CoutL % 16 == 0) ? CoutL : 0;
    for (int nt = 256; nt >= 64; nt -= 64)
        if (CoutL % nt == 0) return nt;
    return 0;
}

bool conv_geometry(int H, int W, int* TH, int* TW) {
    int tw = W < 128 ? W : 128;
    if (tw < 8 || 128 % tw != 0 || W % tw != 0) return false;
    int th = 128 / tw;
    if (th > H) return false;
    *TH = th; *TW = tw;
    return true;
}


// set when a cluster launch was refused on this system (e.g. a partition without co-schedulable SM pairs): the one-CTA kernel
// takes over for the rest of the process
static int g_pair_refused = 0;
static bool pair_env_on() { static const bool on = !getenv("HDIFF_CONV_PAIR") || atoi(getenv("HDIFF_CONV_PAIR")) > 0; return on; }

// Shared-memory plan and kernel modes of a launch (p.N .. p.kblocks, p.NT, p.kb_w already set).  Returns false when nothing fits.
static bool conv_plan(ConvTcParams& p, int CoutL, int Cout, int P_in, int P_out, int ksize, int out_nchw_c, bool chan_sums,
                      bool allow_stage, bool has_res = false) {
    (void)P_out;
    // Shared-memory plan.  Options, each dropped when the ring would get too short:
        //   txm   shifted-operand mode: 3x3, a tile is one 128-pixel row segment (needs >= 3 stages)
        //   wres  resident weights: the packed weight matrix fits in 96 KB
        //   stage staged tensor-store epilogue (2 x 16 KB)
        static const bool txm_off = getenv("HDIFF_CONV_TXM_OFF") != nullptr;
        static const bool wres_on = getenv("HDIFF_CONV_WRES") != nullptr;    // measured: no gain (the operand fetch of the MMAs bounds
                                                                             // the kernel, not the TMA fill), so off by default
        static const int stage_env = getenv("HDIFF_CONV_STAGE") ? atoi(getenv("HDIFF_CONV_STAGE")) : 2;   // 0 off, 1 1x1 only, 2 every eligible conv
        static const int budget_kb = getenv("HDIFF_CONV_BUDGET") ? atoi(getenv("HDIFF_CONV_BUDGET")) : 200;
        const int budget = budget_kb * 1024;
        const long long wbytes = (long long)CoutL * p.kb_w * 128;
        const bool want_txm = !txm_off && ksize == 3 && p.TH == 1 && p.TW == 128;
        const bool want_wres = wres_on && wbytes <= 96 * 1024 && !(pair_env_on() && ksize == 3 && (p.NT == 128 || p.NT == 256));
        const bool want_stage = allow_stage && out_nchw_c == 0 && p.NT % 64 == 0 && Cout % 64 == 0 &&
                                (stage_env == 2 || (stage_env == 1 && ksize == 1));
        const int txm_a_tx = (p.TW + 2) * 128, txm_a_slot = (txm_a_tx + 1023) / 1024 * 1024;
        // CTA pair (see ConvTcParams::pair): every 3x3 layer with 64-, 128- or 256-channel tiles and an even number of pixel tiles.
        // scripts/conv_pair_bench.py (profiles/r02_conv_pair.txt): 128-channel tiles -25 % (128->128 @128^2 0.142 -> 0.107 ms; the 41 KB
        // stage also leaves room for the staged epilogue), 64->128 @256^2 -35 %, 256-channel tiles -10..-15 %, 64-channel tiles -4..-20 %.
        // HDIFF_CONV_PAIR=0 switches the mode off.
        static const int pair_env = getenv("HDIFF_CONV_PAIR") ? atoi(getenv("HDIFF_CONV_PAIR")) : 1;
        const bool pair_shape = p.NT == 64 || p.NT == 128 || p.NT == 256;
        // staging buffers: HDIFF_CONV_STAGE_BUFS=4 gives the 1x1 layers four (44: every staged layer without a residual)
        static const int sbufs_env = getenv("HDIFF_CONV_STAGE_BUFS") ? atoi(getenv("HDIFF_CONV_STAGE_BUFS")) : 2;
        const int nbufs = !has_res && ((sbufs_env == 4 && ksize == 1) || sbufs_env == 44) ? 4 : 2;
        const bool pair = pair_env > 0 && !g_pair_refused && ksize == 3 && pair_shape && out_nchw_c == 0 && p.m_tiles % 2 == 0 && p.m_tiles >= 2;
        const int b_rows = pair ? p.NT / 2 : p.NT;
        auto stages_of = [&](bool txm, bool wres, bool stage) {
            const int sbytes = (txm ? txm_a_slot : kABytes) + (wres ? 0 : (txm ? 3 : 1) * b_rows * 128);
            const long long avail = budget - (stage ? nbufs * kABytes : 0) - (wres ? wbytes : 0);
            return avail <= 0 ? 0 : (int)(avail / sbytes);
        };
        bool txm = false, wres = false, stage = false, found = false;
        for (int t = want_txm ? 1 : 0; t >= 0 && !found; --t)
            for (int w = want_wres ? 1 : 0; w >= 0 && !found; --w)
                for (int g = want_stage ? 1 : 0; g >= 0 && !found; --g)
                    if (stages_of(t, w, g) >= (t ? 3 : 2)) { txm = t; wres = w; stage = g; found = true; }
        if (!found) return false;
        p.txm = txm; p.wres = wres; p.stage_out = stage; p.stage_bufs = nbufs;
        p.pair = pair && !wres && (!chan_sums || stage);      // statistics: only the staged epilogue's (no template flag) in pair mode
        if (pair && !p.pair) {                      // planned with half weight tiles but the pair was dropped: plan again without it
            ConvTcParams q = p; q.m_tiles = 1;      // (an odd tile count switches the pair off)
            if (!conv_plan(q, CoutL, Cout, P_in, P_out, ksize, out_nchw_c, chan_sums, allow_stage, has_res)) return false;
            q.m_tiles = p.m_tiles; p = q;
            return true;
        }
#ifdef HDIFF_LAB
        static const int dbg = getenv("HDIFF_CONV_DBG") ? atoi(getenv("HDIFF_CONV_DBG")) : 0;
        p.dbg = dbg;
#else
        p.dbg = 0;
#endif
        static const int contig_env = getenv("HDIFF_CONV_CONTIG") ? atoi(getenv("HDIFF_CONV_CONTIG")) : -1;
        p.contig = contig_env >= 0 ? contig_env : (chan_sums && stage ? 1 : 0);
        // second issuer.  Mode 1 (alternate tiles): a gain where the issuing thread is the bottleneck (64->64: 0.201 -> 0.178 ms),
        // a small loss at N = 128 (+5 %: the watcher's arrive delays the release of a stage).  Mode 2 (both issuers on every
        // stage, partial accumulators): the same at 64->64, no loss at N = 128 (ring-fill bound: 0.139 ms either way) and a
        // further gain on tiles with many stages (128+64 -> 64: 0.473 -> 0.438 ms); on 1x1 layers (one stage per tile) mode 1 is the better one.  HDIFF_CONV_MMA2=0/1/2 forces a mode.
        static const int mma2_env = getenv("HDIFF_CONV_MMA2") ? atoi(getenv("HDIFF_CONV_MMA2")) : -1;
        int mode = 0;                                   // by shape, from per-layer timings inside a training step
        if (p.NT <= 64) mode = ksize == 1 ? 1 : 2;
        // K = 576 tiles WITH a residual are bound by their epilogue (lab build: MMA stream alone 0.142 ms, + epilogue 0.196, + residual
        // 0.262 at 64->64 @256^2); there the partial accumulators of mode 2 only add epilogue work: one issuer 0.247 ms
        if (p.NT <= 64 && ksize == 3 && has_res && p.nchunk_c * P_in <= 1) mode = 0;
        // ... and WITHOUT a residual by alternate tiles (mode 1: two accumulators, no partial sums for the epilogue to add): 64->64
        // @256^2 0.200 -> 0.171 ms, @128^2 0.063 -> 0.050 ms (profiles/r02_conv_pair.txt)
        else if (p.NT <= 64 && ksize == 3 && p.nchunk_c * P_in <= 1) mode = 1;
        else if (p.NT <= 128 && ksize == 3) mode = (p.txm && P_in == 1) ? 0 : 2;
        p.mma2 = mma2_env >= 0 ? mma2_env : mode;
        if (p.mma2 == 2 && (p.NT > 128 || p.wres)) p.mma2 = p.NT <= 64 ? 1 : 0;
        if (p.pair) p.mma2 = 0;                         // one issuer (the leader's) for both SMs
        p.a_slot = txm ? txm_a_slot : kABytes; p.a_tx = txm ? txm_a_tx : kABytes;
        if (txm) p.kblocks = 3 * P_in * p.nchunk_c;       // stages per tile: one per (tap row, parity row, chunk)
        p.stages = stages_of(txm, wres, stage); if (p.stages > kMaxStages) p.stages = kMaxStages;
    return true;
}

}  // namespace

// ---- host: tensor map helpers (shared with the other tcgen05 translation units) ----
hd_encode_tiled_fn hd_get_encode_tiled() {
    static hd_encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<hd_encode_tiled_fn>(sym);
    });
    return fn;
}

int hd_make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box) {
    hd_encode_tiled_fn enc = hd_get_encode_tiled();
    if (!enc) { hd_set_error("cuTensorMapEncodeTiled entry point unavailable"); return HD_ERR_DRIVER; }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 1; i < rank; ++i) gstr[i - 1] = strides_elems[i - 1] * 2;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static char buf[256];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
        hd_set_error(buf);
        return HD_ERR_DRIVER;
    }
    return HD_OK;
}

// NHWC activation view [N][P*H][P*W][C] as the 5-d tensor (Cv = P*C, W, P, H, N); box = 64 channels x TW x 1 x TH x 1
int hd_make_act_tmap(CUtensorMap* m, const void* base, int C, int P, int N, int H, int W, int box_c, int TW, int TH) {
    uint64_t dims[5] = {(uint64_t)P * C, (uint64_t)W, (uint64_t)P, (uint64_t)H, (uint64_t)N};
    uint64_t PW = (uint64_t)P * W, PH = (uint64_t)P * H;
    uint64_t str[4] = {(uint64_t)P * C, PW * C, (uint64_t)P * PW * C, PH * PW * C};
    uint32_t box[5] = {(uint32_t)box_c, (uint32_t)TW, 1, (uint32_t)TH, 1};
    return hd_make_tmap_bf16(m, base, 5, dims, str, box);
}

extern "C" int hd_conv_tc_supported(int C0, int C1, int P_in, int Cout, int P_out, int H, int W, int k) {
    if (k != 1 && k != 3) return 0;
    if (C0 <= 0 || C0 % 64 != 0 || C1 % 64 != 0) return 0;
    if (!(P_in == 1 || (P_in == 2 && C1 == 0))) return 0;
    if (!(P_out == 1 || P_out == 2)) return 0;
    if (Cout % 16 != 0) return 0;
    if (pick_nt(Cout * P_out * P_out) == 0) return 0;
    int TH, TW;
    return conv_geometry(H, W, &TH, &TW) ? 1 : 0;
}

extern "C" int hd_conv_tc(const void* in0, int C0, const void* in1, int C1, int P_in, const void* w, const float* bias,
                          const float* emb, int64_t emb_stride, const void* res, void* out, int Cout, int P_out,
                          int N, int H, int W, int ksize, int out_nchw_c, double* chan_sums, cudaStream_t stream) {
    HD_REQUIRE(in0 && w && out && N > 0);
    HD_REQUIRE(out_nchw_c >= 0 && out_nchw_c <= 16 && (out_nchw_c == 0 || (P_out == 1 && !emb && !res)));
    if (!hd_conv_tc_supported(C0, C1, P_in, Cout, P_out, H, W, ksize)) { hd_set_error("hd_conv_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    HD_REQUIRE(P_out == 1 || !emb);
    ConvTcParams p{};
    p.N = N; p.H = H; p.W = W;
    conv_geometry(H, W, &p.TH, &p.TW);
    p.tiles_x = W / p.TW; p.tiles_y = (H + p.TH - 1) / p.TH;
    p.m_tiles = N * p.tiles_x * p.tiles_y;
    const int CoutL = Cout * P_out * P_out;
    p.NT = pick_nt(CoutL); p.n_tiles = CoutL / p.NT;
    p.fd_nt = make_fastdiv(p.n_tiles); p.fd_tx = make_fastdiv(p.tiles_x); p.fd_ty = make_fastdiv(p.tiles_y);
    p.fd_txy = make_fastdiv(p.tiles_x * p.tiles_y);
    p.k = ksize; p.pad = ksize / 2; p.P_in = P_in;
    if (P_in == 1) { p.nchunk0 = C0 / 64; p.nchunk_c = (C0 + C1) / 64; }
    else { p.nchunk0 = 2 * C0 / 64; p.nchunk_c = p.nchunk0; }
    p.kblocks = ksize * ksize * P_in * p.nchunk_c;
    const int CinL = (C0 + C1) * P_in * P_in;
    p.kb_w = p.kblocks;
    // statistics in the staged epilogue handle one 64-channel block per tile; other widths with `chan_sums` take the direct-store
    // epilogue and its register butterfly
    HD_REQUIRE(conv_plan(p, CoutL, Cout, P_in, P_out, ksize, out_nchw_c, chan_sums != nullptr, !chan_sums || Cout == 64, res != nullptr));
    const int stage_bytes = p.a_slot + (p.wres ? 0 : (p.txm ? 3 : 1) * (p.pair ? p.NT / 2 : p.NT) * 128);
    p.nprod = p.stages < kProducers ? p.stages : kProducers;
    p.Cout = Cout; p.P_out = P_out;
    p.bias = bias; p.emb = emb; p.emb_stride = emb_stride;
    p.res = (const __nv_bfloat16*)res; p.out = (__nv_bfloat16*)out;
    p.out_nchw = out_nchw_c ? (float*)out : nullptr; p.nchw_c = out_nchw_c;
    HD_REQUIRE(!chan_sums || (P_out == 1 && out_nchw_c == 0));

    p.chan_sums = chan_sums;

    CUtensorMap mA0, mA1, mB, mOut, mRes;
    const int box_w = p.txm ? p.TW + 2 : p.TW;
    int rc = hd_make_act_tmap(&mA0, in0, C0, P_in, N, H, W, 64, box_w, p.TH); if (rc) return rc;
    if (p.stage_out) { rc = hd_make_act_tmap(&mOut, out, Cout, P_out, N, H, W, 64, p.TW, p.TH); if (rc) return rc; }
    else mOut = mA0;
    if (p.stage_out && res) { rc = hd_make_act_tmap(&mRes, res, Cout, P_out, N, H, W, 64, p.TW, p.TH); if (rc) return rc; }
    else mRes = mA0;
    if (C1 > 0) { rc = hd_make_act_tmap(&mA1, in1, C1, 1, N, H, W, 64, box_w, p.TH); if (rc) return rc; }
    else mA1 = mA0;
    {
        uint64_t dims[2] = {(uint64_t)ksize * ksize * CinL, (uint64_t)CoutL};
        uint64_t str[1] = {(uint64_t)ksize * ksize * CinL};
        uint32_t box[2] = {64, (uint32_t)(p.pair ? p.NT / 2 : p.NT)};
        rc = hd_make_tmap_bf16(&mB, w, 2, dims, str, box); if (rc) return rc;
    }
    const size_t smem = (size_t)p.stages * stage_bytes + (p.stage_out ? p.stage_bufs * kABytes : 0) + (p.wres ? (size_t)CoutL * p.kb_w * 128 : 0) + 1024 /*align slack*/ + (2 * kMaxStages + 8) * 8 + 2 * 256 * 4 + 2 * 512 * 4;
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(conv_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            hd_set_error("cudaFuncSetAttribute(conv_tc_kernel)"); return HD_ERR_CUDA;
        }
        hd_mark_on_device(&attr_set);
    }
    int grid = p.m_tiles * p.n_tiles; const int sms = hd_num_sms(); if (grid > sms) grid = sms;
    static const bool verbose = getenv("HDIFF_CONV_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "hd_conv_tc: N%d %dx%d %d+%d(P%d)->%d(P%d) k%d NT%d | pair %d txm %d stage_out %d wres %d mma2 %d stages %d stage_bytes %d smem %zu grid %d\n",
                N, H, W, C0, C1, P_in, Cout, P_out, ksize, p.NT, p.pair, p.txm, p.stage_out, p.wres, p.mma2, p.stages, stage_bytes, smem, grid);
    if (p.pair) {
        grid &= ~1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        const cudaError_t e = res ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true, true>, mA0, mA1, mB, mOut, mRes, p)
                                  : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, true>, mA0, mA1, mB, mOut, mRes, p);
        if (e == cudaErrorInvalidClusterSize || e == cudaErrorInvalidConfiguration || e == cudaErrorLaunchOutOfResources ||
            e == cudaErrorNotSupported) {
            // the launch was refused before anything ran: same layer again on the one-CTA kernel (still this library's tcgen05 path)
            (void)cudaGetLastError();
            g_pair_refused = 1;
            fprintf(stderr, "hdiff_b200: cluster launch of the CTA-pair convolution refused (%s); using the one-CTA kernel\n", cudaGetErrorString(e));
            return hd_conv_tc(in0, C0, in1, C1, P_in, w, bias, emb, emb_stride, res, out, Cout, P_out, N, H, W, ksize, out_nchw_c, chan_sums, stream);
        }
        if (e != cudaSuccess) { hd_set_error(cudaGetErrorString(e)); return HD_ERR_CUDA; }
        HD_CHECK_LAUNCH();
        return HD_OK;
    }
    if (chan_sums && !p.stage_out) {          // (the staged epilogue takes its statistics from the staged tile, no template flag)
        if (res) conv_tc_kernel<true, true, false><<<grid, kThreads, smem, stream>>>(mA0, mA1, mB, mOut, mRes, p);
        else conv_tc_kernel<true, false, false><<<grid, kThreads, smem, stream>>>(mA0, mA1, mB, mOut, mRes, p);
    } else {
        if (res) conv_tc_kernel<false, true, false><<<grid, kThreads, smem, stream>>>(mA0, mA1, mB, mOut, mRes, p);
        else conv_tc_kernel<false, false, false><<<grid, kThreads, smem, stream>>>(mA0, mA1, mB, mOut, mRes, p);
    }
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// timing experiments: (clock64, globaltimer ns) at the start and the end of CTA 0 of the last hd_conv_tc launch with HDIFF_CONV_DBG & 4
#ifdef HDIFF_LAB
extern "C" int hd_conv_dbg_read(long long* out4) {
    HD_REQUIRE(out4);
    if (cudaMemcpyFromSymbol(out4, g_conv_dbg, sizeof(long long) * 4) != cudaSuccess) { hd_set_error("cudaMemcpyFromSymbol"); return HD_ERR_CUDA; }
    return HD_OK;
}
#endif

// 1 if a launch of this shape with `chan_sums` takes the staged epilogue, where the GroupNorm statistics of the output are
// read back out of the staged tile in shared memory (cheap); 0 if it would run the register butterfly of the direct-store
// epilogue (measured slower than a separate hd_gn_stats pass) or the shape is not supported.
extern "C" int hd_conv_tc_stats_staged(int C0, int C1, int P_in, int Cout, int P_out, int H, int W, int ksize) {
    if (P_out != 1 || Cout != 64 || !hd_conv_tc_supported(C0, C1, P_in, Cout, P_out, H, W, ksize)) return 0;
    ConvTcParams p{};
    p.H = H; p.W = W;
    conv_geometry(H, W, &p.TH, &p.TW);
    const int CoutL = Cout * P_out * P_out;
    p.NT = pick_nt(CoutL); p.n_tiles = CoutL / p.NT;
    p.k = ksize; p.P_in = P_in;
    p.nchunk_c = P_in == 1 ? (C0 + C1) / 64 : 2 * C0 / 64;
    p.kblocks = ksize * ksize * P_in * p.nchunk_c;
    p.kb_w = p.kblocks;
    if (!conv_plan(p, CoutL, Cout, P_in, P_out, ksize, 0, true, true)) return 0;
    return p.stage_out ? 1 : 0;
}
