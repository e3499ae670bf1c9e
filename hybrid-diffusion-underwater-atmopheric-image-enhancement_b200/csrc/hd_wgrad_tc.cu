// K3: convolution weight gradient on tcgen05 / TMEM / TMA, bf16 in, fp32 accumulate.
//   dW[co][tap][ci] = sum over pixels of dY(pixel, co) * X(pixel + tap, ci)      (autograd of the nn.Conv2d /
//   nn.ConvTranspose2d call sites DiffusionFreeGuidence/ModelCondition.py:71-72,82-83,96-99,130,144,147)
// GEMM view: the reduction (K) runs over PIXELS, so both operands are "MN-major" in shared memory exactly as
// TMA delivers NHWC boxes ([pixel rows][64 channels], 128-byte swizzle):
//   A (M = 128) = two 64-wide blocks of shifted input X, one per (tap, 64-channel chunk)
//   B (N <= 256) = the dY box, 64-wide blocks of output channels
//   D[(tap, ci)][co] accumulates in TMEM; a CTA owns `G` such 128-row pairs and a contiguous range of
//   pixel tiles (split-K over pixels).  Partials go to an fp32 workspace [split][rows][N]; a second kernel
//   sums them in a fixed order (deterministic) and writes the packed [co][tap][ci] gradient.
#include "hd_tc_common.cuh"
#include <stdlib.h>

int hd_make_act_tmap(CUtensorMap* m, const void* base, int C, int P, int N, int H, int W, int box_c, int TW, int TH);

namespace {

// warp 0 + warps 6..: TMA producers (a stage is up to 12 tensor copies; one issuing thread is the bottleneck, so the
// copies of a stage are dealt round-robin to kProducers threads), warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kProducers = 4;
// warp 6: second MMA issuer (another scheduler than warp 1).  One thread issues a tcgen05.mma every ~58 cycles at best and
// everything else it executes is added on top (scripts/probe_queue.py), so for N <= 128 its gaps are tensor-pipe idle time;
// the two issuers take the even / odd accumulator groups of every stage and release it together.
constexpr int kSecondMma = 6;
constexpr int kThreads = 224 + 32 * (kProducers - 1);
constexpr int kBlkBytes = 64 * 128;    // one [64 pixels][64 channels] bf16 box
constexpr int kMaxStages = 8;

struct WgradTcParams {
    int N, H, W, TH, TW, tiles_x, tiles_y, pix_tiles, tiles_per_split, splits;
    int k, pad, P_in, nchunk0, nchunk_c;      // input side chunking (as in hd_conv_tc)
    int nblocks;                              // kk * P_in * nchunk_c  (64-row blocks of D)
    int G;                                    // 128-row pairs per CTA
    int ngroups;
    int NT, n_tiles, nb;                      // output-channel tile (<= 256), tiles, nb = NT / 64
    int nch_dy, P_dy;                         // 64-channel chunks per parity row of dY view
    int stages;
    int halo;                                 // 3x3, one-row pixel tiles: the input blocks of a tap row come out of ONE (TW+2)-pixel
    int max_units;                            //   halo box at row offsets 0/1/2; max_units = halo boxes per stage (over all groups)
    float* ws;                                // [splits][nblocks_padded * 64][CoutL]
    int rows_padded, CoutL;
};

constexpr int kHaloBytes = 66 * 128;          // [TW + 2 = 66 pixels][64 channels] bf16
constexpr int kHaloSlot = 9 * 1024;

// Blocks of a CTA's group -> halo units.  Block b (global index blk) is the input shifted by tap (ty, tx), parity row py,
// channel chunk cc; all blocks with the same (ty, py, cc) read one halo box at pixel offset tx.
struct GroupLayout {
    int nunits;
    int u_cc[8], u_py[8], u_ty[8];            // per unit (at most 8 per stage)
    int b_unit[16], b_tx[16];                 // per block (at most 2 x 8 groups)
};
__host__ __device__ inline void decode_group(const WgradTcParams& p, int group, GroupLayout* L) {
    L->nunits = 0;
    for (int b = 0; b < 2 * p.G; ++b) {
        int blk = group * 2 * p.G + b;
        if (blk >= p.nblocks) blk = p.nblocks - 1;
        const int cc = blk % p.nchunk_c; const int r2 = blk / p.nchunk_c;
        const int py = r2 % p.P_in; const int tap = r2 / p.P_in;
        const int ty = tap / 3, tx = tap % 3;
        int u = -1;
        for (int i = 0; i < L->nunits; ++i) if (L->u_cc[i] == cc && L->u_py[i] == py && L->u_ty[i] == ty) u = i;
        if (u < 0) { u = L->nunits++; L->u_cc[u] = cc; L->u_py[u] = py; L->u_ty[u] = ty; }
        L->b_unit[b] = u; L->b_tx[b] = tx;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX0, const __grid_constant__ CUtensorMap mapX1,
                const __grid_constant__ CUtensorMap mapDY, const WgradTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = p.halo ? p.max_units * kHaloSlot : p.G * 2 * kBlkBytes, b_bytes = p.nb * kBlkBytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* tfull = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x, split = blockIdx.y, n_tile = blockIdx.z;
    const int t_begin = split * p.tiles_per_split;
    int t_end = t_begin + p.tiles_per_split; if (t_end > p.pix_tiles) t_end = p.pix_tiles;
    const int nsteps = t_end - t_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapX0); tma_prefetch_desc(&mapX1); tma_prefetch_desc(&mapDY);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], kProducers); mbar_init(&empty[s], 2); }
        mbar_init(tfull, 2);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp > kSecondMma) {
        const int prod = warp == 0 ? 0 : warp - kSecondMma;
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            // bytes this producer brings per stage: copies b (0 .. 2G+nb-1) with b % kProducers == prod
            int my_loads = 0;
            for (int b = prod; b < 2 * p.G + p.nb; b += kProducers) ++my_loads;
            // the (tap, parity, chunk) coordinates of this CTA's operand blocks do not depend on the pixel tile: decode
            // them once (the divisions below used to sit in the per-tile loop of this single thread)
            constexpr int kMaxBlk = 8;
            int a_c[kMaxBlk], a_py[kMaxBlk], a_dy[kMaxBlk], a_dx[kMaxBlk];
#pragma unroll
            for (int b = 0; b < kMaxBlk; ++b) {
                int blk = group * 2 * p.G + b;
                if (blk >= p.nblocks) blk = p.nblocks - 1;               // padding rows: duplicate, discarded later
                const int cc = blk % p.nchunk_c; const int r2 = blk / p.nchunk_c;
                a_c[b] = cc; a_py[b] = r2 % p.P_in;
                const int tap = r2 / p.P_in;
                a_dy[b] = tap / p.k - p.pad; a_dx[b] = tap % p.k - p.pad;
            }
            const int dy_c0 = (n_tile * p.nb) % p.nch_dy, dy_py0 = (n_tile * p.nb) / p.nch_dy;
            int tx_i = t_begin % p.tiles_x; int r = t_begin / p.tiles_x;
            int ty_i = r % p.tiles_y; int n = r / p.tiles_y;
            if (p.halo) {
                GroupLayout L;
                decode_group(p, group, &L);
                // copies of a stage: the halo units, then the dY blocks; dealt round-robin to the producers
                uint32_t my_bytes = 0;
                for (int i = prod; i < L.nunits + p.nb; i += kProducers) my_bytes += i < L.nunits ? kHaloBytes : kBlkBytes;
                for (int t = t_begin; t < t_end; ++t) {
                    const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], my_bytes);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (u < L.nunits && (u % kProducers) == prod) {
                            if (L.u_cc[u] < p.nchunk0) tma_load_5d(sa + u * kHaloSlot, &mapX0, &full[stage], L.u_cc[u] * 64, x0 - 1, L.u_py[u], y0 + L.u_ty[u] - 1, n);
                            else tma_load_5d(sa + u * kHaloSlot, &mapX1, &full[stage], (L.u_cc[u] - p.nchunk0) * 64, x0 - 1, L.u_py[u], y0 + L.u_ty[u] - 1, n);
                        }
                    }
                    int cc = dy_c0, py = dy_py0;
                    for (int b = 0; b < p.nb; ++b) {
                        if (((L.nunits + b) % kProducers) == prod)
                            tma_load_5d(sa + a_bytes + b * kBlkBytes, &mapDY, &full[stage], cc * 64, x0, py, y0, n);
                        if (++cc == p.nch_dy) { cc = 0; ++py; }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    if (++tx_i == p.tiles_x) { tx_i = 0; if (++ty_i == p.tiles_y) { ty_i = 0; ++n; } }
                }
            } else
            for (int t = t_begin; t < t_end; ++t) {
                const int x0 = tx_i * p.TW, y0 = ty_i * p.TH;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], (uint32_t)(my_loads * kBlkBytes));
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
#pragma unroll
                for (int b = 0; b < kMaxBlk; ++b) {
                    if (b < 2 * p.G && (b % kProducers) == prod) {
                        if (a_c[b] < p.nchunk0) tma_load_5d(sa + b * kBlkBytes, &mapX0, &full[stage], a_c[b] * 64, x0 + a_dx[b], a_py[b], y0 + a_dy[b], n);
                        else tma_load_5d(sa + b * kBlkBytes, &mapX1, &full[stage], (a_c[b] - p.nchunk0) * 64, x0 + a_dx[b], a_py[b], y0 + a_dy[b], n);
                    }
                }
                int cc = dy_c0, py = dy_py0;
                for (int b = 0; b < p.nb; ++b) {
                    if (((2 * p.G + b) % kProducers) == prod)
                        tma_load_5d(sa + a_bytes + b * kBlkBytes, &mapDY, &full[stage], cc * 64, x0, py, y0, n);
                    if (++cc == p.nch_dy) { cc = 0; ++py; }
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                if (++tx_i == p.tiles_x) { tx_i = 0; if (++ty_i == p.tiles_y) { ty_i = 0; ++n; } }
            }
        }
    } else if (warp == 1 || warp == kSecondMma) {
        const int me = warp == 1 ? 0 : 1;            // this issuer takes the accumulator groups g with g % 2 == me
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, p.NT, 1, 1);
            int stage = 0; uint32_t phase = 0;
            // offset of the first block of each 128-row pair inside the A part of a stage, and the distance to its second block
            uint32_t a_off[8], a_lbo[8];
            if (p.halo) {
                GroupLayout L;
                decode_group(p, group, &L);
                for (int g = 0; g < 8; ++g) {
                    if (g < p.G) {
                        const uint32_t o0 = L.b_unit[2 * g] * kHaloSlot + L.b_tx[2 * g] * 128;
                        const uint32_t o1 = L.b_unit[2 * g + 1] * kHaloSlot + L.b_tx[2 * g + 1] * 128;
                        a_off[g] = o0; a_lbo[g] = o1 - o0;            // host guarantees o1 >= o0
                    } else { a_off[g] = 0; a_lbo[g] = 0; }
                }
            } else {
                for (int g = 0; g < 8; ++g) { a_off[g] = g * 2 * kBlkBytes; a_lbo[g] = kBlkBytes; }
            }
            // descriptor LOW words for ring stage 0 (address >> 4 | LBO >> 4 << 16; the high word — SBO 1024, version, 128-byte
            // swizzle — is one constant): stepping through stages and K slices is then a 32-bit add per operand
            const uint32_t ring = smem_u32(smem);
            uint32_t a_lo[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) a_lo[g] = (((ring + a_off[g]) & 0x3FFFF) >> 4) | (((a_lbo[g] >> 4) & 0x3FFF) << 16);
            const uint32_t b_lo0 = (((ring + (uint32_t)a_bytes) & 0x3FFFF) >> 4) | (((uint32_t)kBlkBytes >> 4) << 16);
            const uint32_t st16 = (uint32_t)stage_bytes >> 4;
            uint32_t soff = 0;                       // (stage * stage_bytes) >> 4
            for (int s = 0; s < nsteps; ++s) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t b_lo = b_lo0 + soff;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (g >= p.G) break;
                    if ((g & 1) != me) continue;
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // K = 16 pixels = 2 groups of 8 rows = 2048 bytes
                        umma_bf16_lo(tmem_base + g * p.NT, a_lo[g] + soff + k * 128, b_lo + k * 128, idesc, (s | k) != 0);
                }
                umma_commit(&empty[stage]);          // (an issuer without groups arrives at once)
                soff += st16;
                if (++stage == p.stages) { stage = 0; phase ^= 1; soff = 0; }
            }
            umma_commit(tfull);
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        if (nsteps > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int g = 0; g < p.G; ++g) {
            const int R = (group * p.G + g) * 128 + row;               // row of D: block = R / 64
            float* dst = p.ws + ((size_t)split * p.rows_padded + R) * p.CoutL + (size_t)n_tile * p.NT;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + g * p.NT;
            for (int c = 0; c < p.NT; c += 16) {
                uint32_t v[16];
                if (nsteps > 0) { tmem_ld16(taddr + c, v); tmem_wait_ld(); }
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0;
                }
                float4* d4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dw[co][tap][ci] = sum_s ws[s][blk * 64 + r][co],  blk = tap * (CinL/64) + ci / 64, r = ci % 64
__global__ void wgrad_reduce_kernel(const float* ws, int splits, int rows_padded, int rows, int CoutL, int kk, int CinL, float* dw) {
    const int64_t total = (int64_t)rows * CoutL;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % CoutL); const int R = (int)(i / CoutL);
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += ws[((size_t)sp * rows_padded + R) * CoutL + co];
        const int tap = R / CinL, ci = R - tap * CinL;
        dw[((int64_t)co * kk + tap) * CinL + ci] = s;
    }
}

bool wgrad_geometry(int H, int W, int* TH, int* TW) {
    int tw = W < 64 ? W : 64;
    if (tw < 8 || 64 % tw != 0 || W % tw != 0) return false;
    int th = 64 / tw;
    if (H % th != 0) return false;
    *TH = th; *TW = tw;
    return true;
}
int pick_nt_w(int CoutL) {
    if (CoutL <= 256) return CoutL % 64 == 0 ? CoutL : 0;
    for (int nt = 256; nt >= 64; nt -= 64) if (CoutL % nt == 0) return nt;
    return 0;
}

struct WgradPlan { WgradTcParams p; };

bool make_plan(int C0, int C1, int P_in, int Cdy, int P_dy, int N, int H, int W, int k, WgradTcParams* out) {
    WgradTcParams p{};
    if (k != 1 && k != 3) return false;
    if (C0 <= 0 || C0 % 64 != 0 || C1 % 64 != 0 || Cdy % 64 != 0) return false;
    if (!(P_in == 1 || (P_in == 2 && C1 == 0))) return false;
    if (!(P_dy == 1 || P_dy == 2)) return false;
    if (!wgrad_geometry(H, W, &p.TH, &p.TW)) return false;
    p.N = N; p.H = H; p.W = W;
    p.tiles_x = W / p.TW; p.tiles_y = H / p.TH; p.pix_tiles = N * p.tiles_x * p.tiles_y;
    p.k = k; p.pad = k / 2; p.P_in = P_in;
    if (P_in == 1) { p.nchunk0 = C0 / 64; p.nchunk_c = (C0 + C1) / 64; }
    else { p.nchunk0 = 2 * C0 / 64; p.nchunk_c = p.nchunk0; }
    p.nblocks = k * k * P_in * p.nchunk_c;
    p.CoutL = Cdy * P_dy * P_dy;
    p.NT = pick_nt_w(p.CoutL); if (!p.NT) return false;
    p.n_tiles = p.CoutL / p.NT; p.nb = p.NT / 64;
    p.nch_dy = P_dy * Cdy / 64; p.P_dy = P_dy;
    const int npairs = (p.nblocks + 1) / 2;
    int G = 512 / p.NT; if (G > 4) G = 4; if (G > npairs) G = npairs;
    // fewest groups first (every group re-reads the dY tiles), then the smallest G that still gives that count: a smaller
    // stage means a deeper shared-memory ring (measured on B200: G=3 beats G=4 by 1.4x when both need the same groups)
    while (G > 1 && (npairs + G - 2) / (G - 1) == (npairs + G - 1) / G) --G;
    { const char* e = getenv("HDIFF_WGRAD_G"); if (e) { int g = atoi(e); if (g >= 1 && g <= G) G = g; } }   // tuning knob
    // keep at least two stages in shared memory
    while (G > 1 && 2 * (G * 2 + p.nb) * kBlkBytes > 200 * 1024) --G;
    p.G = G;
    p.ngroups = (npairs + G - 1) / G;
    p.rows_padded = p.ngroups * G * 128;
    int stage_bytes = (G * 2 + p.nb) * kBlkBytes;
    p.halo = 0; p.max_units = 0;
    static const bool halo_off = getenv("HDIFF_WGRAD_HALO_OFF") != nullptr;
    if (!halo_off && k == 3 && p.TH == 1 && p.TW == 64) {
        // halo mode: the stage shrinks (one box per tap row instead of one per tap), so take the largest G (fewest groups)
        WgradTcParams q = p;
        int Gh = 512 / p.NT; if (Gh > 8) Gh = 8; if (Gh > npairs) Gh = npairs;     // up to 8 accumulators of NT columns
        while (Gh > 1 && (npairs + Gh - 2) / (Gh - 1) == (npairs + Gh - 1) / Gh) --Gh;
        q.G = Gh; q.ngroups = (npairs + Gh - 1) / Gh;
        bool ok = true; int max_units = 0;
        for (int g = 0; g < q.ngroups && ok; ++g) {
            GroupLayout L;
            decode_group(q, g, &L);
            if (L.nunits > 8) ok = false;
            if (L.nunits > max_units) max_units = L.nunits;
            for (int i = 0; i < Gh && ok; ++i) {           // the second block of a 128-row pair must not lie below the first
                const int o0 = L.b_unit[2 * i] * kHaloSlot + L.b_tx[2 * i] * 128, o1 = L.b_unit[2 * i + 1] * kHaloSlot + L.b_tx[2 * i + 1] * 128;
                if (o1 < o0) ok = false;
            }
        }
        const int sb = max_units * kHaloSlot + p.nb * kBlkBytes;
        if (ok && (200 * 1024) / sb >= 3 && sb < stage_bytes) {
            p.G = Gh; p.ngroups = q.ngroups; p.rows_padded = p.ngroups * Gh * 128;
            p.halo = 1; p.max_units = max_units; stage_bytes = sb;
        }
    }
    p.stages = (200 * 1024) / stage_bytes; if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) return false;
    // split-K over pixel tiles: one CTA per SM (a CTA owns all 512 TMEM columns, so more CTAs would only queue up), i.e. one
    // wave of equal-length CTAs and half the fp32 partials of a two-wave split
    const int ctas_per_split = p.ngroups * p.n_tiles;
    static const int waves = getenv("HDIFF_WGRAD_WAVES") ? atoi(getenv("HDIFF_WGRAD_WAVES")) : 1;
    int splits = (waves * hd_num_sms()) / ctas_per_split;          // floor: never spill a few CTAs into another wave
    if (splits > p.pix_tiles) splits = p.pix_tiles;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.pix_tiles + splits - 1) / splits;
    p.splits = (p.pix_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    *out = p;
    return true;
}

}  // namespace

extern "C" int hd_wgrad_tc_supported(int C0, int C1, int P_in, int Cdy, int P_dy, int H, int W, int k) {
    WgradTcParams p;
    return make_plan(C0, C1, P_in, Cdy, P_dy, 1, H, W, k, &p) ? 1 : 0;
}

extern "C" long long hd_wgrad_tc_workspace(int C0, int C1, int P_in, int Cdy, int P_dy, int N, int H, int W, int k) {
    WgradTcParams p;
    if (!make_plan(C0, C1, P_in, Cdy, P_dy, N, H, W, k, &p)) return 0;
    return (long long)p.splits * p.rows_padded * p.CoutL * 4;
}

extern "C" int hd_wgrad_tc(const void* in0, int C0, const void* in1, int C1, int P_in, const void* dy, int Cdy, int P_dy,
                           float* dw, void* workspace, long long workspace_bytes, int N, int H, int W, int k, cudaStream_t stream) {
    HD_REQUIRE(in0 && dy && dw && workspace && N > 0);
    WgradTcParams p;
    if (!make_plan(C0, C1, P_in, Cdy, P_dy, N, H, W, k, &p)) { hd_set_error("hd_wgrad_tc: unsupported shape"); return HD_ERR_UNSUPPORTED; }
    HD_REQUIRE(workspace_bytes >= (long long)p.splits * p.rows_padded * p.CoutL * 4);
    p.ws = (float*)workspace;
    CUtensorMap mX0, mX1, mDY;
    const int box_w = p.halo ? p.TW + 2 : p.TW;
    int rc = hd_make_act_tmap(&mX0, in0, C0, P_in, N, H, W, 64, box_w, p.TH); if (rc) return rc;
    if (C1 > 0) { rc = hd_make_act_tmap(&mX1, in1, C1, 1, N, H, W, 64, box_w, p.TH); if (rc) return rc; }
    else mX1 = mX0;
    rc = hd_make_act_tmap(&mDY, dy, Cdy, P_dy, N, H, W, 64, p.TW, p.TH); if (rc) return rc;
    const int stage_bytes = (p.halo ? p.max_units * kHaloSlot : p.G * 2 * kBlkBytes) + p.nb * kBlkBytes;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (2 * kMaxStages + 2) * 8 + 16;
    static unsigned long long attr_set = 0;
    if (!hd_seen_on_device(&attr_set)) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { hd_set_error("cudaFuncSetAttribute(wgrad_tc_kernel)"); return HD_ERR_CUDA; }
        hd_mark_on_device(&attr_set);
    }
    wgrad_tc_kernel<<<dim3(p.ngroups, p.splits, p.n_tiles), kThreads, smem, stream>>>(mX0, mX1, mDY, p);
    HD_CHECK_LAUNCH();
    const int rows = p.nblocks * 64;
    const int CinL = (C0 + C1) * P_in * P_in;
    const int64_t total = (int64_t)rows * p.CoutL;
    int grid = (int)((total + 255) / 256); if (grid > 148 * 8) grid = 148 * 8;
    wgrad_reduce_kernel<<<grid, 256, 0, stream>>>(p.ws, p.splits, p.rows_padded, rows, p.CoutL, k * k, CinL, dw);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
