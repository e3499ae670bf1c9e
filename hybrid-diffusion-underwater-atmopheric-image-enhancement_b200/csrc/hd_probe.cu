// Hardware probe (test infrastructure for the next kernel generation, not on the product path):
// can a tcgen05.mma shared-memory A operand start at an arbitrary 128-byte ROW of a 128B-swizzled TMA box?
// If yes, one halo tile in shared memory serves all 3 horizontal taps of a 3x3 convolution (and the 9 shifted inputs of
// its weight gradient) instead of one TMA copy per tap.
//   D[128][64] = X[shift .. shift+127][0..63] * W[64][64]^T     (bf16 in, fp32 out)
// mode 0: descriptor start address = base + shift*128, base_offset field 0
// mode 1: same start address, base_offset field = shift & 7
#include "hd_tc_common.cuh"

namespace {
__global__ void __launch_bounds__(128, 1)
probe_shift_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, float* out, int shift, int mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX = smem;                 // [144 rows][128 B]
    uint8_t* sW = smem + 144 * 128;     // [64 rows][128 B]   (18432 is a multiple of 1024)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sW + 64 * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 144 * 128 + 64 * 128);
        tma_load_2d(sX, &mapX, bar, 0, 0);
        tma_load_2d(sW, &mapW, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        const uint32_t a0 = smem_u32(sX) + shift * 128;
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = umma_smem_desc(a0 + k * 32, 16, 1024);
            if (mode == 1) ad |= (uint64_t)(shift & 7) << 49;
            umma_bf16(tmem, ad, umma_smem_desc(smem_u32(sW) + k * 32, 16, 1024), idesc, k != 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) out[row * 64 + c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}
}  // namespace

// x: [144][64] bf16, w: [64][64] bf16 (row = output channel), out: [128][64] fp32
extern "C" int hd_probe_shift(const void* x, const void* w, float* out, int shift, int mode, cudaStream_t stream) {
    HD_REQUIRE(x && w && out && shift >= 0 && shift <= 16);
    CUtensorMap mX, mW;
    uint64_t dx[2] = {64, 144}, sx[1] = {64}; uint32_t bx[2] = {64, 144};
    int rc = hd_make_tmap_bf16(&mX, x, 2, dx, sx, bx); if (rc) return rc;
    uint64_t dw[2] = {64, 64}, sw[1] = {64}; uint32_t bw[2] = {64, 64};
    rc = hd_make_tmap_bf16(&mW, w, 2, dw, sw, bw); if (rc) return rc;
    const size_t smem = 144 * 128 + 64 * 128 + 1024 + 64;
    probe_shift_kernel<<<1, 128, smem, stream>>>(mX, mW, out, shift, mode);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
