// Input pipeline and evaluation metrics of the product pipeline on the GPU.  Integer / byte work, HBM-bound:
//   * hd_resize_bilinear_u8: cv2.resize(img, (DW, DH), interpolation=cv2.INTER_LINEAR) for 8-bit HWC images, BIT-EXACT (OpenCV's
//     fixed-point path: 11-bit coefficients, two-stage rounding), optionally written CHW — albumentations' A.Resize(256, 256) +
//     ToTensorV2 of the reference datasets (utils/utils.py:318-323,441-462);
//   * hd_sq_err_u8: per-image sum of squared differences (PSNR = 10 log10(255^2 / mse), skimage's peak_signal_noise_ratio as
//     imported at utils/rotinas.py:21);
//   * hd_uiqm_u8: UIQM = c1 UICM + c2 UISM + c3 UIConM of metrics/metrics.py:77-299 (getUIQM) for 8-bit RGB images: per-pixel
//     Python loops and sorts in the reference.  For 8-bit input R - G is an integer and (R + G) / 2 - B a half-integer, so the
//     alpha-trimmed means come EXACTLY out of two histograms (no sort); Sobel / EME / UIConM are 8 x 8 block reductions;
//   * hd_rgb2lab_u8: cv2.cvtColor(img, cv2.COLOR_RGB2LAB) for 8-bit images, BIT-EXACT (OpenCV's integer look-up-table path);
//   * hd_ssim_u8: skimage's structural_similarity as called at utils/rotinas.py:926 (uniform window, exact integer window sums);
//   * hd_uciqe_u8: UCIQE of metrics/metrics.py:40-76 (chroma deviation, luminance contrast out of a 65536-bin histogram, saturation).
#include "hd_common.cuh"
#include <math.h>

namespace {

// ---- cv2 INTER_LINEAR, 8-bit (resizeGeneric_<HResizeLinear<uchar,int,short,2048>, VResizeLinear<uchar,int,short,FixedPtCast<22>>>) ----
struct Coef { int s0, s1, a0, a1; };
// dst index d of n_dst over n_src source samples: cv2's (float) coordinate, floor, fraction, 11-bit rounded weights
__device__ __forceinline__ void lin_coef(int d, int n_dst, int n_src, float* frac, int* s) {
    const double scale = (double)n_src / (double)n_dst;
    float f = (float)((d + 0.5) * scale - 0.5);
    const int fl = (int)floorf(f);
    *frac = f - (float)fl;
    *s = fl;
}
__device__ __forceinline__ Coef coef_x(int d, int n_dst, int n_src) {       // horizontal: the weight is zeroed at the borders
    float f; int s; lin_coef(d, n_dst, n_src, &f, &s);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    Coef c; c.s0 = s; c.s1 = min(s + 1, n_src - 1);
    c.a0 = __float2int_rn((1.f - f) * 2048.f); c.a1 = __float2int_rn(f * 2048.f);
    return c;
}
__device__ __forceinline__ Coef coef_y(int d, int n_dst, int n_src) {       // vertical: rows are clamped, the weights are not touched
    float f; int s; lin_coef(d, n_dst, n_src, &f, &s);
    Coef c; c.s0 = min(max(s, 0), n_src - 1); c.s1 = min(max(s + 1, 0), n_src - 1);
    c.a0 = __float2int_rn((1.f - f) * 2048.f); c.a1 = __float2int_rn(f * 2048.f);
    return c;
}
__global__ void resize_u8_kernel(const uint8_t* __restrict__ src, int SH, int SW, int C, uint8_t* __restrict__ dst, int DH, int DW, int chw, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int dx = (int)(i % DW); int64_t r = i / DW;
        const int dy = (int)(r % DH); const int64_t n = r / DH;
        const Coef cx = coef_x(dx, DW, SW), cy = coef_y(dy, DH, SH);
        const uint8_t* img = src + n * (int64_t)SH * SW * C;
        const uint8_t* r0 = img + (int64_t)cy.s0 * SW * C;
        const uint8_t* r1 = img + (int64_t)cy.s1 * SW * C;
        for (int c = 0; c < C; ++c) {
            const int h0 = r0[cx.s0 * C + c] * cx.a0 + r0[cx.s1 * C + c] * cx.a1;      // horizontal pass, scaled by 2048
            const int h1 = r1[cx.s0 * C + c] * cx.a0 + r1[cx.s1 * C + c] * cx.a1;
            int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            if (chw) dst[((n * C + c) * DH + dy) * (int64_t)DW + dx] = (uint8_t)v;
            else dst[((n * DH + dy) * (int64_t)DW + dx) * C + c] = (uint8_t)v;
        }
    }
}

__global__ void sq_err_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t per_image, double* out) {
    __shared__ unsigned long long red[8];
    const int n = blockIdx.y;
    const uint8_t* pa = a + n * per_image; const uint8_t* pb = b + n * per_image;
    unsigned long long s = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_image; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)pa[i] - (int)pb[i];
        s += (unsigned long long)(d * d);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; atomicAdd(out + n, (double)t); }
}

// ---- UIQM ----
constexpr int kRG = 511, kYB = 1021, kHist = kRG + kYB;      // R - G in [-255, 255]; R + G - 2 B in [-510, 510]
struct UiqmWs {                      // per image
    unsigned int hist[kHist];
    unsigned int sobel_max[3];       // float bits (non-negative floats order like unsigned integers)
    unsigned int pad;
    double eme[3], uiconm;
};

__global__ void uiqm_hist_kernel(const uint8_t* __restrict__ img, int64_t npix, UiqmWs* ws) {
    __shared__ unsigned int h[kHist];
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < kHist; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint8_t* p = img + n * npix * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = p[3 * i], g = p[3 * i + 1], b = p[3 * i + 2];
        atomicAdd(&h[r - g + 255], 1u);
        atomicAdd(&h[kRG + r + g - 2 * b + 510], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHist; i += blockDim.x) if (h[i]) atomicAdd(&ws[n].hist[i], h[i]);
}

// scipy.ndimage.sobel(x, axis) with mode='reflect' (a 1-pixel halo: the edge pixel repeats): derivative (-1, 0, 1) along `axis`,
// smoothing (1, 2, 1) along the other one.  mag = hypot(sobel(x, 0), sobel(x, 1))   (metrics.py:120-125)
__device__ __forceinline__ float sobel_mag(const uint8_t* img, int H, int W, int y, int x, int c) {
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    auto at = [&](int yy, int xx) { return (float)img[((int64_t)yy * W + xx) * 3 + c]; };
    const float d0 = (at(yp, xm) + 2.f * at(yp, x) + at(yp, xp)) - (at(ym, xm) + 2.f * at(ym, x) + at(ym, xp));      // along rows
    const float d1 = (at(ym, xp) + 2.f * at(y, xp) + at(yp, xp)) - (at(ym, xm) + 2.f * at(y, xm) + at(yp, xm));      // along columns
    return sqrtf(d0 * d0 + d1 * d1);
}
__global__ void uiqm_sobel_max_kernel(const uint8_t* __restrict__ img, int H, int W, UiqmWs* ws) {
    __shared__ unsigned int smax[3];
    const int n = blockIdx.y;
    if (threadIdx.x < 3) smax[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* p = img + (int64_t)n * H * W * 3;
    float m[3] = {0.f, 0.f, 0.f};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)H * W; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i % W);
#pragma unroll
        for (int c = 0; c < 3; ++c) m[c] = fmaxf(m[c], sobel_mag(p, H, W, y, x, c));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicMax(&smax[c], __float_as_uint(m[c]));
    __syncthreads();
    if (threadIdx.x < 3) atomicMax(&ws[n].sobel_max[threadIdx.x], smax[threadIdx.x]);
}

// one thread per 8 x 8 block: EME of the three edge maps (metrics.py:128-193) and the UIConM term (:234-279)
__global__ void uiqm_block_kernel(const uint8_t* __restrict__ img, int H, int W, UiqmWs* ws) {
    const int n = blockIdx.y;
    const int k1 = W / 8, k2 = H / 8;
    const uint8_t* p = img + (int64_t)n * H * W * 3;
    float scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) scale[c] = 255.0f / __uint_as_float(ws[n].sobel_max[c]);       // mag *= 255.0 / np.max(mag), float32
    double e[3] = {0.0, 0.0, 0.0}, u = 0.0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < k1 * k2; b += gridDim.x * blockDim.x) {
        const int bx = b % k1, by = b / k1;
        float emax[3] = {-INFINITY, -INFINITY, -INFINITY}, emin[3] = {INFINITY, INFINITY, INFINITY};
        float vmax = -INFINITY, vmin = INFINITY;
        for (int yy = 0; yy < 8; ++yy)
            for (int xx = 0; xx < 8; ++xx) {
                const int y = by * 8 + yy, x = bx * 8 + xx;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float v = (float)p[((int64_t)y * W + x) * 3 + c];
                    const float ev = (sobel_mag(p, H, W, y, x, c) * scale[c]) * v;
                    emax[c] = fmaxf(emax[c], ev); emin[c] = fminf(emin[c], ev);
                    vmax = fmaxf(vmax, v); vmin = fminf(vmin, v);
                }
            }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (emin[c] != 0.f && emax[c] != 0.f) e[c] += log((double)(emax[c] / emin[c]));
        const float top = vmax - vmin, bot = vmax + vmin;
        if (bot != 0.f && top != 0.f) { const double r = (double)(top / bot); u += r * log(r); }
    }
    // block-level reduction, then one atomic per quantity and CTA
    __shared__ double red[4][8];
    double q[4] = {e[0], e[1], e[2], u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = q[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        if (threadIdx.x < 3) atomicAdd(&ws[n].eme[threadIdx.x], t); else atomicAdd(&ws[n].uiconm, t);
    }
}

// alpha-trimmed mean and deviation (metrics.py:77-102) of a value set given as a histogram: value of bin i = (i - zero) * step
__device__ void trimmed(const unsigned int* h, int bins, int zero, double step, long long K, double* mu_out, double* s_out) {
    const long long TL = (long long)ceil(0.1 * (double)K), TR = (long long)floor(0.1 * (double)K);
    const long long s = TL + 1, e = K - TR;                   // sum(sorted[s:e])
    double sum = 0.0; long long pos = 0;
    for (int i = 0; i < bins; ++i) {
        const long long c = h[i];
        const long long lo = pos > s ? pos : s, hi = (pos + c) < e ? (pos + c) : e;
        if (hi > lo) sum += (double)(hi - lo) * ((double)(i - zero) * step);
        pos += c;
    }
    const double mu = sum / (double)(K - TL - TR);
    double var = 0.0;
    for (int i = 0; i < bins; ++i) { const double d = (double)(i - zero) * step - mu; var += (double)h[i] * d * d; }
    *mu_out = mu; *s_out = var / (double)K;
}
__global__ void uiqm_final_kernel(const UiqmWs* ws, int N, int H, int W, float* out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long K = (long long)H * W;
    double mu_rg, s_rg, mu_yb, s_yb;
    trimmed(ws[n].hist, kRG, 255, 1.0, K, &mu_rg, &s_rg);
    trimmed(ws[n].hist + kRG, kYB, 510, 0.5, K, &mu_yb, &s_yb);
    const double uicm = -0.0268 * sqrt(mu_rg * mu_rg + mu_yb * mu_yb) + 0.1586 * sqrt(s_rg + s_yb);
    const int k1 = W / 8, k2 = H / 8;
    const double w = 2.0 / ((double)k1 * k2);
    const double uism = 0.299 * (w * ws[n].eme[0]) + 0.587 * (w * ws[n].eme[1]) + 0.144 * (w * ws[n].eme[2]);      // 0.144: as in the reference (:190)
    const double uiconm = (-1.0 / ((double)k1 * k2)) * ws[n].uiconm;
    out[4 * n] = (float)(0.0282 * uicm + 0.2953 * uism + 3.5753 * uiconm);
    out[4 * n + 1] = (float)uicm; out[4 * n + 2] = (float)uism; out[4 * n + 3] = (float)uiconm;
}

// ---- OpenCV's 8-bit RGB -> CIE Lab (cv2.cvtColor(img, cv2.COLOR_RGB2LAB), struct RGB2Lab_b in imgproc/src/color_lab.cpp) and UCIQE ----
// Integer arithmetic on two look-up tables: the sRGB gamma curve scaled to 255 * 8 and the Lab f() curve scaled to 2^15 over 3072 arguments.
// The tables are the published formulae evaluated in double; OpenCV initialises them in SINGLE precision, which lands on the other side of a
// rounding boundary at exactly two arguments that any 8-bit colour can reach (49 and 628, both within 1.4e-4 of a half) — set below.  Pinned
// over all 2^24 colours against cv2 itself by tests/test_metrics.py (oracle restatement on CPU, this kernel on the GPU).
constexpr int kGammaTab = 256, kCbrtTab = 3072, kLabTab = kGammaTab + kCbrtTab;
__device__ uint16_t g_lab_tab[kLabTab];

struct LabTables {
    uint16_t v[kLabTab];
    LabTables() {
        for (int i = 0; i < kGammaTab; ++i) {
            const double x = i / 255.0;
            const double g = x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4);
            v[i] = (uint16_t)llrint(255.0 * 8.0 * g);
        }
        for (int i = 0; i < kCbrtTab; ++i) {
            const double x = i / (255.0 * 8.0);
            const double f = x < 216.0 / 24389.0 ? x * (841.0 / 108.0) + 16.0 / 116.0 : cbrt(x);
            v[kGammaTab + i] = (uint16_t)llrint(32768.0 * f);
        }
        v[kGammaTab + 49] -= 1;
        v[kGammaTab + 628] += 1;
    }
};
static const uint16_t* lab_tables_host() {
    static const LabTables t;                     // function-local static: initialised once, thread-safe
    return t.v;
}
static int lab_tables_upload(cudaStream_t stream) {
    static unsigned long long seen = 0;
    if (!hd_seen_on_device(&seen)) {
        if (cudaMemcpyToSymbolAsync(g_lab_tab, lab_tables_host(), sizeof(uint16_t) * kLabTab, 0, cudaMemcpyHostToDevice, stream) != cudaSuccess) return HD_ERR_CUDA;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return HD_ERR_CUDA;      // once per device: later launches on ANY stream see the tables
        hd_mark_on_device(&seen);
    }
    return HD_OK;
}
__device__ __forceinline__ void lab_tables_to_smem(uint16_t* tab) {
    for (int i = threadIdx.x; i < kLabTab; i += blockDim.x) tab[i] = g_lab_tab[i];
    __syncthreads();
}
// coefficients: round(4096 * sRGB->XYZ(D65) / white point), rows sum to 4096
__device__ __forceinline__ void rgb2lab_px(const uint16_t* tab, int r, int g, int b, int* L, int* A, int* B) {
    const uint16_t* cb = tab + kGammaTab;
    const int R = tab[r], G = tab[g], Bc = tab[b];
    const int fX = cb[(R * 1777 + G * 1541 + Bc * 778 + 2048) >> 12];
    const int fY = cb[(R * 871 + G * 2929 + Bc * 296 + 2048) >> 12];
    const int fZ = cb[(R * 73 + G * 448 + Bc * 3575 + 2048) >> 12];
    const int l = (296 * fY - 1336935 + 16384) >> 15;                       // Lscale = (116*255+50)/100, Lshift = -((16*255*2^15+50)/100)
    const int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    const int bb = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    *L = min(max(l, 0), 255); *A = min(max(a, 0), 255); *B = min(max(bb, 0), 255);
}
__global__ void rgb2lab_u8_kernel(const uint8_t* __restrict__ rgb, int64_t npix, uint8_t* __restrict__ lab) {
    __shared__ uint16_t tab[kLabTab];
    lab_tables_to_smem(tab);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        int L, A, B;
        rgb2lab_px(tab, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], &L, &A, &B);
        lab[3 * i] = (uint8_t)L; lab[3 * i + 1] = (uint8_t)A; lab[3 * i + 2] = (uint8_t)B;
    }
}

// UCIQE (metrics/metrics.py:40-76): c1 * var_chr + c2 * con_lum + c3 * aver_sat on (L, a, b) / 255 in float64.  Every per-pixel quantity is
// formed with the reference's own sequence of IEEE operations (explicit round-to-nearest intrinsics: no fused multiply-add), so only the
// ORDER of the sums differs from numpy's.  The luminance takes at most 256 values, so np.histogram(lum, 65536) + cumsum is reproduced
// exactly from a 256-bin integer histogram by the finalising thread.
struct UciqeWs {                    // per image
    double sum_chr, sum_sat, sum_dev;
    unsigned int pad[2];
    unsigned int hist[256];
};
__device__ __forceinline__ double uciqe_chroma(int A, int B) {
    const double a = __ddiv_rn((double)A, 255.0), b = __ddiv_rn((double)B, 255.0);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
}
__device__ __forceinline__ double block_sum(double v, double* red) {      // result valid in thread 0
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    return t;
}
template <int kPass>
__global__ void uciqe_pass_kernel(const uint8_t* __restrict__ img, int64_t npix, UciqeWs* ws) {
    __shared__ uint16_t tab[kLabTab];
    __shared__ unsigned int h[256];
    __shared__ double red[8];
    const int n = blockIdx.y;
    if (kPass == 1) for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    lab_tables_to_smem(tab);
    const uint8_t* p = img + n * npix * 3;
    const double aver_chr = kPass == 2 ? __ddiv_rn(ws[n].sum_chr, (double)npix) : 0.0;
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        int L, A, B;
        rgb2lab_px(tab, p[3 * i], p[3 * i + 1], p[3 * i + 2], &L, &A, &B);
        const double chr = uciqe_chroma(A, B);
        if (kPass == 1) {
            const double lum = __ddiv_rn((double)L, 255.0);
            s0 += chr;
            s1 += __ddiv_rn(chr, __dsqrt_rn(__dadd_rn(__dmul_rn(chr, chr), __dmul_rn(lum, lum))));      // saturation
            atomicAdd(&h[L], 1u);
        } else {
            const double q = __ddiv_rn(aver_chr, chr);
            s0 += fabs(__dsub_rn(1.0, __dmul_rn(q, q)));
        }
    }
    const double t0 = block_sum(s0, red);
    if (kPass == 1) {
        const double t1 = block_sum(s1, red);
        if (threadIdx.x == 0) { atomicAdd(&ws[n].sum_chr, t0); atomicAdd(&ws[n].sum_sat, t1); }
        __syncthreads();
        for (int i = threadIdx.x; i < 256; i += blockDim.x) if (h[i]) atomicAdd(&ws[n].hist[i], h[i]);
    } else if (threadIdx.x == 0) {
        atomicAdd(&ws[n].sum_dev, t0);
    }
}
__global__ void uciqe_final_kernel(const UciqeWs* ws, int N, int64_t npix, double* out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const unsigned int* h = ws[n].hist;
    const double K = (double)npix;
    const double var_chr = __dsqrt_rn(__ddiv_rn(ws[n].sum_dev, K)), aver_sat = __ddiv_rn(ws[n].sum_sat, K);
    // np.histogram(lum, 65536): uniform bins over [min, max] (an empty range is widened by 0.5 on both sides), edge(i) = i * step + first
    // (np.linspace), index = trunc((v - first) / (last - first) * 65536) corrected against the computed edges
    constexpr int NB = 65536;
    int kmin = 0, kmax = 255;
    while (kmin < 255 && h[kmin] == 0) ++kmin;
    while (kmax > 0 && h[kmax] == 0) --kmax;
    double first = __ddiv_rn((double)kmin, 255.0), last = __ddiv_rn((double)kmax, 255.0);
    if (first == last) { first = __dsub_rn(first, 0.5); last = __dadd_rn(last, 0.5); }
    const double denom = __dsub_rn(last, first), step = __ddiv_rn(denom, (double)NB);
    auto edge = [&](int i) { return i == NB ? last : __dadd_rn(__dmul_rn((double)i, step), first); };
    long long cum = 0; int ilow = -1, ihigh = -1;
    for (int k = kmin; k <= kmax; ++k) {
        if (h[k] == 0) continue;
        const double v = __ddiv_rn((double)k, 255.0);
        int idx = (int)__dmul_rn(__ddiv_rn(__dsub_rn(v, first), denom), (double)NB);
        if (idx == NB) --idx;
        if (v < edge(idx)) --idx;
        if (v >= edge(idx + 1) && idx != NB - 1) ++idx;
        cum += h[k];
        const double cdf = __ddiv_rn((double)cum, K);
        if (ilow < 0 && cdf > 0.0100) ilow = idx;
        if (ihigh < 0 && cdf >= 0.9900) ihigh = idx;
    }
    const double con_lum = __dsub_rn(__ddiv_rn((double)(ihigh - 1), (double)(NB - 1)), __ddiv_rn((double)(ilow - 1), (double)(NB - 1)));
    // coe_metric[0] * var_chr + coe_metric[1] * con_lum + coe_metric[2] * aver_sat, left to right
    out[4 * n] = __dadd_rn(__dadd_rn(__dmul_rn(0.4680, var_chr), __dmul_rn(0.2745, con_lum)), __dmul_rn(0.2576, aver_sat));
    out[4 * n + 1] = var_chr; out[4 * n + 2] = con_lum; out[4 * n + 3] = aver_sat;
}

// ---- SSIM: skimage.metrics.structural_similarity(a, b, channel_axis=2, data_range=255) as called at utils/rotinas.py:926 (uniform 7 x 7
// window, sample covariance, K1 = 0.01, K2 = 0.03, borders of (win - 1) / 2 pixels cropped, mean over pixels and channels).  The crop keeps
// exactly the pixels whose window lies inside the image, so the filter's boundary mode never enters; the five window sums are exact integers
// for 8-bit input. ----
__global__ void ssim_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C, int win, double* out) {
    __shared__ double red[8];
    const int n = blockIdx.y, pad = (win - 1) / 2;
    const int Hi = H - 2 * pad, Wi = W - 2 * pad;
    const int64_t items = (int64_t)Hi * Wi * C;
    const uint8_t* pa = a + (int64_t)n * H * W * C;
    const uint8_t* pb = b + (int64_t)n * H * W * C;
    const double np_ = (double)(win * win), cov_norm = np_ / (np_ - 1.0);
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int64_t r = i / C;
        const int x = (int)(r % Wi), y = (int)(r / Wi);            // top-left corner of the window of interior pixel (y + pad, x + pad)
        int sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int dy = 0; dy < win; ++dy) {
            const int64_t row = ((int64_t)(y + dy) * W + x) * C + c;
            for (int dx = 0; dx < win; ++dx) {
                const int u = pa[row + (int64_t)dx * C], v = pb[row + (int64_t)dx * C];
                sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
            }
        }
        const double ux = sx / np_, uy = sy / np_, uxx = sxx / np_, uyy = syy / np_, uxy = sxy / np_;
        const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
        const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
        acc += (A1 * A2) / (B1 * B2);
    }
    const double t = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + n, t / (double)items);
}

}  // namespace

extern "C" int hd_resize_bilinear_u8(const void* src, int N, int SH, int SW, int C, void* dst, int DH, int DW, int chw, cudaStream_t stream) {
    HD_REQUIRE(src && dst && N > 0 && SH > 0 && SW > 0 && C > 0 && C <= 4 && DH > 0 && DW > 0);
    const int64_t total = (int64_t)N * DH * DW;
    int64_t blocks = (total + 255) / 256; if (blocks > (int64_t)hd_num_sms() * 16) blocks = (int64_t)hd_num_sms() * 16;
    resize_u8_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint8_t*)src, SH, SW, C, (uint8_t*)dst, DH, DW, chw, total);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// sq_err[n] = sum over the image of (a - b)^2 (overwritten)
extern "C" int hd_sq_err_u8(const void* a, const void* b, int N, int64_t per_image, double* sq_err, cudaStream_t stream) {
    HD_REQUIRE(a && b && sq_err && N > 0 && per_image > 0);
    if (cudaMemsetAsync(sq_err, 0, sizeof(double) * N, stream) != cudaSuccess) return HD_ERR_CUDA;
    int64_t bx = (per_image + 256 * 16 - 1) / (256 * 16); if (bx > 64) bx = 64; if (bx < 1) bx = 1;
    sq_err_u8_kernel<<<dim3((unsigned)bx, N), 256, 0, stream>>>((const uint8_t*)a, (const uint8_t*)b, per_image, sq_err);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

extern "C" int64_t hd_uiqm_workspace(int N) { return (int64_t)N * (int64_t)sizeof(UiqmWs); }
// img: [N][H][W][3] uint8 RGB; out: [N][4] = (UIQM, UICM, UISM, UIConM)
extern "C" int hd_uiqm_u8(const void* img, int N, int H, int W, void* workspace, int64_t ws_bytes, float* out, cudaStream_t stream) {
    HD_REQUIRE(img && workspace && out && N > 0 && H >= 8 && W >= 8 && ws_bytes >= hd_uiqm_workspace(N));
    UiqmWs* ws = (UiqmWs*)workspace;
    if (cudaMemsetAsync(ws, 0, (size_t)hd_uiqm_workspace(N), stream) != cudaSuccess) return HD_ERR_CUDA;
    const int64_t npix = (int64_t)H * W;
    int bx = (int)((npix + 256 * 8 - 1) / (256 * 8)); if (bx > 64) bx = 64;
    uiqm_hist_kernel<<<dim3(bx, N), 256, 0, stream>>>((const uint8_t*)img, npix, ws);
    uiqm_sobel_max_kernel<<<dim3(bx, N), 256, 0, stream>>>((const uint8_t*)img, H, W, ws);
    const int nblk = (W / 8) * (H / 8);
    int bb = (nblk + 127) / 128; if (bb > 64) bb = 64;
    uiqm_block_kernel<<<dim3(bb, N), 128, 0, stream>>>((const uint8_t*)img, H, W, ws);
    uiqm_final_kernel<<<(N + 63) / 64, 64, 0, stream>>>(ws, N, H, W, out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// The two look-up tables of the Lab conversion, host memory (no GPU needed): gamma[256], cbrt[3072]
extern "C" int hd_lab_tables_host(uint16_t* gamma, uint16_t* cbrt_tab) {
    HD_REQUIRE(gamma && cbrt_tab);
    const uint16_t* t = lab_tables_host();
    for (int i = 0; i < kGammaTab; ++i) gamma[i] = t[i];
    for (int i = 0; i < kCbrtTab; ++i) cbrt_tab[i] = t[kGammaTab + i];
    return HD_OK;
}
// rgb, lab: [npix][3] uint8; bit-exact with cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
extern "C" int hd_rgb2lab_u8(const void* rgb, int64_t npix, void* lab, cudaStream_t stream) {
    HD_REQUIRE(rgb && lab && npix > 0);
    if (int rc = lab_tables_upload(stream)) return rc;
    int64_t blocks = (npix + 255) / 256; if (blocks > (int64_t)hd_num_sms() * 8) blocks = (int64_t)hd_num_sms() * 8;
    rgb2lab_u8_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint8_t*)rgb, npix, (uint8_t*)lab);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
extern "C" int64_t hd_uciqe_workspace(int N) { return (int64_t)N * (int64_t)sizeof(UciqeWs); }
// img: [N][H][W][3] uint8 RGB; out: [N][4] float64 = (UCIQE, var_chr, con_lum, aver_sat) of metrics/metrics.py:40-76
extern "C" int hd_uciqe_u8(const void* img, int N, int H, int W, void* workspace, int64_t ws_bytes, double* out, cudaStream_t stream) {
    HD_REQUIRE(img && workspace && out && N > 0 && H > 0 && W > 0 && ws_bytes >= hd_uciqe_workspace(N));
    if (int rc = lab_tables_upload(stream)) return rc;
    UciqeWs* ws = (UciqeWs*)workspace;
    if (cudaMemsetAsync(ws, 0, (size_t)hd_uciqe_workspace(N), stream) != cudaSuccess) return HD_ERR_CUDA;
    const int64_t npix = (int64_t)H * W;
    int bx = (int)((npix + 256 * 8 - 1) / (256 * 8)); if (bx > 64) bx = 64;
    uciqe_pass_kernel<1><<<dim3(bx, N), 256, 0, stream>>>((const uint8_t*)img, npix, ws);
    uciqe_pass_kernel<2><<<dim3(bx, N), 256, 0, stream>>>((const uint8_t*)img, npix, ws);
    uciqe_final_kernel<<<(N + 63) / 64, 64, 0, stream>>>(ws, N, npix, out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}

// a, b: [N][H][W][C] uint8; out[n] (float64, overwritten) = structural_similarity(a[n], b[n], win_size=win, channel_axis=2, data_range=255)
extern "C" int hd_ssim_u8(const void* a, const void* b, int N, int H, int W, int C, int win, double* out, cudaStream_t stream) {
    HD_REQUIRE(a && b && out && N > 0 && C > 0 && win >= 3 && (win & 1) && win <= 15 && H >= win && W >= win);
    if (cudaMemsetAsync(out, 0, sizeof(double) * N, stream) != cudaSuccess) return HD_ERR_CUDA;
    const int64_t items = (int64_t)(H - win + 1) * (W - win + 1) * C;
    int64_t bx = (items + 255) / 256; if (bx > 4 * hd_num_sms()) bx = 4 * hd_num_sms();
    ssim_u8_kernel<<<dim3((unsigned)bx, N), 256, 0, stream>>>((const uint8_t*)a, (const uint8_t*)b, H, W, C, win, out);
    HD_CHECK_LAUNCH();
    return HD_OK;
}
