/* hdiff_b200 — C ABI of the B200-native DDPM / classifier-free-guidance hot path.
 *
 * The reference (gusanagy/Hybrid-Diffusion-Underwater-Atmopheric-Image-Enhancement) has no FFI: its
 * boundary is the Python class API.  The Python host package mirrors that API and
 * lowers it onto the entry points below, which are what a maintainer binds (ctypes stub in
 * INTEGRATION.md).  Conventions:
 *   - plain pointers and sizes only; device pointers are owned by the caller; kernels never allocate,
 *     free or retain them; scratch is passed in;
 *   - every launch goes on the `stream` argument (a cudaStream_t passed as void*);
 *   - return 0 on success, negative on error (HD_ERR_*); `hd_last_error()` holds the message;
 *   - `dtype`: 0 = fp32 activations ("fp32 check mode"), 1 = bf16 activations (fp32 accumulate).
 *   - activations are NHWC; a view argument P in {1,2} exposes a tensor through its 2x2
 *     space-to-depth rearrangement ([N][2H][2W][C] seen as [N][H][W][4C], channel = (py*2+px)*C + c).
 * Reference line numbers are relative to the reference repository root.
 */
#ifndef HDIFF_B200_H
#define HDIFF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HD_OK 0
#define HD_ERR_ARG (-1)
#define HD_ERR_UNSUPPORTED (-2)
#define HD_ERR_CUDA (-3)
#define HD_ERR_DRIVER (-4)

typedef void* hd_stream_t;

const char* hd_last_error(void);
int hd_abi_version(void);

/* ---- convolutions: nn.Conv2d / nn.ConvTranspose2d call sites DiffusionFreeGuidence/ModelCondition.py:71-72,
 *      82-83,96-99,130,144,147,219,251 and torch.cat :271; backward = autograd of the same (TrainCondition.py:60).
 *      Logical stride-1 'same' convolution, ksize in {1,3}, on views; w = packed [CoutL][ksize^2][CinL]. ---- */
/* out_nchw_f32 = number of leading output channels stored as fp32 NCHW (0: NHWC in the activation dtype) */
int hd_conv_simt(int dtype, const void* in0, int C0, const void* in1, int C1, int P_in, int in_nchw_f32,
                 const void* w, const float* bias, const float* emb, int64_t emb_stride, const void* res,
                 void* out, int Cout, int P_out, int out_nchw_f32, int N, int H, int W, int ksize, hd_stream_t stream);
int hd_wgrad_simt(int dtype, const void* in0, int C0, const void* in1, int C1, int P_in, int in_nchw_f32,
                  const void* dy, int Cdy, int P_dy, int dy_nchw_f32, float* dw, int N, int H, int W, int ksize,
                  hd_stream_t stream);
/* tcgen05 / TMEM / TMA implicit GEMM, bf16 (K1 forward, K2 data gradient through flipped packed weights) */
int hd_conv_tc_supported(int C0, int C1, int P_in, int Cout, int P_out, int H, int W, int ksize);
int hd_conv_tc(const void* in0, int C0, const void* in1, int C1, int P_in, const void* w, const float* bias,
               const float* emb, int64_t emb_stride, const void* res, void* out, int Cout, int P_out,
               int N, int H, int W, int ksize, int out_nchw_c, double* chan_sums, hd_stream_t stream);
/* chan_sums (optional, accumulates, caller zeroes): [N][Cout][2] fp64 per-image per-channel (sum, sum of squares) of the
 * stored output — the next GroupNorm's statistics without another pass over the tensor (P_out == 1 only). */
/* out_nchw_c > 0: `out` is fp32 NCHW [N][out_nchw_c][H][W] and only the first out_nchw_c output channels are stored
 * (the 3-channel tail, ModelCondition.py:251, run as a 64-channel GEMM with zero-padded weights). */
/* NCHW fp32 [N][Cin<=8][HW] -> NHWC bf16 [N][HW][64], channels >= Cin zero (head input / tail output gradient) */
int hd_pad_nchw(const float* in, int Cin, void* out, int N, int64_t HW, hd_stream_t stream);
/* K3 weight gradient: split over pixels, fp32 partials in `workspace`, deterministic second-stage reduce */
int hd_wgrad_tc_supported(int C0, int C1, int P_in, int Cdy, int P_dy, int H, int W, int ksize);
int64_t hd_wgrad_tc_workspace(int C0, int C1, int P_in, int Cdy, int P_dy, int N, int H, int W, int ksize);
int hd_wgrad_tc(const void* in0, int C0, const void* in1, int C1, int P_in, const void* dy, int Cdy, int P_dy,
                float* dw, void* workspace, int64_t workspace_bytes, int N, int H, int W, int ksize, hd_stream_t stream);

/* ---- AttnBlock core: softmax(q k^T C^-1/2) v, single head, head_dim = C (ModelCondition.py:101-120).
 *      qkv = [N][S][3C] (q | k | v), out/dout = [N][S][C], lse/delta = [N][S] fp32. ---- */
int hd_attn_fwd_simt(int dtype, const void* qkv, void* out, float* lse, int N, int S, int C, hd_stream_t stream);
int hd_attn_bwd_simt(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                     void* dqkv, int N, int S, int C, hd_stream_t stream);
int hd_attn_tc_supported(int S, int C);      /* forward on tcgen05 */
int hd_attn_bwd_tc_supported(int S, int C);  /* backward on tcgen05 */
int hd_attn_fwd_tc(const void* qkv, void* out, float* lse, int N, int S, int C, hd_stream_t stream);
/* tcgen05 flash attention (K5).  Backward = three launches (row statistics, dK/dV pass, dQ pass), deterministic, no
 * atomics; `stats` is caller-provided scratch of N*S*2 floats ((lse*log2e, rowsum(dout*out)) per row). */
int hd_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                   int N, int S, int C, hd_stream_t stream);
/* the 128-channel kernels with the softmax scale supplied (softmax(q k^T * scale) v); used by the multi-head route */
int hd_attn_fwd_tc_scaled(const void* qkv, void* out, float* lse, int N, int S, float scale, hd_stream_t stream);
int hd_attn_bwd_tc_scaled(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                          int N, int S, float scale, hd_stream_t stream);
/* Wide heads (C a multiple of 128 up to 1024; hd_attn_*_tc forward to these when C != 128): the AttnBlocks of the wider
 * UNet of BASELINE.json configs[4] (C = 256 at 64x64, C = 512 at 32x32).  Same arguments and semantics as above. */
int hd_attn_wide_tc_supported(int S, int C);
int hd_attn_fwd_wide_tc(const void* qkv, void* out, float* lse, int N, int S, int C, hd_stream_t stream);
int hd_attn_bwd_wide_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* stats, void* dqkv,
                        int N, int S, int C, hd_stream_t stream);

/* ---- Multi-head self-attention core of the MHA ResBlock: nn.MultiheadAttention(out_ch, 8) on the flattened feature map with
 *      q = k = v (ModelCondition.py:189,203-208 == diffusion/Model.py:290,304-309; DynamicUNet middle blocks :425-431).
 *      qkv [N][S][3C] is the packed in-projection's output (q | k | v, head h = channels [h*C/heads, (h+1)*C/heads) of each);
 *      out / dout [N][S][C]; lse / delta [N][heads][S] fp32 (delta is scratch).  Head dims 4, 8, 16, 32, 64. ---- */
int hd_mha_supported(int C, int heads);
/* tensor-core route for head dims 8..64 (bf16): heads zero-padded to the 128-channel layout of hd_attn_*_tc, N * heads sequences.
 * pack: src [N][S][parts*C] -> dst [N*heads][S][parts*128]; unpack: the inverse; parts = 3 (q|k|v) or 1; scale0 multiplies part 0 */
int hd_mha_pack_heads(const void* src, void* dst, int N, int S, int C, int heads, int parts, float scale0, hd_stream_t stream);
int hd_mha_unpack_heads(const void* src, void* dst, int N, int S, int C, int heads, int parts, float scale0, hd_stream_t stream);
int hd_mha_fwd(int dtype, const void* qkv, void* out, float* lse, int N, int S, int C, int heads, hd_stream_t stream);
int hd_mha_bwd(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
               int N, int S, int C, int heads, hd_stream_t stream);

/* ---- hybrid (image-conditioned) pipeline helpers.  Nearest-neighbour up-sampling of a skip tensor by an integer factors (fy, fx) and its
 *      backward (DynamicUNet's skip fix-up, diffusion/Model.py:506-510), NHWC [N][H][W][C] -> [N][fy H][fx W][C]; image scaling
 *      out = in * scale + shift from uint8 or fp32 (hybrid trainer / sampler, diffusion/Diffusion.py:56-57,221). ---- */
int hd_upsample_nearest(int dtype, const void* in, void* out, int N, int H, int W, int C, int fy, int fx, hd_stream_t stream);
int hd_upsample_nearest_bwd(int dtype, const void* dout, void* din, int N, int H, int W, int C, int fy, int fx, hd_stream_t stream);
int hd_image_affine(const void* in, int src_u8, float* out, float scale, float shift, int64_t n, hd_stream_t stream);

/* ---- input pipeline and evaluation metrics of the product pipeline, 8-bit images.
 *      hd_resize_bilinear_u8: cv2.resize(img, (DW, DH), INTER_LINEAR) bit-exact, src [N][SH][SW][C] -> dst [N][DH][DW][C] or, chw != 0,
 *      [N][C][DH][DW] (albumentations A.Resize(256, 256) + ToTensorV2, utils/utils.py:318-323).
 *      hd_sq_err_u8: sq_err[n] = sum (a - b)^2 per image (PSNR, utils/rotinas.py:21).
 *      hd_uiqm_u8: out[n] = (UIQM, UICM, UISM, UIConM) of metrics/metrics.py:77-299 for [N][H][W][3] RGB images. ---- */
int hd_resize_bilinear_u8(const void* src, int N, int SH, int SW, int C, void* dst, int DH, int DW, int chw, hd_stream_t stream);
int hd_sq_err_u8(const void* a, const void* b, int N, int64_t per_image, double* sq_err, hd_stream_t stream);
int64_t hd_uiqm_workspace(int N);
int hd_uiqm_u8(const void* img, int N, int H, int W, void* workspace, int64_t ws_bytes, float* out, hd_stream_t stream);
/*      hd_rgb2lab_u8: cv2.cvtColor(img, cv2.COLOR_RGB2LAB) on 8-bit pixels, bit-exact, [npix][3] -> [npix][3] (metrics/metrics.py:43).
 *      hd_lab_tables_host: the conversion's two look-up tables written to HOST memory (gamma[256], cbrt[3072]); needs no GPU.
 *      hd_uciqe_u8: out[n] = (UCIQE, var_chr, con_lum, aver_sat) in float64 of metrics/metrics.py:40-76 (uciqe(nargin=1, loc=img)).
 *      The first hd_rgb2lab_u8 / hd_uciqe_u8 call on a device uploads the tables and synchronises `stream` once (not capturable in a CUDA graph). */
/*      hd_ssim_u8: out[n] = skimage.metrics.structural_similarity(a[n], b[n], win_size=win, channel_axis=2, data_range=255) in float64
 *      (utils/rotinas.py:22,926; metrics/metrics.py:642 with win = 3). */
int hd_ssim_u8(const void* a, const void* b, int N, int H, int W, int C, int win, double* out, hd_stream_t stream);
int hd_rgb2lab_u8(const void* rgb, int64_t npix, void* lab, hd_stream_t stream);
int hd_lab_tables_host(uint16_t* gamma, uint16_t* cbrt_tab);
int64_t hd_uciqe_workspace(int N);
int hd_uciqe_u8(const void* img, int N, int H, int W, void* workspace, int64_t ws_bytes, double* out, hd_stream_t stream);

/* ---- GroupNorm(32) + Swish (+ Dropout): ModelCondition.py:128-129,141-143,95,249-250.  Two-source input
 *      (C0 | C1 channels) fuses torch.cat :271 into the normalisation.  sums/gsums = [N][G][2] fp64. ---- */
int hd_gn_stats(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, double* sums,
                hd_stream_t stream);
/* 1 when the hd_gn_* calls of this shape run on the second-generation bf16 kernels (csrc/hd_gn.cu): those never hand dy' from the
 * reduce pass to the apply pass (pass dy_act = NULL, dy_is_act = 0) */
int hd_gn_v2(int dtype, int C0, int C1, int G, int64_t HW, int N);
/* group statistics [N][G][2] from the per-channel sums of one or two source tensors (see hd_conv_tc chan_sums) */
int hd_gn_group_sums(const double* cs0, int C0, const double* cs1, int C1, int N, int G, double* sums, hd_stream_t stream);
int hd_gn_apply(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G, const double* sums,
                const float* gamma, const float* beta, float eps, int act, float p_drop, uint64_t seed, void* out,
                hd_stream_t stream);
int hd_gn_bwd_reduce(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                     const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                     uint64_t seed, const void* dy, double* gsums, float* dgamma, float* dbeta, void* dy_act,
                     hd_stream_t stream);
/* dy_act (optional, may alias dy): receives dy' = dy * dropout mask * act'(z); hd_gn_bwd_apply(dy = dy_act, dy_is_act = 1) then
 * skips the sigmoid and the dropout hash */
int hd_gn_bwd_apply(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                    const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                    uint64_t seed, const void* dy, const double* gsums, const void* add, const void* acc0,
                    const void* acc1, void* dx0, void* dx1, float* cs_total, float* cs_per_n, int64_t cs_ld,
                    int cs_n, int dy_is_act, hd_stream_t stream);
/* for c < cs_n: cs_total[c] += sum_{n,pix} dx, cs_per_n[n*cs_ld + c] += sum_pix dx (either may be NULL): the bias and embedding-add
 * gradients of the convolution that produced the normalised tensor, without another pass over dx */
/* both passes in one cooperative launch, the batch walked in L2-sized image groups (3 HBM tensor passes instead of 5);
 * gsums = [N][G][2] fp64 scratch, counter = 4 bytes of scratch for the grid barrier */
int hd_gn_bwd_fused(int dtype, const void* in0, int C0, const void* in1, int C1, int N, int64_t HW, int G,
                    const double* sums, const float* gamma, const float* beta, float eps, int act, float p_drop,
                    uint64_t seed, const void* dy, double* gsums, float* dgamma, float* dbeta, const void* add,
                    const void* acc0, const void* acc1, void* dx0, void* dx1, unsigned* counter, hd_stream_t stream);
/* bias / embedding-add gradients: per-sample and total column sums (both accumulate) */
int hd_colsum(int dtype, const void* t, int nchw_f32, int N, int64_t HW, int C, float* per_n, int64_t ld_per_n,
              float* total, hd_stream_t stream);

/* ---- embedding path, fp32: TimeEmbedding / ConditionalEmbedding / temb_proj / cond_proj
 *      (ModelCondition.py:27-65,132-139,158-159) ---- */
int hd_linear_fwd(const float* x, int M, int K, int64_t ldx, const float* w, const float* b, float* y, int Nout,
                  int64_t ldy, int in_swish, int accumulate, hd_stream_t stream);
int hd_linear_bwd_x(const float* dy, int M, int Nout, int64_t lddy, const float* w, int K, const float* x_pre,
                    int64_t ldx, float* dx, int64_t lddx, int accumulate, hd_stream_t stream);
int hd_linear_bwd_w(const float* dy, int M, int Nout, int64_t lddy, const float* x, int K, int64_t ldx, int in_swish,
                    float* dw, float* db, hd_stream_t stream);
int hd_embedding_fwd(const float* table, int rows, int dim, const int64_t* idx, int M, float* out, hd_stream_t stream);
int hd_embedding_bwd(const float* dout, int dim, const int64_t* idx, int M, float* dtable, int64_t padding_idx,
                     hd_stream_t stream);

/* ---- parameter layouts: OIHW / IOHW fp32 masters <-> packed GEMM operands ---- */
int hd_gather_pack(int out_dtype, const float* src, const int32_t* ia, const int32_t* ib, int64_t n, void* out,
                   hd_stream_t stream);
int hd_scatter_unpack(const float* packed, const int32_t* inv, int64_t n, float* dst, hd_stream_t stream);

/* ---- diffusion process: extract / q_sample / mse_loss (DiffusionCondition.py:9-16,41-45) and one sampler step
 *      p_mean_variance + CFG mix + noise + NaN check + final clip (DiffusionCondition.py:68-98) ---- */
int hd_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ab, const float* sqrt_1m_ab,
                float* xt, int N, int64_t chw, int T, hd_stream_t stream);   /* t outside [0, T) aborts the launch */
int hd_mse_fwd(const float* pred, const float* noise, float* loss, int64_t n, hd_stream_t stream);
int hd_mse_bwd(const float* pred, const float* noise, const float* g, float* dpred, int64_t n, hd_stream_t stream);
int hd_sampler_step(float* x, const float* eps_c, const float* eps_u, const float* z, float w1, float w, const float* coef,
                    const int* step_ptr, int last_step_clip, int* nan_flag, int64_t n, hd_stream_t stream);
int hd_add_int(int* p, int delta, hd_stream_t stream);

/* 1 if hd_conv_tc with `chan_sums` would take the staged epilogue for this shape (statistics read back out of the staged
 * tile: cheap); the host requests conv-epilogue statistics only then */
int hd_conv_tc_stats_staged(int C0, int C1, int P_in, int Cout, int P_out, int H, int W, int ksize);
/* ---- clip_grad_norm_ + AdamW on the flat buffers (TrainCondition.py:39,61-63) ---- */
int hd_sqnorm(const float* g, int64_t n, double* out, int accumulate, hd_stream_t stream);   /* accumulate: out += instead of out = */
int hd_adamw_flat(float* p, float* g, float* m, float* v, int64_t n, const double* sqnorm, float max_norm, double lr,
                  double b1, double b2, double eps, double wd, int step, hd_stream_t stream);   /* bias corrections formed in double */

#ifdef __cplusplus
}
#endif
#endif
