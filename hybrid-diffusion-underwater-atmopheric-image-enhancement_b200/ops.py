"""Operator layer of the hot path: one Python method per C-ABI entry point.

All activations are NHWC tensors of the compute dtype (bf16, or fp32 in check mode); the
3-channel network input / output stay NCHW fp32 as in the reference API.  A *view* argument
`P` (1 or 2) exposes a tensor through its 2x2 space-to-depth rearrangement, which
turns the reference's strided and transposed convolutions into stride-1 convolutions.

`CudaOps` is the only backend of the product.  It raises if the CUDA library is missing.  Tests
install a torch re-statement of these methods (tests/emu_backend.py) to exercise the host logic on
machines without a GPU; nothing in this package refers to it.
"""
from __future__ import annotations

import torch

from . import _lib

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _p(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CudaOps:
    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        self.use_tc = True          # tcgen05 kernels where the shape is covered
        self.launches = 0           # kernels launched through this object (bench.py reports it)
        self.tc_launches = 0
        self.gn_inplace = __import__("os").environ.get("HDIFF_GN_INPLACE", "1") != "0"   # GroupNorm backward: the reduce pass hands dy' to the apply pass in place
        self.gn_fused = __import__("os").environ.get("HDIFF_GN_FUSED", "0") != "0"   # one cooperative launch per GroupNorm backward (measured slower: see DESIGN.md)
        self._gn_counter = None
        self.mha_tc = __import__("os").environ.get("HDIFF_MHA_TC", "1") != "0"     # 8-head attention on the tcgen05 kernels (padded heads)
        self.prof = None            # bench.py: dict family -> [(start_event, end_event, work)], CUDA events on the launch stream

    # ---- per-launch device timing for bench.py's roofline (off unless `prof` is a dict) ----
    def _t0(self):
        if self.prof is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _t1(self, e0, family, work, hbm_bytes=None):
        """`work`: algorithmic FLOPs (tensor families) or compulsory bytes (HBM families) of the launch; `hbm_bytes`: for a tensor-family
        launch, its compulsory HBM bytes as well (every operand once), so that bench.py can say which roofline binds that launch."""
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.prof.setdefault(family, []).append((e0, e1, work) if hbm_bytes is None else (e0, e1, work, hbm_bytes))

    @staticmethod
    def _nbytes(*tensors):
        return float(sum(t.numel() * t.element_size() for t in tensors if t is not None))

    # ---- convolution family -------------------------------------------------------------
    def conv(self, x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, in_nchw=False, out_nchw=False, alg_frac=1.0,
             Cout_pad=None, chan_sums=None, quiet=False):
        """out = conv_k(view(x0|x1, P_in)) + bias + emb[:, :, None, None] + res, logical stride 1, 'same' padding.
        w: packed [CoutL][k*k][CinL] in the compute dtype.  H, W are the LOGICAL spatial dims.
        `alg_frac`: share of the packed taps that are real reference taps (space-to-depth views carry zeros);
        only used for bench.py's algorithmic-FLOP accounting."""
        dt = _DT[w.dtype]
        C0 = x0.shape[1] if in_nchw else x0.shape[-1]
        C1 = 0 if x1 is None else x1.shape[-1]
        Cout = out.shape[1] if out_nchw else out.shape[-1]
        nchw_c = Cout if out_nchw else 0           # channels stored as fp32 NCHW (the 3-channel tail)
        if Cout_pad is not None:
            Cout = Cout_pad                        # GEMM width of a zero-padded weight matrix
        lib = self.lib
        e0 = self._t0()
        flops = 2.0 * N * H * W * (Cout * P_out * P_out) * k * k * ((C0 + C1) * P_in * P_in) * alg_frac
        if (self.use_tc and dt == BF16 and not in_nchw and nchw_c <= 16
                and lib.hd_conv_tc_supported(C0, C1, P_in, Cout, P_out, H, W, k)):
            rc = lib.hd_conv_tc(_p(x0), C0, _p(x1), C1, P_in, _p(w), _p(bias), _p(emb),
                                0 if emb is None else emb.stride(0), _p(res), _p(out), Cout, P_out,
                                N, H, W, k, nchw_c, _p(chan_sums), _stream())
            _lib.check(rc, "hd_conv_tc")
            self.launches += 1
            self.tc_launches += 1
            self._t1(e0, "conv_tc", flops, self._nbytes(x0, x1, w, res, out) if e0 is not None else None)
            return True if chan_sums is not None else None      # True: the per-channel statistics were produced
        if self.use_tc and dt == BF16 and not quiet:
            _warn_cuda_core("conv", f"C={C0}+{C1} P_in={P_in} Cout={Cout} P_out={P_out} H={H} W={W} k={k}")
        rc = lib.hd_conv_simt(dt, _p(x0), C0, _p(x1), C1, P_in, int(in_nchw), _p(w), _p(bias), _p(emb),
                              0 if emb is None else emb.stride(0), _p(res), _p(out), Cout, P_out, nchw_c,
                              N, H, W, k, _stream())
        _lib.check(rc, "hd_conv_simt")
        self.launches += 1
        self._t1(e0, "conv_simt", flops)

    def wgrad(self, x0, x1, P_in, dy, P_dy, dw, N, H, W, k, dtype, in_nchw=False, dy_nchw=False, workspace=None, alg_frac=1.0,
              quiet=False):
        """dw[CoutL][k*k][CinL] (fp32, overwritten) = sum over pixels of dy (x) shifted input."""
        dt = _DT[dtype]
        C0 = x0.shape[1] if in_nchw else x0.shape[-1]
        C1 = 0 if x1 is None else x1.shape[-1]
        Cdy = dy.shape[1] if dy_nchw else dy.shape[-1]
        lib = self.lib
        e0 = self._t0()
        flops = 2.0 * N * H * W * (Cdy * P_dy * P_dy) * k * k * ((C0 + C1) * P_in * P_in) * alg_frac
        if (self.use_tc and dt == BF16 and not in_nchw and not dy_nchw
                and lib.hd_wgrad_tc_supported(C0, C1, P_in, Cdy, P_dy, H, W, k)):
            need = lib.hd_wgrad_tc_workspace(C0, C1, P_in, Cdy, P_dy, N, H, W, k)
            if workspace is None or workspace.numel() * 4 < need:
                workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=dw.device)
            rc = lib.hd_wgrad_tc(_p(x0), C0, _p(x1), C1, P_in, _p(dy), Cdy, P_dy, _p(dw), _p(workspace),
                                 workspace.numel() * 4, N, H, W, k, _stream())
            _lib.check(rc, "hd_wgrad_tc")
            self.launches += 2
            self.tc_launches += 1
            self._t1(e0, "wgrad_tc", flops, self._nbytes(x0, x1, dy, dw) if e0 is not None else None)
            return
        if self.use_tc and dt == BF16 and not quiet:
            _warn_cuda_core("wgrad", f"C={C0}+{C1} P_in={P_in} Cdy={Cdy} P_dy={P_dy} H={H} W={W} k={k}")
        rc = lib.hd_wgrad_simt(dt, _p(x0), C0, _p(x1), C1, P_in, int(in_nchw), _p(dy), Cdy, P_dy, int(dy_nchw),
                               _p(dw), N, H, W, k, _stream())
        _lib.check(rc, "hd_wgrad_simt")
        self.launches += 1
        self._t1(e0, "wgrad_simt", flops)

    def pad_nchw(self, x, out):
        """fp32 NCHW [N][C<=8][H][W] -> bf16 NHWC [N][H][W][64], zero-padded channels (input of the tcgen05 head conv)."""
        N, C = x.shape[0], x.shape[1]
        _lib.check(self.lib.hd_pad_nchw(_p(x), C, _p(out), N, x.shape[2] * x.shape[3], _stream()), "hd_pad_nchw")
        self.launches += 1

    # ---- attention -----------------------------------------------------------------------
    def attn_fwd(self, qkv, out, lse, N, S, C):
        e0 = self._t0()
        tc = self.use_tc and qkv.dtype == torch.bfloat16 and self.lib.hd_attn_tc_supported(S, C)
        if self.use_tc and qkv.dtype == torch.bfloat16 and self.lib.hd_attn_tc_supported(S, C):
            _lib.check(self.lib.hd_attn_fwd_tc(_p(qkv), _p(out), _p(lse), N, S, C, _stream()), "hd_attn_fwd_tc")
            self.tc_launches += 1
        else:
            if self.use_tc and qkv.dtype == torch.bfloat16:
                _warn_cuda_core("attention", f"S={S} C={C}")
            _lib.check(self.lib.hd_attn_fwd_simt(_DT[qkv.dtype], _p(qkv), _p(out), _p(lse), N, S, C, _stream()), "hd_attn_fwd_simt")
        self.launches += 1
        self._t1(e0, "attn_fwd_tc" if tc else "attn_fwd_simt", 4.0 * N * S * S * C)

    def attn_bwd(self, qkv, out, dout, lse, delta, dqkv, N, S, C):
        e0 = self._t0()
        tc = self.use_tc and qkv.dtype == torch.bfloat16 and self.lib.hd_attn_bwd_tc_supported(S, C)
        if tc:
            stats = torch.empty((N, S, 2), dtype=torch.float32, device=qkv.device)   # (lse*log2e, rowsum(dO*O)) scratch
            _lib.check(self.lib.hd_attn_bwd_tc(_p(qkv), _p(out), _p(dout), _p(lse), _p(stats), _p(dqkv), N, S, C, _stream()), "hd_attn_bwd_tc")
            self.tc_launches += 2
        else:
            _lib.check(self.lib.hd_attn_bwd_simt(_DT[qkv.dtype], _p(qkv), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), N, S, C, _stream()), "hd_attn_bwd_simt")
        self.launches += 3
        self._t1(e0, "attn_bwd_tc" if tc else "attn_bwd_simt", 8.0 * N * S * S * C)

    def _mha_tc(self, qkv, S, C, heads):
        """Tensor-core route: head dims 8..64 run zero-padded on the 128-channel tcgen05 attention kernels (see csrc/hd_mha.cu)."""
        hd = C // heads
        return (self.use_tc and qkv.dtype == torch.bfloat16 and hd % 8 == 0 and hd <= 64 and self.mha_tc
                and self.lib.hd_attn_tc_supported(S, 128) and self.lib.hd_attn_bwd_tc_supported(S, 128))

    def mha_fwd(self, qkv, out, lse, N, S, C, heads):
        """Multi-head self-attention core (nn.MultiheadAttention with q = k = v): qkv [N, S, 3C] -> out [N, S, C], lse [N, heads, S]."""
        e0 = self._t0()
        if self._mha_tc(qkv, S, C, heads):
            hd = C // heads
            pad = torch.empty((N * heads, S, 3 * 128), dtype=qkv.dtype, device=qkv.device)
            opad = torch.empty((N * heads, S, 128), dtype=qkv.dtype, device=qkv.device)
            _lib.check(self.lib.hd_mha_pack_heads(_p(qkv), _p(pad), N, S, C, heads, 3, 1.0, _stream()), "hd_mha_pack_heads")
            _lib.check(self.lib.hd_attn_fwd_tc_scaled(_p(pad), _p(opad), _p(lse), N * heads, S, hd ** -0.5, _stream()), "hd_attn_fwd_tc_scaled")
            _lib.check(self.lib.hd_mha_unpack_heads(_p(opad), _p(out), N, S, C, heads, 1, 1.0, _stream()), "hd_mha_unpack_heads")
            self.launches += 3
            self.tc_launches += 1
            self._t1(e0, "mha_fwd_tc", 4.0 * N * S * S * C)
            return
        _lib.check(self.lib.hd_mha_fwd(_DT[qkv.dtype], _p(qkv), _p(out), _p(lse), N, S, C, heads, _stream()), "hd_mha_fwd")
        self.launches += 1
        self._t1(e0, "mha_fwd", 4.0 * N * S * S * C)

    def mha_bwd(self, qkv, out, dout, lse, delta, dqkv, N, S, C, heads):
        e0 = self._t0()
        if self._mha_tc(qkv, S, C, heads):
            hd = C // heads
            dev, dt = qkv.device, qkv.dtype
            pad = torch.empty((N * heads, S, 3 * 128), dtype=dt, device=dev)
            opad = torch.empty((N * heads, S, 128), dtype=dt, device=dev)
            gpad = torch.empty((N * heads, S, 128), dtype=dt, device=dev)
            dpad = torch.empty((N * heads, S, 3 * 128), dtype=dt, device=dev)
            stats = torch.empty((N * heads, S, 2), dtype=torch.float32, device=dev)
            lib, st = self.lib, _stream()
            _lib.check(lib.hd_mha_pack_heads(_p(qkv), _p(pad), N, S, C, heads, 3, 1.0, st), "hd_mha_pack_heads")
            _lib.check(lib.hd_mha_pack_heads(_p(out), _p(opad), N, S, C, heads, 1, 1.0, st), "hd_mha_pack_heads")
            _lib.check(lib.hd_mha_pack_heads(_p(dout), _p(gpad), N, S, C, heads, 1, 1.0, st), "hd_mha_pack_heads")
            _lib.check(lib.hd_attn_bwd_tc_scaled(_p(pad), _p(opad), _p(gpad), _p(lse), _p(stats), _p(dpad), N * heads, S, hd ** -0.5, st),
                       "hd_attn_bwd_tc_scaled")
            _lib.check(lib.hd_mha_unpack_heads(_p(dpad), _p(dqkv), N, S, C, heads, 3, 1.0, st), "hd_mha_unpack_heads")
            self.launches += 7
            self.tc_launches += 2
            self._t1(e0, "mha_bwd_tc", 8.0 * N * S * S * C)
            return
        _lib.check(self.lib.hd_mha_bwd(_DT[qkv.dtype], _p(qkv), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), N, S, C, heads, _stream()),
                   "hd_mha_bwd")
        self.launches += 2
        self._t1(e0, "mha_bwd", 8.0 * N * S * S * C)

    # ---- hybrid pipeline helpers -----------------------------------------------------------
    def upsample_nearest(self, x, out):
        """NHWC [N, H, W, C] -> out [N, fy H, fx W, C] (F.interpolate(mode="nearest") by integer factors)."""
        N, H, W, C = x.shape
        fy, fx = out.shape[1] // H, out.shape[2] // W
        assert out.shape == (N, fy * H, fx * W, C)
        _lib.check(self.lib.hd_upsample_nearest(_DT[x.dtype], _p(x), _p(out), N, H, W, C, fy, fx, _stream()), "hd_upsample_nearest")
        self.launches += 1

    def upsample_nearest_bwd(self, dout, din):
        N, H, W, C = din.shape
        fy, fx = dout.shape[1] // H, dout.shape[2] // W
        assert dout.shape == (N, fy * H, fx * W, C)
        _lib.check(self.lib.hd_upsample_nearest_bwd(_DT[din.dtype], _p(dout), _p(din), N, H, W, C, fy, fx, _stream()), "hd_upsample_nearest_bwd")
        self.launches += 1

    def image_affine(self, x, out, scale, shift):
        """out (fp32) = x * scale + shift for a uint8 or fp32 image tensor of any shape (contiguous)."""
        assert x.dtype in (torch.uint8, torch.float32) and x.is_contiguous() and out.dtype == torch.float32
        _lib.check(self.lib.hd_image_affine(_p(x), int(x.dtype == torch.uint8), _p(out), float(scale), float(shift), x.numel(), _stream()),
                   "hd_image_affine")
        self.launches += 1

    # ---- GroupNorm family ----------------------------------------------------------------
    def gn_stats(self, x0, x1, N, HW, G, sums):
        e0 = self._t0()
        C0, C1 = x0.shape[-1], 0 if x1 is None else x1.shape[-1]
        _lib.check(self.lib.hd_gn_stats(_DT[x0.dtype], _p(x0), C0, _p(x1), C1, N, HW, G, _p(sums), _stream()), "hd_gn_stats")
        self.launches += 1
        self._t1(e0, "gn_stats", float(N * HW * (C0 + C1) * x0.element_size()))

    def gn_group_sums(self, cs0, cs1, N, G, sums):
        """sums[N][G][2] from per-channel sums cs0 [N][C0][2] (| cs1 [N][C1][2]) left by the producing convolutions."""
        _lib.check(self.lib.hd_gn_group_sums(_p(cs0), cs0.shape[1], _p(cs1), 0 if cs1 is None else cs1.shape[1], N, G, _p(sums),
                                             _stream()), "hd_gn_group_sums")
        self.launches += 1

    def gn_apply(self, x0, x1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, out):
        e0 = self._t0()
        C0, C1 = x0.shape[-1], 0 if x1 is None else x1.shape[-1]
        _lib.check(self.lib.hd_gn_apply(_DT[x0.dtype], _p(x0), C0, _p(x1), C1, N, HW, G, _p(sums), _p(gamma), _p(beta),
                                        eps, int(act), float(p_drop), int(seed), _p(out), _stream()), "hd_gn_apply")
        self.launches += 1
        self._t1(e0, "gn_apply", float(2 * N * HW * (C0 + C1) * x0.element_size()))

    def gn_bwd(self, x0, x1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, dy, gsums, dgamma, dbeta,
               add, acc0, acc1, dx0, dx1, cs_total=None, cs_per_n=None, cs_n=None, overwrite_dy=False):
        """dgamma/dbeta accumulate; dx0/dx1 are overwritten with dx (+ add + acc0/acc1).  cs_total[c] / cs_per_n[n, c]
        (optional, accumulate) receive the column sums of dx for the leading cs_n channels (default: all).
        overwrite_dy: the caller does not need dy afterwards — the reduce pass leaves dy * mask * act'(z) in it and the apply
        pass reads that instead of recomputing the sigmoid and the dropout hash."""
        C0, C1 = x0.shape[-1], 0 if x1 is None else x1.shape[-1]
        dt = _DT[x0.dtype]
        e0 = self._t0()
        a = (dt, _p(x0), C0, _p(x1), C1, N, HW, G, _p(sums), _p(gamma), _p(beta), eps, int(act), float(p_drop), int(seed), _p(dy))
        if self.gn_fused and cs_total is None and cs_per_n is None:
            if self._gn_counter is None or self._gn_counter.device != x0.device:
                self._gn_counter = torch.zeros(1, dtype=torch.int32, device=x0.device)
            _lib.check(self.lib.hd_gn_bwd_fused(*a, _p(gsums), _p(dgamma), _p(dbeta), _p(add), _p(acc0), _p(acc1), _p(dx0), _p(dx1),
                                                _p(self._gn_counter), _stream()), "hd_gn_bwd_fused")
            self.launches += 1
        else:
            inplace = (bool(overwrite_dy) and self.gn_inplace and (act or p_drop > 0)
                       and not self.lib.hd_gn_v2(dt, C0, C1, G, HW, N))       # first-generation kernels only (see csrc/hd_gn.cu)
            _lib.check(self.lib.hd_gn_bwd_reduce(*a, _p(gsums), _p(dgamma), _p(dbeta), _p(dy) if inplace else None, _stream()),
                       "hd_gn_bwd_reduce")
            _lib.check(self.lib.hd_gn_bwd_apply(*a, _p(gsums), _p(add), _p(acc0), _p(acc1), _p(dx0), _p(dx1), _p(cs_total), _p(cs_per_n),
                                                0 if cs_per_n is None else cs_per_n.stride(0),
                                                (C0 + C1) if cs_n is None else int(cs_n), int(inplace), _stream()), "hd_gn_bwd_apply")
            self.launches += 2
        # compulsory bytes (every operand once): x, dy in, dx out, + the optional addends
        nb = (3 * (C0 + C1) + (0 if add is None else C0 + C1) + (0 if acc0 is None else C0) + (0 if acc1 is None else C1))
        self._t1(e0, "gn_bwd", float(nb * N * HW * x0.element_size()))

    def colsum(self, t, N, HW, C, per_n, total, nchw=False):
        """per_n[n, c] += sum_pix t ; total[c] += sum_{n,pix} t   (either may be None)."""
        _lib.check(self.lib.hd_colsum(_DT[t.dtype], _p(t), int(nchw), N, HW, C, _p(per_n),
                                      0 if per_n is None else per_n.stride(0), _p(total), _stream()), "hd_colsum")
        self.launches += 1

    # ---- embedding path (fp32) -----------------------------------------------------------
    def linear_fwd(self, x, w, b, y, in_swish=False, accumulate=False):
        M, K = x.shape
        _lib.check(self.lib.hd_linear_fwd(_p(x), M, K, x.stride(0), _p(w), _p(b), _p(y), w.shape[0], y.stride(0),
                                          int(in_swish), int(accumulate), _stream()), "hd_linear_fwd")
        self.launches += 1

    def linear_bwd_x(self, dy, w, x_pre, dx, accumulate=False):
        M, Nout = dy.shape
        K = w.shape[1]
        _lib.check(self.lib.hd_linear_bwd_x(_p(dy), M, Nout, dy.stride(0), _p(w), K, _p(x_pre),
                                            0 if x_pre is None else x_pre.stride(0), _p(dx), dx.stride(0),
                                            int(accumulate), _stream()), "hd_linear_bwd_x")
        self.launches += 1

    def linear_bwd_w(self, dy, x, dw, db, in_swish=False):
        M, Nout = dy.shape
        K = x.shape[1]
        _lib.check(self.lib.hd_linear_bwd_w(_p(dy), M, Nout, dy.stride(0), _p(x), K, x.stride(0), int(in_swish),
                                            _p(dw), _p(db), _stream()), "hd_linear_bwd_w")
        self.launches += 1

    def embedding_fwd(self, table, idx, out):
        _lib.check(self.lib.hd_embedding_fwd(_p(table), table.shape[0], table.shape[1], _p(idx), idx.numel(), _p(out), _stream()), "hd_embedding_fwd")
        self.launches += 1

    def embedding_bwd(self, dout, idx, dtable, padding_idx=-1):
        _lib.check(self.lib.hd_embedding_bwd(_p(dout), dout.shape[1], _p(idx), idx.numel(), _p(dtable), padding_idx, _stream()), "hd_embedding_bwd")
        self.launches += 1

    # ---- parameter packing ---------------------------------------------------------------
    def gather_pack(self, src, ia, ib, out):
        _lib.check(self.lib.hd_gather_pack(_DT[out.dtype], _p(src), _p(ia), _p(ib), out.numel(), _p(out), _stream()), "hd_gather_pack")
        self.launches += 1

    def scatter_unpack(self, packed, inv, dst):
        _lib.check(self.lib.hd_scatter_unpack(_p(packed), _p(inv), dst.numel(), _p(dst), _stream()), "hd_scatter_unpack")
        self.launches += 1

    # ---- diffusion process ---------------------------------------------------------------
    def q_sample(self, x0, noise, t, sab, s1ab, xt):
        N = x0.shape[0]
        _lib.check(self.lib.hd_q_sample(_p(x0), _p(noise), _p(t), _p(sab), _p(s1ab), _p(xt), N, x0.numel() // N, sab.numel(), _stream()), "hd_q_sample")
        self.launches += 1

    def mse_fwd(self, pred, noise, loss):
        _lib.check(self.lib.hd_mse_fwd(_p(pred), _p(noise), _p(loss), pred.numel(), _stream()), "hd_mse_fwd")
        self.launches += 1

    def mse_bwd(self, pred, noise, g, dpred):
        _lib.check(self.lib.hd_mse_bwd(_p(pred), _p(noise), _p(g), _p(dpred), pred.numel(), _stream()), "hd_mse_bwd")
        self.launches += 1

    def sampler_step(self, x, eps_c, eps_u, z, w, coef, step_ptr, clip_last, nan_flag):
        _lib.check(self.lib.hd_sampler_step(_p(x), _p(eps_c), _p(eps_u), _p(z), float(1.0 + w), float(w), _p(coef), _p(step_ptr),
                                            int(clip_last), _p(nan_flag), x.numel(), _stream()), "hd_sampler_step")
        self.launches += 1

    def add_int(self, p, delta):
        _lib.check(self.lib.hd_add_int(_p(p), int(delta), _stream()), "hd_add_int")
        self.launches += 1

    # ---- optimizer -----------------------------------------------------------------------
    def sqnorm(self, g, out, accumulate=False):
        """out[0] (fp64) = sum g^2 (accumulate: += , for a norm over several ranges of the flat buffer)"""
        _lib.check(self.lib.hd_sqnorm(_p(g), g.numel(), _p(out), int(accumulate), _stream()), "hd_sqnorm")
        self.launches += 1

    def adamw_flat(self, p, g, m, v, sqnorm, max_norm, lr, b1, b2, eps, wd, step):
        _lib.check(self.lib.hd_adamw_flat(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(sqnorm), float(max_norm), float(lr),
                                          float(b1), float(b2), float(eps), float(wd), int(step), _stream()), "hd_adamw_flat")
        self.launches += 1


_warned = set()


def _warn_cuda_core(what, detail):
    """A bf16 shape outside the tcgen05 tiles runs on the CUDA-core kernels (the fp32 check-mode path): correct, but one to
    two orders of magnitude slower.  Say so once per (operator, shape)."""
    key = (what, detail)
    if key in _warned:
        return
    _warned.add(key)
    import warnings
    warnings.warn(f"hdiff_b200: {what} {detail} is outside the tcgen05 kernels' tiles and runs on the CUDA-core kernel "
                  f"(much slower); see DESIGN.md for the covered shapes", RuntimeWarning, stacklevel=3)


_backend = None


def get():
    """The operator backend.  Created on first use; fails loudly when the CUDA library is absent."""
    global _backend
    if _backend is None:
        _backend = CudaOps()
    return _backend


def set_backend(b):
    """Test hook (tests/emu_backend.py installs a torch re-statement for CPU-only host-logic tests)."""
    global _backend
    _backend = b
