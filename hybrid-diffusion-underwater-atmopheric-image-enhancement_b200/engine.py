"""Host side of the UNet hot path: parameter containers with the reference's names, the flat
parameter / gradient buffers, the packed GEMM weight layouts, and the forward / backward schedule
of C-ABI kernel launches.

Reference structure followed (paths relative to the reference repository):
  blocks      DiffusionFreeGuidence/ModelCondition.py:22-164  (Swish, TimeEmbedding, ConditionalEmbedding,
              DownSample, UpSample, AttnBlock, ResBlock_old)  == diffusion/Model.py:18-265
  topology    DiffusionFreeGuidence/ModelCondition.py:213-276 (UNet), call signature diffusion/Train.py:30-31
  backward    PyTorch autograd of the above (triggered at TrainCondition.py:60)

The nn.Module tree below only HOLDS parameters (same names, shapes and default initialisation as
the reference, so state_dicts interchange); none of the torch layer `forward`s is ever called.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops as _ops

GN_GROUPS = 32
GN_EPS = 1e-5
# GroupNorm statistics out of the producing convolution's epilogue: "auto" = only where the kernel takes them from the tile it
# has staged in shared memory for its tensor store (cheap), "1" = also through the register butterfly of the direct-store
# epilogue (measured slower than the separate statistics pass), "0" = never.
CONV_EPILOGUE_STATS = os.environ.get("HDIFF_CONV_STATS", "auto")
# Weight gradients leave the dependency chain of the backward pass (nothing reads them before the optimizer / all-reduce),
# so they CAN be issued on a second stream: the tensor-pipe-bound wgrad kernels then run under the HBM-bound GroupNorm
# backward kernels of the next layer instead of in front of them.  Measured on B200 (cfg2, batch 32): 86.1 -> 84.9 ms per
# step only (the GPU is power-capped and a wgrad CTA leaves room for one GroupNorm CTA per SM), and per-launch event times
# stop meaning anything once kernels overlap, so it is OFF by default (HDIFF_WGRAD_STREAM=1 enables it).
WGRAD_SIDE_STREAM = os.environ.get("HDIFF_WGRAD_STREAM", "0") == "1"


# =============================================================================================
# Parameter containers (names == reference state_dict keys)
# =============================================================================================
class Swish(nn.Module):
    """Place holder so that nn.Sequential indices match the reference (ModelCondition.py:22-24)."""

    def forward(self, x):  # pragma: no cover - never called by the engine
        raise RuntimeError("hdiff_b200 modules are parameter containers; call the UNet")


def sinusoid_table(T: int, d_model: int) -> torch.Tensor:
    """fp32 [T, d_model], (sin, cos) interleaved per frequency (ModelCondition.py:31-38)."""
    assert d_model % 2 == 0
    freq = torch.exp(-(torch.arange(0, d_model, step=2) / d_model * math.log(10000)))
    ang = torch.arange(T).float()[:, None] * freq[None, :]
    return torch.stack([torch.sin(ang), torch.cos(ang)], dim=-1).view(T, d_model)


class TimeEmbedding(nn.Module):
    def __init__(self, T, d_model, dim):
        super().__init__()
        self.timembedding = nn.Sequential(
            nn.Embedding.from_pretrained(sinusoid_table(T, d_model), freeze=False),
            nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))


class ConditionalEmbedding(nn.Module):
    def __init__(self, num_labels, d_model, dim):
        super().__init__()
        assert d_model % 2 == 0
        self.condEmbedding = nn.Sequential(
            nn.Embedding(num_labels + 1, d_model, padding_idx=0),
            nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))


class ImageConditionalEmbedding(nn.Module):
    """DynamicUNet's condition encoder (diffusion/Model.py:110-167): three stride-2 3x3 convolutions (no activation between
    them), global average pool, Linear -> Swish -> Linear.  Parameter container; the engine runs it."""

    def __init__(self, d_model, dim):
        super().__init__()
        channels = d_model // 16
        self.conv1 = nn.Conv2d(3, channels, kernel_size=3, stride=2, padding=1)
        self.conv2 = nn.Conv2d(channels, channels * 2, kernel_size=3, stride=2, padding=1)
        self.conv3 = nn.Conv2d(channels * 2, channels * 4, kernel_size=3, stride=2, padding=1)
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.linear1 = nn.Linear(channels * 4, dim)
        self.activation = Swish()
        self.linear2 = nn.Linear(dim, dim)


class DownSample(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.c1 = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1)
        self.c2 = nn.Conv2d(in_ch, in_ch, 5, stride=2, padding=2)


class UpSample(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.c = nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1)
        self.t = nn.ConvTranspose2d(in_ch, in_ch, 5, 2, 2, 1)


class AttnBlock(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1)
        self.proj = nn.Conv2d(in_ch, in_ch, 1)


MHA_HEADS = 8      # nn.MultiheadAttention(out_ch, num_heads=8), ModelCondition.py:189 == diffusion/Model.py:290


class ResBlock(nn.Module):
    """attn=True, mha=False: ResBlock_old + AttnBlock (GroupNorm, single head, residual; ModelCondition.py:124-164).
    attn=True, mha=True: the reference's live ResBlock (ModelCondition.py:166-211 == diffusion/Model.py:267-312): 8-head
    nn.MultiheadAttention on the flattened map whose output REPLACES h (no norm, no residual)."""

    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False, mha=False):
        super().__init__()
        self.in_ch, self.out_ch, self.p_drop = in_ch, out_ch, float(dropout)
        self.mha = bool(attn and mha)
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(), nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.cond_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1))
        # registration order = state_dict key order: ResBlock_old has shortcut before attn (ModelCondition.py:145-153), the MHA
        # ResBlock attn before shortcut (:189-194)
        if attn and mha:
            self.attn = nn.MultiheadAttention(out_ch, num_heads=MHA_HEADS)       # holds in_proj_weight / in_proj_bias / out_proj.*
            self.shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()
        else:
            self.shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()
            self.attn = AttnBlock(out_ch) if attn else nn.Identity()


# =============================================================================================
# Packed-layout descriptions
# =============================================================================================
class ConvSpec:
    """One logical stride-1 convolution (k in {1,3}) between views, with its packed weights."""
    __slots__ = ("name", "k", "CinL", "CoutL", "P_in", "P_out", "Cout", "Cin", "w_off", "wd_off", "b_off", "n_w", "has_bias",
                 "_bias_inv", "_bias_same", "alg_frac", "real_cout")

    def __init__(self, name, k, Cin, Cout, P_in, P_out):
        self.name, self.k, self.Cin, self.Cout, self.P_in, self.P_out = name, k, Cin, Cout, P_in, P_out
        self.CinL, self.CoutL = Cin * P_in * P_in, Cout * P_out * P_out
        self.n_w = self.CoutL * k * k * self.CinL
        self.w_off = self.wd_off = self.b_off = -1
        self.has_bias = True
        # real reference taps / packed taps: DownSample packs 9 + 25 taps into 36 slots, ConvTranspose 25 into 36
        self.alg_frac = 34.0 / 36.0 if P_in == 2 else (25.0 / 36.0 if P_out == 2 else 1.0)
        self.real_cout = Cout


def _tbl_conv(off, Co, Ci, k):
    return (torch.arange(Co * Ci * k * k, dtype=torch.int64).view(Co, Ci, k, k).permute(0, 2, 3, 1) + off).contiguous()


def _tbl_down(off3, off5, C):
    """DownSample (ModelCondition.py:68-76): conv3x3 s2 p1 + conv5x5 s2 p2 == ONE 3x3 stride-1 convolution
    over the 2x2 space-to-depth view: W'[co][ty][tx][py][px][c] = W5[co][c][2(ty-1)+py+2][..] + W3[co][c][2(ty-1)+py+1][..]."""
    i5 = torch.arange(C * C * 25, dtype=torch.int64).view(C, C, 5, 5) + off5
    i3 = torch.arange(C * C * 9, dtype=torch.int64).view(C, C, 3, 3) + off3
    A = torch.full((C, 3, 3, 2, 2, C), -1, dtype=torch.int64)
    B = torch.full((C, 3, 3, 2, 2, C), -1, dtype=torch.int64)
    for ty in range(3):
        for py in range(2):
            d5y, d3y = 2 * (ty - 1) + py + 2, 2 * (ty - 1) + py + 1
            for tx in range(3):
                for px in range(2):
                    d5x, d3x = 2 * (tx - 1) + px + 2, 2 * (tx - 1) + px + 1
                    if 0 <= d5y <= 4 and 0 <= d5x <= 4:
                        A[:, ty, tx, py, px, :] = i5[:, :, d5y, d5x]
                    if 0 <= d3y <= 2 and 0 <= d3x <= 2:
                        B[:, ty, tx, py, px, :] = i3[:, :, d3y, d3x]
    return A.view(C, 3, 3, 4 * C), B.view(C, 3, 3, 4 * C)


def _tbl_s2(off3, Co, Ci):
    """Conv2d(Ci, Co, 3, stride=2, padding=1) as a 3x3 stride-1 convolution over the 2x2 space-to-depth view:
    W'[co][ty][tx][(py, px, ci)] = W[co][ci][2(ty-1)+py+1][2(tx-1)+px+1] (taps outside the 3x3 kernel stay zero)."""
    i3 = torch.arange(Co * Ci * 9, dtype=torch.int64).view(Co, Ci, 3, 3) + off3
    A = torch.full((Co, 3, 3, 2, 2, Ci), -1, dtype=torch.int64)
    for ty in range(3):
        for py in range(2):
            dy = 2 * (ty - 1) + py + 1
            if not 0 <= dy <= 2:
                continue
            for tx in range(3):
                for px in range(2):
                    dx = 2 * (tx - 1) + px + 1
                    if 0 <= dx <= 2:
                        A[:, ty, tx, py, px, :] = i3[:, :, dy, dx]
    return A.view(Co, 3, 3, 4 * Ci)


def _tbl_convT(off, C):
    """ConvTranspose2d(C, C, 5, 2, 2, 1) (ModelCondition.py:83): out[2Y+py, 2X+px, co] is a 3x3 stride-1
    convolution of the input with W''[(py,px,co)][ty][tx][ci] = Wt[ci][co][2(1-ty)+py+2][2(1-tx)+px+2]."""
    it = torch.arange(C * C * 25, dtype=torch.int64).view(C, C, 5, 5) + off   # [ci][co][ky][kx]
    Ftab = torch.full((2, 2, C, 3, 3, C), -1, dtype=torch.int64)
    for py in range(2):
        for ty in range(3):
            ky = -2 * (ty - 1) + py + 2
            if not 0 <= ky <= 4:
                continue
            for px in range(2):
                for tx in range(3):
                    kx = -2 * (tx - 1) + px + 2
                    if 0 <= kx <= 4:
                        Ftab[py, px, :, ty, tx, :] = it[:, :, ky, kx].t()
    return Ftab.view(4 * C, 3, 3, C)


def _dgrad_tbl(Ftab):
    return Ftab.flip(1, 2).permute(3, 1, 2, 0).contiguous()


# =============================================================================================
# The network
# =============================================================================================
class UNetBase(nn.Module):
    """UNet(T, ch, ch_mult, attn, num_res_blocks, dropout[, num_labels]) running on the hdiff_b200
    kernels.  `attn` lists the levels whose down-path ResBlocks carry an AttnBlock;
    the middle is [attn, no-attn], the up path has none (ModelCondition.py:233-236,242)."""

    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None, compute_dtype=None, mha=False,
                 in_channels=3, middle_attn=(True, False), up_extra=1, image_cond=False):
        """mha: attention ResBlocks are the reference's live nn.MultiheadAttention variant (ModelCondition.py:166-211).
        attn: levels whose down-path ResBlocks carry attention ("all": every level, as the live UNet :226 builds them).
        in_channels / middle_attn / up_extra / image_cond describe DynamicUNet (diffusion/Model.py:382-517): 6-channel head,
        four attention middle blocks, num_res_blocks (not + 1) up blocks per level, image condition encoder."""
        super().__init__()
        attn = list(range(len(ch_mult))) if attn == "all" else list(attn)
        assert all(i < len(ch_mult) for i in attn), 'attn index out of bound'
        if compute_dtype is None:
            compute_dtype = torch.float32 if os.environ.get("HDIFF_COMPUTE", "bf16") == "fp32" else torch.bfloat16
        assert compute_dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = compute_dtype
        self.T, self.ch, self.num_labels = T, ch, num_labels
        self.mha, self.in_channels, self.image_cond, self.up_extra = bool(mha), int(in_channels), bool(image_cond), int(up_extra)
        assert 1 <= self.in_channels <= 8
        tdim = ch * 4
        self.tdim = tdim
        self.time_embedding = TimeEmbedding(T, ch, tdim)
        if image_cond:
            assert ch % 32 == 0
            self.cond_embedding = ImageConditionalEmbedding(ch, tdim)
        elif num_labels is not None:
            self.cond_embedding = ConditionalEmbedding(num_labels, ch, tdim)
        self.head = nn.Conv2d(self.in_channels, ch, kernel_size=3, stride=1, padding=1)
        self.downblocks = nn.ModuleList()
        chs = [ch]
        now_ch = ch
        for i, mult in enumerate(ch_mult):
            out_ch = ch * mult
            for _ in range(num_res_blocks):
                self.downblocks.append(ResBlock(now_ch, out_ch, tdim, dropout, attn=(i in attn), mha=mha))
                now_ch = out_ch
                chs.append(now_ch)
            if i != len(ch_mult) - 1:
                self.downblocks.append(DownSample(now_ch))
                chs.append(now_ch)
        self.middleblocks = nn.ModuleList([ResBlock(now_ch, now_ch, tdim, dropout, attn=bool(a), mha=mha) for a in middle_attn])
        self.upblocks = nn.ModuleList()
        self._up_split = []      # per up ResBlock: (C0 from below, C1 from the skip)
        for i, mult in reversed(list(enumerate(ch_mult))):
            out_ch = ch * mult
            for _ in range(num_res_blocks + up_extra):
                skip = chs.pop()
                self.upblocks.append(ResBlock(skip + now_ch, out_ch, tdim, dropout, attn=False))
                self._up_split.append((now_ch, skip))
                now_ch = out_ch
            if i != 0:
                self.upblocks.append(UpSample(now_ch))
        assert up_extra != 1 or len(chs) == 0
        self.tail = nn.Sequential(nn.GroupNorm(32, now_ch), Swish(), nn.Conv2d(now_ch, 3, 3, stride=1, padding=1))
        self._state = None       # built lazily on the parameters' device
        self._frozen = False
        self.dp_group = None     # set by hdiff_b200.parallel.enable_data_parallel
        self.dp_bucket_bytes = 8 << 20

    def load_state_dict(self, state_dict, strict=True, **kw):
        """As nn.Module.load_state_dict.  A checkpoint of the reference's LIVE ModelCondition.UNet carries
        nn.MultiheadAttention keys (`attn.in_proj_weight`, `attn.out_proj.*`, ModelCondition.py:189); this class is the
        AttnBlock variant (`attn.proj_q/k/v`), so under strict=False (TrainCondition.py:36-38) every attention weight would be
        dropped silently: refuse instead and point at the MHA model."""
        if not getattr(self, "mha", False) and any(".attn.in_proj_weight" in k or ".attn.out_proj." in k for k in state_dict):
            raise RuntimeError("this state_dict holds nn.MultiheadAttention weights (the reference's live MHA ResBlock, "
                               "ModelCondition.py:166-211); build the model with mha=True (hdiff_b200 ... UNet(..., mha=True)) to load it")
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # -----------------------------------------------------------------------------------------
    # flat buffers + packed layouts
    # -----------------------------------------------------------------------------------------
    def _resblocks(self) -> List[ResBlock]:
        return [m for m in list(self.downblocks) + list(self.middleblocks) + list(self.upblocks) if isinstance(m, ResBlock)]

    def _build_state(self):
        dev = self.head.weight.device
        rbs = self._resblocks()
        named = dict(self.named_parameters())
        order = []                               # parameter objects in flat order
        order += [rb.temb_proj[1].weight for rb in rbs]
        order += [rb.temb_proj[1].bias for rb in rbs]
        order += [rb.cond_proj[1].weight for rb in rbs]
        order += [rb.cond_proj[1].bias for rb in rbs]
        seen = {id(p) for p in order}
        # convolution parameters last: their gradients are produced in packed form (gpk) and all-reduced there, so
        # the data-parallel exchange of the flat buffer only covers the prefix [0, n_direct)
        conv_ids = {id(p) for m in self.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.MultiheadAttention))
                    for p in m.parameters()}
        order += [p for p in named.values() if id(p) not in seen and id(p) not in conv_ids]
        n_direct_params = len(order)
        order += [p for p in named.values() if id(p) not in seen and id(p) in conv_ids]
        offs, off = {}, 0
        for p in order:
            offs[id(p)] = off
            off += (p.numel() + 3) // 4 * 4      # keep every parameter 16-byte aligned
        n_flat = off
        n_direct = offs[id(order[n_direct_params])] if n_direct_params < len(order) else n_flat
        flat = torch.zeros(n_flat, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in order:
                o = offs[id(p)]
                flat[o:o + p.numel()].copy_(p.detach().reshape(-1).float())
                p.data = flat[o:o + p.numel()].view(p.shape)
        st = _State()
        st.device, st.flat, st.n_flat, st.offs, st.order = dev, flat, n_flat, offs, order
        st.n_direct = n_direct
        st.ptrs = [p.data_ptr() for p in order]
        st.flat_grad = None
        O = lambda p: offs[id(p)]

        # ---- embedding path offsets (flat fp32, used in place) ----
        st.emb_total = sum(rb.out_ch for rb in rbs)
        st.emb_offs, e = [], 0
        for rb in rbs:
            st.emb_offs.append(e)
            e += rb.out_ch
        st.o_tw, st.o_tb = O(rbs[0].temb_proj[1].weight), O(rbs[0].temb_proj[1].bias)
        st.o_cw, st.o_cb = O(rbs[0].cond_proj[1].weight), O(rbs[0].cond_proj[1].bias)
        # contiguity of the grouped projections (16-byte padding never triggers: out_ch*tdim % 4 == 0)
        assert all(rb.out_ch % 4 == 0 for rb in rbs), "channel counts must be multiples of 4"

        # ---- conv specs and index tables ----
        specs, tabA, tabB, tabD, tabD2, btabA, btabB = [], [], [], [], [], [], []
        inv = torch.full((n_flat,), -1, dtype=torch.int64)
        wcur = [0]
        bcur = [0]

        def add(spec, Ftab, F2=None, bias_a=None, bias_b=None, bias_inv=None):
            spec.w_off = wcur[0]
            tabA.append(Ftab.reshape(-1))
            tabB.append((F2 if F2 is not None else torch.full_like(Ftab, -1)).reshape(-1))
            wcur[0] += spec.n_w
            spec.wd_off = wcur[0]
            tabA.append(_dgrad_tbl(Ftab).reshape(-1))
            tabB.append((_dgrad_tbl(F2) if F2 is not None else torch.full_like(Ftab, -1)).reshape(-1))
            wcur[0] += spec.n_w
            # inverse (gradient) map: packed dW position of every parameter element
            lin = torch.arange(spec.n_w, dtype=torch.int64)
            for tb in (Ftab, F2):
                if tb is not None:
                    f = tb.reshape(-1)
                    m = f >= 0
                    inv[f[m]] = lin[m] + spec.w_off // 2      # dW buffer has no dgrad copies: half the stride
            if bias_a is not None:
                spec.b_off = bcur[0]
                btabA.append(bias_a)
                btabB.append(bias_b if bias_b is not None else torch.full_like(bias_a, -1))
                bcur[0] += bias_a.numel()
                spec._bias_inv = bias_inv
            else:
                spec.has_bias = False
            specs.append(spec)
            return spec

        def conv_spec(name, m: nn.Conv2d, pad_in=None, pad_out=None):
            """pad_in / pad_out: GEMM widths of a zero-padded weight matrix (the 3-channel head / tail run on the
            tcgen05 kernel as 64-channel convolutions; padded rows / columns are zero and receive no gradient)."""
            Co, Ci, k, _ = m.weight.shape
            CoP, CiP = pad_out or Co, pad_in or Ci
            s = ConvSpec(name, k, CiP, CoP, 1, 1)
            s.alg_frac = (Co * Ci) / float(CoP * CiP)
            s.real_cout = Co
            tbl = torch.full((CoP, k, k, CiP), -1, dtype=torch.int64)
            tbl[:Co, :, :, :Ci] = _tbl_conv(O(m.weight), Co, Ci, k)
            bidx = torch.full((CoP,), -1, dtype=torch.int64)
            bidx[:Co] = torch.arange(Co, dtype=torch.int64) + O(m.bias)
            return add(s, tbl, bias_a=bidx, bias_inv=[(O(m.bias), Co)])

        def linear_spec(name, w, bvec):
            """a Linear / packed projection [Co, Ci] run as a 1x1 convolution (its weight IS the packed GEMM layout)"""
            Co, Ci = w.shape
            return add(ConvSpec(name, 1, Ci, Co, 1, 1), _tbl_conv(O(w), Co, Ci, 1),
                       bias_a=torch.arange(Co, dtype=torch.int64) + O(bvec), bias_inv=[(O(bvec), Co)])

        # the image condition encoder's stride-2 convolutions come FIRST in the packed buffers: their gradients are the last
        # ones of a backward pass, and the data-parallel exchange walks the buffer from its end
        st.cond = {}
        if self.image_cond:
            ce = self.cond_embedding
            for nm in ("conv1", "conv2", "conv3"):
                m = getattr(ce, nm)
                Co, Ci = m.weight.shape[:2]
                sp_ = ConvSpec("cond_" + nm, 3, Ci, Co, 2, 1)
                sp_.alg_frac = 9.0 / 36.0
                st.cond[nm] = add(sp_, _tbl_s2(O(m.weight), Co, Ci), bias_a=torch.arange(Co, dtype=torch.int64) + O(m.bias),
                                  bias_inv=[(O(m.bias), Co)])
        st.pad_io = self.compute_dtype == torch.bfloat16
        st.head = conv_spec("head", self.head, pad_in=64 if st.pad_io else None)
        st.tail = conv_spec("tail", self.tail[2], pad_out=64 if st.pad_io else None)
        st.blocks = {}
        st.block_lo = {}
        for mod in list(self.downblocks) + list(self.middleblocks) + list(self.upblocks):
            b = {}
            if isinstance(mod, ResBlock):
                b["conv1"] = conv_spec("conv1", mod.block1[2])
                b["conv2"] = conv_spec("conv2", mod.block2[3])
                if isinstance(mod.shortcut, nn.Conv2d):
                    b["shortcut"] = conv_spec("shortcut", mod.shortcut)
                if mod.mha:
                    b["qkv"] = linear_spec("qkv", mod.attn.in_proj_weight, mod.attn.in_proj_bias)
                    b["proj"] = linear_spec("proj", mod.attn.out_proj.weight, mod.attn.out_proj.bias)
                elif isinstance(mod.attn, AttnBlock):
                    a = mod.attn
                    C = mod.out_ch
                    s = ConvSpec("qkv", 1, C, 3 * C, 1, 1)
                    Ftab = torch.cat([_tbl_conv(O(c.weight), C, C, 1) for c in (a.proj_q, a.proj_k, a.proj_v)], 0)
                    bidx = torch.cat([torch.arange(C, dtype=torch.int64) + O(c.bias) for c in (a.proj_q, a.proj_k, a.proj_v)])
                    b["qkv"] = add(s, Ftab, bias_a=bidx, bias_inv=[(O(c.bias), C) for c in (a.proj_q, a.proj_k, a.proj_v)])
                    b["proj"] = conv_spec("proj", a.proj)
            elif isinstance(mod, DownSample):
                C = mod.c1.weight.shape[0]
                s = ConvSpec("down", 3, C, C, 2, 1)
                A, B = _tbl_down(O(mod.c1.weight), O(mod.c2.weight), C)
                ar = torch.arange(C, dtype=torch.int64)
                b["down"] = add(s, A, B, bias_a=ar + O(mod.c1.bias), bias_b=ar + O(mod.c2.bias),
                                bias_inv=[(O(mod.c1.bias), C), (O(mod.c2.bias), C)])
                s._bias_same = True
            elif isinstance(mod, UpSample):
                C = mod.c.weight.shape[0]
                s = ConvSpec("convT", 3, C, C, 1, 2)
                ar = torch.arange(C, dtype=torch.int64)
                b["convT"] = add(s, _tbl_convT(O(mod.t.weight), C), bias_a=(ar + O(mod.t.bias)).repeat(4),
                                 bias_inv=[(O(mod.t.bias), C)])
                b["conv"] = conv_spec("conv", mod.c)
            st.blocks[id(mod)] = b
            st.block_lo[id(mod)] = min(sp.w_off for sp in b.values()) // 2
        st.specs = specs
        st.n_wpack = wcur[0]
        st.n_dw = wcur[0] // 2
        st.n_bpack = bcur[0]
        # bias gradients live behind the weight gradients in one buffer `gpk`
        for s in specs:
            if s.has_bias:
                pos = st.n_dw + s.b_off
                if getattr(s, "_bias_same", False):
                    for (bo, n) in s._bias_inv:           # both biases read the same column sums
                        inv[bo:bo + n] = torch.arange(n, dtype=torch.int64) + pos
                else:
                    q = pos
                    for (bo, n) in s._bias_inv:
                        inv[bo:bo + n] = torch.arange(n, dtype=torch.int64) + q
                        q += n
        assert st.n_wpack < 2 ** 31 and n_flat < 2 ** 31
        st.ia = torch.cat(tabA).to(torch.int32).to(dev)
        st.ib = torch.cat(tabB).to(torch.int32).to(dev)
        st.bia = torch.cat(btabA).to(torch.int32).to(dev)
        st.bib = torch.cat(btabB).to(torch.int32).to(dev)
        st.inv = inv.to(torch.int32).to(dev)
        st.wpack = torch.empty(st.n_wpack, dtype=self.compute_dtype, device=dev)
        st.bpack = torch.empty(st.n_bpack, dtype=torch.float32, device=dev)
        st.gpk = torch.zeros(st.n_dw + st.n_bpack, dtype=torch.float32, device=dev)
        st.packed_version = None
        st.wgrad_ws = None
        st.chan, st.chan_pool, st.chan_off = {}, None, 0
        st.sum_cout = sum(sp.Cout for sp in specs)
        self._state = st
        return st

    def _get_state(self):
        st = self._state
        if st is None or st.device != self.head.weight.device or st.wpack.dtype != self.compute_dtype \
                or any(p.data_ptr() != q for p, q in zip(st.order, st.ptrs)):
            st = self._build_state()
        return st

    def repack(self):
        """Refresh the packed (compute-dtype) GEMM weights from the fp32 parameters."""
        st = self._get_state()
        ops = _ops.get()
        ops.gather_pack(st.flat, st.ia, st.ib, st.wpack)
        ops.gather_pack(st.flat, st.bia, st.bib, st.bpack)

    class _Frozen:
        def __init__(self, net):
            self.net = net

        def __enter__(self):
            self.net.repack()
            self.prev = self.net._frozen
            self.net._frozen = True

        def __exit__(self, *a):
            self.net._frozen = self.prev

    def frozen_weights(self):
        """Context in which the parameters are known not to change (sampling): pack once."""
        return UNetBase._Frozen(self)

    # -----------------------------------------------------------------------------------------
    # forward schedule
    # -----------------------------------------------------------------------------------------
    def _wv(self, st, spec, dgrad=False):
        o = spec.wd_off if dgrad else spec.w_off
        return st.wpack[o:o + spec.n_w]

    def _bv(self, st, spec):
        return st.bpack[spec.b_off:spec.b_off + spec.CoutL] if spec.has_bias else None

    def _chan_alloc(self, st, N, C):
        """[N][C][2] fp64 zeros out of the per-forward pool (one memset per forward instead of one per convolution)."""
        n = N * C * 2
        if st.chan_pool is None or st.chan_off + n > st.chan_pool.numel():
            st.chan_pool = torch.zeros(max(n, N * 2 * st.sum_cout), dtype=torch.float64, device=st.device)
            st.chan_off = 0
        v = st.chan_pool[st.chan_off: st.chan_off + n].view(N, C, 2)
        st.chan_off += n
        return v

    def _conv(self, st, spec, x0, x1=None, emb=None, res=None, dgrad=False, in_nchw=False, out_nchw=False, want_stats=False,
              quiet=False):
        """want_stats: let the kernel's epilogue leave per-image per-channel (sum, sum of squares) of the output for the
        GroupNorm that reads it next (tcgen05 path only; otherwise _gn_fwd falls back to the statistics kernel)."""
        ops = _ops.get()
        P_in, P_out = (spec.P_out, spec.P_in) if dgrad else (spec.P_in, spec.P_out)
        Cout = spec.Cin if dgrad else spec.Cout
        if in_nchw:
            N, H, W = x0.shape[0], x0.shape[2] // P_in, x0.shape[3] // P_in
        else:
            N, H, W = x0.shape[0], x0.shape[1] // P_in, x0.shape[2] // P_in
        if out_nchw:
            out = torch.empty((N, spec.real_cout, H, W), dtype=torch.float32, device=x0.device)
        else:
            out = torch.empty((N, H * P_out, W * P_out, Cout), dtype=self.compute_dtype, device=x0.device)
        cs = None
        if (want_stats and CONV_EPILOGUE_STATS != "0" and not dgrad and not out_nchw and not in_nchw and P_out == 1
                and self.compute_dtype == torch.bfloat16 and ops.use_tc):
            C0, C1 = x0.shape[-1], 0 if x1 is None else x1.shape[-1]
            if CONV_EPILOGUE_STATS == "1" or ops.lib.hd_conv_tc_stats_staged(C0, C1, P_in, Cout, P_out, H, W, spec.k):
                cs = self._chan_alloc(st, N, Cout)
        got = ops.conv(x0, x1, P_in, self._wv(st, spec, dgrad), None if dgrad else self._bv(st, spec), emb, res, out, P_out,
                       N, H, W, spec.k, in_nchw=in_nchw, out_nchw=out_nchw, alg_frac=spec.alg_frac,
                       Cout_pad=Cout if out_nchw and Cout != spec.real_cout else None, chan_sums=cs, quiet=quiet)
        if cs is not None and got:
            st.chan[id(out)] = (weakref.ref(out), cs)     # keyed by the tensor OBJECT: an address can be reused after a free
        return out

    def _wgrad(self, st, spec, x0, x1, dy, in_nchw=False, dy_nchw=False, bias_done=False, quiet=False):
        """weight gradient into gpk (packed layout) + bias gradient (column sums of dy; `bias_done`: the producer of dy
        already accumulated them, see _gn_bwd)."""
        ops = _ops.get()
        if in_nchw:
            N, H, W = x0.shape[0], x0.shape[2] // spec.P_in, x0.shape[3] // spec.P_in
        else:
            N, H, W = x0.shape[0], x0.shape[1] // spec.P_in, x0.shape[2] // spec.P_in
        dw = st.gpk[spec.w_off // 2: spec.w_off // 2 + spec.n_w]
        side = self._side_stream(st, dy)
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())     # dy, the saved input and the zeroed bias region are ready
            for t in (x0, x1, dy):
                if t is not None:
                    t.record_stream(side)                     # the allocator must not hand the block out while `side` reads it
        with torch.cuda.stream(side) if side is not None else _nullctx():
            ops.wgrad(x0, x1, spec.P_in, dy, spec.P_out, dw, N, H, W, spec.k, self.compute_dtype, in_nchw=in_nchw, dy_nchw=dy_nchw,
                      alg_frac=spec.alg_frac, quiet=quiet)
            if spec.has_bias and not bias_done and id(spec) not in st.bias_done:
                C = spec.Cout                       # physical channels of dy (bias is per physical channel)
                db = st.gpk[st.n_dw + spec.b_off: st.n_dw + spec.b_off + C]
                if dy_nchw:
                    ops.colsum(dy, N, dy.shape[2] * dy.shape[3], C, None, db, nchw=True)
                else:
                    ops.colsum(dy, N, dy.shape[1] * dy.shape[2], C, None, db)

    @staticmethod
    def _side_stream(st, like):
        """The stream the weight-gradient kernels run on (None: the current stream)."""
        if not (WGRAD_SIDE_STREAM and like.is_cuda) or getattr(_ops.get(), "prof", None) is not None:      # per-launch timing: everything on one stream
            return None
        if getattr(st, "side", None) is None:
            st.side = torch.cuda.Stream(device=like.device)
        return st.side

    @staticmethod
    def _join_side(st):
        """The current stream waits for every weight gradient issued so far."""
        side = getattr(st, "side", None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)

    def _gn_fwd(self, x0, x1, gn: nn.GroupNorm, act, p_drop=0.0, seed=0):
        ops = _ops.get()
        N, H, W = x0.shape[:3]
        C = x0.shape[3] + (0 if x1 is None else x1.shape[3])
        sums = torch.empty((N, GN_GROUPS, 2), dtype=torch.float64, device=x0.device)
        st = self._state
        def left_behind(t):
            e = st.chan.get(id(t))
            return e[1] if e is not None and e[0]() is t else None
        cs0 = left_behind(x0)
        cs1 = None if x1 is None else left_behind(x1)
        if cs0 is not None and (x1 is None or cs1 is not None):
            ops.gn_group_sums(cs0, cs1, N, GN_GROUPS, sums)          # statistics left behind by the producing convolutions
        else:
            ops.gn_stats(x0, x1, N, H * W, GN_GROUPS, sums)
        out = torch.empty((N, H, W, C), dtype=self.compute_dtype, device=x0.device)
        ops.gn_apply(x0, x1, N, H * W, GN_GROUPS, sums, gn.weight, gn.bias, GN_EPS, act, p_drop, seed, out)
        return out, sums

    def _gn_bwd(self, st, x0, x1, gn, sums, act, p_drop, seed, dy, add=None, acc0=None, acc1=None, cs_total=None, cs_per_n=None,
                cs_n=None):
        """cs_total / cs_per_n: accumulate the column sums of the returned gradient (bias and embedding-add gradients of
        the convolution that produced x) inside the apply pass instead of re-reading the gradient."""
        ops = _ops.get()
        N, H, W = x0.shape[:3]
        gs = torch.empty((N, GN_GROUPS, 2), dtype=torch.float64, device=x0.device)
        dx0 = torch.empty_like(x0)
        dx1 = None if x1 is None else torch.empty_like(x1)
        # dy is this function's to consume (every caller passes a gradient nobody else reads), unless it doubles as an addend
        own = dy is not add and dy is not acc0 and dy is not acc1
        ops.gn_bwd(x0, x1, N, H * W, GN_GROUPS, sums, gn.weight, gn.bias, GN_EPS, act, p_drop, seed, dy, gs,
                   st.grad_view(gn.weight), st.grad_view(gn.bias), add, acc0, acc1, dx0, dx1, cs_total=cs_total, cs_per_n=cs_per_n,
                   cs_n=cs_n, overwrite_dy=own)
        return dx0, dx1

    def _res_fwd(self, st, rb: ResBlock, idx, x0, x1, emb_all, save, training):
        sp = st.blocks[id(rb)]
        ctx = {}
        a1, sums1 = self._gn_fwd(x0, x1, rb.block1[0], act=1)
        eo = st.emb_offs[idx]
        h1 = self._conv(st, sp["conv1"], a1, emb=emb_all[:, eo:eo + rb.out_ch], want_stats=True)
        p_drop = rb.p_drop if training else 0.0
        seed = 0
        if p_drop > 0:       # host-side counter stream: no device sync, reproducible under torch.manual_seed
            self._drop_calls = getattr(self, "_drop_calls", 0) + 1
            seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._drop_calls * 0xD1B54A32D192ED03) & ((1 << 63) - 1)
        a2, sums2 = self._gn_fwd(h1, None, rb.block2[0], act=1, p_drop=p_drop, seed=seed)
        if "shortcut" in sp:
            s = self._conv(st, sp["shortcut"], x0, x1)
        else:
            s = x0
        h2 = self._conv(st, sp["conv2"], a2, res=s, want_stats=True)
        out = h2
        if rb.mha:
            # nn.MultiheadAttention(q = k = v = h): packed in-projection, 8 heads, out-projection; the result REPLACES h
            # (no norm, no residual: ModelCondition.py:203-208)
            ops = _ops.get()
            N, H, W, C = h2.shape
            qkv = self._conv(st, sp["qkv"], h2)
            o = torch.empty_like(h2)
            lse = torch.empty((N, MHA_HEADS, H * W), dtype=torch.float32, device=h2.device)
            ops.mha_fwd(qkv, o, lse, N, H * W, C, MHA_HEADS)
            out = self._conv(st, sp["proj"], o, want_stats=True)
            if save:
                ctx.update(qkv=qkv, o=o, lse=lse)
        elif "qkv" in sp:
            ops = _ops.get()
            N, H, W, C = h2.shape
            g, sums3 = self._gn_fwd(h2, None, rb.attn.group_norm, act=0)
            qkv = self._conv(st, sp["qkv"], g)
            o = torch.empty_like(h2)
            lse = torch.empty((N, H * W), dtype=torch.float32, device=h2.device)
            ops.attn_fwd(qkv, o, lse, N, H * W, C)
            out = self._conv(st, sp["proj"], o, res=h2, want_stats=True)
            if save:
                ctx.update(g=g, sums3=sums3, qkv=qkv, o=o, lse=lse)
        if save:
            ctx.update(x0=x0, x1=x1, a1=a1, sums1=sums1, h1=h1, a2=a2, sums2=sums2, h2=h2, p_drop=p_drop, seed=seed, idx=idx)
        return out, ctx

    def _bias_view(self, st, spec, n):
        return st.gpk[st.n_dw + spec.b_off: st.n_dw + spec.b_off + n]

    def _claim_bias(self, st, spec):
        """The caller promises to accumulate the column sums of `spec`'s output gradient itself (inside a GroupNorm
        backward apply pass); the later _wgrad(spec, ...) then skips its own pass over that gradient."""
        st.bias_done.add(id(spec))
        return self._bias_view(st, spec, spec.Cout)

    def _res_bwd(self, st, rb: ResBlock, ctx, d_out, d_emb_all, acc0, prev_spec=None):
        """prev_spec: the convolution whose output is this block's (first) input; its bias gradient = column sums of the
        gradient this block returns, accumulated by the final GroupNorm backward."""
        ops = _ops.get()
        sp = st.blocks[id(rb)]
        x0, x1 = ctx["x0"], ctx["x1"]
        if rb.mha:
            N, H, W, C = ctx["h2"].shape
            self._wgrad(st, sp["proj"], ctx["o"], None, d_out)
            d_o = self._conv(st, sp["proj"], d_out, dgrad=True)
            dqkv = torch.empty_like(ctx["qkv"])
            delta = torch.empty((N, MHA_HEADS, H * W), dtype=torch.float32, device=d_out.device)
            ops.mha_bwd(ctx["qkv"], ctx["o"], d_o, ctx["lse"], delta, dqkv, N, H * W, C, MHA_HEADS)
            self._wgrad(st, sp["qkv"], ctx["h2"], None, dqkv)
            d_h2 = self._conv(st, sp["qkv"], dqkv, dgrad=True)
        elif "qkv" in sp:
            N, H, W, C = ctx["h2"].shape
            self._wgrad(st, sp["proj"], ctx["o"], None, d_out)
            d_o = self._conv(st, sp["proj"], d_out, dgrad=True)
            dqkv = torch.empty_like(ctx["qkv"])
            delta = torch.empty((N, H * W), dtype=torch.float32, device=d_out.device)
            ops.attn_bwd(ctx["qkv"], ctx["o"], d_o, ctx["lse"], delta, dqkv, N, H * W, C)
            self._wgrad(st, sp["qkv"], ctx["g"], None, dqkv)
            d_g = self._conv(st, sp["qkv"], dqkv, dgrad=True)
            d_h2, _ = self._gn_bwd(st, ctx["h2"], None, rb.attn.group_norm, ctx["sums3"], 0, 0.0, 0, d_g, add=d_out,
                                   cs_total=self._claim_bias(st, sp["conv2"]))
        else:
            d_h2 = d_out
        self._wgrad(st, sp["conv2"], ctx["a2"], None, d_h2)
        d_a2 = self._conv(st, sp["conv2"], d_h2, dgrad=True)
        # d_h1 = gradient of conv1's output: its column sums are conv1's bias gradient (total) and the gradient of the
        # per-sample embedding add (per image); both come out of the GroupNorm apply pass
        c1 = sp["conv1"]
        C = c1.Cout
        eo = st.emb_offs[ctx["idx"]]
        db1 = st.gpk[st.n_dw + c1.b_off: st.n_dw + c1.b_off + C]
        d_h1, _ = self._gn_bwd(st, ctx["h1"], None, rb.block2[0], ctx["sums2"], 1, ctx["p_drop"], ctx["seed"], d_a2,
                               cs_total=db1, cs_per_n=d_emb_all[:, eo:eo + C])
        self._wgrad(st, c1, ctx["a1"], None, d_h1, bias_done=True)
        d_a1 = self._conv(st, sp["conv1"], d_h1, dgrad=True)
        if "shortcut" in sp:
            self._wgrad(st, sp["shortcut"], x0, x1, d_h2)
            add = self._conv(st, sp["shortcut"], d_h2, dgrad=True)
        else:
            add = d_h2
        cs = None if prev_spec is None else self._claim_bias(st, prev_spec)
        return self._gn_bwd(st, x0, x1, rb.block1[0], ctx["sums1"], 1, 0.0, 0, d_a1, add=add, acc0=acc0, cs_total=cs,
                            cs_n=x0.shape[-1])

    def _cond_fwd(self, st, labels, ctx, save):
        """Image condition encoder (diffusion/Model.py:153-167): conv s2 x3 (CUDA-core kernels: 4..16 channels), global average
        pool, Linear.  Returns c0 [N, 4 * channels] fp32 (the pooled features, input of linear1)."""
        ops = _ops.get()
        sp = st.cond
        lab = labels.contiguous().float()
        N = lab.shape[0]
        y1 = self._conv(st, sp["conv1"], lab, in_nchw=True, quiet=True)
        y2 = self._conv(st, sp["conv2"], y1, quiet=True)
        y3 = self._conv(st, sp["conv3"], y2, quiet=True)
        C3, HW3 = y3.shape[-1], y3.shape[1] * y3.shape[2]
        c0 = torch.zeros((N, C3), dtype=torch.float32, device=lab.device)
        ops.colsum(y3, N, HW3, C3, c0, None)
        c0.mul_(1.0 / HW3)
        if save:
            ctx.update(cond_lab=lab, cond_y1=y1, cond_y2=y2, cond_shape3=tuple(y3.shape))
        return c0

    def _cond_bwd(self, st, ctx, d_c0):
        """backward of _cond_fwd from the gradient of the pooled features"""
        ops = _ops.get()
        sp = st.cond
        N, H3, W3, C3 = ctx["cond_shape3"]
        d_pool = (d_c0 * (1.0 / (H3 * W3))).to(self.compute_dtype).view(N, 1, 1, C3).contiguous()
        d_y3 = torch.empty((N, H3, W3, C3), dtype=self.compute_dtype, device=d_c0.device)
        ops.upsample_nearest(d_pool, d_y3)                                   # mean-pool backward: every pixel gets d / (H W)
        self._wgrad(st, sp["conv3"], ctx["cond_y2"], None, d_y3, quiet=True)
        d_y2 = self._conv(st, sp["conv3"], d_y3, dgrad=True, quiet=True)
        self._wgrad(st, sp["conv2"], ctx["cond_y1"], None, d_y2, quiet=True)
        d_y1 = self._conv(st, sp["conv2"], d_y2, dgrad=True, quiet=True)
        self._wgrad(st, sp["conv1"], ctx["cond_lab"], None, d_y1, in_nchw=True, quiet=True)

    def _run_forward(self, x, t, labels, save, context_zero=False):
        ops = _ops.get()
        st = self._get_state()
        if not self._frozen:
            self.repack()
        training = self.training
        dev = x.device
        st.chan, st.chan_pool, st.chan_off = {}, None, 0
        assert x.dim() == 4 and x.shape[1] == self.in_channels and x.dtype == torch.float32
        x = x.contiguous()
        N = x.shape[0]
        t = t.to(torch.int64).contiguous()
        te = self.time_embedding.timembedding
        f32 = dict(dtype=torch.float32, device=dev)
        ctx = {"N": N}
        # ---- embedding path (fp32) ----
        e0 = torch.empty((N, self.ch), **f32)
        ops.embedding_fwd(te[0].weight, t, e0)
        e1 = torch.empty((N, self.tdim), **f32)
        ops.linear_fwd(e0, te[1].weight, te[1].bias, e1)
        temb = torch.empty((N, self.tdim), **f32)
        ops.linear_fwd(e1, te[3].weight, te[3].bias, temb, in_swish=True)
        rbs = self._resblocks()
        nrb = len(rbs)
        emb_all = torch.empty((N, st.emb_total), **f32)
        w_t = st.flat[st.o_tw: st.o_tw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
        b_t = st.flat[st.o_tb: st.o_tb + st.emb_total]
        ops.linear_fwd(temb, w_t, b_t, emb_all, in_swish=True)
        ctx.update(t=t, e0=e0, e1=e1, temb=temb)
        w_c = st.flat[st.o_cw: st.o_cw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
        b_c = st.flat[st.o_cb: st.o_cb + st.emb_total]
        if self.image_cond and (context_zero or labels is None):
            # DynamicUNet with context_zero (diffusion/Model.py:483-484): cemb = 0, so cond_proj adds its bias only
            cemb = torch.zeros((N, self.tdim), **f32)
            ops.linear_fwd(cemb, w_c, b_c, emb_all, in_swish=True, accumulate=True)
            ctx.update(cemb=cemb, cond_zero=True)
        elif labels is not None:
            if self.image_cond:
                ce = self.cond_embedding
                lin1, lin2 = ce.linear1, ce.linear2
                c0 = self._cond_fwd(st, labels, ctx, save)
            else:
                labels = labels.to(torch.int64).contiguous()
                ce = self.cond_embedding.condEmbedding
                lin1, lin2 = ce[1], ce[3]
                c0 = torch.empty((N, self.ch), **f32)
                ops.embedding_fwd(ce[0].weight, labels, c0)
            c1 = torch.empty((N, self.tdim), **f32)
            ops.linear_fwd(c0, lin1.weight, lin1.bias, c1)
            cemb = torch.empty((N, self.tdim), **f32)
            ops.linear_fwd(c1, lin2.weight, lin2.bias, cemb, in_swish=True)
            ops.linear_fwd(cemb, w_c, b_c, emb_all, in_swish=True, accumulate=True)
            ctx.update(labels=labels, c0=c0, c1=c1, cemb=cemb)
        # ---- head ----
        if st.pad_io:
            xp = torch.empty((N, x.shape[2], x.shape[3], 64), dtype=self.compute_dtype, device=dev)
            ops.pad_nchw(x, xp)
            h = self._conv(st, st.head, xp, want_stats=True)
            ctx["x"] = xp
        else:
            h = self._conv(st, st.head, x, in_nchw=True)
            ctx["x"] = x
        hs = [h]
        bctx = []
        ri = 0
        for mod in self.downblocks:
            if isinstance(mod, ResBlock):
                h, c = self._res_fwd(st, mod, ri, h, None, emb_all, save, training)
                ri += 1
            else:
                sp = st.blocks[id(mod)]["down"]
                xin = h
                h = self._conv(st, sp, xin, want_stats=True)
                c = {"x": xin} if save else None
            bctx.append(c)
            hs.append(h)
        for mod in self.middleblocks:
            h, c = self._res_fwd(st, mod, ri, h, None, emb_all, save, training)
            ri += 1
            bctx.append(c)
        n_skips = len(hs)
        for mod in self.upblocks:
            if isinstance(mod, ResBlock):
                sk = hs.pop()
                sk_idx = len(hs)                       # position of this skip in the list the down path built
                if sk.shape[1:3] != h.shape[1:3]:
                    # DynamicUNet pops fewer skips than it pushed, so a skip can be one level coarser than h:
                    # F.interpolate(skip_h, size=h.shape[2:], mode="nearest") (diffusion/Model.py:506-510)
                    assert h.shape[1] % sk.shape[1] == 0 and h.shape[2] % sk.shape[2] == 0, "nearest skip resize: integer factors only"
                    up = torch.empty((N, h.shape[1], h.shape[2], sk.shape[3]), dtype=sk.dtype, device=dev)
                    ops.upsample_nearest(sk, up)
                    small_shape, sk = tuple(sk.shape), up
                else:
                    small_shape = None
                h, c = self._res_fwd(st, mod, ri, h, sk, emb_all, save, training)
                if save:
                    c["skip_idx"], c["skip_small"] = sk_idx, small_shape
                ri += 1
            else:
                sp = st.blocks[id(mod)]
                xin = h
                u = self._conv(st, sp["convT"], xin)
                h = self._conv(st, sp["conv"], u, want_stats=True)
                c = {"x": xin, "u": u} if save else None
            bctx.append(c)
        assert (len(hs) == 0 or self.up_extra != 1) and ri == nrb
        del n_skips
        a, sums = self._gn_fwd(h, None, self.tail[0], act=1)
        eps = self._conv(st, st.tail, a, out_nchw=True)
        if save:
            ctx.update(bctx=bctx, tail_h=h, tail_a=a, tail_sums=sums)
        return eps, ctx

    # -----------------------------------------------------------------------------------------
    # backward schedule
    # -----------------------------------------------------------------------------------------
    def _run_backward(self, ctx, d_eps):
        ops = _ops.get()
        st = self._get_state()
        dev = d_eps.device
        N = ctx["N"]
        d_eps = d_eps.contiguous().float()
        # fresh gradient buffers
        if st.flat_grad is None or any(p.grad is not None and p.grad.data_ptr() == st.flat_grad.data_ptr() + 4 * st.offs[id(p)]
                                        for p in st.order[:1] + st.order[-1:]):
            st.flat_grad = torch.zeros(st.n_flat, dtype=torch.float32, device=dev)
        else:
            st.flat_grad.zero_()
        st.gpk[st.n_dw:].zero_()
        st.bias_done = set()
        f32 = dict(dtype=torch.float32, device=dev)
        d_emb_all = torch.zeros((N, st.emb_total), **f32)
        reducer = None
        if self.dp_group is not None:
            from . import parallel
            reducer = parallel.GradReducer(self.dp_group, self.dp_bucket_bytes)
            reducer.attach(st.gpk, st.n_dw)
            reducer.before_reduce = lambda: self._join_side(st)      # weight gradients come from the side stream
            self.last_reducer = reducer
        # ---- tail ----
        if st.pad_io:
            d_eps_p = torch.empty((N, d_eps.shape[2], d_eps.shape[3], 64), dtype=self.compute_dtype, device=dev)
            ops.pad_nchw(d_eps, d_eps_p)
            self._wgrad(st, st.tail, ctx["tail_a"], None, d_eps_p)
            d_a = self._conv(st, st.tail, d_eps_p, dgrad=True)
        else:
            self._wgrad(st, st.tail, ctx["tail_a"], None, d_eps, dy_nchw=True)
            d_a = self._conv(st, st.tail, d_eps, dgrad=True, in_nchw=True)
        d_h, _ = self._gn_bwd(st, ctx["tail_h"], None, self.tail[0], ctx["tail_sums"], 1, 0.0, 0, d_a)
        bctx = ctx["bctx"]
        mods = list(self.downblocks) + list(self.middleblocks) + list(self.upblocks)
        n_down = len(self.downblocks)
        dskip = {}                      # index into the down path's skip list -> gradient w.r.t. that skip tensor
        for li in range(len(mods) - 1, -1, -1):
            mod, c = mods[li], bctx[li]
            is_up = li >= n_down + len(self.middleblocks)
            if isinstance(mod, ResBlock):
                # the tensor feeding this layer is hs[li] for down layers and for the first middle block
                acc0 = None
                if not is_up and li <= n_down:
                    acc0 = dskip.pop(li, None)
                # the convolution that produced this block's first input (the module before it; the head for the first)
                if li == 0:
                    prev_spec = st.head
                else:
                    pm = mods[li - 1]
                    psp = st.blocks[id(pm)]
                    prev_spec = psp["down"] if isinstance(pm, DownSample) else psp["conv"] if isinstance(pm, UpSample) \
                        else psp["proj"] if "proj" in psp else psp["conv2"]
                d_h, d_sk = self._res_bwd(st, mod, c, d_h, d_emb_all, acc0, prev_spec=prev_spec)
                if is_up:
                    if c["skip_small"] is not None:          # the skip was resized by nearest interpolation: sum the blocks back
                        small = torch.empty(c["skip_small"], dtype=d_sk.dtype, device=dev)
                        ops.upsample_nearest_bwd(d_sk, small)
                        d_sk = small
                    dskip[c["skip_idx"]] = d_sk
            elif isinstance(mod, DownSample):
                sp = st.blocks[id(mod)]["down"]
                self._wgrad(st, sp, c["x"], None, d_h)
                d_h = self._conv(st, sp, d_h, res=dskip.pop(li, None), dgrad=True)
            else:
                sp = st.blocks[id(mod)]
                self._wgrad(st, sp["conv"], c["u"], None, d_h)
                d_u = self._conv(st, sp["conv"], d_h, dgrad=True)
                self._wgrad(st, sp["convT"], c["x"], None, d_u)
                d_h = self._conv(st, sp["convT"], d_u, dgrad=True)
            bctx[li] = None
            if reducer is not None:
                reducer.ready(st.block_lo[id(mod)])      # weight gradients of this and all later blocks are final
        assert not dskip, "unconsumed skip gradients"
        # ---- head ----
        self._wgrad(st, st.head, ctx["x"], None, d_h, in_nchw=not st.pad_io)
        # ---- data parallel: every packed weight and bias gradient is final here; their exchange runs under the embedding-path
        #      kernels below (only the small directly-written prefix of the flat buffer has to wait for those) ----
        late_conv = self.image_cond and "labels" in ctx      # the condition encoder's convolutions are still to come
        if reducer is not None and not late_conv:
            self._join_side(st)
            reducer.flush(st.gpk[st.n_dw:])
        # ---- embedding path ----
        fg = st.flat_grad
        te = self.time_embedding.timembedding
        temb, e1, e0 = ctx["temb"], ctx["e1"], ctx["e0"]
        gv = st.grad_view
        w_t = st.flat[st.o_tw: st.o_tw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
        ops.linear_bwd_w(d_emb_all, temb, fg[st.o_tw: st.o_tw + st.emb_total * self.tdim], fg[st.o_tb: st.o_tb + st.emb_total], in_swish=True)
        d_temb = torch.empty_like(temb)
        ops.linear_bwd_x(d_emb_all, w_t, temb, d_temb)
        ops.linear_bwd_w(d_temb, e1, gv(te[3].weight), gv(te[3].bias), in_swish=True)
        d_e1 = torch.empty_like(e1)
        ops.linear_bwd_x(d_temb, te[3].weight, e1, d_e1)
        ops.linear_bwd_w(d_e1, e0, gv(te[1].weight), gv(te[1].bias))
        d_e0 = torch.empty_like(e0)
        ops.linear_bwd_x(d_e1, te[1].weight, None, d_e0)
        ops.embedding_bwd(d_e0, ctx["t"], gv(te[0].weight))
        if ctx.get("cond_zero"):
            # cemb = 0: the cond_proj weights get a zero gradient (Swish(0) = 0), their biases the column sums
            ops.linear_bwd_w(d_emb_all, ctx["cemb"], fg[st.o_cw: st.o_cw + st.emb_total * self.tdim], fg[st.o_cb: st.o_cb + st.emb_total], in_swish=True)
        elif "labels" in ctx:
            if self.image_cond:
                lin1, lin2 = self.cond_embedding.linear1, self.cond_embedding.linear2
            else:
                ce = self.cond_embedding.condEmbedding
                lin1, lin2 = ce[1], ce[3]
            cemb, c1, c0 = ctx["cemb"], ctx["c1"], ctx["c0"]
            w_c = st.flat[st.o_cw: st.o_cw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
            ops.linear_bwd_w(d_emb_all, cemb, fg[st.o_cw: st.o_cw + st.emb_total * self.tdim], fg[st.o_cb: st.o_cb + st.emb_total], in_swish=True)
            d_cemb = torch.empty_like(cemb)
            ops.linear_bwd_x(d_emb_all, w_c, cemb, d_cemb)
            ops.linear_bwd_w(d_cemb, c1, gv(lin2.weight), gv(lin2.bias), in_swish=True)
            d_c1 = torch.empty_like(c1)
            ops.linear_bwd_x(d_cemb, lin2.weight, c1, d_c1)
            ops.linear_bwd_w(d_c1, c0, gv(lin1.weight), gv(lin1.bias))
            d_c0 = torch.empty_like(c0)
            ops.linear_bwd_x(d_c1, lin1.weight, None, d_c0)
            if self.image_cond:
                self._cond_bwd(st, ctx, d_c0)
            else:
                ops.embedding_bwd(d_c0, ctx["labels"], gv(ce[0].weight), padding_idx=0)
        # ---- data parallel: the directly written part of the flat buffer (GroupNorm, Linear, embedding tables); then the
        #      compute stream waits for every exchange of this step ----
        self._join_side(st)
        if reducer is not None:
            if late_conv:
                reducer.finish(st.gpk[st.n_dw:], fg[:st.n_direct])
            else:
                reducer.finish(fg[:st.n_direct])
        # ---- packed conv gradients -> parameter layouts ----
        ops.scatter_unpack(st.gpk, st.inv, fg)
        used_cond = "labels" in ctx or bool(ctx.get("cond_zero"))
        grads = []
        for p in st.order:
            if not p.requires_grad:
                grads.append(None)
                continue
            grads.append(fg[st.offs[id(p)]: st.offs[id(p)] + p.numel()].view(p.shape))
        return grads, used_cond


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class _State:
    side = None

    def grad_view(self, p):
        o = self.offs[id(p)]
        return self.flat_grad[o:o + p.numel()]


class _UNetFunction(torch.autograd.Function):
    """ONE autograd node for the whole network.  Its tensor inputs are exactly the parameters this forward uses: the
    unconditional model's `cond_proj.*` (registered for checkpoint compatibility, never used: ModelCondition.py:199-200)
    are not inputs, so autograd — and torch's DistributedDataParallel, which walks the graph to find the parameters that
    will receive a gradient — sees them as unused, as it does in the reference."""

    @staticmethod
    def forward(fctx, net, x, t, labels, context_zero, *params):
        with torch.no_grad():
            eps, ctx = net._run_forward(x.detach(), t, labels, save=True, context_zero=context_zero)
        fctx.net, fctx.hd_ctx, fctx.param_ids = net, ctx, [id(p) for p in params]
        return eps

    @staticmethod
    def backward(fctx, d_eps):
        net, ctx = fctx.net, fctx.hd_ctx
        fctx.hd_ctx = None
        if ctx is None:
            raise RuntimeError("hdiff_b200 UNet: backward through the same forward twice is not supported")
        with torch.no_grad():
            grads, used_cond = net._run_backward(ctx, d_eps)
        st = net._state
        by_id = {id(p): g for p, g in zip(st.order, grads)}
        return (None, None, None, None, None) + tuple(by_id[i] for i in fctx.param_ids)


def _apply(net, x, t, labels, context_zero=False):
    """The autograd node's tensor inputs are the parameters this forward USES (and that require a gradient bookkeeping-wise)."""
    st = net._get_state()
    params = st.order
    skip = set()
    if net.image_cond:
        if context_zero or labels is None:       # cemb = 0: the condition encoder is not evaluated
            skip = {id(q) for q in net.cond_embedding.parameters()}
    elif labels is None:
        skip = {id(q) for rb in net._resblocks() for q in rb.cond_proj.parameters()}
    if skip:
        params = [p for p in params if id(p) not in skip]
    return _UNetFunction.apply(net, x, t, labels, context_zero, *params)


# route UNetBase.forward through the single autograd node
def _forward(self, x, t, labels=None, context_zero=False):
    if self.image_cond:
        pass                                     # DynamicUNet: labels is an image or None (diffusion/Model.py:475-484)
    elif self.num_labels is None:
        assert labels is None, "unconditional UNet takes (x, t)"
    else:
        assert labels is not None, "conditional UNet takes (x, t, labels)"
    st = self._get_state()
    if torch.is_grad_enabled() and any(p.requires_grad for p in st.order):
        return _apply(self, x, t, labels, context_zero)
    eps, _ = self._run_forward(x, t, labels, save=False, context_zero=context_zero)
    return eps


UNetBase.forward = _forward
