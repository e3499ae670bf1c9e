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


class ResBlock(nn.Module):
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False):
        super().__init__()
        self.in_ch, self.out_ch, self.p_drop = in_ch, out_ch, float(dropout)
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(), nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.cond_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1))
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

    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None, compute_dtype=None):
        super().__init__()
        assert all(i < len(ch_mult) for i in attn), 'attn index out of bound'
        if compute_dtype is None:
            compute_dtype = torch.float32 if os.environ.get("HDIFF_COMPUTE", "bf16") == "fp32" else torch.bfloat16
        assert compute_dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = compute_dtype
        self.T, self.ch, self.num_labels = T, ch, num_labels
        tdim = ch * 4
        self.tdim = tdim
        self.time_embedding = TimeEmbedding(T, ch, tdim)
        if num_labels is not None:
            self.cond_embedding = ConditionalEmbedding(num_labels, ch, tdim)
        self.head = nn.Conv2d(3, ch, kernel_size=3, stride=1, padding=1)
        self.downblocks = nn.ModuleList()
        chs = [ch]
        now_ch = ch
        for i, mult in enumerate(ch_mult):
            out_ch = ch * mult
            for _ in range(num_res_blocks):
                self.downblocks.append(ResBlock(now_ch, out_ch, tdim, dropout, attn=(i in attn)))
                now_ch = out_ch
                chs.append(now_ch)
            if i != len(ch_mult) - 1:
                self.downblocks.append(DownSample(now_ch))
                chs.append(now_ch)
        self.middleblocks = nn.ModuleList([ResBlock(now_ch, now_ch, tdim, dropout, attn=True),
                                           ResBlock(now_ch, now_ch, tdim, dropout, attn=False)])
        self.upblocks = nn.ModuleList()
        self._up_split = []      # per up ResBlock: (C0 from below, C1 from the skip)
        for i, mult in reversed(list(enumerate(ch_mult))):
            out_ch = ch * mult
            for _ in range(num_res_blocks + 1):
                skip = chs.pop()
                self.upblocks.append(ResBlock(skip + now_ch, out_ch, tdim, dropout, attn=False))
                self._up_split.append((now_ch, skip))
                now_ch = out_ch
            if i != 0:
                self.upblocks.append(UpSample(now_ch))
        assert len(chs) == 0
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
        conv_ids = {id(p) for m in self.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) for p in m.parameters()}
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
                if isinstance(mod.attn, AttnBlock):
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

    def _conv(self, st, spec, x0, x1=None, emb=None, res=None, dgrad=False, in_nchw=False, out_nchw=False, want_stats=False):
        """want_stats: let the kernel's epilogue leave per-image per-channel (sum, sum of squares) of the output for the
        GroupNorm that reads it next (tcgen05 path only; otherwise _gn_fwd falls back to the statistics kernel)."""
        ops = _ops.get()
        P_in, P_out = (spec.P_out, spec.P_in) if dgrad else (spec.P_in, spec.P_out)
        Cout = spec.Cin if dgrad else spec.Cout
        if in_nchw:
            N, _, H, W = x0.shape
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
                       Cout_pad=Cout if out_nchw and Cout != spec.real_cout else None, chan_sums=cs)
        if cs is not None and got:
            st.chan[id(out)] = (weakref.ref(out), cs)     # keyed by the tensor OBJECT: an address can be reused after a free
        return out

    def _wgrad(self, st, spec, x0, x1, dy, in_nchw=False, dy_nchw=False, bias_done=False):
        """weight gradient into gpk (packed layout) + bias gradient (column sums of dy; `bias_done`: the producer of dy
        already accumulated them, see _gn_bwd)."""
        ops = _ops.get()
        if in_nchw:
            N, _, H, W = x0.shape
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
                      alg_frac=spec.alg_frac)
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
        if not (WGRAD_SIDE_STREAM and like.is_cuda):
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
        if "qkv" in sp:
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
        if "qkv" in sp:
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

    def _run_forward(self, x, t, labels, save):
        ops = _ops.get()
        st = self._get_state()
        if not self._frozen:
            self.repack()
        training = self.training
        dev = x.device
        st.chan, st.chan_pool, st.chan_off = {}, None, 0
        assert x.dim() == 4 and x.shape[1] == 3 and x.dtype == torch.float32
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
        if labels is not None:
            labels = labels.to(torch.int64).contiguous()
            ce = self.cond_embedding.condEmbedding
            c0 = torch.empty((N, self.ch), **f32)
            ops.embedding_fwd(ce[0].weight, labels, c0)
            c1 = torch.empty((N, self.tdim), **f32)
            ops.linear_fwd(c0, ce[1].weight, ce[1].bias, c1)
            cemb = torch.empty((N, self.tdim), **f32)
            ops.linear_fwd(c1, ce[3].weight, ce[3].bias, cemb, in_swish=True)
            w_c = st.flat[st.o_cw: st.o_cw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
            b_c = st.flat[st.o_cb: st.o_cb + st.emb_total]
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
        for mod in self.upblocks:
            if isinstance(mod, ResBlock):
                h, c = self._res_fwd(st, mod, ri, h, hs.pop(), emb_all, save, training)
                ri += 1
            else:
                sp = st.blocks[id(mod)]
                xin = h
                u = self._conv(st, sp["convT"], xin)
                h = self._conv(st, sp["conv"], u, want_stats=True)
                c = {"x": xin, "u": u} if save else None
            bctx.append(c)
        assert len(hs) == 0 and ri == nrb
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
        dskip = {}                      # index into hs -> gradient w.r.t. that skip tensor
        next_skip = 0                   # up blocks consume hs from the end; in reverse order from the start
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
                    dskip[next_skip] = d_sk
                    next_skip += 1
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
        if reducer is not None:
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
        if "labels" in ctx:
            ce = self.cond_embedding.condEmbedding
            cemb, c1, c0 = ctx["cemb"], ctx["c1"], ctx["c0"]
            w_c = st.flat[st.o_cw: st.o_cw + st.emb_total * self.tdim].view(st.emb_total, self.tdim)
            ops.linear_bwd_w(d_emb_all, cemb, fg[st.o_cw: st.o_cw + st.emb_total * self.tdim], fg[st.o_cb: st.o_cb + st.emb_total], in_swish=True)
            d_cemb = torch.empty_like(cemb)
            ops.linear_bwd_x(d_emb_all, w_c, cemb, d_cemb)
            ops.linear_bwd_w(d_cemb, c1, gv(ce[3].weight), gv(ce[3].bias), in_swish=True)
            d_c1 = torch.empty_like(c1)
            ops.linear_bwd_x(d_cemb, ce[3].weight, c1, d_c1)
            ops.linear_bwd_w(d_c1, c0, gv(ce[1].weight), gv(ce[1].bias))
            d_c0 = torch.empty_like(c0)
            ops.linear_bwd_x(d_c1, ce[1].weight, None, d_c0)
            ops.embedding_bwd(d_c0, ctx["labels"], gv(ce[0].weight), padding_idx=0)
        # ---- data parallel: the directly written part of the flat buffer (GroupNorm, Linear, embedding tables); then the
        #      compute stream waits for every exchange of this step ----
        self._join_side(st)
        if reducer is not None:
            reducer.finish(fg[:st.n_direct])
        # ---- packed conv gradients -> parameter layouts ----
        ops.scatter_unpack(st.gpk, st.inv, fg)
        used_cond = "labels" in ctx
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
    def forward(fctx, net, x, t, labels, *params):
        with torch.no_grad():
            eps, ctx = net._run_forward(x.detach(), t, labels, save=True)
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
        return (None, None, None, None) + tuple(by_id[i] for i in fctx.param_ids)


def _apply(net, x, t, labels):
    st = net._get_state()
    params = st.order
    if labels is None:
        skip = {id(q) for rb in net._resblocks() for q in rb.cond_proj.parameters()}
        params = [p for p in params if id(p) not in skip]
    return _UNetFunction.apply(net, x, t, labels, *params)


# route UNetBase.forward through the single autograd node
def _forward(self, x, t, labels=None):
    if self.num_labels is None:
        assert labels is None, "unconditional UNet takes (x, t)"
    else:
        assert labels is not None, "conditional UNet takes (x, t, labels)"
    st = self._get_state()
    if torch.is_grad_enabled() and any(p.requires_grad for p in st.order):
        return _apply(self, x, t, labels)
    eps, _ = self._run_forward(x, t, labels, save=False)
    return eps


UNetBase.forward = _forward
