"""GPU parity tests AT THE BENCHMARK'S OWN SHAPES (BASELINE.json configs[1]/[2]: 256x256, ch=64 [1,2,2,2], attention at
S = 16 384): the grids, waves and image boundaries that `bench.py` runs, not scaled-down stand-ins.

  (a) attention forward / backward at S = 16 384, d = 128 (N = 1 and 3) and at S = 1 024, N = 32 (the middle block of a
      batch of 32) against a row-chunked fp32 restatement of ModelCondition.py:101-120 that never forms an S x S tensor;
  (b) convolution / data gradient / weight gradient / GroupNorm at the layer shapes of the cfg2 step with N >= 4, so that
      persistent CTAs walk many tiles, cross image boundaries and the one-wave wgrad splits see M = 262 144 ... 2 M pixels;
  (c) the whole UNet at 256x256: forward and EVERY parameter gradient against the oracle in fp32 on the same GPU.

Tolerances are the north star's: relative error <= 1e-2 in bf16 per layer (a small multiple for whole-network depth).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_torch as R
from tests.emu_backend import EmuOps


@pytest.fixture(autouse=True)
def _setup():
    import hdiff_b200.ops as hops
    hops.set_backend(None)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.cuda.empty_cache()


def _ops():
    import hdiff_b200.ops as hops
    return hops.get()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# ----------------------------------------------------------------------------------------------------------------
# (a) attention
# ----------------------------------------------------------------------------------------------------------------
def attn_reference_chunked(qkv, dout=None, chunk=2048):
    """softmax(q k^T C^-1/2) v and its gradients in fp32, query rows in chunks: peak memory chunk x S, never S x S
    (AttnBlock.forward, ModelCondition.py:108-117: bmm, scale by C^-0.5, softmax over keys, bmm)."""
    N, S, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.float().view(N, S, 3, C).unbind(2)
    scale = C ** -0.5
    out = torch.empty(N, S, C, device=qkv.device)
    lse = torch.empty(N, S, device=qkv.device)
    grads = None if dout is None else torch.zeros(N, S, 3, C, device=qkv.device)
    for n in range(N):
        for i in range(0, S, chunk):
            s = (q[n, i:i + chunk] @ k[n].t()) * scale
            l = torch.logsumexp(s, dim=-1)
            p = torch.exp(s - l[:, None])
            o = p @ v[n]
            out[n, i:i + chunk], lse[n, i:i + chunk] = o, l
            if dout is not None:
                do = dout[n, i:i + chunk].float()
                grads[n, :, 2] += p.t() @ do
                ds = p * (do @ v[n].t() - (do * o).sum(-1, keepdim=True))
                grads[n, i:i + chunk, 0] = (ds @ k[n]) * scale
                grads[n, :, 1] += (ds.t() @ q[n, i:i + chunk]) * scale
    return out, lse, (None if grads is None else grads.view(N, S, 3 * C))


@pytest.mark.parametrize("S,N", [(16384, 1), (16384, 3), (1024, 32)], ids=["S16384_N1", "S16384_N3", "S1024_N32"])
def test_attention_at_benchmark_sequence_lengths(S, N):
    dev = torch.device("cuda")
    ops = _ops()
    C = 128
    assert ops.lib.hd_attn_tc_supported(S, C) and ops.lib.hd_attn_bwd_tc_supported(S, C)
    torch.manual_seed(S + N)
    qkv = torch.randn(N, S, 3 * C, device=dev)
    qkv[:, :, :C] *= 2.0                                                          # sharper rows than unit-variance scores
    qkv[:, :, C:2 * C] *= torch.linspace(0.6, 1.4, S, device=dev)[None, :, None]  # the running maximum keeps moving
    qkv = qkv.to(torch.bfloat16)
    dout = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    out = torch.full((N, S, C), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.full((N, S), float("nan"), device=dev)
    dqkv = torch.full_like(qkv, float("nan"))
    before = ops.tc_launches
    ops.attn_fwd(qkv, out, lse, N, S, C)
    ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C)
    torch.cuda.synchronize()
    assert ops.tc_launches == before + 3, "must run on the tcgen05 kernels"
    ro, rl, rd = attn_reference_chunked(qkv, dout)
    assert _rel(out.float(), ro) < 1e-2, _rel(out.float(), ro)
    assert float((lse - rl).abs().max()) < 2e-2 * max(1.0, float(rl.abs().max()) * 0.05)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        r = _rel(dqkv[:, :, sl].float(), rd[:, :, sl])
        assert r < 1.5e-2, (name, r)
    # every image of the batch, not only the norm over all of them (a wrong tile of one image hides in a global norm)
    for n in range(N):
        assert _rel(out[n].float(), ro[n]) < 1e-2 and _rel(dqkv[n].float(), rd[n]) < 1.5e-2, n


# ----------------------------------------------------------------------------------------------------------------
# (b) convolution family and GroupNorm at the cfg2 layer shapes
# ----------------------------------------------------------------------------------------------------------------
BENCH_CONVS = [
    # name, C0, C1, P_in, Cout, P_out, N, H, W (logical), k, emb, res      -- the layer of the cfg2 step it is
    ("L0_res_64_64", 64, 0, 1, 64, 1, 4, 256, 256, 3, True, True),        # 3x3 M = N*65536, N = 64, K = 576 (x7 per step)
    ("L0_up_cat128+64_64", 128, 64, 1, 64, 1, 4, 256, 256, 3, True, False),  # first L0 up block: K = 1728 over two sources
    ("L0_up_cat64+64_64", 64, 64, 1, 64, 1, 4, 256, 256, 3, True, False),    # K = 1152 over two sources
    ("L0_dgrad_64_192", 64, 0, 1, 192, 1, 4, 256, 256, 3, False, False),     # data gradient of the K = 1728 layer
    ("L0_shortcut_cat_1x1", 128, 64, 1, 64, 1, 4, 256, 256, 1, False, False),
    ("L1_res_128_128", 128, 0, 1, 128, 1, 6, 128, 128, 3, True, True),      # 3x3 M = N*16384, N = 128, K = 1152
    ("L1_up_cat128+128_128", 128, 128, 1, 128, 1, 4, 128, 128, 3, True, False),
    ("L1_qkv_1x1", 128, 0, 1, 384, 1, 4, 128, 128, 1, False, False),
    ("L1_proj_1x1_res", 128, 0, 1, 128, 1, 4, 128, 128, 1, False, True),
    ("down_256_to_128_c64", 64, 0, 2, 64, 1, 4, 128, 128, 3, False, False),  # DownSample as one 3x3 over the s2d view
    ("down_128_to_64_c128", 128, 0, 2, 128, 1, 4, 64, 64, 3, False, False),
    ("convT_128_to_256_c128", 128, 0, 1, 128, 2, 4, 128, 128, 3, False, False),  # ConvTranspose 5x5 s2: M = 2 M, K = 3200
    ("convT_dgrad_256_to_128_c128", 128, 0, 2, 128, 1, 4, 128, 128, 3, False, False),
    ("L2_res_128_128_64x64", 128, 0, 1, 128, 1, 8, 64, 64, 3, True, True),
    ("L3_res_128_128_32x32", 128, 0, 1, 128, 1, 8, 32, 32, 3, True, True),
]


def _conv_inputs(case, dev):
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, use_emb, use_res = case
    g = torch.Generator(device="cuda").manual_seed(abs(hash(name)) % 2 ** 31)
    bf = torch.bfloat16
    x0 = torch.randn(N, H * P_in, W * P_in, C0, generator=g, device=dev).to(bf)
    x1 = torch.randn(N, H, W, C1, generator=g, device=dev).to(bf) if C1 else None
    CinL, CoutL = (C0 + C1) * P_in * P_in, Cout * P_out * P_out
    w = (torch.randn(CoutL * k * k * CinL, generator=g, device=dev) / (k * k * CinL) ** 0.5).to(bf)
    bias = torch.randn(CoutL, generator=g, device=dev)
    emb = torch.randn(N, CoutL, generator=g, device=dev) if use_emb else None
    res = torch.randn(N, H * P_out, W * P_out, Cout, generator=g, device=dev).to(bf) if use_res else None
    return x0, x1, w, bias, emb, res


@pytest.mark.parametrize("case", BENCH_CONVS, ids=[c[0] for c in BENCH_CONVS])
def test_conv_forward_and_wgrad_at_cfg2_layer_shapes(case):
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, use_emb, use_res = case
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    assert ops.lib.hd_conv_tc_supported(C0, C1, P_in, Cout, P_out, H, W, k), "must be covered by the tcgen05 kernel"
    x0, x1, w, bias, emb, res = _conv_inputs(case, dev)
    out = torch.full((N, H * P_out, W * P_out, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    want_stats = P_out == 1 and bool(ops.lib.hd_conv_tc_stats_staged(C0, C1, P_in, Cout, P_out, H, W, k))
    cs = torch.zeros(N, Cout, 2, dtype=torch.float64, device=dev) if want_stats else None
    before = ops.tc_launches
    ops.conv(x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, chan_sums=cs)
    torch.cuda.synchronize()
    assert ops.tc_launches == before + 1
    ref = torch.empty(N, H * P_out, W * P_out, Cout, dtype=torch.float32, device=dev)
    emu.conv(x0.float(), None if x1 is None else x1.float(), P_in, w.float(), bias, emb, None if res is None else res.float(),
             ref, P_out, N, H, W, k)
    assert _rel(out.float(), ref) < 1e-2, (name, _rel(out.float(), ref))
    for n in range(N):                                   # per image: a mis-indexed tile in one image must not hide
        assert _rel(out[n].float(), ref[n]) < 1e-2, (name, n)
    if cs is not None:                                   # GroupNorm statistics left by the epilogue, on the stored values
        o = out.double().reshape(N, -1, Cout)
        assert torch.allclose(cs[:, :, 0], o.sum(1), rtol=1e-5, atol=5e-2)      # fp32 partial sums per warp, fp64 across
        assert torch.allclose(cs[:, :, 1], (o * o).sum(1), rtol=1e-5, atol=5e-2)
    del ref
    # weight gradient of the same layer
    assert ops.lib.hd_wgrad_tc_supported(C0, C1, P_in, Cout, P_out, H, W, k)
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = torch.randn(N, H * P_out, W * P_out, Cout, generator=g, device=dev).to(torch.bfloat16)
    dw = torch.full((w.numel(),), float("nan"), dtype=torch.float32, device=dev)
    before = ops.tc_launches
    ops.wgrad(x0, x1, P_in, dy, P_out, dw, N, H, W, k, torch.bfloat16)
    torch.cuda.synchronize()
    assert ops.tc_launches == before + 1
    rdw = torch.empty_like(dw)
    emu.wgrad(x0.float(), None if x1 is None else x1.float(), P_in, dy.float(), P_out, rdw, N, H, W, k, torch.float32)
    assert _rel(dw, rdw) < 1e-2, (name, _rel(dw, rdw))


GN_SHAPES = [
    # N, H*W, C0, C1, p_drop       -- where it occurs in the cfg2 step
    (4, 256 * 256, 64, 0, 0.1),     # L0 block2 (dropout)
    (4, 256 * 256, 128, 64, 0.0),   # first L0 up block: GroupNorm over cat(128, 64), groups of 6 channels straddle the sources
    (4, 256 * 256, 64, 64, 0.0),
    (6, 128 * 128, 128, 0, 0.1),
    (4, 128 * 128, 128, 128, 0.0),
    (8, 64 * 64, 128, 0, 0.1),
    (8, 32 * 32, 128, 128, 0.0),
]


@pytest.mark.parametrize("shape", GN_SHAPES, ids=[f"N{s[0]}_HW{s[1]}_C{s[2]}+{s[3]}_p{s[4]}" for s in GN_SHAPES])
def test_groupnorm_swish_at_cfg2_layer_shapes(shape):
    """GroupNorm(32) + Swish (+ dropout) forward and backward at full-size tensors.  With dropout the kernel's own mask
    (recovered from a forward pass on the same seed: dropped <=> output exactly 0 where the no-dropout output is not)
    is applied to the checker, so values and gradients are compared element for element."""
    N, HW, C0, C1, p_drop = shape
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    bf = torch.bfloat16
    C = C0 + C1
    torch.manual_seed(HW + C)
    x0 = (torch.randn(N, HW, 1, C0, device=dev) * 1.7 + 0.4).to(bf)
    x1 = (torch.randn(N, HW, 1, C1, device=dev) * 0.6 - 0.8).to(bf) if C1 else None
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev) * 0.5
    seed = 0x1234567 + HW
    sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
    ops.gn_stats(x0, x1, N, HW, 32, sums)
    rs = torch.empty_like(sums)
    emu.gn_stats(x0, x1, N, HW, 32, rs)
    assert _rel(sums, rs) < 1e-6
    out = torch.empty(N, HW, 1, C, dtype=bf, device=dev)
    ops.gn_apply(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p_drop, seed, out)
    ref = torch.empty(N, HW, 1, C, device=dev)
    emu.gn_apply(x0.float(), None if x1 is None else x1.float(), N, HW, 32, rs, gamma, beta, 1e-5, 1, 0.0, 0, ref)
    keep = None
    if p_drop > 0:
        nodrop = torch.empty_like(out)
        ops.gn_apply(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, 0.0, 0, nodrop)
        keep = ~((out == 0) & (nodrop != 0))
        frac = 1.0 - float(keep.float().mean())
        assert abs(frac - p_drop) < 2e-3, frac                      # the drop rate on 10^7-10^8 elements
        ref = ref * keep / (1.0 - p_drop)
        del nodrop
    assert _rel(out.float(), ref) < 1e-2, _rel(out.float(), ref)
    del ref, out
    # backward: dx (+ add + acc0), dgamma, dbeta, and the column sums handed to the producing convolution
    dy = torch.randn(N, HW, 1, C, device=dev).to(bf)
    add = torch.randn(N, HW, 1, C, device=dev).to(bf)
    acc0 = torch.randn(N, HW, 1, C0, device=dev).to(bf)
    gs = torch.empty_like(sums)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    cs_tot, cs_n = torch.zeros(C0, device=dev), torch.zeros(N, C0, device=dev)
    dx0 = torch.empty_like(x0)
    dx1 = None if x1 is None else torch.empty_like(x1)
    dy_in = dy.clone()
    ops.gn_bwd(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p_drop, seed, dy_in, gs, dg, db, add, acc0, None, dx0, dx1,
               cs_total=cs_tot, cs_per_n=cs_n, cs_n=C0, overwrite_dy=True)
    torch.cuda.synchronize()
    rdg, rdb = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    r0 = torch.empty(N, HW, 1, C0, device=dev)
    r1 = None if x1 is None else torch.empty(N, HW, 1, C1, device=dev)
    rcs_tot, rcs_n = torch.zeros(C0, device=dev), torch.zeros(N, C0, device=dev)
    dyr = dy.float() if keep is None else dy.float() * keep / (1.0 - p_drop)     # dropout sits after the activation
    emu.gn_bwd(x0.float(), None if x1 is None else x1.float(), N, HW, 32, rs, gamma, beta, 1e-5, 1, 0.0, 0, dyr, None, rdg, rdb,
               add.float(), acc0.float(), None, r0, r1, cs_total=rcs_tot, cs_per_n=rcs_n, cs_n=C0)
    assert _rel(dx0.float(), r0) < 1e-2, _rel(dx0.float(), r0)
    if x1 is not None:
        assert _rel(dx1.float(), r1) < 1e-2, _rel(dx1.float(), r1)
    assert _rel(dg, rdg) < 1e-2 and _rel(db, rdb) < 1e-2, (_rel(dg, rdg), _rel(db, rdb))
    # column sums of the bf16-rounded dx over 10^5..10^6 pixels: absolute slack = rounding noise of that many terms
    slack = 4e-3 * float(r0.abs().mean()) * (N * HW) ** 0.5
    assert float((cs_tot - rcs_tot).abs().max()) < slack * 4, float((cs_tot - rcs_tot).abs().max())
    assert float((cs_n - rcs_n).abs().max()) < slack * 4


# ----------------------------------------------------------------------------------------------------------------
# (c) the whole network at 256 x 256
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cond", [False, True], ids=["uncond_cfg2", "cond_cfg3"])
def test_unet_256_forward_and_all_gradients_vs_fp32_oracle(cond):
    """BASELINE.json configs[1] / [2] network at its own resolution (256x256; attention at S = 16 384 and 1 024), batch 2:
    forward and every parameter gradient of the bf16 CUDA path against the oracle in fp32 (TF32 off) on the same GPU."""
    from hdiff_b200.diffusion.Model import UNet as UNetU
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet as UNetC
    import hdiff_b200.ops as hops
    dev = torch.device("cuda")
    cfg = dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(5)
    ref = R.UNet(num_labels=10 if cond else None, **cfg)
    net = UNetC(num_labels=10, **cfg) if cond else UNetU(**cfg)
    net.load_state_dict(ref.state_dict())
    net, ref = net.to(dev).train(), ref.to(dev).train()
    B = 2
    torch.manual_seed(6)
    x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
    t = torch.tensor([17, 803], device=dev)
    lab = torch.tensor([0, 4], device=dev) if cond else None       # one null label (label dropout), one class
    before = hops.get().tc_launches
    e = net(x, t, lab) if cond else net(x, t)
    er = ref(x, t, lab)
    assert _rel(e.detach(), er.detach()) < 3e-2, _rel(e.detach(), er.detach())
    gy = torch.randn_like(er)
    e.backward(gy)
    er.backward(gy)
    assert hops.get().tc_launches - before > 200, "the 256x256 step must run on the tcgen05 kernels"
    pr = dict(ref.named_parameters())
    gscale = max(float(p.grad.norm()) for p in pr.values() if p.grad is not None)
    bad = []
    for k, p in net.named_parameters():
        if pr[k].grad is None:
            assert p.grad is None, k                                # cond_proj of the unconditional model: no gradient
            continue
        a, b = p.grad.double(), pr[k].grad.double()
        if float((a - b).norm()) > 0.12 * float(b.norm()) + 6e-3 * 3e-2 * gscale + 2e-5:
            bad.append((k, _rel(a, b)))
    assert not bad, bad[:8]
