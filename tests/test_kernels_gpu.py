"""GPU parity tests, kernel level: every C-ABI kernel against a plain PyTorch fp32 re-statement of the
same op on seeded inputs (tolerances written next to each check: 1e-5-class for the fp32 check mode,
1e-2 for bf16, as BASELINE.json's north_star states)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from tests.emu_backend import EmuOps, to_logical, from_logical  # torch re-statement used as the checker


@pytest.fixture(autouse=True)
def _no_tf32():
    # the checker must be true fp32: cuDNN / cuBLAS default to TF32 on this GPU
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _ops():
    import hdiff_b200.ops as hops
    hops.set_backend(None)
    return hops.get()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CONV_CASES = [
    # name, C0, C1, P_in, Cout, P_out, N, H, W, k, emb, res
    ("3x3_64_64", 64, 0, 1, 64, 1, 2, 16, 16, 3, True, True),
    ("3x3_128_128_w32", 128, 0, 1, 128, 1, 2, 32, 32, 3, True, False),
    ("3x3_concat_128+64_to_64", 128, 64, 1, 64, 1, 1, 16, 32, 3, False, True),
    ("1x1_concat_shortcut", 64, 128, 1, 128, 1, 2, 16, 16, 1, False, False),
    ("1x1_qkv_128_384", 128, 0, 1, 384, 1, 1, 16, 16, 1, False, False),
    ("down_s2d_64", 64, 0, 2, 64, 1, 2, 16, 16, 3, False, False),
    ("convT_s2d_64", 64, 0, 1, 64, 2, 2, 16, 16, 3, False, False),
    ("convT_s2d_128", 128, 0, 1, 128, 2, 1, 16, 16, 3, False, False),
    ("3x3_wide_w256", 64, 0, 1, 64, 1, 1, 4, 256, 3, False, False),
    ("3x3_ragged_h", 64, 0, 1, 64, 1, 1, 20, 32, 3, False, True),
    # W >= 128: one 128-pixel row segment per tile; 3x3 convolutions run in shifted-operand mode (one halo box per tap row)
    ("3x3_rowseg_64_64", 64, 0, 1, 64, 1, 2, 5, 256, 3, True, True),
    ("3x3_rowseg_concat", 128, 64, 1, 64, 1, 1, 3, 128, 3, True, True),
    ("3x3_rowseg_128_128", 128, 0, 1, 128, 1, 2, 5, 128, 3, True, True),
    ("down_rowseg_s2d", 64, 0, 2, 64, 1, 1, 4, 128, 3, False, False),
    ("convT_rowseg_s2d", 64, 0, 1, 64, 2, 1, 3, 128, 3, False, False),
    ("3x3_rowseg_n192", 64, 0, 1, 192, 1, 1, 2, 256, 3, False, False),
    # several tiles per persistent CTA with a 3-stage ring (fewer stages than TMA-issuing warps)
    ("3x3_rowseg_128_128_many_tiles", 128, 0, 1, 128, 1, 8, 64, 128, 3, True, True),
    # 64 channels, several tiles per CTA, image boundaries inside a CTA's tile range (41 rows per image, ranges of 2): the
    # staged-epilogue statistics flush when the image changes
    ("3x3_rowseg_64_64_many_images", 64, 0, 1, 64, 1, 6, 41, 128, 3, True, True),
    ("1x1_64_64_many_images", 64, 64, 1, 64, 1, 5, 37, 128, 1, False, True),
]


def _conv_inputs(case, dtype, dev):
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, use_emb, use_res = case
    g = torch.Generator(device="cpu").manual_seed(abs(hash(name)) % 2 ** 31)
    x0 = torch.randn(N, H * P_in, W * P_in, C0, generator=g).to(dtype).to(dev)
    x1 = torch.randn(N, H, W, C1, generator=g).to(dtype).to(dev) if C1 else None
    CinL, CoutL = (C0 + C1) * P_in * P_in, Cout * P_out * P_out
    w = (torch.randn(CoutL * k * k * CinL, generator=g) / (k * k * CinL) ** 0.5).to(dtype).to(dev)
    bias = torch.randn(CoutL, generator=g).to(dev)
    emb = torch.randn(N, CoutL, generator=g).to(dev) if use_emb else None
    res = torch.randn(N, H * P_out, W * P_out, Cout, generator=g).to(dtype).to(dev) if use_res else None
    return x0, x1, w, bias, emb, res


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("mode", ["fp32_simt", "bf16_simt", "bf16_tc"])
def test_conv_forward(case, mode):
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, use_emb, use_res = case
    dev = torch.device("cuda")
    dtype = torch.float32 if mode == "fp32_simt" else torch.bfloat16
    ops = _ops()
    ops.use_tc = mode == "bf16_tc"
    if mode == "bf16_tc":
        assert ops.lib.hd_conv_tc_supported(C0, C1, P_in, Cout, P_out, H, W, k), "case must be covered by the tcgen05 kernel"
    x0, x1, w, bias, emb, res = _conv_inputs(case, dtype, dev)
    out = torch.empty(N, H * P_out, W * P_out, Cout, dtype=dtype, device=dev)
    before = ops.tc_launches
    ops.conv(x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k)
    torch.cuda.synchronize()
    assert (ops.tc_launches > before) == (mode == "bf16_tc")
    ref = torch.empty(N, H * P_out, W * P_out, Cout, dtype=torch.float32, device=dev)
    EmuOps().conv(x0.float(), None if x1 is None else x1.float(), P_in, w.float(), bias, emb,
                  None if res is None else res.float(), ref, P_out, N, H, W, k)
    tol = 2e-5 if mode == "fp32_simt" else 1e-2
    assert _rel(out.float(), ref) < tol, (name, mode, _rel(out.float(), ref))
    ops.use_tc = True


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("mode", ["fp32_simt", "bf16_tc"])
def test_conv_wgrad(case, mode):
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, _, _ = case
    dev = torch.device("cuda")
    dtype = torch.float32 if mode == "fp32_simt" else torch.bfloat16
    ops = _ops()
    ops.use_tc = mode == "bf16_tc"
    x0, x1, w, bias, emb, res = _conv_inputs(case, dtype, dev)
    g = torch.Generator(device="cpu").manual_seed(7)
    dy = torch.randn(N, H * P_out, W * P_out, Cout, generator=g).to(dtype).to(dev)
    dw = torch.empty_like(w, dtype=torch.float32)
    ops.wgrad(x0, x1, P_in, dy, P_out, dw, N, H, W, k, dtype)
    torch.cuda.synchronize()
    ref = torch.empty_like(dw)
    EmuOps().wgrad(x0.float(), None if x1 is None else x1.float(), P_in, dy.float(), P_out, ref, N, H, W, k, torch.float32)
    tol = 2e-5 if mode == "fp32_simt" else 1e-2
    assert _rel(dw, ref) < tol, (name, mode, _rel(dw, ref))
    ops.use_tc = True


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] == 1], ids=[c[0] for c in CONV_CASES if c[5] == 1])
def test_conv_epilogue_channel_statistics(case):
    """hd_conv_tc's optional per-image per-channel (sum, sum of squares) of the stored output, and the group statistics
    derived from them, against torch on the kernel's own bf16 output."""
    name, C0, C1, P_in, Cout, P_out, N, H, W, k, use_emb, use_res = case
    dev = torch.device("cuda")
    ops = _ops()
    x0, x1, w, bias, emb, res = _conv_inputs(case, torch.bfloat16, dev)
    out = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=dev)
    cs = torch.zeros(N, Cout, 2, dtype=torch.float64, device=dev)
    got = ops.conv(x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, chan_sums=cs)
    torch.cuda.synchronize()
    assert got is True
    o = out.double().reshape(N, H * W, Cout)
    assert torch.allclose(cs[:, :, 0], o.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(cs[:, :, 1], (o * o).sum(1), rtol=1e-5, atol=1e-3)
    if Cout % 32 == 0:
        sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
        ops.gn_group_sums(cs, None, N, 32, sums)
        ref = torch.empty_like(sums)
        ops.gn_stats(out, None, N, H * W, 32, ref)
        assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-3)


def test_head_tail_layout_convs():
    """3-channel NCHW fp32 boundary: head (NCHW fp32 -> NHWC) and tail (NHWC -> NCHW fp32), forward + wgrad."""
    dev = torch.device("cuda")
    ops = _ops()
    torch.manual_seed(0)
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1e-2)):
        N, H, W, C = 2, 12, 20, 64
        x = torch.randn(N, 3, H, W, device=dev)
        w = (torch.randn(C * 9 * 3, device=dev) / 27 ** 0.5).to(dtype)
        b = torch.randn(C, device=dev)
        out = torch.empty(N, H, W, C, dtype=dtype, device=dev)
        ops.conv(x, None, 1, w, b, None, None, out, 1, N, H, W, 3, in_nchw=True)
        ref = F.conv2d(x, w.float().view(C, 3, 3, 3).permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
        assert _rel(out.float(), ref) < tol
        a = torch.randn(N, H, W, C, device=dev).to(dtype)
        wt = (torch.randn(3 * 9 * C, device=dev) / (9 * C) ** 0.5).to(dtype)
        bt = torch.randn(3, device=dev)
        eps = torch.empty(N, 3, H, W, device=dev)
        ops.conv(a, None, 1, wt, bt, None, None, eps, 1, N, H, W, 3, out_nchw=True)
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt.float().view(3, 3, 3, C).permute(0, 3, 1, 2), bt, padding=1)
        assert _rel(eps, ref) < tol
        dy = torch.randn(N, 3, H, W, device=dev)
        dw = torch.empty(3 * 9 * C, device=dev)
        ops.wgrad(a, None, 1, dy, 1, dw, N, H, W, 3, dtype, dy_nchw=True)
        ref = torch.nn.grad.conv2d_weight(a.float().permute(0, 3, 1, 2), (3, C, 3, 3), dy, padding=1).permute(0, 2, 3, 1).reshape(-1)
        assert _rel(dw, ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 16 * 16, 64, 0), (2, 8 * 8, 128, 64), (1, 32 * 32, 192, 0), (3, 100, 32, 0),
                                   (2, 16 * 16, 512, 512), (3, 777, 256, 0), (5, 7, 64, 32)])
def test_groupnorm_swish_forward_backward(dtype, tol, shape):
    N, HW, C0, C1 = shape
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    torch.manual_seed(1)
    C = C0 + C1
    x0 = (torch.randn(N, HW, 1, C0, device=dev) * 2 + 0.5).to(dtype)
    x1 = (torch.randn(N, HW, 1, C1, device=dev) - 1).to(dtype) if C1 else None
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
    ops.gn_stats(x0, x1, N, HW, 32, sums)
    rs = torch.empty_like(sums)
    emu.gn_stats(x0, x1, N, HW, 32, rs)
    assert _rel(sums, rs) < 1e-6
    for act in (0, 1):
        out = torch.empty(N, HW, 1, C, dtype=dtype, device=dev)
        ops.gn_apply(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, act, 0.0, 0, out)
        ref = torch.empty(N, HW, 1, C, device=dev)
        emu.gn_apply(x0.float(), None if x1 is None else x1.float(), N, HW, 32, rs, gamma, beta, 1e-5, act, 0.0, 0, ref)
        assert _rel(out.float(), ref) < tol
        dy = torch.randn(N, HW, 1, C, device=dev).to(dtype)
        add = torch.randn(N, HW, 1, C, device=dev).to(dtype)
        acc0 = torch.randn(N, HW, 1, C0, device=dev).to(dtype)
        gs = torch.empty_like(sums)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        dx0 = torch.empty_like(x0)
        dx1 = None if x1 is None else torch.empty_like(x1)
        ops.gn_bwd(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, act, 0.0, 0, dy, gs, dg, db, add, acc0, None, dx0, dx1)
        rdg, rdb = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        r0 = torch.empty(N, HW, 1, C0, device=dev)
        r1 = None if x1 is None else torch.empty(N, HW, 1, C1, device=dev)
        emu.gn_bwd(x0.float(), None if x1 is None else x1.float(), N, HW, 32, rs, gamma, beta, 1e-5, act, 0.0, 0, dy.float(),
                   None, rdg, rdb, add.float(), acc0.float(), None, r0, r1)
        assert _rel(dx0.float(), r0) < tol and _rel(dg, rdg) < tol and _rel(db, rdb) < tol
        if x1 is not None:
            assert _rel(dx1.float(), r1) < tol


def test_dropout_stream_is_consistent_between_forward_and_backward():
    dev = torch.device("cuda")
    ops = _ops()
    N, HW, C = 2, 1024, 64
    torch.manual_seed(2)
    x = torch.randn(N, HW, 1, C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
    ops.gn_stats(x, None, N, HW, 32, sums)
    y0 = torch.empty_like(x)
    y1 = torch.empty_like(x)
    ops.gn_apply(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 0, 0.0, 0, y0)
    ops.gn_apply(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 0, 0.25, 1234, y1)
    keep = (y1 != 0)
    frac = float(keep.float().mean())
    assert abs(frac - 0.75) < 0.01                                        # drop rate
    assert torch.allclose(y1[keep], y0[keep] / 0.75, rtol=1e-5, atol=1e-6)  # inverted-dropout scaling
    # backward regenerates the same mask: gradient is zero exactly where the forward dropped
    dy = torch.ones_like(x)
    gs = torch.empty_like(sums)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx = torch.empty_like(x)
    ops.gn_bwd(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 0, 0.25, 1234, dy, gs, dg, db, None, None, None, dx, None)
    assert abs(float(db.sum()) - float(keep.float().sum()) / 0.75) < 1e-2 * float(keep.float().sum())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("S,C", [(64, 64), (256, 128), (1024, 128)])
def test_attention_forward_backward(dtype, tol, S, C):
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    torch.manual_seed(3)
    N = 2
    qkv = torch.randn(N, S, 3 * C, device=dev).to(dtype)
    out = torch.empty(N, S, C, dtype=dtype, device=dev)
    lse = torch.empty(N, S, device=dev)
    ops.attn_fwd(qkv, out, lse, N, S, C)
    ro, rl = torch.empty(N, S, C, device=dev), torch.empty(N, S, device=dev)
    emu.attn_fwd(qkv.float(), ro, rl, N, S, C)
    assert _rel(out.float(), ro) < tol and _rel(lse, rl) < 1e-3
    dout = torch.randn(N, S, C, device=dev).to(dtype)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(N, S, device=dev)
    ops.attn_bwd(qkv, out, dout, lse, delta, dqkv, N, S, C)
    rd = torch.empty(N, S, 3 * C, device=dev)
    emu.attn_bwd(qkv.float(), ro, dout.float(), rl, None, rd, N, S, C)
    assert _rel(dqkv.float(), rd) < tol, _rel(dqkv.float(), rd)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("S,C,heads", [(64, 32, 8), (256, 64, 8), (1024, 128, 8), (200, 256, 8), (100, 512, 8), (77, 64, 4),
                                       (1024, 256, 8), (512, 512, 8), (128, 256, 4)])
def test_multihead_attention_forward_backward(dtype, tol, S, C, heads):
    """hd_mha_*: the attention core of nn.MultiheadAttention(C, heads) with q = k = v (ModelCondition.py:189,203-208), against
    softmax(q_h k_h^T / sqrt(hd)) v_h per head in fp32; sequence lengths that are not multiples of the kernel's tiles included.
    bf16 with S % 128 == 0 and head dim 8..64 takes the tensor-core route (heads zero-padded onto the tcgen05 kernels), the rest
    the CUDA-core kernels."""
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    torch.manual_seed(S + C)
    N = 2
    qkv = (torch.randn(N, S, 3 * C, device=dev) * 1.5).to(dtype)
    out = torch.empty(N, S, C, dtype=dtype, device=dev)
    lse = torch.empty(N, heads, S, device=dev)
    ops.mha_fwd(qkv, out, lse, N, S, C, heads)
    ro, rl = torch.empty(N, S, C, device=dev), torch.empty(N, heads, S, device=dev)
    emu.mha_fwd(qkv.float(), ro, rl, N, S, C, heads)
    assert _rel(out.float(), ro) < tol and _rel(lse, rl) < 1e-4, (_rel(out.float(), ro), _rel(lse, rl))
    dout = torch.randn(N, S, C, device=dev).to(dtype)
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(N, heads, S, device=dev)
    ops.mha_bwd(qkv, out, dout, lse, delta, dqkv, N, S, C, heads)
    rd = torch.empty(N, S, 3 * C, device=dev)
    emu.mha_bwd(qkv.float(), ro, dout.float(), rl, None, rd, N, S, C, heads)
    assert _rel(dqkv.float(), rd) < tol, _rel(dqkv.float(), rd)


@pytest.mark.parametrize("S,C,heads", [(1024, 256, 8), (256, 128, 8), (384, 512, 8)])
def test_multihead_attention_tensor_core_route_matches_cuda_core_kernels(S, C, heads):
    """The two bf16 routes of ops.mha_* (padded heads on tcgen05 vs hd_mha_* on CUDA cores) on the same inputs; pack / unpack
    round trip is the identity on the live channels and writes zeros elsewhere."""
    dev = torch.device("cuda")
    ops = _ops()
    assert ops.mha_tc and ops._mha_tc(torch.empty(1, dtype=torch.bfloat16, device=dev), S, C, heads)
    torch.manual_seed(S * 3 + C)
    N, bf = 3, torch.bfloat16
    qkv = (torch.randn(N, S, 3 * C, device=dev) * 1.5).to(bf)
    pad = torch.full((N * heads, S, 384), float("nan"), dtype=bf, device=dev)
    back = torch.full_like(qkv, float("nan"))
    assert ops.lib.hd_mha_pack_heads(qkv.data_ptr(), pad.data_ptr(), N, S, C, heads, 3, 1.0, None) == 0
    assert ops.lib.hd_mha_unpack_heads(pad.data_ptr(), back.data_ptr(), N, S, C, heads, 3, 1.0, None) == 0
    assert torch.equal(back, qkv)
    hd = C // heads
    pv = pad.view(N, heads, S, 3, 128)
    assert torch.equal(pv[..., :hd], qkv.view(N, S, 3, heads, hd).permute(0, 3, 1, 2, 4)) and bool((pv[..., hd:] == 0).all())
    dout = torch.randn(N, S, C, device=dev).to(bf)
    res = []
    for tc in (True, False):
        ops.mha_tc = tc
        try:
            out = torch.empty(N, S, C, dtype=bf, device=dev)
            lse = torch.empty(N, heads, S, device=dev)
            dqkv = torch.full_like(qkv, float("nan"))
            delta = torch.empty(N, heads, S, device=dev)
            before = ops.tc_launches
            ops.mha_fwd(qkv, out, lse, N, S, C, heads)
            ops.mha_bwd(qkv, out, dout, lse, delta, dqkv, N, S, C, heads)
            assert (ops.tc_launches - before == 3) == tc
            res.append((out.float(), lse, dqkv.float()))
        finally:
            ops.mha_tc = True
    assert _rel(res[0][0], res[1][0]) < 1e-2 and _rel(res[0][1], res[1][1]) < 1e-4 and _rel(res[0][2], res[1][2]) < 1.5e-2, \
        [_rel(a, b) for a, b in zip(res[0], res[1])]


@pytest.mark.parametrize("S,N,qscale", [(128, 2, 1.0), (256, 3, 1.0), (1024, 2, 1.0), (4096, 1, 1.0), (1024, 2, 6.0), (2048, 1, 12.0)])
def test_attention_tcgen05_forward(S, N, qscale):
    """Flash-style tcgen05 forward against softmax(q k^T C^-1/2) v in fp32; `qscale` sharpens the scores so that the
    running maximum moves by more than the lazy-rescale threshold (the O-rescale path runs)."""
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    C = 128
    assert ops.lib.hd_attn_tc_supported(S, C)
    torch.manual_seed(S + int(qscale))
    qkv = torch.randn(N, S, 3 * C, device=dev)
    qkv[:, :, :C] *= qscale
    # ascending key norms along the sequence make later tiles raise the maximum
    qkv[:, :, C:2 * C] *= torch.linspace(0.5, 1.5, S, device=dev)[None, :, None]
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty(N, S, C, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(N, S, device=dev)
    before = ops.tc_launches
    ops.attn_fwd(qkv, out, lse, N, S, C)
    torch.cuda.synchronize()
    assert ops.tc_launches == before + 1
    ro, rl = torch.empty(N, S, C, device=dev), torch.empty(N, S, device=dev)
    emu.attn_fwd(qkv.float(), ro, rl, N, S, C)
    assert _rel(out.float(), ro) < 1e-2, _rel(out.float(), ro)           # bf16 tolerance (north_star)
    assert float((lse - rl).abs().max()) < 2e-2 * max(1.0, float(rl.abs().max()) * 0.05), float((lse - rl).abs().max())


@pytest.mark.parametrize("S,N,qscale", [(128, 2, 1.0), (256, 3, 1.0), (1024, 2, 1.0), (4096, 1, 1.0), (1024, 2, 4.0)])
def test_attention_tcgen05_backward(S, N, qscale):
    """dQ / dK / dV of the tcgen05 backward (two deterministic passes, no atomics) against autograd of the fp32 formula."""
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    C = 128
    assert ops.lib.hd_attn_bwd_tc_supported(S, C)
    torch.manual_seed(7 * S + int(qscale))
    qkv = torch.randn(N, S, 3 * C, device=dev)
    qkv[:, :, :C] *= qscale
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty(N, S, C, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(N, S, device=dev)
    ops.attn_fwd(qkv, out, lse, N, S, C)
    dout = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    dqkv = torch.full_like(qkv, float("nan"))
    before = ops.tc_launches
    ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C)
    torch.cuda.synchronize()
    assert ops.tc_launches == before + 2
    rd = torch.empty(N, S, 3 * C, device=dev)
    emu.attn_bwd(qkv.float(), None, dout.float(), None, None, rd, N, S, C)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        r = _rel(dqkv[:, :, sl].float(), rd[:, :, sl])
        assert r < 1.5e-2, (name, r)


@pytest.mark.parametrize("S,N,C,qscale", [(256, 2, 256, 1.0), (1024, 2, 256, 1.0), (4096, 1, 256, 1.0), (1024, 2, 512, 1.0), (256, 1, 384, 1.0),
                                          (1024, 2, 256, 8.0), (512, 3, 128, 1.0), (1024, 1, 1024, 1.0)])
def test_attention_tcgen05_wide_heads(S, N, C, qscale):
    """Wide-head tcgen05 attention (hd_attn_wide_tc.cu: C = 256 / 512 levels of BASELINE.json configs[4]) forward and backward
    against the fp32 formula.  C = 128 goes through the wide entry points directly (the dispatcher keeps C = 128 on the
    dedicated kernels)."""
    import ctypes
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    assert ops.lib.hd_attn_wide_tc_supported(S, C) and ops.lib.hd_attn_tc_supported(S, C) and ops.lib.hd_attn_bwd_tc_supported(S, C)
    torch.manual_seed(11 * S + C + int(qscale))
    qkv = torch.randn(N, S, 3 * C, device=dev)
    qkv[:, :, :C] *= qscale
    qkv[:, :, C:2 * C] *= torch.linspace(0.5, 1.5, S, device=dev)[None, :, None]
    qkv = qkv.to(torch.bfloat16)
    out = torch.full((N, S, C), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.full((N, S), float("nan"), device=dev)
    dout = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    dqkv = torch.full_like(qkv, float("nan"))
    if C == 128:
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        stats = torch.empty(N, S, 2, device=dev)
        assert ops.lib.hd_attn_fwd_wide_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), N, S, C, st) == 0
        assert ops.lib.hd_attn_bwd_wide_tc(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), stats.data_ptr(),
                                           dqkv.data_ptr(), N, S, C, st) == 0
    else:
        before = ops.tc_launches
        ops.attn_fwd(qkv, out, lse, N, S, C)
        ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C)
        assert ops.tc_launches == before + 3
    torch.cuda.synchronize()
    ro, rl = torch.empty(N, S, C, device=dev), torch.empty(N, S, device=dev)
    emu.attn_fwd(qkv.float(), ro, rl, N, S, C)
    assert _rel(out.float(), ro) < 1e-2, _rel(out.float(), ro)
    assert float((lse - rl).abs().max()) < 2e-2 * max(1.0, float(rl.abs().max()) * 0.05), float((lse - rl).abs().max())
    rd = torch.empty(N, S, 3 * C, device=dev)
    emu.attn_bwd(qkv.float(), None, dout.float(), None, None, rd, N, S, C)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        r = _rel(dqkv[:, :, sl].float(), rd[:, :, sl])
        assert r < 1.5e-2, (name, r)


@pytest.mark.parametrize("env", [{"HDIFF_CONV_MMA2": "0", "HDIFF_CONV_STAGE": "0"}, {"HDIFF_CONV_MMA2": "1"}, {"HDIFF_CONV_MMA2": "2"},
                                 {"HDIFF_CONV_TXM_OFF": "1"}, {"HDIFF_CONV_WRES": "1"}, {"HDIFF_WGRAD_HALO_OFF": "1"}],
                         ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_conv_alternate_kernel_modes(env):
    """The kernel modes that are not the default for a shape (one / two MMA issuers, direct-store epilogue, no shifted
    operands, resident weights, classic wgrad) are chosen through environment switches read once per process: run the
    convolution / wgrad parity cases again in a child process under each of them."""
    import os
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "(test_conv_forward or wgrad) and bf16_tc"], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_groupnorm_backward_in_place_dy():
    """ops.gn_bwd(overwrite_dy=True): with the first-generation kernels (HDIFF_GN_V1=1) the reduce pass leaves
    dy * mask * act'(z) in dy and the apply pass consumes it; same results as the two independent passes up to the bf16
    rounding of that intermediate."""
    dev = torch.device("cuda")
    ops = _ops()
    torch.manual_seed(9)
    N, HW, C = 3, 24 * 24, 128
    x = (torch.randn(N, HW, 1, C, device=dev) * 1.5 + 0.3).to(torch.bfloat16)
    dy = torch.randn(N, HW, 1, C, device=dev).to(torch.bfloat16)
    add = torch.randn(N, HW, 1, C, device=dev).to(torch.bfloat16)
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
    ops.gn_stats(x, None, N, HW, 32, sums)
    res = []
    for inplace in (False, True):
        d = dy.clone()
        gs = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        dx = torch.empty_like(x)
        cs = torch.zeros(C, device=dev)
        ops.gn_bwd(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 1, 0.1, 77, d, gs, dg, db, add, None, None, dx, None, cs_total=cs,
                   overwrite_dy=inplace)
        res.append((dx.float(), dg, db, gs, cs, d))
    if ops.lib.hd_gn_v2(1, C, 0, 32, HW, N):
        # second-generation kernels (csrc/hd_gn.cu) are HBM-bound and never rewrite dy: the flag is accepted and ignored
        assert torch.equal(res[1][5], dy) and torch.equal(res[0][5], dy)
    else:
        assert not torch.equal(res[1][5], dy) and torch.equal(res[0][5], dy)      # the second call did overwrite its dy
    assert _rel(res[1][0], res[0][0]) < 4e-3
    for i in (1, 2, 3):
        assert torch.allclose(res[1][i].double(), res[0][i].double(), rtol=1e-4, atol=1e-4)       # the sums do not see the rounding (fp32 atomics: order)
    assert _rel(res[1][4], res[0][4]) < 2e-3


def test_embedding_path_and_packing_kernels():
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    torch.manual_seed(4)
    M, K, Nn = 5, 96, 40
    x, w, b = torch.randn(M, K, device=dev), torch.randn(Nn, K, device=dev), torch.randn(Nn, device=dev)
    for sw in (False, True):
        y, r = torch.zeros(M, Nn, device=dev), torch.zeros(M, Nn, device=dev)
        ops.linear_fwd(x, w, b, y, in_swish=sw)
        emu.linear_fwd(x, w, b, r, in_swish=sw)
        assert _rel(y, r) < 1e-5
        dy = torch.randn(M, Nn, device=dev)
        dx, rdx = torch.empty(M, K, device=dev), torch.empty(M, K, device=dev)
        ops.linear_bwd_x(dy, w, x if sw else None, dx)
        emu.linear_bwd_x(dy, w, x if sw else None, rdx)
        assert _rel(dx, rdx) < 1e-5
        dw, db, rdw, rdb = (torch.zeros(Nn * K, device=dev), torch.zeros(Nn, device=dev),
                            torch.zeros(Nn * K, device=dev), torch.zeros(Nn, device=dev))
        ops.linear_bwd_w(dy, x, dw, db, in_swish=sw)
        emu.linear_bwd_w(dy, x, rdw, rdb, in_swish=sw)
        assert _rel(dw, rdw) < 1e-5 and _rel(db, rdb) < 1e-5
    table = torch.randn(11, 32, device=dev)
    idx = torch.tensor([0, 3, 3, 10, 0], device=dev)
    out = torch.empty(5, 32, device=dev)
    ops.embedding_fwd(table, idx, out)
    assert torch.equal(out, table[idx])
    dt, rdt = torch.zeros(11 * 32, device=dev), torch.zeros(11 * 32, device=dev)
    ops.embedding_bwd(out, idx, dt, padding_idx=0)
    emu.embedding_bwd(out, idx, rdt, padding_idx=0)
    assert _rel(dt, rdt) < 1e-6 and float(dt[:32].abs().max()) == 0
    src = torch.randn(1000, device=dev)
    ia = torch.randint(-1, 1000, (777,), device=dev, dtype=torch.int32)
    ib = torch.randint(-1, 1000, (777,), device=dev, dtype=torch.int32)
    for dt_ in (torch.float32, torch.bfloat16):
        o, r = torch.empty(777, dtype=dt_, device=dev), torch.empty(777, dtype=dt_, device=dev)
        ops.gather_pack(src, ia, ib, o)
        emu.gather_pack(src, ia, ib, r)
        assert torch.equal(o, r)
    inv = torch.randint(-1, 777, (1000,), device=dev, dtype=torch.int32)
    packed = torch.randn(777, device=dev)
    d, r = torch.ones(1000, device=dev), torch.ones(1000, device=dev)
    ops.scatter_unpack(packed, inv, d)
    emu.scatter_unpack(packed, inv, r)
    assert torch.equal(d, r)


def test_diffusion_elementwise_kernels():
    dev = torch.device("cuda")
    ops, emu = _ops(), EmuOps()
    torch.manual_seed(5)
    N = 3
    x0, nz = torch.randn(N, 3, 16, 16, device=dev), torch.randn(N, 3, 16, 16, device=dev)
    t = torch.tensor([0, 500, 999], device=dev)
    sab, s1 = torch.rand(1000, device=dev), torch.rand(1000, device=dev)
    xt, r = torch.empty_like(x0), torch.empty_like(x0)
    ops.q_sample(x0, nz, t, sab, s1, xt)
    emu.q_sample(x0, nz, t, sab, s1, r)
    assert torch.allclose(xt, r, rtol=1e-6, atol=1e-6)
    loss, rl = torch.empty_like(x0), torch.empty_like(x0)
    ops.mse_fwd(x0, nz, loss)
    emu.mse_fwd(x0, nz, rl)
    assert torch.allclose(loss, rl, rtol=1e-6, atol=1e-6)
    g = torch.randn_like(x0)
    d, rd = torch.empty_like(x0), torch.empty_like(x0)
    ops.mse_bwd(x0, nz, g, d)
    emu.mse_bwd(x0, nz, g, rd)
    assert torch.allclose(d, rd, rtol=1e-6, atol=1e-6)
    coef = torch.rand(10, 3, device=dev)
    for s in (7, 0):
        x, rx = x0.clone(), x0.clone()
        step = torch.tensor([s], dtype=torch.int32, device=dev)
        flag, rflag = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
        ops.sampler_step(x, nz, g, xt, 1.8, coef, step, True, flag)
        emu.sampler_step(rx, nz, g, xt, 1.8, coef, step, True, rflag)
        assert torch.allclose(x, rx, rtol=1e-5, atol=1e-5) and int(flag) == 0
    x = x0.clone()
    x[0, 0, 0, 0] = float("nan")
    ops.sampler_step(x, nz, None, xt, 0.0, coef, step, True, flag)
    assert int(flag) == 1
    ops.add_int(step, -1)
    assert int(step) == -1
    # clip + AdamW against torch
    p = torch.randn(4096, device=dev)
    gr = torch.randn(4096, device=dev) * 3
    tp = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([tp], lr=1e-3, weight_decay=1e-2)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    for step_i in (1, 2, 3):
        tp.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([tp], 1.0)
        opt.step()
        g2 = gr.clone()
        ops.sqnorm(g2, sq)
        ops.adamw_flat(p, g2, m, v, sq, 1.0, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step_i)
        assert torch.allclose(p, tp.detach(), rtol=1e-5, atol=1e-6)
