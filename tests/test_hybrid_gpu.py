"""GPU parity of the §8(f) rows on the CUDA library, against the REFERENCE'S OWN classes (imported from oracle/_ref on the GPU
box): the live MHA UNet (ModelCondition.py:213-276), DynamicUNet (diffusion/Model.py:382-517) and the hybrid sampler
(diffusion/Diffusion.py:182-269).  fp32 check mode pins the algorithm (1e-4 class), bf16 is the product path (<= 3e-2 on
whole-network outputs and gradients, as for the benchmarked UNet)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")]


@pytest.fixture(autouse=True)
def _setup():
    import hdiff_b200.ops as hops
    hops.set_backend(None)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _close(a, b, rtol, atol):
    return float((a.double() - b.double()).norm()) <= rtol * float(b.double().norm()) + atol


def _check_grads(net, ref, tol, yard=None):
    """Per-parameter gradient check against the fp32 reference.  `yard` (bf16 only): the gradients of the SAME reference module
    under torch.autocast(bfloat16) -- stock PyTorch's own bf16 noise on this input.  Some DynamicUNet parameters (the 8x8 middle
    blocks behind four attention layers) sit at 11-13 % in BOTH bf16 evaluations (tests/tools/debug_hybrid_bf16_noise.py), so a
    parameter passes at 4 tol or at 1.5x the yardstick's own error, and the median over parameters must not exceed the
    yardstick's median by more than 20 %."""
    pr = dict(ref.named_parameters())
    gscale = max(float(p.grad.norm()) for p in pr.values() if p.grad is not None)
    worst, n, rels, yrels = ("", 0.0), 0, [], []
    for k, p in net.named_parameters():
        if pr[k].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        n += 1
        assert p.grad is not None, k
        r = _rel(p.grad, pr[k].grad)
        rtol = 4 * tol
        if yard is not None:
            yr = _rel(yard[k], pr[k].grad)
            rels.append(r)
            yrels.append(yr)
            rtol = max(rtol, 1.5 * yr)
        if not _close(p.grad, pr[k].grad, rtol, 2e-3 * tol * gscale + 2e-5) and r > worst[1]:
            worst = (k, r)
    assert worst[0] == "", worst
    if yard is not None:
        med, ymed = sorted(rels)[len(rels) // 2], sorted(yrels)[len(yrels) // 2]
        assert med <= 1.2 * ymed + 1e-3, (med, ymed)
    return n


def _autocast_grads(ref, fwd, gy):
    """Gradients of the reference module under torch.autocast(bfloat16); leaves ref's .grad cleared."""
    for p in ref.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        e = fwd()
    e.float().backward(gy)
    g = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
    for p in ref.parameters():
        p.grad = None
    return g


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_live_mha_unet_vs_reference(dtype, tol):
    """ch = 64 at 32 x 32: 8-head attention with head dims 8 (S = 1024) and 16 (S = 256, 64), packed projections on the conv
    kernels (tcgen05 in bf16)."""
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    dev = torch.device("cuda")
    mc = ref_loader.model_condition()
    cfg = dict(T=1000, num_labels=10, ch=64, ch_mult=[1, 2, 2], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(31)
    ref = mc.UNet(**cfg).to(dev)
    net = UNet(compute_dtype=dtype, mha=True, **cfg)
    net.load_state_dict(ref.state_dict())
    net.to(dev)
    x = torch.rand(2, 3, 32, 32, device=dev) * 2 - 1
    t = torch.tensor([5, 700], device=dev)
    lab = torch.tensor([0, 7], device=dev)
    e, er = net(x, t, lab), ref(x, t, lab)
    assert _rel(e.detach(), er.detach()) < tol, _rel(e.detach(), er.detach())
    gy = torch.randn_like(er)
    e.backward(gy)
    er.backward(gy)
    assert _check_grads(net, ref, tol) > 80


@pytest.mark.parametrize("context_zero", [True, False], ids=["context_zero", "image_condition"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_dynamic_unet_vs_reference(dtype, tol, context_zero):
    from hdiff_b200.diffusion.Model import DynamicUNet
    dev = torch.device("cuda")
    dm = ref_loader.diffusion_model()
    cfg = dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(32)
    ref = dm.DynamicUNet(**cfg).to(dev)
    with torch.no_grad():
        ref.tail[-1].weight.mul_(3e4)
    net = DynamicUNet(compute_dtype=dtype, **cfg)
    net.load_state_dict(ref.state_dict())
    net.to(dev)
    x = torch.rand(2, 6, 64, 64, device=dev) * 2 - 1
    x[:, 2] += 0.3
    t = torch.tensor([9, 600], device=dev)
    lab = torch.rand(2, 3, 64, 64, device=dev) * 2 - 1
    e, er = net(x, t, lab, context_zero=context_zero), ref(x, t, lab, context_zero=context_zero)
    assert _rel(e.detach(), er.detach()) < tol, _rel(e.detach(), er.detach())
    assert [p.requires_grad for p in net.parameters()] == [p.requires_grad for p in ref.parameters()]
    gy = torch.randn_like(er)
    yard = _autocast_grads(ref, lambda: ref(x, t, lab, context_zero=context_zero), gy) if dtype == torch.bfloat16 else None
    e.backward(gy)
    er.backward(gy)
    assert _check_grads(net, ref, tol, yard) > 100


def test_hybrid_sampler_graph_replay_vs_reference():
    """30-step ancestral chain and 10-step DDIM chain of the hybrid sampler in fp32 check mode (CUDA-graph replay of the step,
    6-channel input assembled inside the captured region) against the reference class on the same seeds."""
    from hdiff_b200.diffusion.Model import DynamicUNet
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionSampler
    dev = torch.device("cuda")
    dm = ref_loader.diffusion_model()
    RefSampler = ref_loader.hybrid_sampler_class()
    for ddim, T in ((False, 30), (True, 1000)):
        cfg = dict(T=T, ch=32, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
        torch.manual_seed(33)
        ref = dm.DynamicUNet(**cfg).to(dev).eval()
        with torch.no_grad():
            ref.tail[-1].weight.mul_(3e4)
        net = DynamicUNet(compute_dtype=torch.float32, **cfg)
        net.load_state_dict(ref.state_dict())
        net.to(dev).eval()
        img = torch.randint(0, 256, (2, 3, 32, 32), device=dev, dtype=torch.uint8)
        kw = dict(ddim=True, unconditional_guidance_scale=2.0, ddim_step=10) if ddim else {}
        torch.manual_seed(3)
        with torch.no_grad():
            want = RefSampler(ref, 1e-4, 0.02, T).to(dev)(img, **kw)
        torch.manual_seed(3)
        got = GaussianDiffusionSampler(net, 1e-4, 0.02, T).to(dev)(img, **kw)
        assert float((got - want).abs().max()) < 5e-3, (ddim, float((got - want).abs().max()))


def test_nearest_upsample_and_image_affine_kernels():
    import hdiff_b200.ops as hops
    ops = hops.get()
    dev = torch.device("cuda")
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(3, 5, 7, 64, device=dev).to(dtype)
        out = torch.empty(3, 10, 21, 64, dtype=dtype, device=dev)
        ops.upsample_nearest(x, out)
        assert torch.equal(out, x.repeat_interleave(2, 1).repeat_interleave(3, 2))
        dout = torch.randn_like(out)
        din = torch.empty_like(x)
        ops.upsample_nearest_bwd(dout, din)
        want = dout.float().view(3, 5, 2, 7, 3, 64).sum((2, 4))
        assert _rel(din.float(), want) < (1e-6 if dtype == torch.float32 else 4e-3)
    u8 = torch.randint(0, 256, (2, 3, 9, 11), device=dev, dtype=torch.uint8)
    o = torch.empty(u8.shape, device=dev)
    ops.image_affine(u8, o, 2.0 / 255.0, -1.0)
    assert torch.allclose(o, (u8.float() / 255.0) * 2 - 1, atol=1e-6)
