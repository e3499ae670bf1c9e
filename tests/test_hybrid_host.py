"""CPU tests (operator test double) of the §8(f) rows against the REFERENCE'S OWN classes imported by path
(oracle/ref_loader.py): the live MHA UNet of DiffusionFreeGuidence/ModelCondition.py:213-276 and the hybrid pipeline's
DynamicUNet of diffusion/Model.py:382-517.  Skipped where the reference sources are absent (neither /root/reference nor
oracle/_ref)."""
import pytest
import torch

import hdiff_b200.ops as hops
from oracle import ref_loader
from tests.emu_backend import EmuOps

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


@pytest.fixture(autouse=True)
def emu():
    prev = hops._backend
    hops.set_backend(EmuOps())
    yield
    hops.set_backend(prev)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _close(a, b, rtol, atol=2e-5):
    return float((a - b).norm()) <= rtol * float(b.norm()) + atol


def _compare_grads(net, ref, rtol=3e-4):
    """every parameter gradient; the absolute floor (gradients that are mathematically zero: a conv bias feeding a GroupNorm)
    scales with the largest gradient of the model"""
    pr = dict(ref.named_parameters())
    atol = 1e-5 * max(float(p.grad.norm()) for p in pr.values() if p.grad is not None)
    checked = 0
    for k, p in net.named_parameters():
        rg = pr[k].grad
        if rg is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        assert _close(p.grad, rg, rtol, atol), (k, _rel(p.grad, rg))
        checked += 1
    return checked


def test_live_mha_unet_matches_the_reference_class():
    """ModelCondition.UNet as the reference file defines it: nn.MultiheadAttention(C, 8) in every down ResBlock and the first
    middle block, output replaces h (ModelCondition.py:166-211,226,234)."""
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    mc = ref_loader.model_condition()
    cfg = dict(T=50, num_labels=5, ch=32, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(11)
    ref = mc.UNet(**cfg)
    net = UNet(compute_dtype=torch.float32, mha=True, **cfg)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    net.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 8, 8)
    t = torch.tensor([3, 41])
    lab = torch.tensor([2, 0])
    e_ref = ref(x, t, lab)
    e = net(x, t, lab)
    assert _rel(e.detach(), e_ref.detach()) < 2e-5, _rel(e.detach(), e_ref.detach())
    gy = torch.randn_like(e_ref)
    e_ref.backward(gy)
    e.backward(gy)
    assert _compare_grads(net, ref) > 50


def test_mha_checkpoint_is_refused_by_the_attnblock_model_and_loaded_by_the_mha_model():
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    mc = ref_loader.model_condition()
    cfg = dict(T=20, num_labels=3, ch=32, ch_mult=[1, 1], num_res_blocks=1, dropout=0.0)
    sd = mc.UNet(**cfg).state_dict()
    with pytest.raises(RuntimeError, match="mha=True"):
        UNet(compute_dtype=torch.float32, **cfg).load_state_dict(sd, strict=False)
    UNet(compute_dtype=torch.float32, mha=True, **cfg).load_state_dict(sd)


@pytest.mark.parametrize("context_zero", [True, False], ids=["context_zero", "image_condition"])
def test_dynamic_unet_matches_the_reference_class(context_zero):
    """diffusion/Model.py DynamicUNet: 6-channel head, MHA middle blocks, nearest-resized skips, red / blue gate, and with
    context_zero=False the image condition encoder (three stride-2 convolutions + pool + MLP)."""
    from hdiff_b200.diffusion.Model import DynamicUNet
    dm = ref_loader.diffusion_model()
    cfg = dict(T=50, ch=32, ch_mult=[1, 2, 2], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(12)
    ref = dm.DynamicUNet(**cfg)
    net = DynamicUNet(compute_dtype=torch.float32, **cfg)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    # same init recipe (xavier head, 1e-5-gain tail): statistics, then take the reference's values
    assert float(net.tail[-1].weight.abs().max()) < 1e-4 and float(net.head.bias.abs().max()) == 0.0
    net.load_state_dict(ref.state_dict())
    with torch.no_grad():                                     # a tail that is not ~0, so that gradients are informative
        ref.tail[-1].weight.mul_(1e5)
        net.tail[-1].weight.mul_(1e5)
    x = torch.randn(2, 6, 16, 16)
    x[:, 2] += 0.5                                             # blue > red: "subaquatic" (even middle blocks train)
    t = torch.tensor([3, 41])
    lab = torch.randn(2, 3, 16, 16)
    e_ref = ref(x, t, lab, context_zero=context_zero)
    e = net(x, t, lab, context_zero=context_zero)
    assert _rel(e.detach(), e_ref.detach()) < 3e-5, _rel(e.detach(), e_ref.detach())
    assert [p.requires_grad for p in net.parameters()] == [p.requires_grad for p in ref.parameters()]
    frozen = [all(not p.requires_grad for p in b.parameters()) for b in net.middleblocks]
    assert frozen == [False, True, False, True]
    gy = torch.randn_like(e_ref)
    e_ref.backward(gy)
    e.backward(gy)
    assert _compare_grads(net, ref) > 60
    # the other side of the gate
    x2 = x.clone()
    x2[:, 0] += 2.0
    net.zero_grad(); ref.zero_grad()
    net(x2, t, lab, context_zero=context_zero).backward(gy)
    ref(x2, t, lab, context_zero=context_zero).backward(gy)
    frozen = [all(not p.requires_grad for p in b.parameters()) for b in net.middleblocks]
    assert frozen == [True, False, True, False]
    _compare_grads(net, ref)


def _dynamic_pair(seed=21, T=50):
    from hdiff_b200.diffusion.Model import DynamicUNet
    dm = ref_loader.diffusion_model()
    cfg = dict(T=T, ch=32, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(seed)
    ref = dm.DynamicUNet(**cfg)
    with torch.no_grad():
        ref.tail[-1].weight.mul_(3e4)              # the reference initialises the tail at gain 1e-5: make eps matter
    net = DynamicUNet(compute_dtype=torch.float32, **cfg)
    net.load_state_dict(ref.state_dict())
    return net, ref


def test_hybrid_trainer_matches_the_restated_reference_forward():
    """diffusion/Diffusion.py:54-96 around the reference's own DynamicUNet: uint8 images, same seeds, loss and every gradient."""
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer
    from oracle import ref_torch as R
    net, ref = _dynamic_pair()
    tr = GaussianDiffusionTrainer(net, 1e-4, 0.02, 50)
    g = torch.Generator().manual_seed(4)
    gt = torch.randint(0, 256, (2, 3, 16, 16), generator=g, dtype=torch.uint8)
    inp = torch.randint(0, 256, (2, 3, 16, 16), generator=g, dtype=torch.uint8)
    torch.manual_seed(7)
    out = tr(gt, inp, 0)
    assert len(out) == 5 and out[2] == 0 and out[3] == 0 and out[4] == 0
    torch.manual_seed(7)
    mse_ref, _ = R.hybrid_trainer_forward(ref, tr.sqrt_alphas_bar, tr.sqrt_one_minus_alphas_bar, 50, gt, inp)
    assert _rel(out[1].detach(), mse_ref.detach()) < 3e-5
    out[0].sum().backward()
    mse_ref.sum().backward()
    assert _compare_grads(net, ref) > 40
    # the plain signature of the same class still works (diffusion/Train.py:41,51)
    from hdiff_b200.diffusion.Model import UNet
    plain = GaussianDiffusionTrainer(UNet(T=50, ch=32, ch_mult=[1, 1], attn=[], num_res_blocks=1, dropout=0.0, compute_dtype=torch.float32),
                                     1e-4, 0.02, 50)
    assert plain(torch.rand(2, 3, 8, 8) * 2 - 1).shape == (2, 3, 8, 8)


@pytest.mark.parametrize("ddim", [False, True], ids=["ancestral", "ddim"])
def test_hybrid_sampler_matches_the_reference_class(ddim):
    """The reference's GaussianDiffusionSampler (diffusion/Diffusion.py:182-269, cut out of its module by AST) around its own
    DynamicUNet against ours on the same seeds: ancestral chain, and the DDIM chain with a guidance scale != 1."""
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionSampler
    T = 1000 if ddim else 12                       # the reference's DDIM branch hard-codes 1000 (:246-247)
    net, ref = _dynamic_pair(seed=22, T=T)
    net.eval(); ref.eval()
    RefSampler = ref_loader.hybrid_sampler_class()
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (2, 3, 16, 16), generator=g, dtype=torch.uint8)
    kw = dict(ddim=True, unconditional_guidance_scale=2.5, ddim_step=4) if ddim else {}
    torch.manual_seed(9)
    with torch.no_grad():
        want = RefSampler(ref, 1e-4, 0.02, T)(img, **kw)
    torch.manual_seed(9)
    got = GaussianDiffusionSampler(net, 1e-4, 0.02, T)(img, **kw)
    assert got.shape == want.shape and float(got.abs().max()) <= 1.0
    assert float((got - want).abs().max()) < 2e-4, float((got - want).abs().max())
