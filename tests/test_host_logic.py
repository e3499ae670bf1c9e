"""CPU tests of the host logic (kernel schedule, packed layouts, views, flat buffers) with the torch test
double of the operator layer (tests/emu_backend.py) standing in for the CUDA library.  The reference
numbers come from the golden fixtures minted from the reference itself and from the oracle."""
import os

import pytest
import torch

import hdiff_b200.ops as hops
from hdiff_b200.diffusion.Model import UNet as UNetU
from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet as UNetC
from hdiff_b200.DiffusionFreeGuidence.DiffusionCondition import GaussianDiffusionTrainer, GaussianDiffusionSampler
from oracle import ref_torch as R
from tests.emu_backend import EmuOps


@pytest.fixture(autouse=True)
def emu():
    prev = hops._backend
    hops.set_backend(EmuOps())
    yield
    hops.set_backend(prev)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _close(a, b, rtol, atol=2e-5):
    """relative error in the 2-norm, with an absolute floor: several reference gradients are
    mathematically zero (a bias feeding a per-channel GroupNorm, proj_k.bias under softmax)."""
    d = float((a - b).norm())
    return d <= rtol * float(b.norm()) + atol


def test_state_dict_keys_match_oracle():
    for cfg in (dict(T=100, ch=32, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0),
                dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1)):
        a = UNetC(num_labels=10, compute_dtype=torch.float32, **cfg)
        b = R.UNet(num_labels=10, **cfg)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(sa[k].shape == sb[k].shape for k in sa)
        u = UNetU(compute_dtype=torch.float32, **cfg)
        assert list(u.state_dict().keys()) == list(R.UNet(**cfg).state_dict().keys())


@pytest.mark.parametrize("tag", ["cond", "uncond"])
def test_unet_tiny_forward_backward_vs_reference_golden(golden_dir, tag):
    g = _load(golden_dir, "unet_tiny.pt")
    cfg = g["cfg"]
    if tag == "cond":
        net = UNetC(num_labels=10, compute_dtype=torch.float32, **cfg)
        net.load_state_dict(g["sd"])
    else:
        net = UNetU(compute_dtype=torch.float32, **cfg)
        net.load_state_dict({k: v for k, v in g["sd"].items() if not k.startswith("cond_embedding.")})
    net.train()
    rec = g[tag]
    eps = net(g["x"], g["t"]) if rec["labels"] is None else net(g["x"], g["t"], rec["labels"])
    assert _rel(eps.detach(), rec["eps"]) < 2e-5
    (eps ** 2).sum().backward()
    params = dict(net.named_parameters())
    for k, gref in rec["grads"].items():
        assert params[k].grad is not None, k
        assert _close(params[k].grad, gref, 1e-4), (k, _rel(params[k].grad, gref))
    for k, sq in rec["grad_sqnorm"].items():
        mine = float((params[k].grad.double() ** 2).sum())
        assert abs(mine - sq) <= 4e-4 * sq + 1e-9, (k, mine, sq)
    # parameters the reference leaves without gradient stay without gradient
    for k, p in params.items():
        if k not in rec["grad_sqnorm"]:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k


def test_trainer_and_sampler_vs_reference_golden(golden_dir):
    g = _load(golden_dir, "unet_tiny.pt")
    net = UNetC(num_labels=10, compute_dtype=torch.float32, **g["cfg"])
    net.load_state_dict(g["sd"])
    net.train()
    torch.manual_seed(g["trainer"]["seed"])
    loss = GaussianDiffusionTrainer(net, 1e-4, 0.02, g["cfg"]["T"])(g["x"], g["trainer"]["labels"])
    assert _rel(loss.detach(), g["trainer"]["loss"]) < 1e-4
    net.eval()
    s = g["sampler"]
    torch.manual_seed(s["seed"])
    xT = torch.randn(2, 3, 16, 16)          # the fixture drew x_T after seeding (oracle/make_golden.py)
    assert torch.equal(xT, s["xT"])
    x0 = GaussianDiffusionSampler(net, 1e-4, 0.02, s["T"], w=s["w"])(xT, s["labels"])
    assert float((x0 - s["x0"]).abs().max()) < 1e-4


def test_identity_denoiser_matches_golden(golden_dir):
    g = _load(golden_dir, "diffusion_identity.pt")

    class Id(torch.nn.Module):
        def forward(self, x, t, labels=None):
            return x[:, :3]

    torch.manual_seed(0)
    x = torch.rand(2, 3, 4, 4) * 2 - 1
    assert torch.equal(x, g["trainer_x0"])
    loss = GaussianDiffusionTrainer(Id(), 1e-4, 0.02, 1000)(x, torch.tensor([1, 2]))
    assert torch.allclose(loss, g["trainer_loss"], atol=1e-6)
    torch.manual_seed(1)
    xT = torch.randn(2, 3, 4, 4)
    x0 = GaussianDiffusionSampler(Id(), 1e-4, 0.02, 10, w=1.8)(xT, torch.tensor([1, 2]))
    assert torch.allclose(x0, g["sampler_x0"], atol=1e-5)


def test_wider_topology_vs_oracle():
    """Three levels, two res blocks, attention at level 1, concat widths that straddle GroupNorm groups."""
    cfg = dict(T=50, ch=32, ch_mult=[1, 2, 2], attn=[1], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(3)
    ref = R.UNet(num_labels=5, **cfg)
    net = UNetC(num_labels=5, compute_dtype=torch.float32, **cfg)
    net.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 16, 16)
    t = torch.tensor([1, 40])
    lab = torch.tensor([3, 0])
    e_ref = ref(x, t, lab)
    e = net(x, t, lab)
    assert _rel(e.detach(), e_ref.detach()) < 2e-5
    gy = torch.randn_like(e_ref)
    e_ref.backward(gy)
    e.backward(gy)
    pr = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        if pr[k].grad is None:
            continue
        assert _close(p.grad, pr[k].grad, 2e-4), (k, _rel(p.grad, pr[k].grad))


def test_second_backward_without_zero_grad_accumulates():
    cfg = dict(T=20, ch=32, ch_mult=[1, 1], attn=[], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(5)
    net = UNetU(compute_dtype=torch.float32, **cfg)
    x = torch.randn(1, 3, 8, 8)
    t = torch.tensor([4])
    net(x, t).sum().backward()
    g1 = net.head.weight.grad.clone()
    net(x, t).sum().backward()
    assert torch.allclose(net.head.weight.grad, 2 * g1, rtol=1e-5, atol=1e-7)


def test_bf16_schedule_with_padded_head_and_tail_vs_oracle():
    """bf16 compute mode routes the 3-channel head / tail through 64-channel GEMM operands (zero-padded packed weights,
    channel-padded NHWC input, fp32 NCHW store of the first 3 output channels).  The index tables of that path are
    checked here against the oracle (tolerance = bf16 storage of activations and weights)."""
    cfg = dict(T=50, ch=32, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(9)
    ref = R.UNet(num_labels=None, **cfg)
    net = UNetU(compute_dtype=torch.bfloat16, **cfg)
    net.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 16, 16)
    t = torch.tensor([7, 33])
    e_ref = ref(x, t)
    e = net(x, t)
    assert e.dtype == torch.float32 and e.shape == e_ref.shape
    assert _rel(e.detach(), e_ref.detach()) < 2e-2
    gy = torch.randn_like(e_ref)
    e_ref.backward(gy)
    e.backward(gy)
    pr = dict(ref.named_parameters())
    gscale = max(float(p.grad.norm()) for p in pr.values() if p.grad is not None)
    for k in ("head.weight", "head.bias", "tail.2.weight", "tail.2.bias", "tail.0.weight"):
        p = dict(net.named_parameters())[k]
        assert p.grad.shape == pr[k].grad.shape
        assert _close(p.grad, pr[k].grad, 5e-2, 1e-3 * gscale), (k, _rel(p.grad, pr[k].grad))


def test_ddim_sampler_vs_reference_golden(golden_dir):
    """GaussianDiffusionSampler(..., w)(x_T, labels, ddim=True, ddim_step=n): the hybrid sampler's DDIM branch
    (diffusion/Diffusion.py:241-269, eta = 0, guidance scale 1 + w) as one linear update per step through hd_sampler_step."""
    from tests.test_oracle import _DdimModel
    g = torch.load(os.path.join(golden_dir, "ddim_reference.pt"))
    for run in g["runs"]:
        sa = GaussianDiffusionSampler(_DdimModel(), g["beta_1"], g["beta_T"], g["T"], w=run["scale"] - 1.)
        sa.model.cfg_batched = False                 # the stand-in tells its two branches apart by whole-batch labels
        y0 = sa(run["xT"], torch.tensor([1, 2]), ddim=True, ddim_step=run["ddim_step"])
        assert torch.allclose(y0, run["y0"], atol=2e-5), float((y0 - run["y0"]).abs().max())
    # the table itself: seq order, t + 1 indexing, no noise term
    coef, stride = sa.ddim_tables(20)
    assert stride == 50 and coef.shape == (20, 3) and float(coef[:, 2].abs().max()) == 0.0
    ab = sa.alphas_bar
    assert abs(float(coef[0, 0]) - float((ab[0].float().sqrt() / ab[1].float().sqrt()))) < 1e-7


def test_ddim_unconditional_and_default_steps_vs_oracle():
    """DDIM without labels (no guidance pair) and with a real (tiny) UNet on the operator test double, against the oracle."""
    cfg = dict(T=40, ch=32, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(5)
    ref = R.UNet(**cfg).eval()
    net = UNetU(compute_dtype=torch.float32, **cfg)
    net.load_state_dict(ref.state_dict())
    net.eval()
    xT = torch.randn(2, 3, 8, 8)
    sa = GaussianDiffusionSampler(net, 1e-4, 0.02, cfg["T"])
    y0 = sa(xT, ddim=True, ddim_step=8)
    with torch.no_grad():
        r0 = R.ddim_sample(ref, 1e-4, 0.02, cfg["T"], xT, None, 1., 8)
    assert torch.allclose(y0, r0, atol=2e-4), float((y0 - r0).abs().max())
    # conditional, guided
    torch.manual_seed(6)
    refc = R.UNet(num_labels=4, **cfg).eval()
    netc = UNetC(num_labels=4, compute_dtype=torch.float32, **cfg)
    netc.load_state_dict(refc.state_dict())
    netc.eval()
    lab = torch.tensor([2, 4])
    y0 = GaussianDiffusionSampler(netc, 1e-4, 0.02, cfg["T"], w=1.5)(xT, lab, ddim=True, ddim_step=10)
    with torch.no_grad():
        r0 = R.ddim_sample(refc, 1e-4, 0.02, cfg["T"], xT, lab, 2.5, 10)
    assert torch.allclose(y0, r0, atol=2e-4), float((y0 - r0).abs().max())


def _clone_net(cls, cfg, src, **kw):
    n = cls(compute_dtype=torch.float32, **kw, **cfg)
    n.load_state_dict(src.state_dict())
    return n


def test_flat_adamw_skips_never_used_parameters_like_torch():
    """torch.optim.AdamW skips a parameter whose grad is None: no weight decay, no state.  The unconditional model's
    cond_proj.* are such parameters (ModelCondition.py:199-200); FlatAdamW must leave them bit-identical, and update
    everything else like clip_grad_norm_ + AdamW."""
    from hdiff_b200.optim import FlatAdamW
    cfg = dict(T=20, ch=32, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(3)
    a = UNetU(compute_dtype=torch.float32, **cfg)
    b = _clone_net(UNetU, cfg, a)
    before = {k: v.clone() for k, v in a.state_dict().items()}
    oa = FlatAdamW(a, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
    ob = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=1e-2)
    x, t = torch.randn(2, 3, 8, 8), torch.tensor([3, 11])
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for _ in range(3):
        oa.zero_grad(); ob.zero_grad()
        (a(x, t) ** 2).sum().backward()
        for k in pa:                                  # the SAME gradients for both optimisers: Adam normalises, so the rounding
            pb[k].grad = None if pa[k].grad is None else pa[k].grad.detach().clone()   # noise of exactly-zero gradients would flip signs
        oa.step()
        torch.nn.utils.clip_grad_norm_([p for p in b.parameters() if p.grad is not None], 1.0)
        ob.step()
    for k in pa:
        if "cond_proj" in k:
            assert pa[k].grad is None and torch.equal(pa[k].detach(), before[k]), k          # untouched: no decay
        assert torch.allclose(pa[k].detach(), pb[k].detach(), rtol=2e-5, atol=2e-7), (k, float((pa[k] - pb[k]).abs().max()))
    # an lr scheduler drives param_groups, as in TrainCondition.py:41-44
    oa.param_groups[0]["lr"] = 5e-4
    assert oa.lr == 5e-4


def test_flat_adamw_uses_accumulated_gradients():
    """Two micro-batches without zero_grad: autograd sums into the first gradient buffer; FlatAdamW must step on the SUM
    (it used to read only the last backward's buffer)."""
    from hdiff_b200.optim import FlatAdamW
    cfg = dict(T=20, ch=32, ch_mult=[1, 1], attn=[], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(4)
    a = UNetC(num_labels=3, compute_dtype=torch.float32, **cfg)
    b = _clone_net(UNetC, cfg, a, num_labels=3)
    oa = FlatAdamW(a, lr=1e-3, weight_decay=0.0, max_grad_norm=0.5)
    ob = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=0.0)
    xs = [torch.randn(1, 3, 8, 8) for _ in range(2)]
    t, lab = torch.tensor([7]), torch.tensor([2])
    for x in xs:
        (a(x, t, lab) ** 2).sum().backward()
    first = None
    for x in xs:                                      # reference sum, formed by hand from two separate backward passes
        for p in b.parameters():
            p.grad = None
        (b(x, t, lab) ** 2).sum().backward()
        cur = [None if p.grad is None else p.grad.clone() for p in b.parameters()]
        first = cur if first is None else [u if v is None else v if u is None else u + v for u, v in zip(first, cur)]
    for p, g, q in zip(b.parameters(), first, a.parameters()):
        assert (g is None) == (q.grad is None)
        if g is not None:
            assert torch.allclose(q.grad, g, rtol=1e-4, atol=1e-6)            # a accumulated the sum of the two micro-batches
        p.grad = None if q.grad is None else q.grad.detach().clone()          # same gradients into both optimisers
    oa.step()
    torch.nn.utils.clip_grad_norm_(b.parameters(), 0.5)
    ob.step()
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-5, atol=2e-7), k


def test_ddim_argument_errors_are_explicit():
    sa = GaussianDiffusionSampler(torch.nn.Identity(), 1e-4, 0.02, 1000)
    x = torch.randn(1, 3, 4, 4)
    with pytest.raises(ValueError, match="ddim_step"):
        sa(x, ddim=True)                                   # the reference divides 1000 / None here (Diffusion.py:246)
    with pytest.raises(ValueError, match="alphas_bar"):
        sa(x, ddim=True, ddim_step=1000)                   # stride 1 reads alphas_bar[T] (Diffusion.py:251)


def test_live_reference_checkpoint_keys_are_rejected_with_a_clear_message():
    """A checkpoint of the reference's live MHA UNet carries attn.in_proj_weight / attn.out_proj.* keys; loading it into the
    AttnBlock-style model must fail loudly, not drop every attention weight under strict=False (TrainCondition.py:36-38)."""
    cfg = dict(T=20, ch=32, ch_mult=[1, 1], attn=[0], num_res_blocks=1, dropout=0.0)
    net = UNetC(num_labels=3, compute_dtype=torch.float32, **cfg)
    sd = {k: v for k, v in net.state_dict().items() if ".attn." not in k}
    sd["downblocks.0.attn.in_proj_weight"] = torch.zeros(96, 32)
    with pytest.raises(RuntimeError, match="MultiheadAttention"):
        net.load_state_dict(sd, strict=False)
