"""CPU tests: the oracle restatement (oracle/ref_torch.py) against
  (1) golden fixtures minted from the reference itself (oracle/make_golden.py), and
  (2) the reference imported by file path, when /root/reference exists (build container).
fp32 CPU, same torch ops in the same order => bit-exact unless stated."""
import os

import pytest
import torch

from oracle import ref_loader, ref_torch as R


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_schedule_tables_match_reference_golden(golden_dir):
    g = _load(golden_dir, "schedule_tables.pt")
    for (b1, bT, T), tabs in g.items():
        mine = R.schedule_tables(b1, bT, T)
        for k, v in tabs.items():
            assert torch.equal(mine[k], v), (k, T)


def test_schedule_known_answers():
    # SURVEY.md §8(c) known-answer values (f64), beta_1=1e-4 beta_T=0.02 T=1000
    tab = R.schedule_tables(1e-4, 0.02, 1000)
    assert tab["betas"][0].item() == 9.9999997473787516e-05
    assert tab["betas"][999].item() == 0.019999999552965164
    assert abs(tab["sqrt_alphas_bar"][500].item() - 0.27892051694294445) < 1e-15
    assert abs(tab["coeff1"][999].item() - 1.0101525443218162) < 1e-15
    assert abs(tab["coeff2"][0].item() - 0.010000499911173551) < 1e-15
    assert tab["posterior_var"][0].item() == 0.0
    assert abs(tab["posterior_var"][1].item() - 5.4531875419108355e-05) < 1e-18


def test_time_embedding_table_known_answer():
    tab = R.sinusoid_table(1000, 64)
    ref = torch.tensor([0.84147096, 0.54030234, 0.68156135, 0.73176098])
    assert torch.allclose(tab[1, :4], ref, atol=1e-7)
    assert torch.equal(tab[0, 0::2], torch.zeros(32)) and torch.equal(tab[0, 1::2], torch.ones(32))


def test_identity_denoiser_trainer_and_sampler(golden_dir):
    g = _load(golden_dir, "diffusion_identity.pt")

    class Id(torch.nn.Module):
        def forward(self, x, t, labels=None):
            return x[:, :3]

    torch.manual_seed(0)
    x = torch.rand(2, 3, 4, 4) * 2 - 1
    assert torch.equal(x, g["trainer_x0"])
    loss = R.GaussianDiffusionTrainer(Id(), 1e-4, 0.02, 1000)(x, torch.tensor([1, 2]))
    assert torch.equal(loss, g["trainer_loss"])
    assert abs(loss.sum().item() - 2.16325021) < 1e-5          # SURVEY.md §8(c) KAT
    torch.manual_seed(1)
    xT = torch.randn(2, 3, 4, 4)
    x0 = R.GaussianDiffusionSampler(Id(), 1e-4, 0.02, 10, w=1.8)(xT, torch.tensor([1, 2]))
    assert torch.equal(x0, g["sampler_x0"])


def _block(name):
    tdim = 64
    return {
        "resblock_32_64_attn": lambda: R.ResBlock(32, 64, tdim, 0.1, attn=True),
        "resblock_64_64": lambda: R.ResBlock(64, 64, tdim, 0.1, attn=False),
        "downsample_32": lambda: R.DownSample(32),
        "upsample_32": lambda: R.UpSample(32),
    }[name]()


@pytest.mark.parametrize("name", ["resblock_32_64_attn", "resblock_64_64", "downsample_32", "upsample_32", "attn_64"])
def test_blocks_match_reference_golden(golden_dir, name):
    g = _load(golden_dir, "blocks.pt")[name]
    if name == "attn_64":
        mod = R.AttnBlock(64)
        mod.load_state_dict({k[2:]: v for k, v in g["sd"].items()})
    else:
        mod = _block(name)
        mod.load_state_dict(g["sd"])
    mod.eval()
    x = g["x"].clone().requires_grad_(True)
    extra = [e.clone().requires_grad_(True) for e in g["extra"]]
    y = mod(x, *extra)
    assert torch.equal(y, g["y"])
    y.backward(g["gy"])
    assert torch.allclose(x.grad, g["gx"], rtol=0, atol=1e-6)
    named = dict(mod.named_parameters())
    for k, v in g["gparams"].items():
        k2 = k[2:] if name == "attn_64" else k
        assert torch.allclose(named[k2].grad, v, rtol=1e-5, atol=1e-6), k


def test_unet_tiny_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "unet_tiny.pt")
    cfg = g["cfg"]
    for tag, nl in (("cond", 10), ("uncond", None)):
        net = R.UNet(num_labels=nl, **cfg)
        sd = g["sd"] if nl else {k: v for k, v in g["sd"].items() if not k.startswith("cond_embedding.")}
        net.load_state_dict(sd)
        assert set(net.state_dict().keys()) == set(sd.keys())
        net.train()
        eps = net(g["x"], g["t"], g[tag]["labels"])
        assert torch.equal(eps, g[tag]["eps"]), tag
        (eps ** 2).sum().backward()
        named = dict(net.named_parameters())
        for k, v in g[tag]["grads"].items():
            assert torch.allclose(named[k].grad, v, rtol=1e-4, atol=1e-5), (tag, k)
        for k, v in g[tag]["grad_sqnorm"].items():
            assert abs(float((named[k].grad.double() ** 2).sum()) - v) <= 1e-4 * max(v, 1e-6), (tag, k)
        if nl is None:   # cond_proj never receives a gradient in the unconditional model
            assert all(p.grad is None for k, p in named.items() if ".cond_proj." in k)
    # trainer + CFG sampler
    net = R.UNet(num_labels=10, **cfg)
    net.load_state_dict(g["sd"])
    torch.manual_seed(g["trainer"]["seed"])
    loss = R.GaussianDiffusionTrainer(net, 1e-4, 0.02, cfg["T"])(g["x"], g["trainer"]["labels"])
    assert torch.equal(loss, g["trainer"]["loss"])
    net.eval()
    torch.manual_seed(g["sampler"]["seed"])
    xT = torch.randn(2, 3, 16, 16)
    assert torch.equal(xT, g["sampler"]["xT"])
    with torch.no_grad():
        x0 = R.GaussianDiffusionSampler(net, 1e-4, 0.02, g["sampler"]["T"], w=g["sampler"]["w"])(xT, g["sampler"]["labels"])
    assert torch.equal(x0, g["sampler"]["x0"])


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_state_dict_and_forward():
    cfg = dict(T=50, ch=32, ch_mult=[1, 2], attn=[0], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(5)
    ref = ref_loader.assemble_unet(num_labels=4, **cfg)
    mine = R.UNet(num_labels=4, **cfg)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 8, 8)
    t = torch.tensor([0, 49])
    lab = torch.tensor([1, 0])
    assert torch.equal(ref(x, t, lab), mine(x, t, lab))
    # the reference's live (MHA) UNet shares every key except the attention sub-modules
    mc = ref_loader.model_condition()
    live = mc.UNet(T=50, num_labels=4, ch=32, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
    live_keys = {k for k in live.state_dict() if ".attn." not in k}
    assert live_keys == {k for k in mine.state_dict() if ".attn." not in k}


class _DdimModel(torch.nn.Module):
    """The DDIM fixture's stand-in network behind the label interface: labels != 0 -> its conditional branch, 0 -> the other."""

    def forward(self, x, t, labels=None):
        from oracle.make_golden import ddim_dummy_eps
        tt = (t.float() / 1000.).view(-1, 1, 1, 1)
        return ddim_dummy_eps(x, tt, cond=labels is None or bool((labels != 0).all()))


def test_ddim_restatement_matches_reference_golden(golden_dir):
    """oracle.ddim_sample == the hybrid sampler's DDIM branch run from the reference class (diffusion/Diffusion.py:241-269)."""
    g = torch.load(os.path.join(golden_dir, "ddim_reference.pt"))
    for run in g["runs"]:
        torch.manual_seed(run["seed"])
        xT = torch.randn_like(run["xT"])
        assert torch.equal(xT, run["xT"])
        y0 = R.ddim_sample(_DdimModel(), g["beta_1"], g["beta_T"], g["T"], xT, torch.tensor([1, 2]), run["scale"], run["ddim_step"])
        assert torch.allclose(y0, run["y0"], atol=2e-6), float((y0 - run["y0"]).abs().max())
