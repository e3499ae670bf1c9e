"""ORACLE support (test infrastructure): mint golden fixtures from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Writes small `.pt` files under tests/golden/.  Every tensor in them was produced by the
reference's own classes (loaded by `oracle/ref_loader.py`), fp32 on CPU, torch 2.11.
The CPU tests check `oracle/ref_torch.py` against them; the GPU tests check the CUDA path
against them and against the oracle.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class Identity3(nn.Module):
    """Identity denoiser on the first 3 channels (idea: diffusion/Diffusion.py:373-375)."""

    def forward(self, x, t, labels=None):
        return x[:, :3]


def clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def schedule():
    dc = ref_loader.diffusion_condition()
    out = {}
    for (b1, bT, T) in [(1e-4, 0.02, 1000), (1e-4, 0.028, 500), (1e-4, 0.02, 10)]:
        tr = dc.GaussianDiffusionTrainer(Identity3(), b1, bT, T)
        sa = dc.GaussianDiffusionSampler(Identity3(), b1, bT, T, w=1.8)
        out[(b1, bT, T)] = {
            "betas": tr.betas.clone(), "sqrt_alphas_bar": tr.sqrt_alphas_bar.clone(),
            "sqrt_one_minus_alphas_bar": tr.sqrt_one_minus_alphas_bar.clone(),
            "coeff1": sa.coeff1.clone(), "coeff2": sa.coeff2.clone(),
            "posterior_var": sa.posterior_var.clone(),
        }
    torch.save(out, os.path.join(OUT, "schedule_tables.pt"))


def diffusion_identity():
    dc = ref_loader.diffusion_condition()
    out = {}
    torch.manual_seed(0)
    x = torch.rand(2, 3, 4, 4) * 2 - 1
    labels = torch.tensor([1, 2])
    tr = dc.GaussianDiffusionTrainer(Identity3(), 1e-4, 0.02, 1000)
    out["trainer_x0"] = x.clone()
    out["trainer_loss"] = tr(x, labels).clone()          # consumes randint, randn after seed 0 + rand
    # sampler, 10 steps, identity denoiser, w = 1.8
    torch.manual_seed(1)
    xT = torch.randn(2, 3, 4, 4)
    sa = dc.GaussianDiffusionSampler(Identity3(), 1e-4, 0.02, 10, w=1.8)
    out["sampler_xT"] = xT.clone()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints every step (:88)
        out["sampler_x0"] = sa(xT, labels).clone()
    torch.save(out, os.path.join(OUT, "diffusion_identity.pt"))


def blocks():
    mc = ref_loader.model_condition()
    out = {}
    torch.manual_seed(10)
    B, H = 2, 8
    tdim = 64

    def run(name, mod, x, *extra):
        mod.eval()  # dropout off: the golden is deterministic
        x = x.clone().requires_grad_(True)
        extra = [e.clone().requires_grad_(True) for e in extra]
        y = mod(x, *extra)
        g = torch.randn_like(y)
        y.backward(g)
        out[name] = {
            "sd": clone_sd(mod), "x": x.detach().clone(), "extra": [e.detach().clone() for e in extra],
            "y": y.detach().clone(), "gy": g, "gx": x.grad.clone(),
            "gextra": [None if e.grad is None else e.grad.clone() for e in extra],
            "gparams": {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None},
        }

    temb, cemb = torch.randn(B, tdim), torch.randn(B, tdim)
    run("resblock_32_64_attn", mc.ResBlock_old(32, 64, tdim, 0.1, attn=True), torch.randn(B, 32, H, H), temb, cemb)
    run("resblock_64_64", mc.ResBlock_old(64, 64, tdim, 0.1, attn=False), torch.randn(B, 64, H, H), temb, cemb)
    run("downsample_32", mc.DownSample(32), torch.randn(B, 32, H, H), temb, cemb)
    run("upsample_32", mc.UpSample(32), torch.randn(B, 32, H, H), temb, cemb)

    class A(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = mc.AttnBlock(64)

        def forward(self, x):
            return self.a(x)
    run("attn_64", A(), torch.randn(B, 64, H, H))
    torch.save(out, os.path.join(OUT, "blocks.pt"))


GRAD_KEYS = ("head.weight", "head.bias", "tail.0.weight", "tail.2.weight", "tail.2.bias",
             "time_embedding.timembedding.0.weight", "time_embedding.timembedding.3.weight",
             "cond_embedding.condEmbedding.0.weight", "cond_embedding.condEmbedding.1.bias",
             "downblocks.0.block1.0.weight", "downblocks.0.block1.2.weight", "downblocks.0.temb_proj.1.weight",
             "downblocks.0.cond_proj.1.bias", "downblocks.1.c1.weight", "downblocks.1.c2.weight",
             "downblocks.2.attn.proj_q.weight", "downblocks.2.attn.proj.bias", "downblocks.2.attn.group_norm.weight",
             "middleblocks.0.attn.proj_v.weight", "middleblocks.1.block2.3.weight",
             "upblocks.0.shortcut.weight", "upblocks.0.block1.0.bias", "upblocks.2.t.weight", "upblocks.2.c.bias",
             "upblocks.4.block2.0.weight")


def unet_tiny():
    """Tiny UNet (ch=32, two levels) through the reference blocks: forward, selected grads,
    the reference conditional trainer and a 5-step CFG sampler.  One state_dict is stored (the
    conditional one); the unconditional net loads it minus `cond_embedding.*`."""
    dc = ref_loader.diffusion_condition()
    out = {}
    cfg = dict(T=100, ch=32, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(20)
    cond = ref_loader.assemble_unet(num_labels=10, **cfg)
    sd = clone_sd(cond)
    out["cfg"] = cfg
    out["sd"] = sd
    x = torch.rand(2, 3, 16, 16) * 2 - 1
    t = torch.tensor([3, 77])
    out["x"], out["t"] = x, t
    for tag, net, labels in (("cond", cond, torch.tensor([0, 4])),
                             ("uncond", ref_loader.assemble_unet(num_labels=None, **cfg), None)):
        if labels is None:
            net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("cond_embedding.")})
        net.train()
        eps = net(x, t, labels)
        (eps ** 2).sum().backward()
        grads = dict(net.named_parameters())
        rec = {"labels": labels, "eps": eps.detach().clone(),
               "grads": {k: grads[k].grad.clone() for k in GRAD_KEYS if k in grads and grads[k].grad is not None},
               "grad_sqnorm": {k: float((p.grad.double() ** 2).sum()) for k, p in grads.items() if p.grad is not None}}
        out[tag] = rec
    # reference conditional trainer + CFG sampler around the assembled net
    cond.zero_grad()
    torch.manual_seed(21)
    tr = dc.GaussianDiffusionTrainer(cond, 1e-4, 0.02, cfg["T"])
    lab = torch.tensor([2, 9])
    out["trainer"] = {"seed": 21, "labels": lab, "loss": tr(x, lab).detach().clone()}
    cond.eval()
    sa = dc.GaussianDiffusionSampler(cond, 1e-4, 0.02, 5, w=1.8)
    torch.manual_seed(22)
    xT = torch.randn(2, 3, 16, 16)
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        x0 = sa(xT, lab).clone()
    out["sampler"] = {"seed": 22, "labels": lab, "xT": xT, "x0": x0, "T": 5, "w": 1.8}
    torch.save(out, os.path.join(OUT, "unet_tiny.pt"))


class DdimDummy(nn.Module):
    """Deterministic stand-in for the hybrid pipeline's network in the DDIM fixture: input = cat(conditioning image, y_t)
    (diffusion/Diffusion.py:253).  The reference calls it as model(input, t) and, for guidance, model(input, t,
    context_zero=True) (:254,258): the two calls are told apart by the keyword being passed at all."""

    def forward(self, inp, t, context_zero=None):
        y = inp[:, 3:]
        tt = (t.float() / 1000.).view(-1, 1, 1, 1)
        return ddim_dummy_eps(y, tt, cond=context_zero is None)


def ddim_dummy_eps(y, tt, cond):
    """Roughly 'the noise is most of y_t', so that the trajectory stays inside (-1, 1) and the final clip hides nothing."""
    return 0.9 * y + 0.1 * torch.tanh(y) + 0.05 * tt if cond else 0.8 * y - 0.05 * tt


def ddim_reference():
    """The hybrid sampler's DDIM branch (diffusion/Diffusion.py:241-269), run from the reference class itself."""
    Sampler = ref_loader.hybrid_sampler_class()
    out = {"beta_1": 1e-4, "beta_T": 0.02, "T": 1000, "runs": []}
    sa = Sampler(DdimDummy(), 1e-4, 0.02, 1000)
    torch.manual_seed(30)
    img = torch.randint(0, 256, (2, 3, 8, 8)).float()
    for seed, scale, steps in ((31, 1, 100), (32, 1.8, 100), (33, 2.5, 20)):
        torch.manual_seed(seed)
        with torch.no_grad():
            y0 = sa(img, ddim=True, unconditional_guidance_scale=scale, ddim_step=steps).clone()
        torch.manual_seed(seed)
        xT = torch.randn_like(img)               # the reference's first draw (:243)
        out["runs"].append({"seed": seed, "scale": scale, "ddim_step": steps, "xT": xT, "y0": y0})
    torch.save(out, os.path.join(OUT, "ddim_reference.pt"))


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    schedule()
    diffusion_identity()
    blocks()
    unet_tiny()
    ddim_reference()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
