"""ORACLE (test infrastructure, not product code).

Plain fp32 PyTorch restatement of the reference's DDPM / classifier-free-guidance
hot path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the product package must never do so.

Parity status: the reference ships no golden vectors or tests for this path
(SURVEY.md F8), so the restatement is pinned against the reference's *own code*,
imported by file path in the build container (`oracle/ref_loader.py`), through
  * `oracle/make_golden.py` -> `tests/golden/*.pt` (outputs of the reference itself), and
  * `tests/test_oracle_vs_reference.py` (bit-exact comparison when /root/reference exists).

Every class cites the reference lines it restates (paths relative to /root/reference).
Parameter / buffer names are kept identical so `state_dict()`s interchange.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# Building blocks (DiffusionFreeGuidence/ModelCondition.py:22-164, diffusion/Model.py:18-265)
# --------------------------------------------------------------------------------------
class Swish(nn.Module):
    """x * sigmoid(x)  (ModelCondition.py:22-24)."""

    def forward(self, x):
        return x * torch.sigmoid(x)


def sinusoid_table(T: int, d_model: int) -> torch.Tensor:
    """fp32 [T, d_model] table with (sin, cos) interleaved per frequency
    (ModelCondition.py:31-38)."""
    assert d_model % 2 == 0
    freq = torch.exp(-(torch.arange(0, d_model, step=2) / d_model * math.log(10000)))
    ang = torch.arange(T).float()[:, None] * freq[None, :]
    return torch.stack([torch.sin(ang), torch.cos(ang)], dim=-1).view(T, d_model)


class TimeEmbedding(nn.Module):
    """Trainable sinusoid table -> Linear -> Swish -> Linear (ModelCondition.py:27-49)."""

    def __init__(self, T, d_model, dim):
        super().__init__()
        self.timembedding = nn.Sequential(
            nn.Embedding.from_pretrained(sinusoid_table(T, d_model), freeze=False),
            nn.Linear(d_model, dim),
            Swish(),
            nn.Linear(dim, dim),
        )

    def forward(self, t):
        return self.timembedding(t)


class ConditionalEmbedding(nn.Module):
    """Label table (row 0 = padding = null condition) -> Linear -> Swish -> Linear
    (ModelCondition.py:52-65)."""

    def __init__(self, num_labels, d_model, dim):
        super().__init__()
        assert d_model % 2 == 0
        self.condEmbedding = nn.Sequential(
            nn.Embedding(num_labels + 1, d_model, padding_idx=0),
            nn.Linear(d_model, dim),
            Swish(),
            nn.Linear(dim, dim),
        )

    def forward(self, labels):
        return self.condEmbedding(labels)


class DownSample(nn.Module):
    """conv3x3 s2 p1 + conv5x5 s2 p2, summed (ModelCondition.py:68-76)."""

    def __init__(self, in_ch):
        super().__init__()
        self.c1 = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1)
        self.c2 = nn.Conv2d(in_ch, in_ch, 5, stride=2, padding=2)

    def forward(self, x, temb=None, cemb=None):
        return self.c1(x) + self.c2(x)


class UpSample(nn.Module):
    """ConvTranspose2d(5, s2, p2, op1) then conv3x3 (ModelCondition.py:79-89)."""

    def __init__(self, in_ch):
        super().__init__()
        self.c = nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1)
        self.t = nn.ConvTranspose2d(in_ch, in_ch, 5, 2, 2, 1)

    def forward(self, x, temb=None, cemb=None):
        return self.c(self.t(x))


class AttnBlock(nn.Module):
    """Single-head spatial self-attention with residual (ModelCondition.py:92-120).

    `use_sdpa` (class switch, default off = the reference's arithmetic): bench.py's same-GPU "stock PyTorch" leg swaps
    the materialised bmm / softmax / bmm for F.scaled_dot_product_attention (same function; the library's flash kernel)."""

    use_sdpa = False

    def __init__(self, in_ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1)
        self.proj = nn.Conv2d(in_ch, in_ch, 1)

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x)
        q = self.proj_q(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        k = self.proj_k(h).reshape(B, C, H * W)
        v = self.proj_v(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        if AttnBlock.use_sdpa:
            o = F.scaled_dot_product_attention(q.contiguous()[:, None], k.transpose(1, 2).contiguous()[:, None], v.contiguous()[:, None],
                                               scale=int(C) ** (-0.5))[:, 0]
            return x + self.proj(o.reshape(B, H, W, C).permute(0, 3, 1, 2))
        w = F.softmax(torch.bmm(q, k) * (int(C) ** (-0.5)), dim=-1)
        h = torch.bmm(w, v).view(B, H, W, C).permute(0, 3, 1, 2)
        return x + self.proj(h)


class ResBlock(nn.Module):
    """`ResBlock_old` (ModelCondition.py:124-164): the variant whose `attn` flag attaches an
    `AttnBlock`.  `cemb=None` skips the cond_proj add, as `ResBlock.forward` does
    (ModelCondition.py:199-200); `cond_proj` stays registered either way."""

    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False):
        super().__init__()
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(),
                                    nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.cond_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1))
        self.shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()
        self.attn = AttnBlock(out_ch) if attn else nn.Identity()

    def forward(self, x, temb, cemb=None):
        h = self.block1(x)
        h = h + self.temb_proj(temb)[:, :, None, None]
        if cemb is not None:
            h = h + self.cond_proj(cemb)[:, :, None, None]
        h = self.block2(h)
        h = h + self.shortcut(x)
        return self.attn(h)


class UNet(nn.Module):
    """The `UNet(T, ch, ch_mult, attn, num_res_blocks, dropout)` that diffusion/Train.py:30-31
    calls (undefined in the reference), assembled in the topology of
    ModelCondition.py:213-276 from the reference's blocks:
      * down-path ResBlocks at level i get an AttnBlock iff i in `attn`;
      * middle = [attn, no-attn] (ModelCondition.py:233-236);
      * up-path ResBlocks never get attention (ModelCondition.py:242).
    `num_labels=None` gives the unconditional model (no cond_embedding, cemb=None)."""

    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None):
        super().__init__()
        tdim = ch * 4
        self.time_embedding = TimeEmbedding(T, ch, tdim)
        if num_labels is not None:
            self.cond_embedding = ConditionalEmbedding(num_labels, ch, tdim)
        self.head = nn.Conv2d(3, ch, 3, stride=1, padding=1)
        self.downblocks = nn.ModuleList()
        widths = [ch]
        cur = ch
        last = len(ch_mult) - 1
        for level, mult in enumerate(ch_mult):
            for _ in range(num_res_blocks):
                self.downblocks.append(ResBlock(cur, ch * mult, tdim, dropout, attn=(level in attn)))
                cur = ch * mult
                widths.append(cur)
            if level != last:
                self.downblocks.append(DownSample(cur))
                widths.append(cur)
        self.middleblocks = nn.ModuleList([ResBlock(cur, cur, tdim, dropout, attn=True),
                                           ResBlock(cur, cur, tdim, dropout, attn=False)])
        self.upblocks = nn.ModuleList()
        for level in range(last, -1, -1):
            for _ in range(num_res_blocks + 1):
                self.upblocks.append(ResBlock(widths.pop() + cur, ch * ch_mult[level], tdim, dropout, attn=False))
                cur = ch * ch_mult[level]
            if level != 0:
                self.upblocks.append(UpSample(cur))
        assert not widths
        self.tail = nn.Sequential(nn.GroupNorm(32, cur), Swish(), nn.Conv2d(cur, 3, 3, stride=1, padding=1))

    def forward(self, x, t, labels=None):
        temb = self.time_embedding(t)
        cemb = self.cond_embedding(labels) if labels is not None else None
        h = self.head(x)
        skips = [h]
        for blk in self.downblocks:
            h = blk(h, temb, cemb)
            skips.append(h)
        for blk in self.middleblocks:
            h = blk(h, temb, cemb)
        for blk in self.upblocks:
            if isinstance(blk, ResBlock):
                h = torch.cat([h, skips.pop()], dim=1)
            h = blk(h, temb, cemb)
        assert not skips
        return self.tail(h)


# --------------------------------------------------------------------------------------
# Diffusion process (DiffusionFreeGuidence/DiffusionCondition.py:9-98;
# unconditional twins: diffusion/Diffusion.py:286-368, commented "Old CODE")
# --------------------------------------------------------------------------------------
def extract(v, t, x_shape):
    """gather(v, t) -> fp32 -> [B,1,1,...] (DiffusionCondition.py:9-16)."""
    out = torch.gather(v, index=t, dim=0).float().to(t.device)
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


def schedule_tables(beta_1: float, beta_T: float, T: int) -> dict:
    """float64 tables; linspace is evaluated in fp32 and only then cast
    (DiffusionCondition.py:26-35, 58-66)."""
    betas = torch.linspace(beta_1, beta_T, T).double()
    alphas = 1. - betas
    alphas_bar = torch.cumprod(alphas, dim=0)
    alphas_bar_prev = F.pad(alphas_bar, [1, 0], value=1)[:T]
    coeff1 = torch.sqrt(1. / alphas)
    return {
        "betas": betas,
        "sqrt_alphas_bar": torch.sqrt(alphas_bar),
        "sqrt_one_minus_alphas_bar": torch.sqrt(1. - alphas_bar),
        "coeff1": coeff1,
        "coeff2": coeff1 * (1. - alphas) / torch.sqrt(1. - alphas_bar),
        "posterior_var": betas * (1. - alphas_bar_prev) / (1. - alphas_bar),
    }


class GaussianDiffusionTrainer(nn.Module):
    """Algorithm 1 (DiffusionCondition.py:19-46).  `labels=None` is the unconditional twin
    (diffusion/Diffusion.py:304-314).  RNG order: randint then randn_like."""

    def __init__(self, model, beta_1, beta_T, T):
        super().__init__()
        self.model = model
        self.T = T
        tab = schedule_tables(beta_1, beta_T, T)
        for k in ("betas", "sqrt_alphas_bar", "sqrt_one_minus_alphas_bar"):
            self.register_buffer(k, tab[k])

    def forward(self, x_0, labels=None):
        t = torch.randint(self.T, size=(x_0.shape[0],), device=x_0.device)
        noise = torch.randn_like(x_0)
        x_t = (extract(self.sqrt_alphas_bar, t, x_0.shape) * x_0
               + extract(self.sqrt_one_minus_alphas_bar, t, x_0.shape) * noise)
        pred = self.model(x_t, t) if labels is None else self.model(x_t, t, labels)
        return F.mse_loss(pred, noise, reduction='none')


class GaussianDiffusionSampler(nn.Module):
    """Algorithm 2 with classifier-free guidance (DiffusionCondition.py:49-98).  `labels=None`
    is the unconditional twin (diffusion/Diffusion.py:351-368).  The reference prints the step
    index every iteration (`:88`); the oracle does not."""

    def __init__(self, model, beta_1, beta_T, T, w=0.):
        super().__init__()
        self.model = model
        self.T = T
        self.w = w
        tab = schedule_tables(beta_1, beta_T, T)
        for k in ("betas", "coeff1", "coeff2", "posterior_var"):
            self.register_buffer(k, tab[k])

    def predict_xt_prev_mean_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return extract(self.coeff1, t, x_t.shape) * x_t - extract(self.coeff2, t, x_t.shape) * eps

    def p_mean_variance(self, x_t, t, labels=None):
        var = extract(torch.cat([self.posterior_var[1:2], self.betas[1:]]), t, x_t.shape)
        if labels is None:
            eps = self.model(x_t, t)
        else:
            eps = self.model(x_t, t, labels)
            non_eps = self.model(x_t, t, torch.zeros_like(labels))
            eps = (1. + self.w) * eps - self.w * non_eps
        return self.predict_xt_prev_mean_from_eps(x_t, t, eps), var

    def forward(self, x_T, labels=None):
        x_t = x_T
        for time_step in reversed(range(self.T)):
            t = x_t.new_ones([x_T.shape[0], ], dtype=torch.long) * time_step
            mean, var = self.p_mean_variance(x_t, t, labels)
            noise = torch.randn_like(x_t) if time_step > 0 else 0
            x_t = mean + torch.sqrt(var) * noise
            assert torch.isnan(x_t).int().sum() == 0, "nan in tensor."
        return torch.clip(x_t, -1, 1)


def ddim_sample(model, beta_1, beta_T, T, x_T, labels=None, guidance_scale=1., ddim_step=100):
    """Deterministic DDIM sampling with classifier-free guidance: the `ddim=True` branch of the hybrid sampler,
    diffusion/Diffusion.py:241-269, restated line by line for a label-conditional network.

    What differs from the reference lines, and why:
      * the reference draws y_T itself (`randn_like(input_image)`, :243) and concatenates the conditioning image to the
        network input (:253); here x_T is an argument and the network takes (x, t[, labels]) — SURVEY §8(f): the
        image-conditional DynamicUNet is not on the path;
      * the guidance pair is eps(x, t, labels) and eps(x, t, 0) (the null label), mixed exactly as :258-259;
      * `1000` (:246-247) is written as T.
    `c1 * randn_like` (:264-265) is kept: it multiplies by zero (eta = 0) but advances the generator as the reference does.
    Pinned by tests/golden/ddim_reference.pt, minted from the reference class itself (oracle/make_golden.py)."""
    alphas_bar = torch.cumprod(1. - torch.linspace(beta_1, beta_T, T).double(), dim=0).to(x_T.device)      # :188-190
    y_t = x_T
    step = int(T / ddim_step)                                                                                # :246-247
    seq = range(0, T, step)
    seq_next = [-1] + list(seq[:-1])
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = (torch.ones(y_t.shape[0]) * i).to(y_t.device).long()
        next_t = (torch.ones(y_t.shape[0]) * j).to(y_t.device).long()
        at = extract(alphas_bar, (t + 1).long(), y_t.shape)                                                  # :251
        at_next = extract(alphas_bar, (next_t + 1).long(), y_t.shape)                                        # :252
        eps = model(y_t, t) if labels is None else model(y_t, t, labels)                                     # :254
        if guidance_scale != 1 and labels is not None:                                                       # :257-259
            eps_unconditional = model(y_t, t, torch.zeros_like(labels))
            eps = eps_unconditional + guidance_scale * (eps - eps_unconditional)
        y0_pred = (y_t - eps * (1 - at).sqrt()) / at.sqrt()                                                  # :261
        eta = 0
        c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()                                    # :263
        c2 = ((1 - at_next) - c1 ** 2).sqrt()                                                                # :264
        y_t = at_next.sqrt() * y0_pred + c1 * torch.randn_like(y_t) + c2 * eps                               # :265
    return torch.clip(y_t, -1, 1)                                                                            # :267


# --------------------------------------------------------------------------------------
# Caller contract (TrainCondition.py:53-63, diffusion/Train.py:49-55): one training step
# --------------------------------------------------------------------------------------
def train_step(trainer, optimizer, x_0, labels=None, grad_clip=1.0):
    """zero_grad -> loss -> backward -> clip -> AdamW.  Loss scaling follows the caller:
    `.sum()/1000.` unconditional (diffusion/Train.py:51), `.sum()/b**2` conditional
    (TrainCondition.py:59)."""
    optimizer.zero_grad()
    if labels is None:
        loss = trainer(x_0).sum() / 1000.
    else:
        loss = trainer(x_0, labels).sum() / x_0.shape[0] ** 2.
    loss.backward()
    torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), grad_clip)
    optimizer.step()
    return loss


def hybrid_trainer_forward(model, sqrt_alphas_bar, sqrt_one_minus_alphas_bar, T, gt_images, input_image):
    """ORACLE restatement of the hybrid trainer's forward up to the noise MSE (diffusion/Diffusion.py:54-96; the class itself
    cannot be constructed here: its ctor loads DINOv2 through torch.hub and moves it to CUDA, :49-54).  Line for line:
    :56-57 scaling, :61-62 t and noise, :63-65 q_sample, :67 six-channel input, :71-74 the 2 % coin (both branches leave the
    model's default context_zero=True in force), :89 unreduced MSE, :93-94 y_0 reconstruction.  Parity of this restatement is
    unpinned (no reference run possible); the model inside it is the reference's own DynamicUNet."""
    input_image = (input_image.float() / 255.0) * 2 - 1
    gt_images = (gt_images.float() / 255.0) * 2 - 1
    t = torch.randint(T, size=(gt_images.shape[0],), device=gt_images.device)
    noise = torch.randn_like(gt_images, dtype=torch.float32)
    y_t = (extract(sqrt_alphas_bar, t, gt_images.shape) * gt_images +
           extract(sqrt_one_minus_alphas_bar, t, gt_images.shape) * noise)
    inp = torch.cat([input_image, y_t], dim=1).float()
    if torch.rand(1) < 0.02:
        noise_pred = model(inp, t, gt_images, context_zero=True)
    else:
        noise_pred = model(inp, t, gt_images)
    mse_loss = F.mse_loss(noise_pred, noise, reduction='none')
    y_0_pred = 1 / extract(sqrt_alphas_bar, t, gt_images.shape) * (
        y_t - extract(sqrt_one_minus_alphas_bar, t, gt_images.shape) * noise_pred).float() / 255.0
    return mse_loss, y_0_pred
