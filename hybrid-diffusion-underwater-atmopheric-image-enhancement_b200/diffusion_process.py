"""DDPM forward-noising trainer (Algorithm 1) and classifier-free-guidance ancestral sampler
(Algorithm 2) on the hdiff_b200 kernels.

Reference followed (paths relative to the reference repository):
  extract                      DiffusionFreeGuidence/DiffusionCondition.py:9-16   (dup diffusion/Diffusion.py:16-23)
  GaussianDiffusionTrainer     DiffusionFreeGuidence/DiffusionCondition.py:19-46  (uncond twin diffusion/Diffusion.py:304-314)
  GaussianDiffusionSampler     DiffusionFreeGuidence/DiffusionCondition.py:49-98  (uncond twin diffusion/Diffusion.py:351-368)

  DDIM (eta = 0) with guidance  diffusion/Diffusion.py:241-269 (the hybrid sampler's `ddim=True` branch), SURVEY §8(f) rank 2

Differences kept deliberately: the per-step `print` (:88) is dropped and the per-step NaN
assertion (:96, a device->host sync) becomes a device flag checked once after the last step.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops as _ops


def extract(v, t, x_shape):
    """gather(v, t) -> fp32 -> [B,1,1,...]  (DiffusionCondition.py:9-16)."""
    out = torch.gather(v, index=t, dim=0).float().to(t.device)
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


def schedule_tables(beta_1, beta_T, T):
    """float64 tables; linspace is evaluated in fp32 and only then cast (DiffusionCondition.py:26-35,58-66)."""
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
        "alphas_bar": alphas_bar,
    }


class _MseFn(torch.autograd.Function):
    """F.mse_loss(pred, noise, reduction='none') with its gradient, one kernel each (K6)."""

    @staticmethod
    def forward(ctx, pred, noise):
        pred = pred.contiguous()
        loss = torch.empty_like(pred)
        _ops.get().mse_fwd(pred, noise, loss)
        ctx.save_for_backward(pred, noise)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, noise = ctx.saved_tensors
        g = g.expand_as(pred).contiguous()
        d = torch.empty_like(pred)
        _ops.get().mse_bwd(pred, noise, g, d)
        return d, None


class GaussianDiffusionTrainer(nn.Module):
    """GaussianDiffusionTrainer(model, beta_1, beta_T, T).forward(x_0[, labels]) -> unreduced loss [B,3,H,W].
    RNG order as in the reference: randint, then randn_like, both on x_0's device generator."""

    def __init__(self, model, beta_1, beta_T, T):
        super().__init__()
        self.model = model
        self.T = T
        tab = schedule_tables(beta_1, beta_T, T)
        self.register_buffer('betas', tab["betas"])
        self.register_buffer('sqrt_alphas_bar', tab["sqrt_alphas_bar"])
        self.register_buffer('sqrt_one_minus_alphas_bar', tab["sqrt_one_minus_alphas_bar"])
        # what extract() hands to the arithmetic: the f64 entries rounded to fp32
        self.register_buffer('_sab32', tab["sqrt_alphas_bar"].float(), persistent=False)
        self.register_buffer('_s1ab32', tab["sqrt_one_minus_alphas_bar"].float(), persistent=False)

    def forward(self, x_0, labels=None):
        assert x_0.dtype == torch.float32
        x_0 = x_0.contiguous()
        t = torch.randint(self.T, size=(x_0.shape[0], ), device=x_0.device)
        noise = torch.randn_like(x_0)
        x_t = torch.empty_like(x_0)
        _ops.get().q_sample(x_0, noise, t, self._sab32, self._s1ab32, x_t)
        pred = self.model(x_t, t) if labels is None else self.model(x_t, t, labels)
        return _MseFn.apply(pred, noise)


class GaussianDiffusionSampler(nn.Module):
    """GaussianDiffusionSampler(model, beta_1, beta_T, T[, w]).forward(x_T[, labels]) -> x_0 in [-1, 1].
    With labels, each step evaluates the conditional and the null-label (0) network as ONE batch of 2B."""

    def __init__(self, model, beta_1, beta_T, T, w=0.):
        super().__init__()
        self.model = model
        self.T = T
        self.w = w
        tab = schedule_tables(beta_1, beta_T, T)
        self.register_buffer('betas', tab["betas"])
        self.register_buffer('coeff1', tab["coeff1"])
        self.register_buffer('coeff2', tab["coeff2"])
        self.register_buffer('posterior_var', tab["posterior_var"])
        var = torch.cat([tab["posterior_var"][1:2], tab["betas"][1:]])          # DiffusionCondition.py:74
        coef = torch.stack([tab["coeff1"].float(), tab["coeff2"].float(), torch.sqrt(var.float())], dim=1).contiguous()
        self.register_buffer('_coef', coef, persistent=False)                    # [T][3] fp32
        self.register_buffer('alphas_bar', tab["alphas_bar"])
        self._ddim = {}                                                          # ddim_step -> (coef table [n][3] fp32, stride)

    def ddim_tables(self, ddim_step):
        """Coefficients of the deterministic DDIM update of diffusion/Diffusion.py:241-269 (eta = 0):
            seq = range(0, T, T // ddim_step), visited from its end; at step i (previous entry j, -1 before the first):
            at = alphas_bar[i + 1], at_next = alphas_bar[j + 1]                      (:251-252, the reference indexes at t + 1)
            y0 = (y_t - eps sqrt(1 - at)) / sqrt(at);  y_{next} = sqrt(at_next) y0 + sqrt(1 - at_next) eps   (:261-265, c1 = 0)
        which is the linear update  y_next = k1 y_t - k2 eps  that hd_sampler_step applies, with
            k1 = sqrt(at_next) / sqrt(at),   k2 = sqrt(at_next) sqrt(1 - at) / sqrt(at) - sqrt(1 - at_next)
        and no noise term.  Returns (table [len(seq)][3] fp32 in seq order, stride); the reference hard-codes 1000 for T."""
        if ddim_step is None:
            raise ValueError("ddim=True needs ddim_step (number of DDIM steps, e.g. 100; the reference divides by it, Diffusion.py:246)")
        ddim_step = int(ddim_step)
        if ddim_step not in self._ddim:
            if ddim_step < 1 or ddim_step > self.T:
                raise ValueError(f"ddim_step must be in [1, T={self.T}]")
            stride = int(self.T / ddim_step)
            seq = list(range(0, self.T, stride))
            seq_next = [-1] + seq[:-1]
            if seq[-1] + 1 >= self.T:
                raise ValueError(f"ddim_step={ddim_step} visits t={seq[-1]} and reads alphas_bar[t + 1] (Diffusion.py:251), which is out "
                                 f"of range for T={self.T}; use a ddim_step whose stride T // ddim_step is >= 2")
            ab = self.alphas_bar
            at = ab[torch.tensor([i + 1 for i in seq], device=ab.device)].float()            # extract(...).float()
            an = ab[torch.tensor([j + 1 for j in seq_next], device=ab.device)].float()
            k1 = an.sqrt() / at.sqrt()
            k2 = an.sqrt() * (1 - at).sqrt() / at.sqrt() - (1 - an).sqrt()
            self._ddim[ddim_step] = (torch.stack([k1, k2, torch.zeros_like(k1)], dim=1).contiguous(), stride)
        return self._ddim[ddim_step]

    def predict_xt_prev_mean_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return extract(self.coeff1, t, x_t.shape) * x_t - extract(self.coeff2, t, x_t.shape) * eps

    def _eval_model(self, x, t):
        """the unconditional network evaluation of one step (the hybrid sampler prepends its conditioning image here)"""
        return self.model(x, t)

    def _eps_pair(self, x, t, labels):
        """(eps_cond, eps_uncond) — batched as 2B when the model accepts it."""
        B = x.shape[0]
        if getattr(self.model, "cfg_batched", True):
            e = self.model(torch.cat([x, x], 0), torch.cat([t, t], 0), torch.cat([labels, torch.zeros_like(labels)], 0))
            return e[:B], e[B:]
        return self.model(x, t, labels), self.model(x, t, torch.zeros_like(labels))

    use_cuda_graph = True      # capture one time step (2B-batch UNet forward + fused update) and replay it T-1 times

    def _one_step(self, x, labels, step, nan_flag, coef=None, stride=1):
        """One ancestral step, in place on x.  The time index lives in device memory (`step`), so the same launch
        sequence serves every t (DiffusionCondition.py:86-95 with the per-step print / host NaN check removed).
        `coef` / `stride`: another update table indexed by `step`, with the network evaluated at t = step * stride (DDIM)."""
        ops = _ops.get()
        B = x.shape[0]
        t = (step.to(torch.int64) * stride).expand(B).contiguous()
        if labels is None:
            eps_c, eps_u = self._eval_model(x, t), None
        else:
            eps_c, eps_u = self._eps_pair(x, t, labels)
            eps_u = eps_u.contiguous()
        eps_c = eps_c.contiguous()
        assert eps_c.shape == x.shape
        # ancestral: ignored by the kernel at t == 0 (the reference adds no noise there).  DDIM: the reference draws
        # `c1 * randn_like` with c1 = 0 every step (Diffusion.py:264-265) — the draw is kept so that the generator advances as
        # the reference's does, the coefficient table's noise column is zero.
        z = torch.randn_like(x)
        ops.sampler_step(x, eps_c, eps_u, z, self.w, self._coef if coef is None else coef, step, True, nan_flag)
        ops.add_int(step, -1)

    def run_steps(self, x, labels, step, nan_flag, n_steps, coef=None, stride=1):
        """Advance `n_steps` time steps starting at the index held in `step` (bench.py times a bounded number of steps)."""
        frozen = self.model.frozen_weights() if hasattr(self.model, "frozen_weights") else _Null()
        graphable = self.use_cuda_graph and x.is_cuda and n_steps > 2 and hasattr(self.model, "frozen_weights")
        with torch.no_grad(), frozen:
            if not graphable:
                for _ in range(n_steps):
                    self._one_step(x, labels, step, nan_flag, coef, stride)
                return
            # the captured step is bound to these buffers; callers that keep them (bench.py, chunked sampling) reuse it
            key = (x.data_ptr(), tuple(x.shape), None if labels is None else labels.data_ptr(), step.data_ptr(), nan_flag.data_ptr(),
                   None if coef is None else coef.data_ptr(), stride)
            cached = getattr(self, "_graph", None)
            if cached is None or cached[0] != key:
                self._one_step(x, labels, step, nan_flag, coef, stride)    # eager: also initialises every lazily built state
                n_steps -= 1
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._one_step(x, labels, step, nan_flag, coef, stride)
                self._graph = cached = (key, graph)
            for _ in range(n_steps):
                cached[1].replay()

    def forward(self, x_T, labels=None, ddim=False, ddim_step=None):
        """ddim=True: the deterministic DDIM sampler of the hybrid pipeline (diffusion/Diffusion.py:241-269, eta = 0) over
        `ddim_step` of the T time steps, with the same classifier-free guidance (its `unconditional_guidance_scale` is
        1 + w) — ten times fewer network evaluations at ddim_step = 100."""
        assert x_T.dtype == torch.float32
        dev = x_T.device
        x = x_T.clone().contiguous()
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        if ddim:
            coef, stride = self.ddim_tables(ddim_step)
            coef = coef.to(dev)
            n = coef.shape[0]
            step = torch.full((1,), n - 1, dtype=torch.int32, device=dev)
            self.run_steps(x, labels, step, nan_flag, n, coef, stride)
            self._graph = None
            assert int(nan_flag.item()) == 0, "nan in tensor."
            return x
        step = torch.full((1,), self.T - 1, dtype=torch.int32, device=dev)
        self.run_steps(x, labels, step, nan_flag, self.T)
        self._graph = None                       # bound to this call's buffers
        assert int(nan_flag.item()) == 0, "nan in tensor."
        return x


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
