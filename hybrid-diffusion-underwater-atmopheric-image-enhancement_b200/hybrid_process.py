"""The hybrid (image-conditioned) pipeline's trainer and sampler on the hdiff_b200 kernels.

Reference followed (paths relative to the reference repository):
  GaussianDiffusionTrainer   diffusion/Diffusion.py:26-180   forward(gt_images, input_image, stage) -> [loss, mse_loss, ...]
  GaussianDiffusionSampler   diffusion/Diffusion.py:182-269  forward(input_image, ddim, unconditional_guidance_scale, ddim_step)

What is reproduced: uint8 -> [-1, 1] scaling of both images (:56-57; the sampler's is /255 only, :221), t / noise draw order,
q_sample, the 6-channel model input cat([input_image, y_t]) (:67), the 2 % `context_zero=True` coin drawn from the CPU
generator (:71) — note that the reference's other branch calls `model(input, t, gt_images)` and so ALSO runs with the
model's default `context_zero=True`; that is kept —, the unreduced noise MSE (:89), the y_0 reconstruction (:93-94), the
ancestral chain (:226-238) and the DDIM chain with guidance (:241-269).

Out of scope: the perceptual / MS-SSIM / colour terms (:160-167) are third-party networks
(`torch.hub` DINOv2, kornia) — `extra_losses` takes callables `(y_0_pred, gt_images) -> tensor` for callers that have them;
their slots in the returned list are 0 otherwise.
"""
from __future__ import annotations

import torch

from . import ops as _ops
from .diffusion_process import GaussianDiffusionSampler as _Sampler
from .diffusion_process import GaussianDiffusionTrainer as _Trainer
from .diffusion_process import _MseFn, extract


def _scaled(img, scale, shift):
    """fp32 copy of a uint8 / float image batch, x * scale + shift, one kernel"""
    src = img if img.dtype == torch.uint8 else img.float()
    src = src.contiguous()
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    _ops.get().image_affine(src, out, scale, shift)
    return out


class HybridGaussianDiffusionTrainer(_Trainer):
    """forward(gt_images, input_image, stage) -> [loss, mse_loss, perceptual_dino, msssim, col_loss]; `loss` and `mse_loss`
    are the unreduced [B, 3, H, W] noise MSE (plus the extra terms, if any were supplied)."""

    def __init__(self, model, beta_1, beta_T, T, perceptual_vgg: str = "vgg16", perceptual_dino: str = "dinov2_vits14", extra_losses=None):
        super().__init__(model, beta_1, beta_T, T)
        self.num = 0
        self.stage = 0
        self.extra_losses = dict(extra_losses or {})       # name in {"perceptual_dino", "msssim", "col_loss"} -> (callable, weight)
        self.last_y_0_pred = None

    def forward(self, gt_images, input_image, stage=0):
        self.stage = stage
        input_image = _scaled(input_image, 2.0 / 255.0, -1.0)               # (x.float() / 255) * 2 - 1
        gt_images = _scaled(gt_images, 2.0 / 255.0, -1.0)
        B = gt_images.shape[0]
        t = torch.randint(self.T, size=(B,), device=gt_images.device)
        noise = torch.randn_like(gt_images, dtype=torch.float32)
        y_t = torch.empty_like(gt_images)
        _ops.get().q_sample(gt_images, noise, t, self._sab32, self._s1ab32, y_t)
        inp = torch.cat([input_image, y_t], dim=1).float()
        if torch.rand(1) < 0.02:
            noise_pred = self.model(inp, t, gt_images, context_zero=True)
        else:
            noise_pred = self.model(inp, t, gt_images)
        mse_loss = _MseFn.apply(noise_pred, noise)
        loss = mse_loss
        terms = {"perceptual_dino": 0, "msssim": 0, "col_loss": 0}
        if self.extra_losses:
            y_0_pred = 1 / extract(self.sqrt_alphas_bar, t, gt_images.shape) * (
                y_t - extract(self.sqrt_one_minus_alphas_bar, t, gt_images.shape) * noise_pred).float() / 255.0
            self.last_y_0_pred = y_0_pred
            for name, (fn, weight) in self.extra_losses.items():
                terms[name] = fn(y_0_pred, gt_images) * weight
                loss = loss + terms[name]
        return [loss, mse_loss, terms["perceptual_dino"], terms["msssim"], terms["col_loss"]]


class HybridGaussianDiffusionSampler(_Sampler):
    """forward(input_image, ddim=False, unconditional_guidance_scale=1, ddim_step=None) -> y_0 in [-1, 1].  The model sees
    cat([input_image / 255, y_t]); one time step (network + fused update) is captured in a CUDA graph and replayed."""

    def __init__(self, model, beta_1, beta_T, T):
        super().__init__(model, beta_1, beta_T, T, w=0.)
        self._img = None

    def _eval_model(self, x, t):
        inp = torch.cat([self._img, x], dim=1)
        return self.model(inp, t)

    def forward(self, input_image, ddim=False, unconditional_guidance_scale=1, ddim_step=None):
        img = _scaled(input_image, 1.0 / 255.0, 0.0)                        # the sampler does NOT map to [-1, 1] (:221)
        self._img = img
        y_T = torch.randn_like(img)
        # guidance (:257-259) evaluates the network a second time with context_zero=True; the first evaluation already ran with
        # the model's default context_zero=True on the same input, so eps_u == eps and eps_u + s (eps - eps_u) == eps exactly
        # in eval mode: the second evaluation is skipped.
        try:
            return super().forward(y_T, None, ddim=ddim, ddim_step=ddim_step)
        finally:
            self._img = None
