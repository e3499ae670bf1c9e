"""Drop-in for `diffusion/Diffusion.py`.  The reference module holds two generations of the same two class names:

  * the plain unconditional trainer / sampler (the "Old CODE" block :286-368 that diffusion/Train.py:41,78 constructs):
    `GaussianDiffusionTrainer(model, beta_1, beta_T, T).forward(x_0)`, `GaussianDiffusionSampler(...).forward(x_T)`;
  * the live hybrid pipeline (:26-269): `forward(gt_images, input_image, stage)` and
    `forward(input_image, ddim, unconditional_guidance_scale, ddim_step)` around `DynamicUNet`.

`GaussianDiffusionTrainer` / `GaussianDiffusionSampler` here serve both callers: the hybrid signature is recognised by its
second positional argument (an image `input_image`, not a 1-d label vector) / by an image-conditioned model; `HybridGaussianDiffusion*` name it explicitly."""
from ..diffusion_process import extract  # noqa: F401
from ..diffusion_process import GaussianDiffusionSampler as _PlainSampler
from ..hybrid_process import HybridGaussianDiffusionTrainer, HybridGaussianDiffusionSampler


class GaussianDiffusionTrainer(HybridGaussianDiffusionTrainer):
    def forward(self, x_0, input_image=None, stage=0):
        # forward(x_0) / forward(x_0, labels[B] int): the plain / label-conditioned trainer; forward(gt_images, input_image[B,3,H,W],
        # stage): the hybrid one
        if input_image is None or (hasattr(input_image, "dim") and input_image.dim() == 1):
            return super(HybridGaussianDiffusionTrainer, self).forward(x_0, input_image)
        return HybridGaussianDiffusionTrainer.forward(self, x_0, input_image, stage)


class GaussianDiffusionSampler(HybridGaussianDiffusionSampler):
    def __init__(self, model, beta_1, beta_T, T, w=0.):
        _PlainSampler.__init__(self, model, beta_1, beta_T, T, w=w)
        self._img = None

    def forward(self, x, labels=None, ddim=False, unconditional_guidance_scale=1, ddim_step=None):
        if getattr(self.model, "image_cond", False):
            return HybridGaussianDiffusionSampler.forward(self, x, ddim=ddim, unconditional_guidance_scale=unconditional_guidance_scale,
                                                          ddim_step=ddim_step)
        return _PlainSampler.forward(self, x, labels, ddim=ddim, ddim_step=ddim_step)

    def _eval_model(self, x, t):
        if self._img is None:
            return self.model(x, t)
        return HybridGaussianDiffusionSampler._eval_model(self, x, t)
