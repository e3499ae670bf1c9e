"""Drop-in for the plain unconditional trainer / sampler of `diffusion/Diffusion.py` (the "Old CODE"
block :286-368 that diffusion/Train.py:41,78 constructs): `forward(x_0)` / `forward(x_T)`."""
from ..diffusion_process import extract, GaussianDiffusionTrainer, GaussianDiffusionSampler  # noqa: F401
