"""Drop-in for `DiffusionFreeGuidence/DiffusionCondition.py` (extract :9, GaussianDiffusionTrainer :19-46,
GaussianDiffusionSampler :49-98): `forward(x_0, labels)` / `forward(x_T, labels)` with guidance weight `w`."""
from ..diffusion_process import extract, GaussianDiffusionTrainer, GaussianDiffusionSampler  # noqa: F401
