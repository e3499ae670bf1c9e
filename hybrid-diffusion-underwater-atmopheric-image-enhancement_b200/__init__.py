"""hdiff_b200 — B200-native (sm_100a) DDPM / classifier-free-guidance hot path behind the reference's
Python API.  Import through the alias package `hdiff_b200`:

    from hdiff_b200.diffusion.Model import UNet
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer, GaussianDiffusionSampler
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    from hdiff_b200.DiffusionFreeGuidence.DiffusionCondition import GaussianDiffusionTrainer, GaussianDiffusionSampler
"""
