"""Importable alias of the product package.

The product directory is named after the reference repository
(`hybrid-diffusion-underwater-atmopheric-image-enhancement_b200/`), which is not a valid Python
identifier; `import hdiff_b200` exposes its modules (`hdiff_b200.ops`, `hdiff_b200.diffusion.Model`,
`hdiff_b200.DiffusionFreeGuidence.ModelCondition`, ...)."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "hybrid-diffusion-underwater-atmopheric-image-enhancement_b200")
__path__.insert(0, _PKG_DIR)
PKG_DIR = _PKG_DIR
