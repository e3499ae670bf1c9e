"""ORACLE support (test infrastructure): put the reference's own source files where the GPU box can import them.

`/root/reference` exists only in the build container.  This recipe copies, byte for byte, the few files of the hot path
into `oracle/_ref/` — git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored,
so the directory travels to the GPU box with the snapshot like a built `.so`.  `oracle/ref_loader.py` imports from
/root/reference when it exists and from `oracle/_ref/` otherwise, so `bench.py --impl reference` and `cpu_baseline` time
the reference's OWN classes there (`cpu_baseline.kind = "reference"`).  `__graft_entry__.build()` runs this.

    python oracle/make_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("HDIFF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "DiffusionFreeGuidence/DiffusionCondition.py",      # extract, GaussianDiffusionTrainer, GaussianDiffusionSampler
    "DiffusionFreeGuidence/ModelCondition.py",          # blocks + live UNet (one-token SyntaxError repaired in memory by ref_loader)
    "diffusion/Model.py",                               # blocks + DynamicUNet
    "diffusion/Diffusion.py",                           # hybrid trainer / sampler (cut out by AST, never imported whole)
    "metrics/metrics.py",                               # image-quality metrics (cut out by AST)
    "LICENSE",
]


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: nothing copied (oracle/_ref keeps what it has)")
        return False
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copies of reference source files, made by oracle/make_ref.py for the GPU box. Not tracked by git.\n")
    if verbose:
        print(f"copied {len(FILES)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 0)
