"""ORACLE support (test infrastructure): import the *unmodified* reference by file path.

Looks in /root/reference (the build container) and then in `oracle/_ref/` (verbatim copies made by
`oracle/make_ref.py`; git-ignored, shipped to the GPU box with the snapshot).  Used by `oracle/make_golden.py`
and the tests that pin `oracle/ref_torch.py` against the reference (they skip when neither tree exists), and by
`bench.py`'s CPU baseline / `--impl reference`, which time the reference's own classes through it.

Facts handled here:
  F3  DiffusionFreeGuidence/ModelCondition.py:289 has a one-token SyntaxError in the unused
      `DynamicUNet` class; it is repaired in memory, the file on disk is never touched.
  F1  `UNet(T, ch, ch_mult, attn, num_res_blocks, dropout)` is not defined in the reference;
      `assemble_unet` builds it from the reference's own block classes.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """/root/reference in the build container; on the GPU box the verbatim copies that oracle/make_ref.py left in oracle/_ref."""
    env = os.environ.get("HDIFF_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if os.path.isfile(os.path.join(cand, "DiffusionFreeGuidence", "DiffusionCondition.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "DiffusionFreeGuidence", "DiffusionCondition.py"))


def _load_by_path(name: str, rel: str, patch=None) -> types.ModuleType:
    path = os.path.join(REF_ROOT, rel)
    if patch is None:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
        return mod
    with open(path, "r") as f:
        src = patch(f.read())
    mod = types.ModuleType(name)
    mod.__file__ = path
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def _repair_model_condition(src: str) -> str:
    import re
    fixed, n = re.subn(r"pa\s+dding=1", "padding=1", src)
    assert n == 1, f"expected exactly one broken token, found {n}"
    return fixed


_cache = {}


def diffusion_condition():
    """DiffusionFreeGuidence/DiffusionCondition.py, as is."""
    if "dc" not in _cache:
        _cache["dc"] = _load_by_path("_ref_DiffusionCondition", "DiffusionFreeGuidence/DiffusionCondition.py")
    return _cache["dc"]


def model_condition():
    """DiffusionFreeGuidence/ModelCondition.py with the :289 token repaired in memory."""
    if "mc" not in _cache:
        if "telnetlib" not in sys.modules:
            try:
                import telnetlib  # noqa: F401  (ModelCondition.py:4; gone in 3.13)
            except Exception:
                shim = types.ModuleType("telnetlib")
                shim.PRAGMA_HEARTBEAT = None
                sys.modules["telnetlib"] = shim
        _cache["mc"] = _load_by_path("_ref_ModelCondition", "DiffusionFreeGuidence/ModelCondition.py",
                                     patch=_repair_model_condition)
    return _cache["mc"]


def diffusion_model():
    """diffusion/Model.py, as is (DynamicUNet :382-517, the image ConditionalEmbedding :110-167, the MHA ResBlock :267-312)."""
    if "dm" not in _cache:
        if "telnetlib" not in sys.modules:
            try:
                import telnetlib  # noqa: F401
            except Exception:
                sys.modules["telnetlib"] = types.ModuleType("telnetlib")
        _cache["dm"] = _load_by_path("_ref_diffusion_Model", "diffusion/Model.py")
    return _cache["dm"]


def assemble_unet(T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None) -> nn.Module:
    """The F1 composition, made only of reference classes (ResBlock_old, AttnBlock, DownSample,
    UpSample, TimeEmbedding, ConditionalEmbedding) in the topology of ModelCondition.py:213-276."""
    mc = model_condition()

    class RefUNet(nn.Module):
        def __init__(self):
            super().__init__()
            tdim = ch * 4
            self.time_embedding = mc.TimeEmbedding(T, ch, tdim)
            if num_labels is not None:
                self.cond_embedding = mc.ConditionalEmbedding(num_labels, ch, tdim)
            self.head = nn.Conv2d(3, ch, kernel_size=3, stride=1, padding=1)
            self.downblocks = nn.ModuleList()
            chs = [ch]
            now_ch = ch
            for i, mult in enumerate(ch_mult):
                out_ch = ch * mult
                for _ in range(num_res_blocks):
                    self.downblocks.append(mc.ResBlock_old(now_ch, out_ch, tdim, dropout, attn=(i in attn)))
                    now_ch = out_ch
                    chs.append(now_ch)
                if i != len(ch_mult) - 1:
                    self.downblocks.append(mc.DownSample(now_ch))
                    chs.append(now_ch)
            self.middleblocks = nn.ModuleList([
                mc.ResBlock_old(now_ch, now_ch, tdim, dropout, attn=True),
                mc.ResBlock_old(now_ch, now_ch, tdim, dropout, attn=False),
            ])
            self.upblocks = nn.ModuleList()
            for i, mult in reversed(list(enumerate(ch_mult))):
                out_ch = ch * mult
                for _ in range(num_res_blocks + 1):
                    self.upblocks.append(mc.ResBlock_old(chs.pop() + now_ch, out_ch, tdim, dropout, attn=False))
                    now_ch = out_ch
                if i != 0:
                    self.upblocks.append(mc.UpSample(now_ch))
            assert len(chs) == 0
            self.tail = nn.Sequential(nn.GroupNorm(32, now_ch), mc.Swish(),
                                      nn.Conv2d(now_ch, 3, 3, stride=1, padding=1))

        def forward(self, x, t, labels=None):
            temb = self.time_embedding(t)
            if labels is not None:
                cemb = self.cond_embedding(labels)
            else:
                # unconditional: ResBlock_old.forward needs a tensor; a zero cemb through
                # Swish->Linear would still add the cond_proj bias, so bypass cond_proj
                # exactly as ResBlock.forward(cemb=None) does (ModelCondition.py:199-200).
                cemb = None
            h = self.head(x)
            hs = [h]
            for layer in self.downblocks:
                h = _call(layer, h, temb, cemb)
                hs.append(h)
            for layer in self.middleblocks:
                h = _call(layer, h, temb, cemb)
            for layer in self.upblocks:
                if isinstance(layer, mc.ResBlock_old):
                    h = torch.cat([h, hs.pop()], dim=1)
                h = _call(layer, h, temb, cemb)
            assert len(hs) == 0
            return self.tail(h)

    def _call(layer, h, temb, cemb):
        if isinstance(layer, mc.ResBlock_old) and cemb is None:
            y = layer.block1(h)
            y = y + layer.temb_proj(temb)[:, :, None, None]
            y = layer.block2(y)
            y = y + layer.shortcut(h)
            return layer.attn(y)
        return layer(h, temb, cemb)

    return RefUNet()


def hybrid_sampler_class():
    """`GaussianDiffusionSampler` of diffusion/Diffusion.py:181-269 (the hybrid pipeline's sampler, with the DDIM branch
    :241-269) and its `extract` (:16-23).  The module itself cannot be imported here (it needs kornia / lpips and reaches for
    torch.hub + CUDA in another class), so the two definitions are cut out of the unmodified source by AST and executed on
    their own."""
    if "hybrid" not in _cache:
        import ast
        import torch.nn.functional as F
        path = os.path.join(REF_ROOT, "diffusion", "Diffusion.py")
        with open(path, "r") as f:
            tree = ast.parse(f.read())
        keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "extract")
                or (isinstance(n, ast.ClassDef) and n.name == "GaussianDiffusionSampler")]
        assert len(keep) == 2, [getattr(n, "name", None) for n in keep]
        ns = {"torch": torch, "nn": nn, "F": F}
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
        _cache["hybrid"] = ns["GaussianDiffusionSampler"]
    return _cache["hybrid"]
