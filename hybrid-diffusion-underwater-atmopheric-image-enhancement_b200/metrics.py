"""Input pipeline and evaluation metrics on the GPU for 8-bit images.

Reference followed (paths relative to the reference repository):
  resize_u8      albumentations A.Resize(256, 256) [= cv2.resize(..., INTER_LINEAR)] + ToTensorV2, utils/utils.py:318-323,441-462
  psnr_u8        skimage.metrics.peak_signal_noise_ratio as imported at utils/rotinas.py:21 (data_range 255)
  uiqm_u8        metrics/metrics.py:77-299 (getUIQM = 0.0282 UICM + 0.2953 UISM + 3.5753 UIConM)
  ssim_u8        skimage.metrics.structural_similarity(a, b, channel_axis=2, data_range=255) as imported at utils/rotinas.py:22 and called
                 at :926 (metrics/metrics.py:642 passes win_size=3)
  rgb2lab_u8     cv2.cvtColor(img, cv2.COLOR_RGB2LAB) on 8-bit images, metrics/metrics.py:43
  uciqe_u8       metrics/metrics.py:40-76 (uciqe(nargin=1, loc=img) = 0.4680 var_chr + 0.2745 con_lum + 0.2576 aver_sat)
All of them call the C-ABI library (csrc/hd_metrics.cu); there is no CPU fallback."""
from __future__ import annotations

import torch

from . import _lib
from .ops import _p, _stream


def _check_u8(x, dims):
    assert x.is_cuda and x.dtype == torch.uint8 and x.dim() == dims and x.is_contiguous(), "contiguous CUDA uint8 tensor expected"


def resize_u8(images, height=256, width=256, chw=True):
    """images: uint8 [N, H, W, C] (HWC, as cv2 / albumentations hold them) -> uint8 [N, C, height, width] (chw=True: what
    ToTensorV2 returns) or [N, height, width, C].  Bit-exact with cv2.resize(img, (width, height), interpolation=cv2.INTER_LINEAR)."""
    _check_u8(images, 4)
    N, H, W, C = images.shape
    out = torch.empty((N, C, height, width) if chw else (N, height, width, C), dtype=torch.uint8, device=images.device)
    _lib.check(_lib.load().hd_resize_bilinear_u8(_p(images), N, H, W, C, _p(out), height, width, int(chw), _stream()), "hd_resize_bilinear_u8")
    return out


def psnr_u8(a, b):
    """PSNR per image (data_range 255) of two uint8 batches [N, ...] -> float64 [N]"""
    assert a.shape == b.shape
    _check_u8(a, a.dim()); _check_u8(b, b.dim())
    N = a.shape[0]
    per = a.numel() // N
    sq = torch.empty(N, dtype=torch.float64, device=a.device)
    _lib.check(_lib.load().hd_sq_err_u8(_p(a), _p(b), N, per, _p(sq), _stream()), "hd_sq_err_u8")
    return 10.0 * torch.log10(255.0 ** 2 / (sq / per))


def uiqm_u8(images):
    """images: uint8 RGB [N, H, W, 3] -> float32 [N, 4] = (UIQM, UICM, UISM, UIConM) per image"""
    _check_u8(images, 4)
    N, H, W, C = images.shape
    assert C == 3
    lib = _lib.load()
    nbytes = int(lib.hd_uiqm_workspace(N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=images.device)
    out = torch.empty((N, 4), dtype=torch.float32, device=images.device)
    _lib.check(lib.hd_uiqm_u8(_p(images), N, H, W, _p(ws), nbytes, _p(out), _stream()), "hd_uiqm_u8")
    return out


def rgb2lab_u8(images):
    """images: uint8 RGB [..., 3] -> uint8 Lab [..., 3], bit-exact with cv2.cvtColor(img, cv2.COLOR_RGB2LAB)"""
    _check_u8(images, images.dim())
    assert images.shape[-1] == 3
    out = torch.empty_like(images)
    _lib.check(_lib.load().hd_rgb2lab_u8(_p(images), images.numel() // 3, _p(out), _stream()), "hd_rgb2lab_u8")
    return out


def uciqe_u8(images):
    """images: uint8 RGB [N, H, W, 3] -> float64 [N, 4] = (UCIQE, var_chr, con_lum, aver_sat) per image"""
    _check_u8(images, 4)
    N, H, W, C = images.shape
    assert C == 3
    lib = _lib.load()
    nbytes = int(lib.hd_uciqe_workspace(N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=images.device)
    out = torch.empty((N, 4), dtype=torch.float64, device=images.device)
    _lib.check(lib.hd_uciqe_u8(_p(images), N, H, W, _p(ws), nbytes, _p(out), _stream()), "hd_uciqe_u8")
    return out


def ssim_u8(a, b, win_size=7):
    """SSIM per image of two uint8 batches [N, H, W, C] -> float64 [N] (uniform window, sample covariance, data_range 255, cropped borders)"""
    assert a.shape == b.shape
    _check_u8(a, 4); _check_u8(b, 4)
    N, H, W, C = a.shape
    out = torch.empty(N, dtype=torch.float64, device=a.device)
    _lib.check(_lib.load().hd_ssim_u8(_p(a), _p(b), N, H, W, C, int(win_size), _p(out), _stream()), "hd_ssim_u8")
    return out
