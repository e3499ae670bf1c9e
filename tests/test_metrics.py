"""Input pipeline / metrics: the oracle restatements pinned on CPU, the CUDA kernels against them on GPU."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, ref_metrics

cv2 = pytest.importorskip("cv2")


def test_resize_restatement_is_bit_exact_with_cv2():
    rng = np.random.default_rng(0)
    for (sh, sw, dh, dw) in [(300, 400, 256, 256), (512, 512, 256, 256), (100, 77, 256, 256), (256, 256, 256, 256), (480, 640, 128, 96)]:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref_metrics.resize_bilinear_u8(src, dh, dw), want), (sh, sw, dh, dw)


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_reference_uiqm_extraction_runs():
    ns = ref_metrics.reference_uiqm()
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (32, 40, 3), dtype=np.uint8)
    v = ns["getUIQM"](img)
    assert np.isfinite(v)
    assert abs(v - (0.0282 * ns["_uicm"](img.astype(np.float32)) + 0.2953 * ns["_uism"](img.astype(np.float32))
                    + 3.5753 * ns["_uiconm"](img.astype(np.float32), 8))) < 1e-9


@pytest.mark.gpu
def test_resize_kernel_is_bit_exact_with_cv2():
    from hdiff_b200 import metrics
    rng = np.random.default_rng(2)
    for (sh, sw, dh, dw) in [(300, 400, 256, 256), (720, 1280, 256, 256), (100, 77, 256, 256), (256, 256, 256, 256), (333, 517, 128, 96)]:
        src = rng.integers(0, 256, (3, sh, sw, 3), dtype=np.uint8)
        want = np.stack([cv2.resize(s, (dw, dh), interpolation=cv2.INTER_LINEAR) for s in src])
        x = torch.from_numpy(src).cuda()
        got = metrics.resize_u8(x, dh, dw, chw=False).cpu().numpy()
        assert np.array_equal(got, want), (sh, sw, dh, dw, np.abs(got.astype(int) - want).max())
        got_chw = metrics.resize_u8(x, dh, dw, chw=True).cpu().numpy()
        assert np.array_equal(got_chw, want.transpose(0, 3, 1, 2))          # ToTensorV2


@pytest.mark.gpu
def test_psnr_kernel():
    from hdiff_b200 import metrics
    g = torch.Generator().manual_seed(3)
    a = torch.randint(0, 256, (4, 64, 48, 3), generator=g, dtype=torch.uint8)
    b = (a.int() + torch.randint(-9, 10, a.shape, generator=g)).clamp(0, 255).to(torch.uint8)
    want = [10 * np.log10(255.0 ** 2 / np.mean((a[i].numpy().astype(np.float64) - b[i].numpy()) ** 2)) for i in range(4)]
    got = metrics.psnr_u8(a.cuda(), b.cuda()).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_uiqm_kernel_vs_reference_getUIQM():
    """The reference's own getUIQM (Python loops, sorts, scipy Sobel) on natural-looking and on random 8-bit images."""
    from hdiff_b200 import metrics
    ns = ref_metrics.reference_uiqm()
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:96, 0:128]
    smooth = np.stack([127 + 100 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], -1)
    imgs = np.stack([np.clip(smooth + rng.normal(0, 12, smooth.shape), 0, 255).astype(np.uint8),
                     rng.integers(0, 256, (96, 128, 3), dtype=np.uint8),
                     np.clip(smooth * 0.5 + 40, 0, 255).astype(np.uint8)])
    got = metrics.uiqm_u8(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i, img in enumerate(imgs):
        x = img.astype(np.float32)
        want = (ns["getUIQM"](img), ns["_uicm"](x), ns["_uism"](x), ns["_uiconm"](x, 8))
        assert np.allclose(got[i], want, rtol=2e-4, atol=1e-5), (i, got[i], want)
