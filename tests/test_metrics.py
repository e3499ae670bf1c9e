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


def _all_colours():
    v = np.arange(256, dtype=np.uint8)
    return np.ascontiguousarray(np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(4096, 4096, 3))


def test_rgb2lab_restatement_is_bit_exact_with_cv2_on_all_colours():
    allc = _all_colours()
    want = cv2.cvtColor(allc, cv2.COLOR_RGB2LAB)
    assert np.array_equal(ref_metrics.rgb2lab_u8(allc), want)


def test_library_lab_tables_equal_the_oracle_tables():
    """host-only entry point (no GPU): the tables the kernels use are the ones pinned above"""
    import ctypes
    from hdiff_b200 import _lib
    gamma = (ctypes.c_uint16 * 256)()
    cb = (ctypes.c_uint16 * 3072)()
    assert _lib.load().hd_lab_tables_host(gamma, cb) == 0
    g, c = ref_metrics.lab_tables()
    assert np.array_equal(np.frombuffer(gamma, np.uint16), g) and np.array_equal(np.frombuffer(cb, np.uint16), c)


@pytest.mark.gpu
def test_rgb2lab_kernel_is_bit_exact_with_cv2_on_all_colours():
    from hdiff_b200 import metrics
    allc = _all_colours()
    want = cv2.cvtColor(allc, cv2.COLOR_RGB2LAB)
    got = metrics.rgb2lab_u8(torch.from_numpy(allc).cuda()).cpu().numpy()
    assert np.array_equal(got, want)


def _uciqe_images():
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:120, 0:160]
    smooth = np.stack([127 + 100 * np.sin(xx / 11.0 + c) * np.cos(yy / 8.0 - c) for c in range(3)], -1)
    water = np.clip(smooth * np.array([0.25, 0.6, 0.8]) + np.array([10, 60, 90]) + rng.normal(0, 6, smooth.shape), 0, 255)
    return np.stack([np.clip(smooth + rng.normal(0, 12, smooth.shape), 0, 255).astype(np.uint8),
                     rng.integers(0, 256, (120, 160, 3), dtype=np.uint8),
                     water.astype(np.uint8),
                     np.full((120, 160, 3), (40, 90, 160), dtype=np.uint8)])           # one colour: the empty-range histogram branch


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_uciqe_restatement_vs_reference_uciqe():
    """the kernels' algorithm (256-bin histogram instead of np.histogram's 65536 bins) against the reference's own function"""
    uciqe = ref_metrics.reference_uciqe()
    rng = np.random.default_rng(6)
    extra = [np.clip(rng.normal(rng.integers(30, 220), rng.integers(1, 60), (48, 64, 3)), 0, 255).astype(np.uint8) for _ in range(12)]
    for img in list(_uciqe_images()) + extra:
        with np.errstate(all="ignore"):
            want = uciqe(nargin=1, loc=img)
        got = ref_metrics.uciqe_restated(img)
        assert abs(got[0] - want) <= 1e-14 * abs(want), (got, want)


@pytest.mark.gpu
@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_uciqe_kernel_vs_reference_uciqe():
    """The reference's own uciqe (cv2 Lab conversion, numpy float64, np.histogram with 65536 bins): the per-pixel arithmetic and the
    histogram bins are reproduced exactly, only the order of the float64 sums differs."""
    from hdiff_b200 import metrics
    uciqe = ref_metrics.reference_uciqe()
    imgs = _uciqe_images()
    got = metrics.uciqe_u8(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i, img in enumerate(imgs):
        with np.errstate(all="ignore"):
            want = uciqe(nargin=1, loc=img)
        # image 3 is one colour: var_chr = sqrt(mean |1 - (mean(chroma) / chroma)^2|) is the square root of pure rounding noise there
        # (1e-16 -> 1e-8 in either implementation), everywhere else the sums agree to the last few bits
        tol = dict(rtol=1e-11, atol=0) if i != 3 else dict(rtol=0, atol=1e-7)
        assert np.isclose(got[i, 0], want, **tol), (i, got[i], want)
        lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB) / 255
        chr_ = np.sqrt(lab[..., 1] ** 2 + lab[..., 2] ** 2)
        assert np.isclose(got[i, 3], np.mean(chr_ / np.sqrt(chr_ ** 2 + lab[..., 0] ** 2)), rtol=1e-12)
        assert np.isclose(got[i, 1], np.sqrt(np.mean(abs(1 - np.square(np.mean(chr_) / chr_)))), **tol)
        assert np.isclose(got[i, 2], ref_metrics.uciqe_restated(img)[2], rtol=0, atol=1e-15), (i, got[i, 2])     # the histogram bins are exact


def _ssim_pairs():
    rng = np.random.default_rng(7)
    imgs = _uciqe_images()[:3]
    noisy = [np.clip(im.astype(np.int64) + rng.integers(-s, s + 1, im.shape), 0, 255).astype(np.uint8) for im, s in zip(imgs, (4, 40, 12))]
    return list(zip(imgs, noisy)) + [(imgs[0], imgs[0]), (imgs[0], imgs[2])]


def test_ssim_restatement_against_the_window_definition():
    """SSIM has no pinned reference here (scikit-image is absent: parity unpinned); the restatement is at least checked against a direct
    evaluation of the definition — per interior pixel, plain sums over its win x win window — and on identical images."""
    for win in (7, 3):
        for a, b in _ssim_pairs()[:2] + _ssim_pairs()[3:4]:
            a, b = a[:40, :48], b[:40, :48]
            pad = (win - 1) // 2
            tot = []
            for c in range(3):
                x, y = a[..., c].astype(np.float64), b[..., c].astype(np.float64)
                vals = []
                for i in range(pad, x.shape[0] - pad):
                    for j in range(pad, x.shape[1] - pad):
                        wx, wy = x[i - pad:i + pad + 1, j - pad:j + pad + 1], y[i - pad:i + pad + 1, j - pad:j + pad + 1]
                        ux, uy = wx.mean(), wy.mean()
                        n = win * win
                        vx, vy = ((wx - ux) ** 2).sum() / (n - 1), ((wy - uy) ** 2).sum() / (n - 1)
                        vxy = ((wx - ux) * (wy - uy)).sum() / (n - 1)
                        c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
                        vals.append((2 * ux * uy + c1) * (2 * vxy + c2) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2)))
                tot.append(np.mean(vals))
            assert abs(ref_metrics.ssim_restated(a, b, win) - np.mean(tot)) < 1e-10
    a = _ssim_pairs()[0][0]
    assert abs(ref_metrics.ssim_restated(a, a) - 1.0) < 1e-12


@pytest.mark.gpu
def test_ssim_kernel_vs_restatement():
    from hdiff_b200 import metrics
    pairs = _ssim_pairs()
    a = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    b = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    for win in (7, 3):
        got = metrics.ssim_u8(a, b, win_size=win).cpu().numpy()
        want = [ref_metrics.ssim_restated(p, q, win) for p, q in pairs]
        assert np.allclose(got, want, rtol=1e-10, atol=1e-12), (win, got, want)
    assert abs(got[3] - 1.0) < 1e-12
