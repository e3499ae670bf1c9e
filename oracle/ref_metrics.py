"""ORACLE (test infrastructure; nothing in the product package imports this): CPU restatements for the input-pipeline /
metric kernels of csrc/hd_metrics.cu.

  resize_bilinear_u8   numpy restatement of OpenCV's 8-bit INTER_LINEAR resize (the algorithm lives in the third-party
                       dependency opencv-python — albumentations' A.Resize calls cv2.resize, utils/utils.py:318-323 —: 11-bit
                       fixed-point weights `cvRound(w * 2048)`, horizontal pass in int, vertical pass
                       `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`).  Pinned bit-for-bit against cv2.resize
                       itself (opencv-python 4.13 in this image) by tests/test_metrics.py.
  rgb2lab_u8           numpy restatement of OpenCV's 8-bit RGB -> Lab (cv2.cvtColor(img, cv2.COLOR_RGB2LAB), the call at
                       metrics/metrics.py:43; algorithm in the third-party dependency opencv-python, `RGB2Lab_b` of
                       imgproc/src/color_lab.cpp: sRGB gamma table scaled by 255 * 8, 12-bit XYZ coefficients divided by the D65
                       white point, Lab f() table scaled by 2^15, `CV_DESCALE` roundings).  Pinned bit-for-bit against cv2 itself
                       over ALL 2^24 colours by tests/test_metrics.py.
  uciqe_restated       the algorithm of csrc/hd_metrics.cu's UCIQE kernels in numpy float64: per-pixel terms as in the reference,
                       np.histogram(lum, 65536) + cumsum reproduced from a 256-bin integer histogram of L.  Pinned against the
                       reference's own uciqe by tests/test_metrics.py.
  ssim_restated        scikit-image's structural_similarity for the call the reference makes (utils/rotinas.py:926:
                       SSIM(res, gt, channel_axis=2, data_range=255); metrics/metrics.py:642 with win_size=3).  PARITY UNPINNED: the algorithm
                       lives in the third-party dependency scikit-image (reference environment: CLEDiff_bkp.yaml), which is NOT installed in
                       this image, and the reference holds no golden values for it.  Restated from the published algorithm (Wang et al. 2004
                       as implemented in skimage/metrics/_structural_similarity.py): float64, scipy.ndimage.uniform_filter (the function
                       skimage itself calls), sample covariance, K1 = 0.01, K2 = 0.03, borders of (win - 1) // 2 cropped, mean over pixels
                       per channel, then over channels.
  reference_uciqe      the reference's OWN uciqe (metrics/metrics.py:40-76), cut out of the unmodified file by AST.
  reference_uiqm       the reference's OWN getUIQM (metrics/metrics.py:77-299), cut out of the unmodified file by AST (the module
                       imports torchvision's Inception and skimage at the top and cannot be imported whole here).
"""
import math
import os

import numpy as np


def resize_bilinear_u8(src, dh, dw):
    sh, sw, _ = src.shape

    def coefs(dn, sn):
        scale = sn / dn
        ofs = np.zeros(dn, np.int64)
        a = np.zeros((dn, 2), np.int64)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            ofs[d] = s
            a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        return ofs, a

    xofs, xa = coefs(dw, sw)
    for d in range(dw):                       # horizontal: the weight is zeroed at the borders
        if xofs[d] < 0:
            xofs[d], xa[d] = 0, (2048, 0)
        if xofs[d] >= sw - 1:
            xofs[d], xa[d] = sw - 1, (2048, 0)
    yofs, ya = coefs(dh, sh)
    S = src.astype(np.int64)
    x1 = np.minimum(xofs + 1, sw - 1)
    Hh = S[:, xofs, :] * xa[:, 0][None, :, None] + S[:, x1, :] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yofs, 0, sh - 1), np.clip(yofs + 1, 0, sh - 1)      # vertical: rows are clamped
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (Hh[y0] >> 4)) >> 16) + ((b1 * (Hh[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def lab_tables():
    """-> (gamma[256], cbrt[3072]) int64.  The published formulae in float64; OpenCV fills its tables in single precision, which falls on
    the other side of a rounding boundary at two reachable arguments (49, 628)."""
    x = np.arange(256) / 255.0
    gamma = np.rint(255.0 * 8 * np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)).astype(np.int64)
    y = np.arange(3072) / (255.0 * 8)
    cb = np.rint(32768 * np.where(y < 216 / 24389, y * (841 / 108) + 16 / 116, np.cbrt(y))).astype(np.int64)
    cb[49] -= 1
    cb[628] += 1
    return gamma, cb


def rgb2lab_u8(img):
    """img: uint8 RGB [..., 3] -> uint8 Lab [..., 3]"""
    gamma, cb = lab_tables()
    m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    c = np.rint(4096 * m / np.array([0.950456, 1.0, 1.088754])[:, None]).astype(np.int64)

    def descale(v, n):
        return (v + (1 << (n - 1))) >> n

    r, g, b = gamma[img[..., 0]], gamma[img[..., 1]], gamma[img[..., 2]]
    fx = cb[descale(r * c[0, 0] + g * c[0, 1] + b * c[0, 2], 12)]
    fy = cb[descale(r * c[1, 0] + g * c[1, 1] + b * c[1, 2], 12)]
    fz = cb[descale(r * c[2, 0] + g * c[2, 1] + b * c[2, 2], 12)]
    lscale, lshift = (116 * 255 + 50) // 100, -((16 * 255 * (1 << 15) + 50) // 100)
    out = np.stack([descale(lscale * fy + lshift, 15), descale(500 * (fx - fy) + 128 * (1 << 15), 15),
                    descale(200 * (fy - fz) + 128 * (1 << 15), 15)], -1)
    return np.clip(out, 0, 255).astype(np.uint8)


def uciqe_restated(img):
    """img: HWC uint8 RGB -> (UCIQE, var_chr, con_lum, aver_sat), following metrics/metrics.py:40-76"""
    lab = rgb2lab_u8(img)
    f = np.float64
    with np.errstate(all="ignore"):
        lum, a, b = lab[..., 0] / f(255), lab[..., 1] / f(255), lab[..., 2] / f(255)
        chroma = np.sqrt(a * a + b * b)
        aver_sat = np.mean(chroma / np.sqrt(chroma * chroma + lum * lum))
        var_chr = np.sqrt(np.mean(np.abs(1 - np.square(np.mean(chroma) / chroma))))
    h = np.bincount(lab[..., 0].ravel(), minlength=256)
    nb = 65536
    occ = np.nonzero(h)[0]
    first, last = f(occ[0]) / f(255), f(occ[-1]) / f(255)
    if first == last:
        first, last = first - 0.5, last + 0.5
    denom = last - first
    step = denom / f(nb)

    def edge(i):
        return last if i == nb else f(i) * step + first

    total, cum, ilow, ihigh = int(h.sum()), 0, None, None
    for k in occ:
        v = f(k) / f(255)
        idx = int(((v - first) / denom) * f(nb))
        if idx == nb:
            idx -= 1
        if v < edge(idx):
            idx -= 1
        if v >= edge(idx + 1) and idx != nb - 1:
            idx += 1
        cum += int(h[k])
        cdf = f(cum) / f(total)
        if ilow is None and cdf > 0.0100:
            ilow = idx
        if ihigh is None and cdf >= 0.9900:
            ihigh = idx
    con_lum = f(ihigh - 1) / f(nb - 1) - f(ilow - 1) / f(nb - 1)
    return 0.4680 * var_chr + 0.2745 * con_lum + 0.2576 * aver_sat, var_chr, con_lum, aver_sat


def ssim_restated(im1, im2, win_size=7, data_range=255.0):
    """im1, im2: HWC uint8 -> float (PARITY UNPINNED, see the module header)"""
    from scipy.ndimage import uniform_filter
    k1, k2 = 0.01, 0.03
    npx = win_size ** 2
    cov_norm = npx / (npx - 1)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (win_size - 1) // 2
    per_channel = []
    for ch in range(im1.shape[2]):
        x, y = im1[..., ch].astype(np.float64), im2[..., ch].astype(np.float64)
        ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
        uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        a1, a2, b1, b2 = 2 * ux * uy + c1, 2 * vxy + c2, ux ** 2 + uy ** 2 + c1, vx + vy + c2
        smap = (a1 * a2) / (b1 * b2)
        per_channel.append(smap[pad:smap.shape[0] - pad, pad:smap.shape[1] - pad].mean(dtype=np.float64))
    return float(np.mean(per_channel))


_cache = {}


def reference_uciqe():
    """-> the reference's uciqe(nargin, loc) (loc: HWC uint8 RGB image), executed from the unmodified source (needs cv2)"""
    if "uciqe" not in _cache:
        import ast
        import cv2
        from . import ref_loader
        path = os.path.join(ref_loader.REF_ROOT, "metrics", "metrics.py")
        with open(path, "r") as f:
            tree = ast.parse(f.read())
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "uciqe"][:1]
        assert keep
        ns = {"np": np, "cv2": cv2}
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
        _cache["uciqe"] = ns["uciqe"]
    return _cache["uciqe"]


def reference_uiqm():
    """-> the reference's getUIQM(x) (x: HWC uint8 / float image), with its helpers, executed from the unmodified source"""
    if "uiqm" not in _cache:
        import ast
        from scipy import ndimage
        from . import ref_loader
        path = os.path.join(ref_loader.REF_ROOT, "metrics", "metrics.py")
        with open(path, "r") as f:
            tree = ast.parse(f.read())
        want = {"mu_a", "s_a", "_uicm", "sobel", "eme", "_uism", "plip_g", "plip_theta", "plip_cross", "plip_diag", "plip_multiplication",
                "plip_phiInverse", "plip_phi", "_uiconm", "getUIQM"}
        keep, seen = [], set()
        for n in tree.body:                   # the FIRST definition of each name (the file re-defines eme further down for nmetrics)
            if isinstance(n, ast.FunctionDef) and n.name in want and n.name not in seen:
                keep.append(n)
                seen.add(n.name)
        assert seen == want, sorted(want - seen)
        ns = {"np": np, "math": math, "ndimage": ndimage}
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
        _cache["uiqm"] = ns
    return _cache["uiqm"]
