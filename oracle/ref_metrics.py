"""ORACLE (test infrastructure; nothing in the product package imports this): CPU restatements for the input-pipeline /
metric kernels of csrc/hd_metrics.cu.

  resize_bilinear_u8   numpy restatement of OpenCV's 8-bit INTER_LINEAR resize (the algorithm lives in the third-party
                       dependency opencv-python — albumentations' A.Resize calls cv2.resize, utils/utils.py:318-323 —: 11-bit
                       fixed-point weights `cvRound(w * 2048)`, horizontal pass in int, vertical pass
                       `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`).  Pinned bit-for-bit against cv2.resize
                       itself (opencv-python 4.13 in this image) by tests/test_metrics.py.
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


_cache = {}


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
