"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the SLIC k-means stage as ``csrc/slic.cu`` states it.

PARITY UNPINNED against the reference: the reference calls scikit-image's ``slic(img, n_segments=100, compactness=10,
start_label=0)`` (utils/image_to_graph/image_to_graph_superpixel.py:31); scikit-image is neither vendored nor
version-pinned (requirements.txt:12) nor installed here, so there are no reference labels to compare with.  What this
file pins is the algorithm the device kernels claim to implement (Achanta et al., TPAMI 2012, with scikit-image's
documented conventions), step by step:

  1. ``step = sqrt(H W / n_segments)``; grid of ``ny = round(H / step)`` x ``nx = round(W / step)`` centres (at least 1
     each), ``K = ny nx``; spatial scale ``S = max(H / ny, W / nx)``.
  2. sRGB -> CIELAB (D65 white, the constants of ``rgb_to_lab`` in csrc/slic.cu), divided by ``compactness``.
  3. centre ``k = gy nx + gx`` starts at ``((gy + 0.5) H / ny, (gx + 0.5) W / nx)`` with the colour of the pixel under it.
  4. ``iters`` (10) Lloyd iterations: every pixel (centre at ``(y + 0.5, x + 0.5)``) takes the nearest of the centres of
     the 3 x 3 grid cells around its own cell ``(y ny // H, x nx // W)`` under
     ``d = |dLab|^2 + |dyx|^2 / S^2`` (ties: lowest centre index); a centre moves to the mean colour / position of its
     pixels (sums in 2^-20 fixed point, so the order of summation is immaterial); a centre without pixels stays.
  5. one more assignment gives the labels (values in ``0 .. K - 1``; the connectivity post-pass is separate).

The device ranks candidates by |c|^2 - 2 p.c (the pixel's own |p|^2 is common to them) as one fp32 FMA chain; this
restatement evaluates the distance itself in float64, so pixels within rounding of a
tie can differ - the test states the agreement it requires.
"""
from __future__ import annotations

import math

import numpy as np


def slic_grid(H: int, W: int, n_segments: int):
    step = math.sqrt(H * W / float(max(n_segments, 1)))
    ny = max(1, int(math.floor(H / step + 0.5)))
    nx = max(1, int(math.floor(W / step + 0.5)))
    return ny, nx, np.float32(max(H / ny, W / nx))


def rgb_to_lab_scaled(img_u8: np.ndarray, compactness: float) -> np.ndarray:
    c = img_u8.astype(np.float32) * np.float32(1.0 / 255.0)
    lin = np.where(c > np.float32(0.04045), np.power((c + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)),
                   c / np.float32(12.92)).astype(np.float32)
    r, g, b = lin[..., 0], lin[..., 1], lin[..., 2]
    X = (np.float32(0.412453) * r + np.float32(0.357580) * g + np.float32(0.180423) * b) / np.float32(0.95047)
    Y = np.float32(0.212671) * r + np.float32(0.715160) * g + np.float32(0.072169) * b
    Z = (np.float32(0.019334) * r + np.float32(0.119193) * g + np.float32(0.950227) * b) / np.float32(1.08883)

    def f(t):
        return np.where(t > np.float32(0.008856), np.cbrt(t), np.float32(7.787) * t + np.float32(16.0 / 116.0)).astype(np.float32)

    fx, fy, fz = f(X), f(Y), f(Z)
    inv_c = np.float32(1.0 / compactness)
    return np.stack([(np.float32(116.0) * fy - np.float32(16.0)) * inv_c, np.float32(500.0) * (fx - fy) * inv_c,
                     np.float32(200.0) * (fy - fz) * inv_c], -1).astype(np.float32)


def slic_kmeans(img_u8: np.ndarray, n_segments: int = 100, compactness: float = 10.0, iters: int = 10):
    """``uint8 [H, W, 3]`` -> ``(labels int32 [H, W], centres float32 [K, 5] = (L, a, b, y, x))``."""
    H, W, _ = img_u8.shape
    ny, nx, S = slic_grid(H, W, n_segments)
    K = ny * nx
    lab = rgb_to_lab_scaled(img_u8, compactness)
    cen = np.zeros((K, 5), np.float32)
    for k in range(K):
        gy, gx = divmod(k, nx)
        cy, cx = np.float32((gy + 0.5) * H / ny), np.float32((gx + 0.5) * W / nx)
        cen[k, :3] = lab[min(int(cy), H - 1), min(int(cx), W - 1)]
        cen[k, 3], cen[k, 4] = cy, cx
    ys, xs = np.mgrid[0:H, 0:W]
    gy, gx = ys * ny // H, xs * nx // W
    yc, xc = ys + 0.5, xs + 0.5
    fixed = np.rint(lab.astype(np.float64) * 2.0 ** 20).astype(np.int64)
    labels = None
    for it in range(iters + 1):
        best = np.full((H, W), np.inf)
        labels = (gy * nx + gx).astype(np.int64)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy, xx = gy + dy, gx + dx
                ok = (yy >= 0) & (yy < ny) & (xx >= 0) & (xx < nx)
                k = np.where(ok, yy * nx + xx, 0)
                c = cen[k].astype(np.float64)
                d = ((lab.astype(np.float64) - c[..., :3]) ** 2).sum(-1) \
                    + ((yc - c[..., 3]) / float(S)) ** 2 + ((xc - c[..., 4]) / float(S)) ** 2
                take = ok & (d < best)                       # candidates in ascending k: strict < keeps the lowest
                best = np.where(take, d, best)
                labels = np.where(take, k, labels)
        if it == iters:
            break
        flat = labels.ravel()
        n = np.bincount(flat, minlength=K)
        for j in range(3):
            s = np.bincount(flat, weights=None, minlength=K) * 0
            s = np.zeros(K, np.int64)
            np.add.at(s, flat, fixed[..., j].ravel())
            upd = n > 0
            cen[upd, j] = (s[upd].astype(np.float64) / 2.0 ** 20 / n[upd]).astype(np.float32)
        sy = np.zeros(K, np.int64); np.add.at(sy, flat, (2 * ys + 1).ravel())
        sx = np.zeros(K, np.int64); np.add.at(sx, flat, (2 * xs + 1).ravel())
        upd = n > 0
        cen[upd, 3] = (sy[upd] * 0.5 / n[upd]).astype(np.float32)
        cen[upd, 4] = (sx[upd] * 0.5 / n[upd]).astype(np.float32)
    return labels.astype(np.int32), cen
