"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the resize the reference applies to every image:
``Image.open(path).convert('RGB').resize((r, r))`` (reference utils/image_to_graph/image_to_graph_optimized.py:65-70,
utils/dataloader.py:34 via ImageFolder).  ``Image.resize`` with no ``resample`` argument is BICUBIC for RGB
images and runs Pillow's ``ImagingResample`` (Pillow is a third-party dependency of the reference,
requirements.txt:3, unpinned; 12.2.0 is what this image has): a separable two-pass convolution, horizontal
pass first, 8-bit intermediate, coefficients computed in double precision per output pixel and quantised to
22-bit fixed point.  Pinned: ``tests/test_oracle_golden.py`` checks this restatement bit for bit against
Pillow itself on the shipped JPEGs' shapes and on random shapes (down- and up-scaling, one unchanged axis).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Pillow: Resample.c


def _bicubic(x: float) -> float:
    """Keys cubic, a = -0.5 (Pillow's ``bicubic_filter``), support 2."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the full box ``(0, in_size)``:
    ``bounds int32 [out, 2]`` (first tap, tap count) and ``kk int32 [out, ksize]`` fixed-point weights."""
    support0 = 2.0
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            v = v * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One separable pass along ``axis`` of ``uint8 [H, W, C]``: int32 accumulation from the rounding
    constant ``1 << 21``, arithmetic shift, clamp to a byte."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.tensordot(kk[o, :n].astype(np.int64), img[lo:lo + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        # Pillow accumulates in 32-bit ints; the sum of |weights| * 255 stays far below 2^31
        out[o] = _clip8(acc.astype(np.int32))
    return np.moveaxis(out, 0, axis)


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``np.asarray(Image.fromarray(img).resize((out_w, out_h)))`` for ``uint8 [H, W, 3]``.  An axis whose
    size does not change is not filtered (Pillow's ``need_horizontal`` / ``need_vertical``)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    if W != out_w:
        img = _pass(img, *precompute_coeffs(W, out_w), axis=1)
    if H != out_h:
        img = _pass(img, *precompute_coeffs(H, out_h), axis=0)
    return np.ascontiguousarray(img)
