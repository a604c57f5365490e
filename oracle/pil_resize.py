"""ORACLE (test infrastructure only): numpy restatement of Pillow's 8-bit bilinear resize, the arithmetic behind
detectron2's ResizeTransform.apply_image for uint8 images (reached from cubercnn/data/dataset_mapper.py:26-28 through
T.ResizeShortestEdge: `Image.fromarray(img).resize((new_w, new_h), Image.BILINEAR)`).

Algorithm (Pillow src/libImaging/Resample.c, pinned version: the Pillow of this image, tests compare bit for bit):
precompute_coeffs (triangle filter, support scaled by the down-scale factor = antialiasing, double arithmetic),
normalize_coeffs_8bpc (fixed point, PRECISION_BITS = 22), a horizontal pass over the rows the vertical pass needs, then
a vertical pass; every pass accumulates int32 from 1 << 21 and stores clip8(acc >> 22) as uint8.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size, out_size):
    """-> bounds (out,2) int32 [xmin, count], kk (out, ksize) int32 fixed-point weights (box = the whole axis)."""
    in0, in1 = np.float32(0.0), np.float32(in_size)
    scale = float(np.float32(in1 - in0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = float(in0) + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        if ww != 0.0:
            for x in range(xmax):
                w[x] /= ww
        for x in range(ksize):
            p = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + p) if w[x] < 0 else int(0.5 + p)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """img (H,W,C) uint8; resample along `axis` (1 = horizontal, 0 = vertical)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((len(bounds),) + src.shape[1:], np.uint8)
    for i, (lo, n) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[lo + x] * int(kk[i, x])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, new_h, new_w):
    """img (H,W,C) uint8 -> (new_h,new_w,C) uint8 == np.asarray(Image.fromarray(img).resize((new_w,new_h), BILINEAR))."""
    h, w = img.shape[:2]
    out = img
    if new_w != w:
        bh, kh = precompute_coeffs(w, new_w)
        if new_h != h:                    # Pillow only resamples the rows the vertical pass will read
            bv, _ = precompute_coeffs(h, new_h)
            first, last = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
            tmp = np.zeros((h, new_w, img.shape[2]), np.uint8)
            tmp[first:last] = _pass(img[first:last], bh, kh, 1)
            out = tmp
        else:
            out = _pass(img, bh, kh, 1)
    if new_h != h:
        bv, kv = precompute_coeffs(h, new_h)
        out = _pass(out, bv, kv, 0)
    return out
