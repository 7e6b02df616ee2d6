"""Oracle: peak-centred patch extraction, numpy restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

a1  PatchDataset.__getitem__ with transform=None (reference data.py:211-250):
    integer peak (cy, cx) -> whole-image bilinear translate by (W/2-cx, H/2-cy)
    (an exact integer shift) -> center_crop(P+2*pad) -> center_crop(P).  For integer
    sites this is bit-equal to float32(img)[cy-P/2:cy+P/2, cx-P/2:cx+P/2]
    (verified against the reference in tests/golden/make_golden.py).

a2  AdaptiveLatticeDataset.__getitem__ with transform=None (data.py:478-560):
    float site (cy, cx) -> integer ROI of P+max(16,2*pad) around round(c), zero padded
    at the image border -> sub-pixel bilinear translate by (x_int-cx, y_int-cy) with
    zeros outside -> crop P -> per-patch min-max to [0,1].
"""
from __future__ import annotations

import numpy as np


def global_index_to_site(counts, idx):
    """Linear walk over per-image site lists (data.py:212-220): -> (img_idx, local)."""
    img = 0
    while img < len(counts) and idx >= counts[img]:
        idx -= counts[img]
        img += 1
    if img >= len(counts):
        raise IndexError("index out of range")
    return img, idx


def gather_integer(images, sites, P):
    """images: list of 2-D float arrays (float64 like the reference caches them);
    sites: int array [N,3] of (img_idx, cy, cx).  -> float32 [N,1,P,P]."""
    out = np.empty((len(sites), 1, P, P), dtype=np.float32)
    h = P // 2
    for n, (i, cy, cx) in enumerate(np.asarray(sites, dtype=np.int64)):
        out[n, 0] = images[i][cy - h:cy + h, cx - h:cx + h].astype(np.float32)
    return out


def _bilinear_zero(img, ys, xs):
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    fy = ys - y0; fx = xs - x0
    H, W = img.shape
    out = np.zeros(np.broadcast(ys, xs).shape, dtype=np.float64)
    for dy, dx, w in ((0, 0, (1 - fy) * (1 - fx)), (0, 1, (1 - fy) * fx),
                      (1, 0, fy * (1 - fx)), (1, 1, fy * fx)):
        yy = y0 + dy; xx = x0 + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        out += np.where(ok, w * img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0.0)
    return out


def gather_subpixel(image, cy, cx, P, pad, normalise=True):
    """One a2 patch (transform=None), float64 maths; matches the reference to the
    float32 grid rounding of torchvision's affine (<= 2e-5 abs before min-max)."""
    H, W = image.shape
    roi = P + max(16, 2 * pad)
    yi, xi = int(round(cy)), int(round(cx))
    # ROI pixel (r, c) is image pixel (yi - roi//2 + r, xi - roi//2 + c), zero outside.
    # Output pixel (r, c) of the translated ROI samples ROI position
    # (r - shift_y, c - shift_x) with shift = roi/2 - rel_c = yi - cy (x likewise).
    off = (roi - P) // 2
    r = np.arange(P) + off
    ys = (r - (yi - cy))[:, None] + (yi - roi // 2)
    xs = (r - (xi - cx))[None, :] + (xi - roi // 2)
    # zero-padding of the ROI window itself
    img = image.astype(np.float32).astype(np.float64)
    patch = _bilinear_zero(img, ys + 0 * xs, xs + 0 * ys)
    if normalise:
        lo, hi = patch.min(), patch.max()
        patch = (patch - lo) / (hi - lo) if hi > lo else np.zeros_like(patch)
    return patch.astype(np.float32)[None]
