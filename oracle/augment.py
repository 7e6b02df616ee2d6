"""Oracle: patch augmentation and the paired random rotation, numpy restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

a3  default_transform (reference data.py:78-116): scale U(0.9,1.1) through TF.affine (bilinear,
    zeros outside), optional TF.rotate (bilinear, fill=0), h/v flips, torch.roll by (shift_y, shift_x).
a2' PairedAdaptiveLatticeDataset.__getitem__ (data.py:617-735): sub-pixel ROI crop of P+2*pad ->
    transform(rotation=False) -> TF.rotate by angle~U(0,360) -> centre crops of both -> per-patch
    min-max of both -> (patch, rotated, radians(angle)).

The arithmetic that is not in /root/reference is torchvision's (pinned 0.24.1 in uv.lock, 0.26.0 in the
image): TF.affine / TF.rotate on tensors = _get_inverse_affine_matrix -> _gen_affine_grid ->
grid_sample(bilinear, zeros, align_corners=False); with fill=0 (TF.rotate) the sampled image is
multiplied by the identically sampled all-ones mask (_apply_grid_transform).  Restated here in float64:
output pixel (i, j) of an [S, S] image samples source position
    x = m00*xb + m01*yb + m02 + (S-1)/2,   y = m10*xb + m11*yb + m12 + (S-1)/2,
    xb = j - (S-1)/2, yb = i - (S-1)/2,
with m = [1/s, 0, 0; 0, 1/s, 0] for the scale and [cos a, -sin a, 0; sin a, cos a, 0] for TF.rotate(angle=a).
The random draws are inputs (the reference takes them from Python's `random`): `draw_params` replays the
reference's draw ORDER so that a test seeding `random` the same way gets the same numbers.
"""
from __future__ import annotations

import math
import random

import numpy as np

from oracle.patch import _bilinear_zero


def draw_params(rotation=False, flip_prob=0.5, jitter_amount=4, rng=random):
    """The draws of one default_transform call, in the reference's order (data.py:85-114)."""
    p = {"scale": rng.uniform(0.9, 1.1)}
    p["angle"] = rng.uniform(0, 360) if rotation else None
    p["hflip"] = rng.random() < flip_prob
    p["vflip"] = rng.random() < flip_prob
    if jitter_amount > 0:
        p["shift_x"] = rng.randint(-jitter_amount, jitter_amount)
        p["shift_y"] = rng.randint(-jitter_amount, jitter_amount)
    else:
        p["shift_x"] = p["shift_y"] = 0
    return p


def _affine(img, m, masked):
    S = img.shape[-1]
    c = (S - 1) / 2.0
    yb, xb = np.meshgrid(np.arange(S) - c, np.arange(S) - c, indexing="ij")
    xs = m[0] * xb + m[1] * yb + m[2] + c
    ys = m[3] * xb + m[4] * yb + m[5] + c
    out = _bilinear_zero(img.astype(np.float64), ys, xs)
    if masked:                       # fill=0: value * sampled mask (torchvision _apply_grid_transform)
        out = out * _bilinear_zero(np.ones_like(img, dtype=np.float64), ys, xs)
    return out


def scale_affine(img, scale):
    """TF.affine(angle=0, translate=0, scale=s, shear=0), fill=None (data.py:86-93)."""
    return _affine(img, [1.0 / scale, 0.0, 0.0, 0.0, 1.0 / scale, 0.0], masked=False)


def rotate(img, angle_deg):
    """TF.rotate(angle, bilinear, expand=False, fill=0) (data.py:97-103, 698-704)."""
    r = math.radians(-angle_deg)
    return _affine(img, [math.cos(r), math.sin(r), 0.0, -math.sin(r), math.cos(r), 0.0], masked=True)


def default_transform(patch, p):
    """patch [S,S] float; p from draw_params."""
    out = scale_affine(patch, p["scale"])
    if p.get("angle") is not None:
        out = rotate(out, p["angle"])
    if p["hflip"]:
        out = out[:, ::-1]
    if p["vflip"]:
        out = out[::-1, :]
    return np.roll(out, (p["shift_y"], p["shift_x"]), axis=(0, 1))


def roi_crop(image, cy, cx, P, pad):
    """The [P+2*pad]^2 `patch_big` of data.py:640-691: integer ROI window of P+max(16,2*pad) around
    round(c) (zero outside the image AND outside the window), bilinear sub-pixel shift, centre crop."""
    S = P + 2 * pad
    roi = P + max(16, 2 * pad)
    yi, xi = int(round(cy)), int(round(cx))          # Python round: half to even
    y0, x0 = yi - roi // 2, xi - roi // 2
    H, W = image.shape
    win = np.zeros((roi, roi), dtype=np.float64)
    ya, yb_, xa, xb_ = max(0, y0), min(H, y0 + roi), max(0, x0), min(W, x0 + roi)
    if yb_ > ya and xb_ > xa:
        win[ya - y0:yb_ - y0, xa - x0:xb_ - x0] = image[ya:yb_, xa:xb_].astype(np.float32)
    off = (roi - S) // 2
    r = np.arange(S) + off
    ys = (r - (yi - cy))[:, None] + 0.0 * r[None, :]
    xs = (r - (xi - cx))[None, :] + 0.0 * r[:, None]
    return _bilinear_zero(win, ys, xs)


def _centre(a, P):
    o = (a.shape[-1] - P) // 2
    return a[o:o + P, o:o + P]


def _minmax(a):
    lo, hi = a.min(), a.max()
    return (a - lo) / (hi - lo) if hi > lo else np.zeros_like(a)


def adaptive_item(image, cy, cx, P, pad, p=None):
    """AdaptiveLatticeDataset.__getitem__ (data.py:478-560); p=None means transform=None."""
    big = roi_crop(image, cy, cx, P, pad)
    if p is not None:
        big = default_transform(big, p)
    return _minmax(_centre(big, P)).astype(np.float32)[None]


def paired_item(image, cy, cx, P, pad, p, angle_deg):
    """PairedAdaptiveLatticeDataset.__getitem__ (data.py:617-735); p=None means transform=None."""
    big = roi_crop(image, cy, cx, P, pad)
    if p is not None:
        big = default_transform(big, p)
    rot = rotate(big, angle_deg)
    return (_minmax(_centre(big, P)).astype(np.float32)[None],
            _minmax(_centre(rot, P)).astype(np.float32)[None], np.radians(angle_deg))
