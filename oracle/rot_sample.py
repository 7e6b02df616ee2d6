"""Oracle: rotate + bilinear sample (affine_grid + grid_sample), numpy restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with explicit forward AND hand-derived backward, what the reference
gets from ``F.affine_grid(rot, size, align_corners=False)`` followed by
``F.grid_sample(x, grid, mode="bilinear", padding_mode="reflection",
align_corners=False)`` at its three call sites:

  * reference src/livae/model.py:250-258  (RotationSTN.forward, x by R(theta))
  * reference src/livae/model.py:465-470  (RVAE.forward, recon by R(-theta))
  * reference src/livae/train.py:670-677  (rotate_to_canonical, x by R(theta))

The rotation matrix is [[c, -s, 0], [s, c, 0]] (model.py:230-231, 250-251):
pure rotation, third column hard zero.  ``tx, ty`` are accepted as a
generalisation (normalised-coordinate translation) and default to 0.

Semantics (ATen GridSampler.h / AffineGridGenerator.cpp, restated in
SURVEY.md section 8 row a6):
  base   xs_j = (2j+1)/W - 1,  ys_i = (2i+1)/H - 1
  grid   gx = c*xs - s*ys + tx,  gy = s*xs + c*ys + ty
  unnorm ix = ((gx+1)*W - 1)/2
  reflect about [-0.5, W-0.5]: v=|ix+0.5|, extra=fmod(v,W), flips=floor(v/W),
          even -> extra-0.5, odd -> W-extra-0.5   (sign flips tracked for bwd)
  clip   to [0, W-1]  (gradient 0 where clipped, ATen clip_coordinates_set_grad)
  bilinear 4 taps, out-of-range corners contribute 0.
"""
from __future__ import annotations

import numpy as np


def _coords(c, s, tx, ty, H, W, dtype):
    """Source coordinates (after reflect+clip) and d(coord)/d(grid) multipliers."""
    B = c.shape[0]
    xs = ((2.0 * np.arange(W, dtype=dtype) + 1.0) / W - 1.0)[None, None, :]
    ys = ((2.0 * np.arange(H, dtype=dtype) + 1.0) / H - 1.0)[None, :, None]
    c = c.reshape(B, 1, 1).astype(dtype)
    s = s.reshape(B, 1, 1).astype(dtype)
    tx = np.broadcast_to(np.asarray(tx, dtype=dtype).reshape(-1, 1, 1), (B, 1, 1))
    ty = np.broadcast_to(np.asarray(ty, dtype=dtype).reshape(-1, 1, 1), (B, 1, 1))
    gx = c * xs - s * ys + tx
    gy = s * xs + c * ys + ty

    def one(g, n):
        u = ((g + 1.0) * n - 1.0) / 2.0          # unnormalise, align_corners=False
        mult = np.full_like(u, n / 2.0)
        v = u + 0.5                                # reflect about [-0.5, n-0.5]
        neg = v < 0
        v = np.abs(v)
        mult = np.where(neg, -mult, mult)
        extra = np.fmod(v, n)
        flips = np.floor(v / n)
        odd = (flips.astype(np.int64) % 2) == 1
        r = np.where(odd, n - extra - 0.5, extra - 0.5)
        mult = np.where(odd, -mult, mult)
        lo = r <= 0                                # clip, grad 0 when clipped
        hi = r >= (n - 1)
        r = np.where(lo, 0.0, np.where(hi, n - 1.0, r))
        mult = np.where(lo | hi, 0.0, mult)
        return r.astype(dtype), mult.astype(dtype)

    ix, mx = one(gx, W)
    iy, my = one(gy, H)
    return ix, iy, mx, my, xs, ys


def rot_sample_fwd(img, c, s, tx=0.0, ty=0.0, dtype=np.float64):
    """img [B,C,H,W]; c,s [B] -> out [B,C,H,W]."""
    img = np.asarray(img, dtype=dtype)
    B, C, H, W = img.shape
    ix, iy, _, _, _, _ = _coords(np.asarray(c), np.asarray(s), tx, ty, H, W, dtype)
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    fx = ix - x0
    fy = iy - y0
    out = np.zeros_like(img)
    bidx = np.arange(B)[:, None, None]
    for dy, dx, w in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                      (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xx = x0 + dx
        yy = y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xc = np.clip(xx, 0, W - 1)
        yc = np.clip(yy, 0, H - 1)
        for ch in range(C):
            out[:, ch] += np.where(ok, w * img[bidx, ch, yc, xc], 0.0)
    return out


def rot_sample_bwd(img, c, s, gout, tx=0.0, ty=0.0, dtype=np.float64):
    """Backward of rot_sample_fwd.

    Returns (grad_img [B,C,H,W], grad_c [B], grad_s [B], grad_tx [B], grad_ty [B]).
    grad wrt (c, s) is the contraction of grad_grid with the base grid:
      dL/dc = sum(dL/dgx * xs + dL/dgy * ys),  dL/ds = sum(-dL/dgx * ys + dL/dgy * xs)
    which is what affine_grid's backward (base_grid^T @ grad_grid) gives for the
    [[c,-s],[s,c]] parametrisation (model.py:250-252).
    """
    img = np.asarray(img, dtype=dtype)
    gout = np.asarray(gout, dtype=dtype)
    B, C, H, W = img.shape
    ix, iy, mx, my, xs, ys = _coords(np.asarray(c), np.asarray(s), tx, ty, H, W, dtype)
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    fx = ix - x0
    fy = iy - y0
    gimg = np.zeros_like(img)
    gix = np.zeros((B, H, W), dtype=dtype)
    giy = np.zeros((B, H, W), dtype=dtype)
    bidx = np.broadcast_to(np.arange(B)[:, None, None], (B, H, W))
    taps = ((0, 0, (1 - fx) * (1 - fy), -(1 - fy), -(1 - fx)),
            (0, 1, fx * (1 - fy), (1 - fy), -fx),
            (1, 0, (1 - fx) * fy, -fy, (1 - fx)),
            (1, 1, fx * fy, fy, fx))
    for dy, dx, w, dwx, dwy in taps:
        xx = x0 + dx
        yy = y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xc = np.clip(xx, 0, W - 1)
        yc = np.clip(yy, 0, H - 1)
        for ch in range(C):
            g = gout[:, ch]
            np.add.at(gimg[:, ch], (bidx[ok], yc[ok], xc[ok]), (w * g)[ok])
            v = np.where(ok, img[bidx, ch, yc, xc], 0.0)
            gix += v * dwx * g
            giy += v * dwy * g
    ggx = gix * mx
    ggy = giy * my
    gc = (ggx * xs + ggy * ys).reshape(B, -1).sum(1)
    gs = (-ggx * ys + ggy * xs).reshape(B, -1).sum(1)
    gtx = ggx.reshape(B, -1).sum(1)
    gty = ggy.reshape(B, -1).sum(1)
    return gimg, gc, gs, gtx, gty
