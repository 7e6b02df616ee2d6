"""Oracle-side PROTOTYPE (CPU, torch float64) of the phase-folded form of the decoder blocks d1-d3
`Upsample(x2, bilinear, align_corners=False) -> ReflectionPad2d(1) -> Conv2d(Cin, Cout, 3)` (reference
model.py:357-368), the next kernel item in DESIGN.md section 7.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
it pins the algebra the future sm_100a kernels must implement; nothing in the product imports it.

1-D facts (n source samples x[0..n-1], padded up-sampled axis p = 0..2n+1, u = p - 1):
  up[2i]   = 0.25 x[i-1] + 0.75 x[i]      (x[-1] := x[0])
  up[2i+1] = 0.75 x[i]   + 0.25 x[i+1]    (x[n]  := x[n-1])
  PU = reflect-padded up:  PU[0] = up[1], PU[2n+1] = up[2n-2]
  E  = the same formulas evaluated on the REPLICATE-extended x:  E[p] = PU[p] except E[0] = x[0], E[2n+1] = x[n-1]
  D  = PU - E: zero except D[0] = 0.25 (x[1] - x[0]) and D[2n+1] = 0.25 (x[n-2] - x[n-1])

Output sample 2i+q (q = 0, 1) of a 3-tap correlation with w over E reads x[i-1], x[i], x[i+1] with the folded
weights  A[q] @ w,
  A[0] = [[.75, .25, 0], [.25, .75, .75], [0, 0, .25]],   A[1] = [[.25, 0, 0], [.75, .75, .25], [0, .25, .75]]
(rows: source tap a = -1, 0, +1; columns: conv tap k).  In 2-D the operator is (E_y + D_y) x (E_x + D_x)
  = E x E  +  D_y x PU_x  +  E_y x D_x,
i.e. ONE 3x3 convolution Cin -> 4*Cout of the replicate-padded LOW-resolution tensor (four output phases; the
same FLOPs as the original layer, 4x fewer pixels, 4x wider N), plus corrections on the outermost output row /
column of each side only:
  rows Y = 0 / 2H-1:    conv1d over X of  w[:, :, 0 / 2, :]  with  PU_x( 0.25 (x[1] - x[0])   /  0.25 (x[H-2] - x[H-1]) )
  cols X = 0 / 2W-1:    conv1d over Y of  w[:, :, :, 0 / 2]  with  E_y ( 0.25 (x[:,1] - x[:,0]) / 0.25 (x[:,W-2] - x[:,W-1]) )
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

A = torch.tensor([[[0.75, 0.25, 0.0], [0.25, 0.75, 0.75], [0.0, 0.0, 0.25]],
                  [[0.25, 0.0, 0.0], [0.75, 0.75, 0.25], [0.0, 0.25, 0.75]]], dtype=torch.float64)


def reference_block(x, w, b=None):
    """the layer as the reference runs it (model.py:357-359)"""
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), w, b)


def fold_weights(w):
    """w [Co,Ci,3,3] -> Wf [4*Co, Ci, 3, 3], output channel (py*2+px)*Co + co"""
    a = A.to(w.dtype)
    wf = torch.einsum("pak,qbl,oikl->pqoiab", a, a, w)
    return wf.reshape(4 * w.shape[0], w.shape[1], 3, 3)


def unfold_weight_grad(gwf, co):
    """adjoint of fold_weights: gradient w.r.t. Wf -> gradient w.r.t. w"""
    a = A.to(gwf.dtype)
    g = gwf.reshape(2, 2, co, gwf.shape[1], 3, 3)
    return torch.einsum("pak,qbl,pqoiab->oikl", a, a, g)


def _up1d_pad(d, mode):
    """1-D Upsample(x2, bilinear) + pad(1) along the last axis; d [B,C,n] -> [B,C,2n+2]"""
    up = F.interpolate(d, scale_factor=2, mode="linear", align_corners=False)
    return F.pad(up, (1, 1), mode=mode)


def folded_block(x, w, b=None):
    """the same layer as ONE convolution over the low-resolution tensor + border corrections"""
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    t = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), fold_weights(w))              # [B, 4Co, H, W]
    y = t.reshape(B, 2, 2, Co, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, Co, 2 * H, 2 * W).clone()
    # D_y x PU_x: top and bottom output rows
    for row, ky, d in ((0, 0, x[:, :, 1, :] - x[:, :, 0, :]), (2 * H - 1, 2, x[:, :, H - 2, :] - x[:, :, H - 1, :])):
        y[:, :, row, :] += F.conv1d(_up1d_pad(0.25 * d, "reflect"), w[:, :, ky, :])
    # E_y x D_x: left and right output columns
    for col, kx, d in ((0, 0, x[:, :, :, 1] - x[:, :, :, 0]), (2 * W - 1, 2, x[:, :, :, W - 2] - x[:, :, :, W - 1])):
        y[:, :, :, col] += F.conv1d(_up1d_pad(0.25 * d, "replicate"), w[:, :, :, kx])
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y
