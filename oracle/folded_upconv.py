"""Oracle side (CPU, torch float64) of the phase-folded form of the decoder blocks
`Upsample(x2, bilinear, align_corners=False) -> ReflectionPad2d(1) -> Conv2d(Cin, Cout, 3)` (reference
model.py:357-368).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): it pins the algebra of csrc/upfold.cu on the
CPU against torch's own ops (tests/test_folded_upconv_cpu.py); nothing in the product imports it.  Two decompositions:
the replicate-padded one below (round 1's design study) and, at the end of the file, the zero-padded one with border
strips that the kernels implement, with the kernel's index maps restated one to one.

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


# ------------------------------------------------------------------------------------------------------------------
# The decomposition csrc/upfold.cu implements (round 2): ZERO padding for the folded convolution F0 (TMA out-of-bounds
# fill; no padded copy of x) and the whole border rule in two batches of 4-row strips.  With the 1-D operators
#   R  [2n+2, n]: true up-sampling + reflect padding (index clamping, positions -1 / 2n mirrored onto 1 / 2n-2),
#   R0 [2n+2, n]: the interior formula on the zero-extended input (what F0 implies),
# P - P0 = (Ry - R0y) (x) Rx + R0y (x) (Rx - R0x); Ry - R0y is non-zero on padded rows {-1, 0, 2h-1, 2h} only:
#   row -1: 0.5 x[0] + 0.25 x[1]      row 0: 0.25 x[0]      row 2h-1: 0.25 x[h-1]      row 2h: 0.25 x[h-2] + 0.5 x[h-1]
# The functions below restate the kernel's index maps (up_true_w, up_zero_w, the strip table) one to one.
# ------------------------------------------------------------------------------------------------------------------
def up_true_w(X: int, j: int, n: int) -> float:
    """csrc/upfold.cu up_true_w: weight of x[j] in the true padded up-sampled value at padded position X in -1..2n"""
    Xc = 1 if X < 0 else (2 * n - 2 if X >= 2 * n else X)
    m = Xc >> 1
    if Xc & 1:
        return (0.75 if m == j else 0.0) + (0.25 if min(m + 1, n - 1) == j else 0.0)
    return (0.25 if max(m - 1, 0) == j else 0.0) + (0.75 if m == j else 0.0)


def up_zero_w(X: int, j: int) -> float:
    """csrc/upfold.cu up_zero_w: the same for the interior formula on the zero-extended input"""
    m = (X + 2) // 2 - 1
    if X - 2 * m:
        return (0.75 if m == j else 0.0) + (0.25 if m + 1 == j else 0.0)
    return (0.25 if m - 1 == j else 0.0) + (0.75 if m == j else 0.0)


def up_matrices(n, dtype=torch.float64):
    R = torch.tensor([[up_true_w(t - 1, j, n) for j in range(n)] for t in range(2 * n + 2)], dtype=dtype)
    R0 = torch.tensor([[up_zero_w(t - 1, j) for j in range(n)] for t in range(2 * n + 2)], dtype=dtype)
    return R, R0


def strips(x):
    """x [B,C,h,w] -> (s_tb [2,B,C,4,2w+2], s_lr [2,B,C,4,2h+2]) as upfold_strips_kernel writes them (NCHW here): side 0 =
    top / left with live rows 0,1; side 1 = bottom / right with live rows 2,3; the lr strips transposed"""
    B, C, h, w = x.shape
    Rx, _ = up_matrices(w, x.dtype)
    _, R0y = up_matrices(h, x.dtype)
    s_tb = x.new_zeros(2, B, C, 4, 2 * w + 2)
    s_lr = x.new_zeros(2, B, C, 4, 2 * h + 2)
    for side in (0, 1):
        for r in ((0, 1) if side == 0 else (2, 3)):
            outer = (r == 0) if side == 0 else (r == 3)          # the strip row outside the image
            c0, c1 = (0.5, 0.25) if outer else (0.25, 0.0)
            l0, l1 = (0, 1) if side == 0 else (h - 1, h - 2)
            s_tb[side, :, :, r] = (c0 * x[:, :, l0] + c1 * x[:, :, l1]) @ Rx.T
            l0, l1 = (0, 1) if side == 0 else (w - 1, w - 2)
            s_lr[side, :, :, r] = (c0 * x[:, :, :, l0] + c1 * x[:, :, :, l1]) @ R0y.T
    return s_tb, s_lr


def folded_block_strips(x, w, b=None):
    """the layer as csrc/upfold.cu computes it: zero-padded folded convolution + strip corrections on the two outermost
    output rows / columns"""
    B, C, h, ww = x.shape
    Co = w.shape[0]
    t = F.conv2d(x, fold_weights(w), padding=1)
    y = t.reshape(B, 2, 2, Co, h, ww).permute(0, 3, 4, 1, 5, 2).reshape(B, Co, 2 * h, 2 * ww)
    s_tb, s_lr = strips(x)
    corr_tb = F.conv2d(s_tb.reshape(2 * B, C, 4, 2 * ww + 2), w).reshape(2, B, Co, 2, 2 * ww)
    corr_lr = F.conv2d(s_lr.reshape(2 * B, C, 4, 2 * h + 2), w.transpose(2, 3)).reshape(2, B, Co, 2, 2 * h)
    y = y.clone()
    y[:, :, :2] += corr_tb[0]
    y[:, :, -2:] += corr_tb[1]
    y[:, :, :, :2] += corr_lr[0].transpose(2, 3)
    y[:, :, :, -2:] += corr_lr[1].transpose(2, 3)
    return y if b is None else y + b.view(1, -1, 1, 1)
