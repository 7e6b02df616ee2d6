"""Oracle: the algebra of csrc/upconv_c1.cu (decoder d4 = Upsample x2 -> ReflectionPad2d(1) -> Conv3x3(C -> 1),
reference model.py:369-372) restated in numpy, loop by loop, with the SAME closed-form border weights the kernel
uses.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); tests/test_folded_upconv_cpu.py checks it against torch
autograd of the reference composition, so the kernel's index logic is pinned on the CPU as well as on the GPU.

  V[i,j,tap] = sum_c x[c,i,j] w[c,tap];      y[Y,X] = sum_tap upP(V[.,.,tap])[Y+ky, X+kx]
  S[i,j,tap] = sum_{m,n} vy[ky][m] vx[kx][n] g[2i-2+m, 2j-2+n]     (zero outside the image)
  gx[c,i,j]  = sum_tap S[i,j,tap] w[c,tap];  gw[c,tap] = sum_{i,j} x[c,i,j] S[i,j,tap]
"""
from __future__ import annotations

import numpy as np


def adj_weights(i: int, n: int) -> np.ndarray:
    """v[k][m]: weight of gradient row 2i-2+m (m = 0..5) on source row i at conv tap offset k (0..2): padded rows
    2i..2i+3 carry .25 .75 .75 .25 (rows 0 / n-1 absorb the clamped tap), rows 1 / n-2 also read the reflected pad
    row -- upconv_c1.cu:adj_weights"""
    W = [0.25, 0.75, 0.75, 0.25]
    if i == 0:
        W[0], W[1] = 0.75, 1.0
    if i == n - 1:
        W[2], W[3] = 1.0, 0.75
    v = np.zeros((3, 6))
    for k in range(3):
        for m in range(6):
            r = m + k - 2
            if 0 <= r < 4:
                v[k, m] = W[r]
    if i == 1:
        v[0, 0] += 0.25
    if i == n - 2:
        v[2, 5] += 0.25
    return v


def pad_src(p: int, n: int):
    """padded index p of the [2n+2] axis -> (a, b, f): value = (1-f) x[a] + f x[b]  -- upconv_c1.cu:pad_src"""
    u = p - 1
    if u < 0:
        u = -u
    if u >= 2 * n:
        u = 4 * n - 2 - u
    a = (u - 1) >> 1 if u > 0 else 0
    f = 0.0 if u == 0 else (0.25 if (u & 1) else 0.75)
    return a, min(a + 1, n - 1), f


def forward(x, w, b=0.0):
    """x [C,H,W], w [C,3,3] -> pre-activation y [2H,2W]"""
    C, H, W = x.shape
    V = np.einsum("cij,ct->ijt", x, w.reshape(C, 9))
    y = np.full((2 * H, 2 * W), float(b))
    for Y in range(2 * H):
        for ky in range(3):
            ra, rb, fy = pad_src(Y + ky, H)
            for X in range(2 * W):
                for kx in range(3):
                    ca, cb, fx = pad_src(X + kx, W)
                    t = ky * 3 + kx
                    top = (1 - fx) * V[ra, ca, t] + fx * V[ra, cb, t]
                    bot = (1 - fx) * V[rb, ca, t] + fx * V[rb, cb, t]
                    y[Y, X] += (1 - fy) * top + fy * bot
    return y


def backward(x, w, g):
    """g [2H,2W] -> (gx [C,H,W], gw [C,3,3], gb)"""
    C, H, W = x.shape
    gp = np.zeros((2 * H + 6, 2 * W + 6))
    gp[2:2 * H + 2, 2:2 * W + 2] = g                      # gp[r + 2] = g[r], zero outside
    S = np.zeros((H, W, 9))
    for i in range(H):
        vy = adj_weights(i, H)
        for j in range(W):
            vx = adj_weights(j, W)
            win = gp[2 * i:2 * i + 6, 2 * j:2 * j + 6]    # gradient rows 2i-2 .. 2i+3
            for ky in range(3):
                for kx in range(3):
                    S[i, j, ky * 3 + kx] = vy[ky] @ win @ vx[kx]
    gx = np.einsum("ijt,ct->cij", S, w.reshape(C, 9))
    gw = np.einsum("cij,ijt->ct", x, S).reshape(C, 3, 3)
    return gx, gw, g.sum()
