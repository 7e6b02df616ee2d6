"""CPU: the phase-folded form of the decoder blocks (oracle/folded_upconv.py, the design of DESIGN.md section 7 item 1)
equals Upsample -> ReflectionPad2d -> Conv2d (reference model.py:357-368) exactly -- values, data gradient and
weight gradient -- including the border rows / columns where the reflected pad differs from the replicate formula."""
import pytest
import torch

from oracle import folded_upconv as FU


@pytest.mark.parametrize("shape", [(2, 5, 3, 4, 4), (1, 8, 4, 16, 16), (2, 3, 2, 7, 5), (1, 4, 1, 2, 9)])
def test_folded_block_is_exact(shape):
    B, Ci, Co, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, Ci, H, W, dtype=torch.float64, generator=g, requires_grad=True)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64, generator=g, requires_grad=True)
    b = torch.randn(Co, dtype=torch.float64, generator=g)
    gy = torch.randn(B, Co, 2 * H, 2 * W, dtype=torch.float64, generator=g)
    ref = FU.reference_block(x, w, b)
    gx_ref, gw_ref = torch.autograd.grad((ref * gy).sum(), (x, w))
    got = FU.folded_block(x, w, b)
    gx, gw = torch.autograd.grad((got * gy).sum(), (x, w))
    assert (got - ref).abs().max() < 1e-12
    assert (gx - gx_ref).abs().max() < 1e-12 and (gw - gw_ref).abs().max() < 1e-11
    # without the corrections only the outermost output ring differs
    t = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 1, 1, 1), mode="replicate"), FU.fold_weights(w))
    y0 = t.reshape(B, 2, 2, Co, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, Co, 2 * H, 2 * W) + b.view(1, -1, 1, 1)
    diff = (y0 - ref).abs()
    assert diff[:, :, 1:-1, 1:-1].max() < 1e-12 and diff.max() > 1e-6


def test_fold_unfold_are_adjoint():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(3, 5, 3, 3, dtype=torch.float64, generator=g)
    gwf = torch.randn(12, 5, 3, 3, dtype=torch.float64, generator=g)
    lhs = (FU.fold_weights(w) * gwf).sum()
    rhs = (w * FU.unfold_weight_grad(gwf, 3)).sum()
    assert abs(float(lhs - rhs)) < 1e-10


@pytest.mark.parametrize("shape", [(3, 4, 4), (2, 6, 9), (4, 5, 4), (1, 8, 8)])
def test_upconv_c1_algebra_matches_autograd(shape):
    """oracle/upconv_c1.py (the V / S formulation and closed-form border weights of csrc/upconv_c1.cu) against torch
    autograd of Upsample -> ReflectionPad2d -> Conv2d(C -> 1): forward, data, weight and bias gradient."""
    import numpy as np

    from oracle import upconv_c1 as U
    C, H, W = shape
    g0 = torch.Generator().manual_seed(7 * C + H + W)
    x = torch.randn(1, C, H, W, dtype=torch.float64, generator=g0, requires_grad=True)
    w = torch.randn(1, C, 3, 3, dtype=torch.float64, generator=g0, requires_grad=True)
    b = torch.tensor([0.3], dtype=torch.float64, requires_grad=True)
    gy = torch.randn(1, 1, 2 * H, 2 * W, dtype=torch.float64, generator=g0)
    ref = FU.reference_block(x, w, b)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad((ref * gy).sum(), (x, w, b))
    y = U.forward(x[0].detach().numpy(), w[0].detach().numpy(), 0.3)
    assert np.abs(y - ref[0, 0].detach().numpy()).max() < 1e-12
    gx, gw, gb = U.backward(x[0].detach().numpy(), w[0].detach().numpy(), gy[0, 0].numpy())
    assert np.abs(gx - gx_ref[0].numpy()).max() < 1e-12
    assert np.abs(gw - gw_ref[0].numpy()).max() < 1e-12
    assert abs(gb - float(gb_ref)) < 1e-12


@pytest.mark.parametrize("shape", [(2, 5, 3, 4, 4), (1, 8, 4, 16, 16), (2, 3, 2, 7, 5), (1, 4, 2, 2, 9), (1, 2, 2, 8, 3)])
def test_zero_padded_fold_plus_strips_is_exact(shape):
    """the decomposition csrc/upfold.cu implements (zero-padded folded convolution + 4-row border strips, its index maps
    restated in oracle/folded_upconv.py) equals Upsample -> ReflectionPad2d -> Conv2d: values, data and weight gradient"""
    B, Ci, Co, H, W = shape
    g = torch.Generator().manual_seed(3 + sum(shape))
    x = torch.randn(B, Ci, H, W, dtype=torch.float64, generator=g, requires_grad=True)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64, generator=g, requires_grad=True)
    b = torch.randn(Co, dtype=torch.float64, generator=g)
    gy = torch.randn(B, Co, 2 * H, 2 * W, dtype=torch.float64, generator=g)
    ref = FU.reference_block(x, w, b)
    gx_ref, gw_ref = torch.autograd.grad((ref * gy).sum(), (x, w))
    got = FU.folded_block_strips(x, w, b)
    gx, gw = torch.autograd.grad((got * gy).sum(), (x, w))
    assert (got - ref).abs().max() < 1e-12
    assert (gx - gx_ref).abs().max() < 1e-12 and (gw - gw_ref).abs().max() < 1e-11


@pytest.mark.parametrize("n", [2, 3, 8, 13])
def test_border_operator_difference_lives_on_four_rows(n):
    """R - R0 (true minus zero-extended up-sampling) is non-zero on padded rows {-1, 0, 2n-1, 2n} only, with the
    coefficients of the strip table in csrc/upfold.cu; both operators against torch's own up-sampling + padding"""
    R, R0 = FU.up_matrices(n)
    x = torch.randn(1, 1, n, dtype=torch.float64, generator=torch.Generator().manual_seed(n))
    up = torch.nn.functional.interpolate(x, scale_factor=2, mode="linear", align_corners=False)
    assert (torch.nn.functional.pad(up, (1, 1), mode="reflect")[0, 0] - R @ x[0, 0]).abs().max() < 1e-14
    d = R - R0
    want = torch.zeros_like(d)
    want[0, 0] += 0.5; want[0, 1] += 0.25; want[1, 0] += 0.25
    want[2 * n, n - 1] += 0.25; want[2 * n + 1, n - 2] += 0.25; want[2 * n + 1, n - 1] += 0.5
    assert (d - want).abs().max() < 1e-14
