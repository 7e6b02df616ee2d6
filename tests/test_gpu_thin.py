"""GPU: thin 1-channel layers and bf16 helper kernels of the tensor-core path (csrc/thin.cu,
csrc/aux.cu bf16 variants, Linear-as-GEMM in livae/tc.py) against torch-CPU fp32 ops on the same
bf16-rounded operands.  fp32 outputs: 1e-4; bf16 outputs: 5e-3 (one rounding)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _bf(t):
    return t.to(BF).to(torch.float32)


def _call(*a):
    from livae._lib import call
    call(*a)


@pytest.mark.parametrize("B,P", [(3, 32), (2, 128), (5, 48)])
def test_stn_conv1_fwd_and_wgrad(B, P):
    rng = np.random.default_rng(B * P)
    x = torch.tensor(rng.random((B, 1, P, P)).astype(np.float32))
    w = torch.tensor((rng.standard_normal((16, 1, 5, 5)) * 0.2).astype(np.float32), requires_grad=True)
    b = torch.tensor((rng.standard_normal(16) * 0.1).astype(np.float32), requires_grad=True)
    full = torch.relu(F.conv2d(x, w, b, padding=2))
    y, ind = F.max_pool2d(full, 2, 2, return_indices=True)
    a1 = torch.empty(B, P // 2, P // 2, 16, dtype=BF, device="cuda")
    idx = torch.empty(B, P // 2, P // 2, 16, dtype=torch.uint8, device="cuda")
    _call("livae_thin_conv1c_fwd", 0, x.cuda(), w.detach().cuda(), b.detach().cuda(), B, P, P, a1, idx)
    assert rel_l2(a1.float().cpu(), _nhwc(y.detach())) < 5e-3
    # backward: pooled PRE-activation gradient (already masked by y > 0), routed by idx
    g = _bf(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))) * (y.detach() > 0)
    (y * g).sum().backward()
    gw = torch.empty(16, 1, 5, 5, device="cuda"); gb = torch.empty(16, device="cuda")
    # route by the REFERENCE argmax (position 0..3 inside the 2x2 window): the kernel's own argmax may
    # differ wherever bf16 rounding of image / weights ties the pool, which is not the wgrad's business
    iy, ix = ind // P, ind % P
    ref_idx = _nhwc(((iy % 2) * 2 + (ix % 2)).to(torch.uint8)).cuda()
    # compared where the pooled output is positive: elsewhere all four ReLU outputs tie at zero, the reference
    # reports position 0, the kernel the position of the largest raw sum, and the ReLU mask removes the gradient
    live = _nhwc(y.detach() > 0).cuda()
    agree = (ref_idx == idx)[live].float().mean().item()
    assert agree > 0.97, agree
    _call("livae_thin_conv1c_wgrad", 0, x.cuda(), _nhwc(g).cuda().to(BF), ref_idx, B, P, P, gw, gb)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4
    assert rel_l2(gb.cpu(), b.grad) < 1e-4


@pytest.mark.parametrize("B,P", [(3, 32), (2, 128)])
def test_encoder_c1_fwd_wgrad_dgrad(B, P):
    rng = np.random.default_rng(B + P)
    x = torch.tensor(rng.random((B, 1, P, P)).astype(np.float32), requires_grad=True)
    w = torch.tensor((rng.standard_normal((32, 1, 4, 4)) * 0.25).astype(np.float32), requires_grad=True)
    b = torch.tensor((rng.standard_normal(32) * 0.1).astype(np.float32), requires_grad=True)
    y = torch.relu(F.conv2d(x, w, b, stride=2, padding=1))
    h1 = torch.empty(B, P // 2, P // 2, 32, dtype=BF, device="cuda")
    _call("livae_thin_conv1c_fwd", 1, x.detach().cuda(), w.detach().cuda(), b.detach().cuda(), B, P, P, h1, None)
    assert rel_l2(h1.float().cpu(), _nhwc(y.detach())) < 5e-3
    g = _bf(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))) * (y.detach() > 0)
    (y * g).sum().backward()
    gd = _nhwc(g).cuda().to(BF)
    gw = torch.empty(32, 1, 4, 4, device="cuda"); gb = torch.empty(32, device="cuda")
    _call("livae_thin_conv1c_wgrad", 1, x.detach().cuda(), gd, None, B, P, P, gw, gb)
    gx = torch.empty(B, 1, P, P, device="cuda")
    _call("livae_thin_conv1c_dgrad", gd, w.detach().cuda(), B, P, P, gx)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4 and rel_l2(gb.cpu(), b.grad) < 1e-4
    assert rel_l2(gx.cpu(), x.grad) < 1e-4


@pytest.mark.parametrize("B,P", [(2, 32), (1, 128)])
def test_decoder_d4_fwd_bwd(B, P):
    rng = np.random.default_rng(7 * B + P)
    u = _bf(torch.tensor(rng.standard_normal((B, 32, P + 2, P + 2)).astype(np.float32))).requires_grad_(True)
    w = torch.tensor((rng.standard_normal((1, 32, 3, 3)) * 0.1).astype(np.float32), requires_grad=True)
    b = torch.tensor([0.05], requires_grad=True)
    y = torch.sigmoid(F.conv2d(u, w, b))
    ud = _nhwc(u.detach()).cuda().to(BF)
    r = torch.empty(B, 1, P, P, device="cuda")
    _call("livae_thin_convc1_fwd", ud, w.detach().cuda(), b.detach().cuda(), B, P + 2, P + 2, 2, r)
    assert rel_l2(r.cpu(), y.detach()) < 2e-5
    g1 = torch.tensor(rng.standard_normal((B, 1, P, P)).astype(np.float32))
    g2 = torch.tensor(rng.standard_normal((B, 1, P, P)).astype(np.float32))
    (y * (g1 + g2)).sum().backward()
    gpre = torch.empty_like(r)
    _call("livae_sigmoid_bwd", r, g1.cuda(), g2.cuda(), r.numel(), gpre)
    want_gpre = (g1 + g2) * y.detach() * (1 - y.detach())
    assert rel_l2(gpre.cpu(), want_gpre) < 1e-5
    gw = torch.empty(1, 32, 3, 3, device="cuda"); gb = torch.empty(1, device="cuda")
    _call("livae_thin_convc1_wgrad", ud, gpre, B, P + 2, P + 2, gw, gb)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4 and rel_l2(gb.cpu(), b.grad) < 1e-4
    gu = torch.empty(B, P + 2, P + 2, 32, dtype=BF, device="cuda")
    _call("livae_thin_conv1c_fwd", 2, gpre, w.detach().cuda(), None, B, P, P, gu, None)
    assert rel_l2(gu.float().cpu(), _nhwc(u.grad)) < 5e-3


def test_pool_unpool_upsample_decfc_bf16():
    from livae import ops
    rng = np.random.default_rng(3)
    B, Cc, H = 3, 32, 8
    full = _bf(torch.tensor(rng.standard_normal((B, Cc, H, H)).astype(np.float32)))
    y, ind = F.max_pool2d(full, 2, 2, return_indices=True)
    pooled = torch.empty(B, H // 2, H // 2, Cc, dtype=BF, device="cuda")
    idx = torch.empty(B, H // 2, H // 2, Cc, dtype=torch.uint8, device="cuda")
    fd = _nhwc(full).cuda().to(BF)
    _call("livae_maxpool_bf16", fd, B, H, H, Cc, pooled, idx)
    assert torch.equal(pooled.float().cpu(), _nhwc(y))
    g = _bf(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32)))
    want = F.max_unpool2d(g, ind, 2, 2)
    gfull = torch.empty(B, H, H, Cc, dtype=BF, device="cuda")
    _call("livae_unpool_bf16", _nhwc(g).cuda().to(BF), idx, B, H, H, Cc, gfull)
    assert torch.equal(gfull.float().cpu(), _nhwc(want))
    # upsample + reflection pad, bf16 vector kernels
    x = _bf(torch.tensor(rng.standard_normal((B, Cc, 6, 6)).astype(np.float32))).requires_grad_(True)
    up = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    gy = _bf(torch.tensor(rng.standard_normal(tuple(up.shape)).astype(np.float32)))
    (up * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().to(BF)
    out = torch.empty(B, 14, 14, Cc, dtype=BF, device="cuda")
    _call("livae_upsample_pad_fwd_bf16", xd, B, 6, 6, Cc, out)
    assert rel_l2(out.float().cpu(), _nhwc(up.detach())) < 5e-3
    mask = torch.tensor(rng.standard_normal((B, 6, 6, Cc)).astype(np.float32)).cuda().to(BF)
    gx = torch.empty(B, 6, 6, Cc, dtype=BF, device="cuda")
    _call("livae_upsample_pad_bwd_bf16", _nhwc(gy).cuda().to(BF), B, 6, 6, Cc, mask, gx)
    assert rel_l2(gx.float().cpu(), _nhwc(x.grad) * (mask.float().cpu() > 0)) < 5e-3
    # decoder fc, bf16 out / pre-activation bf16 gradient in; latent sizes on each of the three kernel variants
    for Ld, q in ((3, 2), (16, 2), (20, 1)):
        z = torch.tensor(rng.standard_normal((B, Ld)).astype(np.float32), requires_grad=True)
        w = torch.tensor(rng.standard_normal((256 * q * q, Ld)).astype(np.float32), requires_grad=True)
        b = torch.tensor(rng.standard_normal(256 * q * q).astype(np.float32), requires_grad=True)
        pre = F.linear(z, w, b).view(B, 256, q, q)
        y0 = torch.empty(B, q, q, 256, dtype=BF, device="cuda")
        _call("livae_decfc_fwd_bf16", z.detach().cuda(), w.detach().cuda(), b.detach().cuda(), B, Ld, 256, q * q, y0)
        assert rel_l2(y0.float().cpu(), _nhwc(torch.relu(pre).detach())) < 5e-3
        gp = _bf(torch.tensor(rng.standard_normal((B, 256, q, q)).astype(np.float32)))
        (pre * gp).sum().backward()
        gw = torch.empty_like(w).cuda(); gb = torch.empty(256 * q * q, device="cuda"); gz = torch.empty(B, Ld, device="cuda")
        _call("livae_decfc_bwd_bf16", z.detach().cuda(), w.detach().cuda(), _nhwc(gp).cuda().to(BF), B, Ld, 256, q * q,
              gw, gb, gz)
        assert rel_l2(gw.cpu(), w.grad) < 1e-5 and rel_l2(gb.cpu(), b.grad) < 1e-5 and rel_l2(gz.cpu(), z.grad) < 1e-5


@pytest.mark.parametrize("B,Cc,H,W", [(2, 256, 8, 8), (2, 32, 64, 64), (1, 64, 37, 21), (3, 8, 2, 2), (1, 24, 5, 7)])
def test_upsample_pad_bf16_shapes(B, Cc, H, W):
    """Upsample(x2, bilinear) + ReflectionPad2d(1) (model.py:357-370) and its adjoint at the decoder's shapes,
    across row-chunk boundaries (H > 16), ragged sizes and a channel count that takes the flat kernel (24)."""
    rng = np.random.default_rng(B * Cc + H)
    x = _bf(torch.tensor(rng.standard_normal((B, Cc, H, W)).astype(np.float32))).requires_grad_(True)
    up = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    gy = _bf(torch.tensor(rng.standard_normal(tuple(up.shape)).astype(np.float32)))
    (up * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().to(BF)
    out = torch.full((B, 2 * H + 2, 2 * W + 2, Cc), float("nan"), dtype=BF, device="cuda")
    _call("livae_upsample_pad_fwd_bf16", xd, B, H, W, Cc, out)
    assert torch.equal(out.float().cpu(), _nhwc(up.detach()).to(BF).float())      # same fp32 formula, one rounding
    for mask in (None, torch.tensor(rng.standard_normal((B, H, W, Cc)).astype(np.float32)).cuda().to(BF)):
        gx = torch.full((B, H, W, Cc), float("nan"), dtype=BF, device="cuda")
        _call("livae_upsample_pad_bwd_bf16", _nhwc(gy).cuda().to(BF), B, H, W, Cc, mask, gx)
        want = _nhwc(x.grad) if mask is None else _nhwc(x.grad) * (mask.float().cpu() > 0)
        assert rel_l2(gx.float().cpu(), want) < 5e-3
        if Cc % 8 == 0 and (Cc // 8) & (Cc // 8 - 1) == 0 and W * Cc // 8 <= 256:
            # fused bias gradient: column sums of the same values (fp32, before the bf16 rounding of gx)
            gx2 = torch.empty_like(gx); gb = torch.full((Cc,), float("nan"), device="cuda")
            _call("livae_upsample_pad_bwd_bias_bf16", _nhwc(gy).cuda().to(BF), B, H, W, Cc, mask, gx2, gb)
            assert torch.equal(gx2, gx)
            assert rel_l2(gb.cpu(), want.sum((0, 1, 2))) < 1e-4
    g2 = torch.tensor(rng.standard_normal((B * H * W, Cc)).astype(np.float32)).cuda().to(BF)
    gb = torch.full((Cc,), float("nan"), device="cuda")
    _call("livae_colsum_bf16", g2, B * H * W, Cc, gb)
    assert rel_l2(gb.cpu(), g2.float().cpu().sum(0)) < 1e-5


@pytest.mark.parametrize("B,H,W", [(3, 8, 8), (2, 32, 32), (1, 5, 9)])
def test_vae_last_deconv_thin_kernels(B, H, W):
    """ConvTranspose2d(32 -> 1, k4, s2, p1) + Sigmoid (model.py:95-96) and its backward against torch-CPU fp32 on the
    same bf16-rounded input: fp32 outputs 1e-4, the bf16 data gradient 5e-3."""
    rng = np.random.default_rng(B * H + W)
    x = _bf(torch.tensor(rng.standard_normal((B, 32, H, W)).astype(np.float32))).requires_grad_(True)
    w = torch.tensor((rng.standard_normal((32, 1, 4, 4)) * 0.2).astype(np.float32), requires_grad=True)
    b = torch.tensor(rng.standard_normal(1).astype(np.float32), requires_grad=True)
    pre = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    y = torch.sigmoid(pre)
    xd = _nhwc(x.detach()).cuda().to(BF)
    out = torch.full((B, 1, 2 * H, 2 * W), float("nan"), device="cuda")
    _call("livae_thin_convt_c1_fwd", xd, w.detach().cuda(), b.detach().cuda(), B, H, W, 32, 2, out)
    assert rel_l2(out.cpu(), y.detach()) < 1e-4
    gpre = torch.tensor(rng.standard_normal(tuple(pre.shape)).astype(np.float32))
    (pre * gpre).sum().backward()
    gw = torch.full((32, 1, 4, 4), float("nan"), device="cuda"); gb = torch.full((1,), float("nan"), device="cuda")
    _call("livae_thin_convt_c1_wgrad", xd, gpre.cuda(), B, H, W, 32, gw, gb)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4 and rel_l2(gb.cpu(), b.grad) < 1e-4
    mask = torch.tensor(rng.standard_normal((B, H, W, 32)).astype(np.float32)).cuda().to(BF)
    for mk in (None, mask):
        gx = torch.full((B, H, W, 32), float("nan"), dtype=BF, device="cuda")
        _call("livae_thin_convt_c1_dgrad", gpre.cuda(), w.detach().cuda(), mk, B, H, W, 32, gx)
        want = _nhwc(x.grad) if mk is None else _nhwc(x.grad) * (mk.float().cpu() > 0)
        assert rel_l2(gx.float().cpu(), want) < 5e-3


@pytest.mark.parametrize("B,Cc,hw,N", [(5, 32, 8, 32), (130, 256, 2, 4), (64, 32, 32, 32)])
def test_linear_as_tensor_core_gemm(B, Cc, hw, N):
    """nn.Linear over an NHWC-flattened map (model.py:210-213, 321-324) via livae.tc._linear_fwd/_linear_bwd"""
    from livae import tc
    rng = np.random.default_rng(B + Cc + hw)
    K = Cc * hw * hw
    x = _bf(torch.tensor(rng.standard_normal((B, Cc, hw, hw)).astype(np.float32))).requires_grad_(True)
    w = _bf(torch.tensor((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))).requires_grad_(True)
    b = torch.tensor(rng.standard_normal(N).astype(np.float32) * 0.1)
    y = F.linear(x.flatten(1), w, b)
    Npad = (N + 15) // 16 * 16
    bp = torch.zeros(Npad); bp[:N] = b
    xd = _nhwc(x.detach()).cuda().to(BF).view(B, -1)
    got = tc._linear_fwd(xd, w.detach().cuda(), Cc, hw, hw, bp.cuda(), Npad, 0)
    assert rel_l2(got[:, :N].cpu(), y.detach()) < 1e-4
    g = _bf(torch.tensor(rng.standard_normal((B, N)).astype(np.float32)))
    (y * g).sum().backward()
    gp = torch.zeros(B, Npad); gp[:, :N] = g
    mask = torch.tensor(rng.standard_normal((B, K)).astype(np.float32)).cuda().to(BF)
    gw, gb, gx = tc._linear_bwd(xd, w.detach().cuda(), Cc, hw, hw, gp.cuda().to(BF), Npad, mask)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4
    assert rel_l2(gb.cpu(), g.sum(0)) < 1e-4
    want_gx = _nhwc(x.grad).reshape(B, -1) * (mask.float().cpu() > 0)
    assert rel_l2(gx.float().cpu(), want_gx) < 5e-3


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 64, 64), (3, 24, 40), (2, 8, 8), (2, 20, 36), (300, 16, 16)])
def test_decoder_d4_fused_backward(B, H, W):
    """livae_upconv_c1_bwd == autograd of Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv3x3(32 -> 1)
    (model.py:369-372): data gradient w.r.t. the low-resolution input (times its ReLU mask, bf16: 5e-3), the column
    sums of that (bias gradient of the layer below), weight and bias gradient (fp32 accumulation of tf32 products:
    1e-3).  B = 300 exercises the persistent loop (more tiles than CTAs)."""
    rng = np.random.default_rng(B * 1000 + H + W)
    x = _bf(torch.tensor(rng.standard_normal((B, 32, H, W)).astype(np.float32))).clamp_min(0).requires_grad_(True)
    w = torch.tensor((rng.standard_normal((1, 32, 3, 3)) * 0.1).astype(np.float32), requires_grad=True)
    bias = torch.zeros(1, requires_grad=True)
    up = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    y = F.conv2d(up, w, bias)
    g = torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    (y * g).sum().backward()
    xl = _nhwc(x.detach())
    want = _nhwc(x.grad) * (xl > 0)
    gx = torch.empty(B, H, W, 32, dtype=BF, device="cuda")
    gbl = torch.full((32,), 7.0, device="cuda")
    gw = torch.full((1, 32, 3, 3), 7.0, device="cuda"); gb = torch.full((1,), 7.0, device="cuda")
    _call("livae_upconv_c1_bwd", g.cuda().contiguous(), w.detach().cuda(), xl.cuda().to(BF), B, H, W, gx, gbl, gw, gb)
    assert rel_l2(gx.float().cpu(), want) < 5e-3
    assert rel_l2(gbl.cpu(), gx.float().cpu().sum((0, 1, 2))) < 1e-4
    assert rel_l2(gw.cpu(), w.grad) < 1e-3
    assert rel_l2(gb.cpu(), bias.grad) < 1e-4
    # borders are where the closed-form weights differ: check the outer two rings separately
    ring = torch.zeros(H, W, dtype=torch.bool)
    ring[:2] = ring[-2:] = True; ring[:, :2] = True; ring[:, -2:] = True
    assert rel_l2(gx.float().cpu()[:, ring], want[:, ring]) < 5e-3


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 64, 64), (3, 24, 40), (2, 8, 8), (2, 20, 36), (1, 2, 2)])
def test_decoder_d4_fused_forward(B, H, W):
    """livae_upconv_c1_fwd == sigmoid(Conv3x3(ReflectionPad2d(1)(Upsample(x2, bilinear)(x)))) (model.py:369-372) in
    fp32 from the bf16 low-resolution map: 2e-5 on the output."""
    rng = np.random.default_rng(B * 1000 + H + W)
    x = _bf(torch.tensor(rng.standard_normal((B, 32, H, W)).astype(np.float32))).clamp_min(0)
    w = torch.tensor((rng.standard_normal((1, 32, 3, 3)) * 0.1).astype(np.float32))
    bias = torch.tensor([0.05])
    up = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    pre = F.conv2d(up.double(), w.double(), bias.double()).float()
    for act, want in ((2, torch.sigmoid(pre)), (0, pre)):
        out = torch.empty(B, 1, 2 * H, 2 * W, device="cuda")
        _call("livae_upconv_c1_fwd", _nhwc(x).cuda().to(BF), w.cuda(), bias.cuda(), B, H, W, act, out)
        assert rel_l2(out.cpu(), want) < 2e-5, (act, rel_l2(out.cpu(), want))
        assert (out.cpu() - want).abs().max() < 1e-4 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("B,h,Cin,Cout", [(3, 16, 128, 64), (2, 32, 64, 32), (5, 8, 64, 32), (2, 24, 128, 32)])
def test_upfold_block_matches_upsample_pad_conv(B, h, Cin, Cout):
    """decoder blocks d2 / d3 phase-folded onto the low-resolution input (csrc/upfold.cu) against the reference
    composition Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv2d(3x3) -> ReLU (model.py:356-368) on the same
    bf16-rounded operands: values, data gradient (incl. the ReLU mask of the layer below) and weight gradient, with
    the border rows / columns checked on their own"""
    from livae import ops
    rng = np.random.default_rng(B * h + Cin)
    x = _bf(torch.tensor(np.maximum(rng.standard_normal((B, Cin, h, h)), 0).astype(np.float32))).requires_grad_(True)
    w = torch.tensor((rng.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(9 * Cin)).astype(np.float32), requires_grad=True)
    b = torch.tensor((rng.standard_normal(Cout) * 0.1).astype(np.float32))
    up = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    pre = F.conv2d(up, w, b)
    y_ref = torch.relu(pre)
    assert ops.upfold_supported(B, h, h, Cin, Cout)
    xd = _nhwc(x.detach()).cuda().to(BF)
    wd = w.detach().cuda()
    y, s_tb, s_lr = ops.upfold_fwd(xd, wd, b.cuda())
    got = y.float().cpu()
    want = _nhwc(y_ref.detach())
    assert rel_l2(got, want) < 5e-3
    for sl in (np.s_[:, :2], np.s_[:, -2:], np.s_[:, :, :2], np.s_[:, :, -2:]):       # the corrected border rows / columns
        assert rel_l2(got[sl], want[sl]) < 5e-3
    # backward from a pre-activation gradient that already carries this layer's ReLU mask
    g = _bf(torch.tensor(rng.standard_normal(tuple(pre.shape)).astype(np.float32))) * (pre.detach() > 0)
    (pre * g).sum().backward()
    gw, gx = ops.upfold_bwd(xd, wd, _nhwc(g).cuda().to(BF), s_tb, s_lr)
    assert rel_l2(gw.cpu(), w.grad) < 2e-3
    gx_ref = _nhwc(x.grad * (x.detach() > 0))
    gxc = gx.float().cpu()
    assert rel_l2(gxc, gx_ref) < 6e-3
    for sl in (np.s_[:, :2], np.s_[:, -2:], np.s_[:, :, :2], np.s_[:, :, -2:]):
        assert rel_l2(gxc[sl], gx_ref[sl]) < 8e-3
