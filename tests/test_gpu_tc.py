"""GPU: tcgen05/TMA convolution engine against torch-CPU fp32 convolution on the SAME bf16-rounded
operands.  With fp32 output the only difference is accumulation order (1e-4); with bf16 output
the result is additionally rounded once (parity class 1e-2, north_star "bf16 GEMM inputs")."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


TC_CASES = [
    # (B, Cin, H, W, Cout, k, stride, pad)
    (2, 64, 16, 16, 64, 3, 1, 1),     # SW128 rows, one chunk
    (2, 32, 16, 16, 64, 3, 1, 1),     # SW64 rows
    (2, 16, 16, 16, 32, 5, 1, 2),     # SW32 rows: STN conv2 (model.py:207)
    (1, 128, 18, 18, 64, 3, 1, 0),    # decoder-style 3x3 p0 on an up-padded map, 2 chunks (model.py:359-363)
    (2, 256, 18, 18, 128, 3, 1, 0),   # decoder d1 (model.py:359)
    (1, 64, 66, 66, 32, 3, 1, 0),     # decoder d3 (model.py:367)
    (2, 32, 32, 32, 64, 4, 2, 1),     # encoder c2: stride 2 via tensor-map elementStrides (model.py:292)
    (2, 64, 32, 32, 128, 4, 2, 1),    # encoder c3
    (2, 128, 16, 16, 256, 4, 2, 1),   # encoder c4: 8x8 outputs, two images per tile, N = 256
    (8, 64, 8, 8, 128, 4, 2, 1),      # 4x4 outputs, eight images per tile
]


@pytest.fixture(params=[1, 0], ids=["halo", "per_tap"], autouse=True)
def halo_mode(request):
    """every tensor-core convolution test runs on both kernels: haloed tile + row-shifted taps
    (default) and one TMA box per tap"""
    from livae import _lib
    _lib.lib().livae_tc_set_halo_mode(request.param)
    _lib.lib().livae_tc_set_wgrad_halo(request.param)
    yield
    _lib.lib().livae_tc_set_halo_mode(1)
    _lib.lib().livae_tc_set_wgrad_halo(1)


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_conv_forward(case):
    from livae import ops
    B, Ci, H, W, Co, k, s, p = case
    assert ops.tc_conv_supported(B, H, W, Ci, Co, k, k, s, p)
    rng = np.random.default_rng(sum(case))
    x = _bf(torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32)))
    w = _bf(torch.tensor((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)))
    b = torch.tensor(rng.standard_normal(Co).astype(np.float32) * 0.1)
    want = F.conv2d(x, w, b, stride=s, padding=p)
    xd = _nhwc(x).cuda().to(torch.bfloat16)
    wp = ops.tc_pack_weights(w.cuda(), Co, Ci, k, k, 0)
    y32 = ops.tc_conv(xd, wp, b.cuda(), k, k, s, p, 0, out_f32=True)
    assert rel_l2(y32.cpu(), _nhwc(want)) < 1e-4
    y16 = ops.tc_conv(xd, wp, b.cuda(), k, k, s, p, 1, out_f32=False)
    assert rel_l2(y16.float().cpu(), _nhwc(torch.relu(want))) < 5e-3


def test_tc_conv_as_stride1_dgrad_with_relu_mask():
    """data gradient of a 3x3 p0 conv = forward conv over gy with flipped/transposed weights, pad 2"""
    from livae import ops
    rng = np.random.default_rng(1)
    B, Ci, H, W, Co, k = 2, 64, 18, 18, 32, 3
    x = torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32), requires_grad=True)
    w = _bf(torch.tensor((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)))
    gy = _bf(torch.tensor(rng.standard_normal((B, Co, H - 2, W - 2)).astype(np.float32)))
    (F.conv2d(x, w) * gy).sum().backward()
    wp = ops.tc_pack_weights(w.cuda(), Co, Ci, k, k, 1)
    assert tuple(wp.shape) == (9, Ci, Co)
    mask = torch.tensor(rng.standard_normal((B, H, W, Ci)).astype(np.float32)).cuda().to(torch.bfloat16)
    gx = ops.tc_conv(_nhwc(gy).cuda().to(torch.bfloat16), wp, None, k, k, 1, k - 1, 0, out_f32=True, relu_mask=mask)
    want = _nhwc(x.grad) * (mask.float().cpu() > 0)
    assert rel_l2(gx.cpu(), want) < 1e-4


DGRAD_CASES = [
    # (B, Cin, H, W, Cout, k, stride, pad): gradient w.r.t. the [B,Cin,H,W] input
    (2, 64, 18, 18, 32, 3, 1, 0),     # decoder 3x3 p0 over an up-padded map: ragged 18x18 tiles
    (1, 128, 34, 34, 64, 3, 1, 0),    # decoder d2
    (3, 16, 16, 16, 32, 5, 1, 2),     # STN conv2 (N = 16)
    (2, 32, 32, 32, 64, 4, 2, 1),     # encoder c2: four output-parity phases
    (2, 64, 16, 16, 128, 4, 2, 1),    # encoder c3
    (3, 128, 8, 8, 256, 4, 2, 1),     # encoder c4: 4x4 phase grids, several images per tile, ragged batch
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_tc_conv_dgrad(case):
    from livae import ops
    B, Ci, H, W, Co, k, s, p = case
    rng = np.random.default_rng(sum(case) + 1)
    x = torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32), requires_grad=True)
    w = _bf(torch.tensor((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)))
    y = F.conv2d(x, w, stride=s, padding=p)
    gy = _bf(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32)))
    (y * gy).sum().backward()
    wp = ops.tc_pack_weights(w.cuda(), Co, Ci, k, k, 2)
    mask = torch.tensor(rng.standard_normal((B, H, W, Ci)).astype(np.float32)).cuda().to(torch.bfloat16)
    gx = ops.tc_conv_dgrad(_nhwc(gy).cuda().to(torch.bfloat16), wp, None, H, W, k, k, s, p, 0, out_f32=True,
                           relu_mask=mask)
    want = _nhwc(x.grad) * (mask.float().cpu() > 0)
    assert rel_l2(gx.cpu(), want) < 1e-4
    # the same kernel is nn.ConvTranspose2d forward (+bias, +ReLU, bf16 out)
    if s == 2:
        b = torch.tensor(rng.standard_normal(Ci).astype(np.float32) * 0.1)
        want_t = torch.relu(F.conv_transpose2d(gy, w, b, stride=2, padding=p))
        got_t = ops.tc_conv_dgrad(_nhwc(gy).cuda().to(torch.bfloat16), wp, b.cuda(), H, W, k, k, s, p, 1)
        assert rel_l2(got_t.float().cpu(), _nhwc(want_t)) < 5e-3


WGRAD_CASES = [
    (2, 64, 18, 18, 32, 3, 1, 0),     # decoder d3-like: 2 taps per 128-row group, N = 32 (SW64 gy)
    (2, 128, 18, 18, 64, 3, 1, 0),    # decoder d2-like: 1 tap per group
    (2, 256, 18, 18, 128, 3, 1, 0),   # decoder d1: 2 groups per tap, 18 groups over 5 CTAs-sets
    (3, 16, 16, 16, 32, 5, 1, 2),     # STN conv2: 8 taps per group (SW32 x boxes), padded last group
    (2, 32, 32, 32, 64, 4, 2, 1),     # encoder c2 (stride 2 via elementStrides)
    (2, 64, 16, 16, 128, 4, 2, 1),    # encoder c3
    (4, 128, 16, 16, 256, 4, 2, 1),   # encoder c4: N = 256, G = 2
    (16, 64, 4, 4, 128, 4, 2, 1),     # tiny maps: 16 images per 64-pixel tile
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_tc_conv_wgrad(case):
    from livae import ops
    B, Ci, H, W, Co, k, s, p = case
    rng = np.random.default_rng(sum(case) + 2)
    x = _bf(torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32)))
    w = torch.zeros(Co, Ci, k, k, requires_grad=True)
    b = torch.zeros(Co, requires_grad=True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    gy = _bf(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32)))
    (y * gy).sum().backward()
    gw, gb = ops.tc_conv_wgrad(_nhwc(x).cuda().to(torch.bfloat16), _nhwc(gy).cuda().to(torch.bfloat16), k, k, s, p)
    assert rel_l2(gw.cpu(), w.grad) < 1e-4
    assert rel_l2(gb.cpu(), b.grad) < 1e-4


def test_cast_roundtrip():
    from livae import ops
    x = torch.randn(1000003).cuda()
    b = ops.cast(x, torch.bfloat16)
    assert torch.equal(b, x.to(torch.bfloat16))
    assert torch.equal(ops.cast(b, torch.float32), b.float())


@pytest.mark.parametrize("B,H", [(3, 16), (2, 64), (5, 32)])
def test_conv5pool_space_to_depth(B, H):
    """STN conv2 (model.py:207-209) as a 3x3 convolution over 2x2 pixel blocks with the max-pool in the
    epilogue (csrc/conv_s2d.cu) against conv2d + relu + max_pool2d on the same bf16-rounded operands"""
    import torch.nn.functional as F
    from livae import ops
    rng = np.random.default_rng(B * H)
    bfr = lambda t: t.to(torch.bfloat16).to(torch.float32)
    x = bfr(torch.tensor(rng.random((B, 16, H, H)).astype(np.float32))).requires_grad_(True)
    w = bfr(torch.tensor((rng.standard_normal((32, 16, 5, 5)) * 0.08).astype(np.float32))).requires_grad_(True)
    b = torch.tensor((rng.standard_normal(32) * 0.1).astype(np.float32), requires_grad=True)
    full = torch.relu(F.conv2d(x, w, b, padding=2))
    y, ind = F.max_pool2d(full, 2, 2, return_indices=True)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    xd = nhwc(x.detach()).cuda().to(torch.bfloat16)
    got, idx = ops.conv5pool_fwd(xd, w.detach().cuda(), b.detach().cuda())
    assert rel_l2(got.float().cpu(), nhwc(y.detach())) < 5e-3
    ref_idx = nhwc((((ind // H) % 2) * 2 + (ind % H) % 2).to(torch.uint8))
    agree = (idx.cpu() == ref_idx).float().mean().item()
    assert agree > 0.97, agree
    # backward with the REFERENCE routing
    g = bfr(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))) * (y.detach() > 0)
    (y * g).sum().backward()
    gw, gb, gx = ops.conv5pool_bwd(xd, w.detach().cuda(), nhwc(g).cuda().to(torch.bfloat16), ref_idx.cuda())
    assert rel_l2(gw.cpu(), w.grad) < 1e-3, rel_l2(gw.cpu(), w.grad)
    assert rel_l2(gb.cpu(), b.grad) < 1e-3
    want_gx = nhwc(x.grad) * (nhwc(x.detach()) > 0)
    assert rel_l2(gx.float().cpu(), want_gx) < 5e-3, rel_l2(gx.float().cpu(), want_gx)


@pytest.mark.parametrize("B,H,Cin,Cout", [(3, 32, 32, 64), (2, 16, 64, 128), (4, 64, 32, 64)])
def test_dgrad_stride2_block_form(B, H, Cin, Cout):
    """data gradient of conv 4x4 s2 p1 as one 3x3 block convolution (csrc/conv_s2d.cu) vs autograd"""
    import torch.nn.functional as F
    from livae import ops
    if not ops.dgrad_s2blk_supported(H, H, Cin, Cout):
        pytest.skip("shape not eligible")
    rng = np.random.default_rng(B + H + Cin)
    bfr = lambda t: t.to(torch.bfloat16).to(torch.float32)
    x = torch.tensor(rng.standard_normal((B, Cin, H, H)).astype(np.float32), requires_grad=True)
    w = bfr(torch.tensor((rng.standard_normal((Cout, Cin, 4, 4)) / np.sqrt(16 * Cin)).astype(np.float32)))
    y = F.conv2d(x, w, None, stride=2, padding=1)
    g = bfr(torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32)))
    (y * g).sum().backward()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    mask = torch.tensor(rng.standard_normal((B, H, H, Cin)).astype(np.float32)).cuda().to(torch.bfloat16)
    gx = ops.dgrad_s2blk(nhwc(g).cuda().to(torch.bfloat16), w.cuda(), H, H, relu_mask=mask)
    want = nhwc(x.grad) * (mask.float().cpu() > 0)
    assert rel_l2(gx.float().cpu(), want) < 5e-3, rel_l2(gx.float().cpu(), want)


@pytest.mark.parametrize("B,K,J", [(70, 1024, 32), (33, 520, 16), (256, 4096, 64), (5, 8, 16)])
def test_linear_dgrad_skinny_kernel(B, K, J):
    """csrc/skinny.cu: gx = (mask > 0) * (g @ w^T) for the wide Linear layers, against fp32 matmul of the same bf16
    operands; ragged B / K (not multiples of the 32 x 512 block tile)"""
    from livae._lib import call
    g = torch.Generator(device="cuda").manual_seed(B + K + J)
    gg = torch.randn(B, J, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(K, J, device="cuda", generator=g) / J ** 0.5).to(torch.bfloat16)
    mask = torch.randn(B, K, device="cuda", generator=g).clamp_min(0).to(torch.bfloat16)
    out = torch.full((B, K), 7.0, device="cuda", dtype=torch.bfloat16)
    call("livae_linear_dgrad", gg, w, mask, B, K, J, out)
    want = (gg.float() @ w.float().t()) * (mask.float() > 0)
    assert float((out.float() - want).abs().max()) <= 1e-2 * float(want.abs().max())
    assert bool(((out.float() == 0) | (mask.float() > 0)).all())
    out2 = torch.empty_like(out)
    call("livae_linear_dgrad", gg, w, None, B, K, J, out2)
    want2 = gg.float() @ w.float().t()
    assert float((out2.float() - want2).abs().max()) <= 1e-2 * float(want2.abs().max())
