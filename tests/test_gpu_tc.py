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


def test_cast_roundtrip():
    from livae import ops
    x = torch.randn(1000003).cuda()
    b = ops.cast(x, torch.bfloat16)
    assert torch.equal(b, x.to(torch.bfloat16))
    assert torch.equal(ops.cast(b, torch.float32), b.float())
