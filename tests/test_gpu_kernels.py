"""GPU parity tests, one kernel family at a time, through the C ABI (ctypes -> liblivae_sm100.so).
Checker: oracle/ (numpy / torch-CPU restatement pinned to the reference) and plain torch-CPU fp32
ops for the dense layers.  Tolerances: bit-exact for the patch gather; 1e-4 relative for fp32
kernels (north_star)."""
import hashlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import patch as OP
from oracle import rot_sample as ORS
from tests.golden.make_golden import synth_image
from tests.util import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from livae import ops as o
    return o


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(dtype).cuda().contiguous()


# ------------------------------------------------------------------ a1: patch gather (bit exact)
def test_patch_gather_bit_exact_golden(ops):
    g = load_golden("patch_gather.npz")
    for HW, P in ((2048, 128), (1024, 64)):
        img = synth_image(HW, 100 + HW)                     # float64, as the reference caches it
        sites = g[f"sites{HW}"]
        s3 = np.concatenate([np.zeros((len(sites), 1), dtype=np.int64), sites], 1).astype(np.int32)
        want = OP.gather_integer([img], s3, P)
        for dt in (torch.float64, torch.float32):
            imgs = torch.from_numpy(img).to(dt).cuda()[None].contiguous()
            got = ops.patch_gather(imgs, torch.from_numpy(s3).cuda(), P).cpu().numpy()
            if dt == torch.float64:
                assert hashlib.sha256(got.tobytes()).hexdigest() == str(g[f"sha{HW}"])
            assert np.array_equal(got, want)


def test_patch_gather_multi_image_borders_empty(ops):
    rng = np.random.default_rng(5)
    imgs = rng.random((3, 96, 80))
    P = 32
    sites = np.array([[0, 16, 16], [1, 80, 64], [2, 0, 0], [2, 95, 79], [1, 50, 3]], dtype=np.int32)
    got = ops.patch_gather(dev(imgs, torch.float64), torch.from_numpy(sites).cuda(), P).cpu().numpy()
    pad = np.zeros((3, 96 + 2 * P, 80 + 2 * P))
    pad[:, P:-P, P:-P] = imgs
    for n, (i, cy, cx) in enumerate(sites):
        want = pad[i, cy + P - P // 2:cy + P + P // 2, cx + P - P // 2:cx + P + P // 2].astype(np.float32)
        assert np.array_equal(got[n, 0], want)
    empty = ops.patch_gather(dev(imgs, torch.float64), torch.zeros((0, 3), dtype=torch.int32).cuda(), P)
    assert empty.shape == (0, 1, P, P)


def test_patch_gather_subpixel_golden_and_oracle(ops):
    """a2: AdaptiveLatticeDataset.__getitem__ (data.py:478-560, transform=None): float sites, zero padding at the
    image border, per-patch min-max.  Against the reference's own output (golden, 5e-5: its affine grid is fp32)
    and against the float64 oracle (1e-6 before the min-max)."""
    g = load_golden("patch_gather.npz")
    img = synth_image(512, 300)
    sites = g["sub_sites"]
    n = len(sites)
    idx = torch.zeros(n, dtype=torch.int32).cuda()
    yx = torch.from_numpy(np.ascontiguousarray(sites, dtype=np.float64)).cuda()
    for dt in (torch.float64, torch.float32):
        imgs = torch.from_numpy(img).to(dt).cuda()[None].contiguous()
        raw = ops.patch_gather_subpixel(imgs, idx, yx, 64)
        want_raw = np.stack([OP.gather_subpixel(img, cy, cx, 64, 8, normalise=False) for cy, cx in sites])
        assert np.abs(raw.cpu().numpy() - want_raw).max() < 1e-6
        got = ops.patch_minmax_(raw).cpu().numpy()
        assert np.abs(got - g["sub_out"]).max() < 5e-5
    # large coordinates (fp32 ulp of 4096 is 5e-4 pixels: sites are float64) and a second image
    rng = np.random.default_rng(9)
    big = rng.random((2, 300, 4200))
    s2 = np.array([[150.37, 4100.61], [149.5, 4150.25], [0.2, 0.7]])
    i2 = np.array([1, 0, 1], dtype=np.int32)
    raw = ops.patch_gather_subpixel(torch.from_numpy(big).cuda(), torch.from_numpy(i2).cuda(), torch.from_numpy(s2).cuda(), 32)
    want = np.stack([OP.gather_subpixel(big[i], cy, cx, 32, 8, normalise=False) for i, (cy, cx) in zip(i2, s2)])
    assert np.abs(raw.cpu().numpy() - want).max() < 1e-6
    assert ops.patch_gather_subpixel(torch.from_numpy(big).cuda(), torch.zeros(0, dtype=torch.int32).cuda(),
                                     torch.zeros((0, 2), dtype=torch.float64).cuda(), 32).shape == (0, 1, 32, 32)


def test_patch_minmax(ops):
    rng = np.random.default_rng(6)
    p = rng.random((5, 1, 64, 64)).astype(np.float32) * 7 - 3
    p[3] = 2.5                                              # constant patch -> zeros (data.py:558)
    want = np.stack([(q - q.min()) / (q.max() - q.min()) if q.max() > q.min() else np.zeros_like(q) for q in p])
    got = ops.patch_minmax_(dev(p)).cpu().numpy()
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ a6: rotate + sample
@pytest.mark.parametrize("shape", [(8, 1, 16, 16), (8, 1, 32, 32), (3, 2, 24, 40), (2, 1, 128, 128), (2, 1, 18, 22)])
def test_rot_sample_fwd_bwd_vs_oracle(ops, shape):
    B, C, H, W = shape
    rng = np.random.default_rng(11 + H)
    x = rng.random(shape).astype(np.float32)
    go = rng.standard_normal(shape).astype(np.float32)
    th = np.array([0.3, 2.5, -1.7, np.pi / 4, 0.05, -3.0, 1.2, 6.0])[:B].astype(np.float32)
    c, s = np.cos(th).astype(np.float32), np.sin(th).astype(np.float32)
    for sgn in (1.0, -1.0):
        want = ORS.rot_sample_fwd(x, c, sgn * s)
        gx, gc, gs, _, _ = ORS.rot_sample_bwd(x, c, sgn * s, go)
        xi = dev(x).requires_grad_(True)
        cs = dev(np.stack([c, s], 1)).requires_grad_(True)
        out = ops.rot_sample(xi, cs, sgn)
        out.backward(dev(go))
        assert np.abs(out.detach().cpu().numpy() - want).max() < 2e-5
        assert np.abs(xi.grad.cpu().numpy() - gx).max() < 1e-4 * max(1.0, np.abs(gx).max())
        assert rel_l2(cs.grad[:, 0].cpu(), gc) < 1e-4
        assert rel_l2(cs.grad[:, 1].cpu(), sgn * gs) < 1e-4


def test_rot_sample_golden_reference(ops):
    """directly against outputs of the reference's F.affine_grid + F.grid_sample"""
    g = load_golden("rot_sample.npz")
    th = g["thetas"].astype(np.float32)
    cs = dev(np.stack([np.cos(th), np.sin(th)], 1))
    for HW in (16, 32):
        rng = np.random.default_rng(7 + HW)
        x = rng.random((len(th), 1, HW, HW)).astype(np.float32)
        go = rng.standard_normal((len(th), 1, HW, HW)).astype(np.float32)
        xi = dev(x).requires_grad_(True)
        csr = cs.clone().requires_grad_(True)
        out = ops.rot_sample(xi, csr, 1.0)
        out.backward(dev(go))
        assert np.abs(out.detach().cpu().numpy() - g[f"out{HW}"]).max() < 2e-5
        assert np.abs(xi.grad.cpu().numpy() - g[f"gx{HW}"]).max() < 5e-5
        gen = [4, 5, 6, 7]   # generic angles (multiples of pi/2 sit on derivative discontinuities)
        assert rel_l2(csr.grad[gen, 0].cpu(), g[f"gc{HW}"][gen]) < 1e-4
        assert rel_l2(csr.grad[gen, 1].cpu(), g[f"gs{HW}"][gen]) < 1e-4


def test_rot_sample_identity_and_empty(ops):
    x = torch.rand(2, 1, 32, 32).cuda()
    cs = torch.tensor([[1.0, 0.0], [1.0, 0.0]]).cuda()
    assert torch.allclose(ops.rot_sample(x, cs, 1.0), x, atol=1e-6)
    e = ops.rot_sample(torch.zeros(0, 1, 32, 32).cuda(), torch.zeros(0, 2).cuda(), 1.0)
    assert e.shape == (0, 1, 32, 32)


# ------------------------------------------------------------------ heads, reparam, losses
def test_stn_head_and_angle(ops):
    rng = np.random.default_rng(2)
    v = rng.standard_normal((64, 2)).astype(np.float32)
    v[0] = [3e-7, -2e-7]                                    # below the F.normalize eps
    vt = torch.tensor(v, requires_grad=True)
    u = F.normalize(vt, dim=1, eps=1e-6)
    th = torch.atan2(u[:, 1:2], u[:, 0:1])
    gcs = torch.tensor(rng.standard_normal((64, 2)).astype(np.float32))
    gth = torch.tensor(rng.standard_normal((64, 1)).astype(np.float32))
    ((u * gcs).sum() + (th * gth).sum()).backward()
    vd = dev(v).requires_grad_(True)
    cs, theta = ops.stn_head(vd)
    ((cs * gcs.cuda()).sum() + (theta * gth.cuda()).sum()).backward()
    assert torch.allclose(cs.detach().cpu(), u.detach(), atol=1e-6)
    assert torch.allclose(theta.detach().cpu(), th.detach(), atol=1e-6)
    assert rel_l2(vd.grad[1:].cpu(), vt.grad[1:]) < 1e-4
    t = torch.tensor(rng.uniform(-3, 3, (33, 1)).astype(np.float32), requires_grad=True)
    w = torch.tensor(rng.standard_normal((33, 2)).astype(np.float32))
    (torch.cat([torch.cos(t), torch.sin(t)], 1) * w).sum().backward()
    td = t.detach().cuda().requires_grad_(True)
    cs2 = ops.angle_to_cs(td)
    (cs2 * w.cuda()).sum().backward()
    assert torch.allclose(cs2.detach().cpu(), torch.cat([torch.cos(t), torch.sin(t)], 1).detach(), atol=1e-6)
    assert rel_l2(td.grad.cpu(), t.grad) < 1e-5


def test_reparam_elbo_cycle(ops):
    rng = np.random.default_rng(3)
    B, Ld, P = 6, 5, 20
    mu = torch.tensor(rng.standard_normal((B, Ld)).astype(np.float32), requires_grad=True)
    lv = torch.tensor((0.5 * rng.standard_normal((B, Ld))).astype(np.float32), requires_grad=True)
    eps = torch.tensor(rng.standard_normal((B, Ld)).astype(np.float32))
    r = torch.tensor(rng.random((B, 1, P, P)).astype(np.float32), requires_grad=True)
    x = torch.tensor(rng.random((B, 1, P, P)).astype(np.float32), requires_grad=True)
    wz = torch.tensor(rng.standard_normal((B, Ld)).astype(np.float32))
    z = mu + eps * torch.exp(0.5 * lv)
    s0 = ((r - x) ** 2).sum()
    s1 = (-0.5 * (1 + lv - mu ** 2 - lv.exp())).sum()
    (0.7 * s0 / B + 3.0 * s1 / B + (z * wz).sum()).backward()
    mud, lvd = mu.detach().cuda().requires_grad_(True), lv.detach().cuda().requires_grad_(True)
    rd, xd = r.detach().cuda().requires_grad_(True), x.detach().cuda().requires_grad_(True)
    zd = ops.reparam(mud, lvd, eps.cuda())
    sums = ops.elbo_sums(rd, xd, mud, lvd)
    (0.7 * sums[0] / B + 3.0 * sums[1] / B + (zd * wz.cuda()).sum()).backward()
    assert torch.allclose(zd.detach().cpu(), z.detach(), atol=1e-6)
    assert abs(float(sums[0]) - float(s0)) < 1e-5 * float(s0)
    assert abs(float(sums[1]) - float(s1)) < 1e-5 * abs(float(s1)) + 1e-6
    for a, b in ((mud, mu), (lvd, lv), (rd, r), (xd, x)):
        assert rel_l2(a.grad.cpu(), b.grad) < 1e-5
    # large deterministic reduction (multi-CTA path), twice -> identical bits
    a = torch.rand(64, 1, 128, 128).cuda(); b = torch.rand(64, 1, 128, 128).cuda()
    v1 = ops.elbo_sums(a, b)[0].item(); v2 = ops.elbo_sums(a, b)[0].item()
    assert v1 == v2
    assert abs(v1 - float(((a.double() - b.double()) ** 2).sum())) < 1e-5 * v1
    # cycle loss
    th = torch.tensor(rng.uniform(-3, 3, (B, 1)).astype(np.float32), requires_grad=True)
    thr = torch.tensor(rng.uniform(-3, 3, (B, 1)).astype(np.float32), requires_grad=True)
    ang = torch.tensor(rng.uniform(0, 6.28, (B,)).astype(np.float32))
    want = (1 - torch.cos((thr - th).reshape(-1) + ang)).mean()
    (2.0 * want).backward()
    thd, thrd = th.detach().cuda().requires_grad_(True), thr.detach().cuda().requires_grad_(True)
    got = ops.cycle_loss(thd.reshape(-1), thrd.reshape(-1), ang.cuda())
    (2.0 * got).backward()
    assert abs(float(got) - float(want)) < 1e-6
    assert rel_l2(thd.grad.cpu(), th.grad) < 1e-5 and rel_l2(thrd.grad.cpu(), thr.grad) < 1e-5


# ------------------------------------------------------------------ dense layers (engine 0, fp32)
def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


CONV_CASES = [
    # (B, Cin, H, W, Cout, k, stride, pad, act, pool)
    (3, 1, 16, 16, 16, 5, 1, 2, "relu", True),      # STN conv1 (model.py:204-206)
    (2, 16, 8, 8, 32, 5, 1, 2, "relu", True),       # STN conv2 (model.py:207-209)
    (3, 1, 16, 16, 32, 4, 2, 1, "relu", False),     # encoder c1 (model.py:290)
    (2, 32, 8, 8, 64, 4, 2, 1, "relu", False),      # encoder c2..c4 (model.py:292-296)
    (2, 24, 10, 10, 40, 3, 1, 0, "relu", False),    # decoder 3x3 p0 (model.py:359-367)
    (2, 32, 10, 12, 1, 3, 1, 0, "sigmoid", False),  # decoder last layer (model.py:371-372)
    (2, 64, 8, 8, 128, 4, 2, 1, "relu", False),     # encoder c3/c4 widths: 64-wide N tiles in dgrad
    (5, 128, 4, 4, 256, 4, 2, 1, "relu", False),    # encoder c4 at P=32 (odd batch)
    (3, 64, 6, 6, 48, 3, 1, 0, "relu", False),      # decoder widths, stride 1
    (5, 8, 4, 4, 7, 4, 1, 0, "none", False),        # Linear over a flattened map (model.py:302-303)
    (6, 256, 2, 2, 2, 2, 1, 0, "none", False),      # encoder heads at P=32 (model.py:302-303)
    (5, 32, 1, 1, 2, 1, 1, 0, "none", False),       # Linear(32, 2) (model.py:213)
]
_ACT = {"none": 0, "relu": 1, "sigmoid": 2}


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd(ops, case):
    B, Ci, H, W, Co, k, s, p, act, pool = case
    rng = np.random.default_rng(sum(case[:8]))
    x = torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32), requires_grad=True)
    w = torch.tensor((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32), requires_grad=True)
    b = torch.tensor(rng.standard_normal(Co).astype(np.float32) * 0.1, requires_grad=True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    y = torch.relu(y) if act == "relu" else torch.sigmoid(y) if act == "sigmoid" else y
    if pool:
        y = F.max_pool2d(y, 2, 2)
    gy = torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().requires_grad_(True)
    wd = w.detach().cuda().requires_grad_(True)
    bd = b.detach().cuda().requires_grad_(True)
    yd = ops.conv2d(xd, wd, bd, k, k, s, p, _ACT[act], pool)
    (yd * _nhwc(gy).cuda()).sum().backward()
    assert rel_l2(yd.detach().cpu(), _nhwc(y.detach())) < 1e-5
    assert rel_l2(xd.grad.cpu(), _nhwc(x.grad)) < 1e-4
    assert rel_l2(wd.grad.cpu(), w.grad) < 1e-4
    assert rel_l2(bd.grad.cpu(), b.grad) < 1e-4


@pytest.mark.parametrize("case", [(2, 16, 4, 4, 8, "relu"), (3, 8, 8, 8, 1, "sigmoid"), (1, 32, 2, 2, 16, "relu")])
def test_conv_transpose2d_fwd_bwd(ops, case):
    B, Ci, H, W, Co, act = case
    rng = np.random.default_rng(sum(case[:5]))
    x = torch.tensor(rng.standard_normal((B, Ci, H, W)).astype(np.float32), requires_grad=True)
    w = torch.tensor((rng.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Ci * 4)).astype(np.float32), requires_grad=True)
    b = torch.tensor(rng.standard_normal(Co).astype(np.float32) * 0.1, requires_grad=True)
    y = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    y = torch.relu(y) if act == "relu" else torch.sigmoid(y)
    gy = torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().requires_grad_(True)
    wd = w.detach().cuda().requires_grad_(True)
    bd = b.detach().cuda().requires_grad_(True)
    yd = ops.conv_transpose2d(xd, wd, bd, 4, 4, 2, 1, _ACT[act])
    (yd * _nhwc(gy).cuda()).sum().backward()
    assert rel_l2(yd.detach().cpu(), _nhwc(y.detach())) < 1e-5
    assert rel_l2(xd.grad.cpu(), _nhwc(x.grad)) < 1e-4
    assert rel_l2(wd.grad.cpu(), w.grad) < 1e-4
    assert rel_l2(bd.grad.cpu(), b.grad) < 1e-4


def test_linear_nhwc_matches_nchw_flatten(ops):
    rng = np.random.default_rng(9)
    B, Cc, H, W, N = 4, 6, 3, 5, 7
    x = torch.tensor(rng.standard_normal((B, Cc, H, W)).astype(np.float32), requires_grad=True)
    w = torch.tensor(rng.standard_normal((N, Cc * H * W)).astype(np.float32) * 0.1, requires_grad=True)
    b = torch.tensor(rng.standard_normal(N).astype(np.float32), requires_grad=True)
    y = torch.relu(F.linear(x.flatten(1), w, b))
    gy = torch.tensor(rng.standard_normal((B, N)).astype(np.float32))
    (y * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().requires_grad_(True)
    wd, bd = w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    yd = ops.linear_nhwc(xd, wd, bd, 1)
    (yd * gy.cuda()).sum().backward()
    assert rel_l2(yd.detach().cpu(), y.detach()) < 1e-5
    assert rel_l2(xd.grad.cpu(), _nhwc(x.grad)) < 1e-4
    assert rel_l2(wd.grad.cpu(), w.grad) < 1e-4 and rel_l2(bd.grad.cpu(), b.grad) < 1e-4


def test_encoder_chain_strict(ops):
    """4 strided convs + two Linear heads composed through autograd, identical inputs on both
    sides: every activation, activation gradient and weight gradient within 1e-5 (model.py:289-324)"""
    torch.manual_seed(0)
    B, P, Ld = 6, 32, 2
    chans = [(1, 32), (32, 64), (64, 128), (128, 256)]
    ws = [torch.randn(co, ci, 4, 4) / (ci * 16) ** 0.5 for ci, co in chans]
    bs = [torch.randn(co) * 0.1 for ci, co in chans]
    q = P // 16
    wm = torch.randn(Ld, 256 * q * q) * 0.05; wl = torch.randn(Ld, 256 * q * q) * 0.05
    x = torch.rand(B, 1, P, P)
    gm = torch.randn(B, Ld); gl = torch.randn(B, Ld)
    xc = x.clone().requires_grad_(True)
    hs = []; h = xc
    pw = [w.clone().requires_grad_(True) for w in ws]
    for w, b in zip(pw, bs):
        h = torch.relu(F.conv2d(h, w, b, stride=2, padding=1)); h.retain_grad(); hs.append(h)
    mu = F.linear(h.flatten(1), wm); lv = F.linear(h.flatten(1), wl)
    ((mu * gm).sum() + (lv * gl).sum()).backward()
    xd = x.cuda().reshape(B, P, P, 1).requires_grad_(True)
    hd = xd; hds = []
    dw = [w.cuda().requires_grad_(True) for w in ws]
    for w, b in zip(dw, bs):
        hd = ops.conv2d(hd, w, b.cuda(), 4, 4, 2, 1, 1); hd.retain_grad(); hds.append(hd)
    mud = ops.linear_nhwc(hd, wm.cuda(), None); lvd = ops.linear_nhwc(hd, wl.cuda(), None)
    ((mud * gm.cuda()).sum() + (lvd * gl.cuda()).sum()).backward()
    assert rel_l2(mud.detach().cpu(), mu.detach()) < 1e-5
    for i in range(4):
        assert rel_l2(hds[i].detach().cpu(), hs[i].detach().permute(0, 2, 3, 1)) < 1e-5
        assert rel_l2(hds[i].grad.cpu(), hs[i].grad.permute(0, 2, 3, 1)) < 1e-5
        assert rel_l2(dw[i].grad.cpu(), pw[i].grad) < 1e-5
    assert rel_l2(xd.grad.cpu().reshape(B, 1, P, P), xc.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 3, 2, 2), (2, 5, 4, 6), (1, 8, 8, 8)])
def test_upsample_pad_fwd_bwd(ops, shape):
    B, Cc, H, W = shape
    rng = np.random.default_rng(H * W)
    x = torch.tensor(rng.standard_normal(shape).astype(np.float32), requires_grad=True)
    y = F.pad(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    gy = torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gy).sum().backward()
    xd = _nhwc(x.detach()).cuda().requires_grad_(True)
    yd = ops.upsample_pad(xd)
    (yd * _nhwc(gy).cuda()).sum().backward()
    assert torch.allclose(yd.detach().cpu(), _nhwc(y.detach()), atol=1e-6)
    assert rel_l2(xd.grad.cpu(), _nhwc(x.grad)) < 1e-5


def test_decoder_fc(ops):
    rng = np.random.default_rng(4)
    B, Ld, Cc, q = 5, 3, 8, 2
    z = torch.tensor(rng.standard_normal((B, Ld)).astype(np.float32), requires_grad=True)
    w = torch.tensor(rng.standard_normal((Cc * q * q, Ld)).astype(np.float32), requires_grad=True)
    b = torch.tensor(rng.standard_normal(Cc * q * q).astype(np.float32), requires_grad=True)
    y = torch.relu(F.linear(z, w, b)).view(B, Cc, q, q)
    gy = torch.tensor(rng.standard_normal((B, Cc, q, q)).astype(np.float32))
    (y * gy).sum().backward()
    zd, wd, bd = (t.detach().cuda().requires_grad_(True) for t in (z, w, b))
    yd = ops.decoder_fc(zd, wd, bd, Cc, q)
    (yd * _nhwc(gy).cuda()).sum().backward()
    assert torch.allclose(yd.detach().cpu(), _nhwc(y.detach()), atol=1e-6)
    for a, c in ((zd, z), (wd, w), (bd, b)):
        assert rel_l2(a.grad.cpu(), c.grad) < 1e-5


def test_l2norm_clip_and_adamw(ops):
    rng = np.random.default_rng(8)
    n = 100003
    g = torch.tensor(rng.standard_normal(n).astype(np.float32))
    p = torch.tensor(rng.standard_normal(n).astype(np.float32))
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=1e-2)
    gd, pd = g.cuda(), p.cuda()
    m, v = torch.zeros_like(pd), torch.zeros_like(pd)
    step = torch.zeros(1, device="cuda")
    for it in range(3):
        pr.grad = g.clone() * (it + 1)
        nrm = torch.nn.utils.clip_grad_norm_([pr], 20.0)
        opt.step()
        gi = gd * (it + 1)
        out = ops.l2norm_clip_(gi, 20.0, apply=True)
        assert abs(float(out[0]) - float(nrm)) < 1e-4 * float(nrm)
        ops.adamw_(pd, gi, m, v, step, 1e-3, (0.9, 0.999), 1e-8, 1e-2, decoupled=True)
    assert float(step) == 3.0
    assert rel_l2(pd.cpu(), pr.detach()) < 1e-6
    assert (pd.cpu() - pr.detach()).abs().max() < 1e-5


def test_cpu_tensors_fail_loudly(ops):
    with pytest.raises(RuntimeError):
        ops.rot_sample(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2), 1.0)


@pytest.mark.parametrize("shape,win", [((3, 1, 128, 128), 11), ((2, 2, 37, 50), 11), ((1, 1, 16, 16), 5), ((4, 1, 64, 64), 11),
                                       ((2, 1, 41, 36), 11), ((1, 2, 12, 8), 11)])
def test_ssim_box_matches_reference_formula(shape, win):
    """fused box-filter SSIM (csrc/ssim.cu) against the reference's avg_pool2d formulation (train.py:606-667)"""
    import torch.nn.functional as F
    from livae.train import _ssim_dev
    rng = np.random.default_rng(shape[2] + win)
    a = torch.tensor(rng.random(shape).astype(np.float32))
    b = (a + 0.1 * torch.tensor(rng.standard_normal(shape).astype(np.float32))).clamp(0, 1)
    ap = lambda t: F.avg_pool2d(t, win, stride=1, padding=win // 2)
    mu1, mu2 = ap(a.double()), ap(b.double())
    s1, s2, s12 = ap(a.double() ** 2) - mu1 ** 2, ap(b.double() ** 2) - mu2 ** 2, ap(a.double() * b.double()) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    want = (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))).mean().item()
    got = _ssim_dev(a.cuda(), b.cuda(), win).item()
    assert abs(got - want) < 2e-5, (got, want)
    assert abs(_ssim_dev(a.cuda(), a.cuda(), win).item() - 1.0) < 1e-5
