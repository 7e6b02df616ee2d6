"""GPU parity: the device-side dataset items (livae/data.py -> gather_roi / augment / rotate_crop kernels through
the C ABI) against (1) the outputs of the REFERENCE's own default_transform and Paired/AdaptiveLatticeDataset
items (tests/golden/augment.npz, Python `random` seeded identically) and (2) the float64 oracle
(oracle/augment.py).  Tolerance 1e-4 abs on [0,1] data (fp32 bilinear grid maths of torchvision; the test images
are white noise, the worst case for resampling error), bit-exact for the pure crop / flip / roll paths."""
import random

import numpy as np
import pytest
import torch

from oracle import augment as OA
from tests.golden.make_golden import synth_image
from tests.util import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def data():
    from livae import data as d
    return d


def test_default_transform_matches_reference(data):
    g = load_golden("augment.npz")
    patch = torch.from_numpy(synth_image(48, 400).astype(np.float32))[None].cuda()
    for k, rot in enumerate((False, True, False, True)):
        random.seed(900 + k)
        got = data.default_transform(patch, rotation=rot).cpu().numpy()
        assert got.shape == g[f"dt{k}"].shape
        assert np.abs(got - g[f"dt{k}"]).max() < TOL, (k, np.abs(got - g[f"dt{k}"]).max())


@pytest.mark.parametrize("tag,P,pad", [("a", 64, 8), ("b", 32, 16)])
def test_dataset_items_match_reference(data, tag, P, pad):
    g = load_golden("augment.npz")
    img = synth_image(256, 410 + P)
    sites = g[f"sites_{tag}"]
    n = len(sites)
    src = data.DevicePatchSource([img], [sites], P, pad, transform=data.default_transform)
    # paired items with the transform: one batch == n consecutive __getitem__ calls of the reference
    random.seed(1000 + P)
    x, r, a = src.paired_batch(np.arange(n))
    assert x.shape == (n, 1, P, P) and r.shape == (n, 1, P, P) and a.shape == (n,)
    assert np.abs(a.cpu().numpy() - g[f"pair_{tag}_angle"]).max() < 1e-6
    assert np.abs(x.cpu().numpy() - g[f"pair_{tag}_x"]).max() < TOL
    assert np.abs(r.cpu().numpy() - g[f"pair_{tag}_r"]).max() < TOL
    # ... and item by item in a different batching (the draw order is per item)
    random.seed(1000 + P)
    for i in range(n):
        xi, ri, ai = src.paired_batch([i])
        assert np.abs(xi.cpu().numpy()[0] - g[f"pair_{tag}_x"][i]).max() < TOL
        assert np.abs(ri.cpu().numpy()[0] - g[f"pair_{tag}_r"][i]).max() < TOL
    # adaptive items with the transform
    random.seed(3000 + P)
    ad = src.adaptive_batch(np.arange(n)).cpu().numpy()
    assert np.abs(ad - g[f"adapt_{tag}"]).max() < TOL
    # transform=None
    src_nt = data.DevicePatchSource([img], [sites], P, pad, transform=None)
    random.seed(2000 + P)
    x, r, a = src_nt.paired_batch(np.arange(n))
    assert np.abs(x.cpu().numpy() - g[f"pairnt_{tag}_x"]).max() < TOL
    assert np.abs(r.cpu().numpy() - g[f"pairnt_{tag}_r"]).max() < TOL
    with pytest.raises(IndexError):
        src.paired_batch([n])


def test_kernels_against_oracle_and_exact_paths(data):
    from livae import ops
    rng = np.random.default_rng(11)
    S, P, N = 40, 24, 7
    big = rng.random((N, 1, S, S)).astype(np.float32)
    d = torch.from_numpy(big).cuda()
    # identity crop: bit exact; min-max of a constant patch is all zeros (data.py:720-721)
    got = ops.rotate_crop(d, P).cpu().numpy()
    o = (S - P) // 2
    assert np.array_equal(got, big[:, :, o:o + P, o:o + P])
    const = torch.full((2, 1, S, S), 0.25, device="cuda")
    assert float(ops.rotate_crop(const, P, None, normalise=True).abs().max()) == 0.0
    # flips + roll only (flag bit 2): bit exact permutation
    flags = torch.tensor([4, 5, 6, 7, 5, 6, 7], dtype=torch.int32).cuda()
    shift = torch.from_numpy(rng.integers(-4, 5, size=(N, 2)).astype(np.int32)).cuda()
    got = ops.augment(d, torch.ones(N, device="cuda"), flags, shift).cpu().numpy()
    for n in range(N):
        w = big[n, 0]
        f = int(flags[n])
        if f & 1:
            w = w[:, ::-1]
        if f & 2:
            w = w[::-1, :]
        w = np.roll(w, tuple(int(v) for v in shift[n].cpu()), axis=(0, 1))
        assert np.array_equal(got[n, 0], w)
    # scale + flips + roll and rotation against the float64 oracle
    scale = rng.uniform(0.9, 1.1, N).astype(np.float32)
    got = ops.augment(d, torch.from_numpy(scale).cuda(), flags & 3, shift).cpu().numpy()
    ang = rng.uniform(0, 360, N)
    ang[0], ang[1] = 0.0, 90.0
    gotr = ops.rotate_crop(d, P, torch.from_numpy(ang).cuda(), normalise=False).cpu().numpy()
    for n in range(N):
        p = {"scale": float(scale[n]), "angle": None, "hflip": bool(int(flags[n]) & 1),
             "vflip": bool(int(flags[n]) & 2), "shift_y": int(shift[n, 0]), "shift_x": int(shift[n, 1])}
        assert np.abs(got[n, 0] - OA.default_transform(big[n, 0], p)).max() < 2e-5
        assert np.abs(gotr[n, 0] - OA._centre(OA.rotate(big[n, 0], ang[n]), P)).max() < 2e-5
    # empty batch
    assert ops.rotate_crop(d[:0], P).shape == (0, 1, P, P)
    assert ops.augment(d[:0], torch.ones(0, device="cuda"), flags[:0], shift[:0]).shape == (0, 1, S, S)


def test_loader_feeds_train_step_shapes(data):
    img = synth_image(512, 77)
    rng = np.random.default_rng(1)
    sites = rng.uniform(64, 448, size=(50, 2))
    src = data.DevicePatchSource([img, img], [sites, sites[:10]], 32, 8, transform=data.default_transform)
    assert len(src) == 60
    loader = data.DevicePatchLoader(src, 16, mode="paired", seed=3)
    assert len(loader) == 3
    batches = list(loader)
    assert len(batches) == 3
    for x, r, a in batches:
        assert x.shape == (16, 1, 32, 32) and r.shape == x.shape and a.shape == (16,)
        assert float(x.min()) == 0.0 and float(x.max()) == 1.0
    ints = data.DevicePatchSource([img], [np.round(sites)], 32, 8)
    p = ints.patch_batch([0, 1]).cpu().numpy()
    cy, cx = np.round(sites[0]).astype(int)
    assert np.array_equal(p[0, 0], img[cy - 16:cy + 16, cx - 16:cx + 16].astype(np.float32))
