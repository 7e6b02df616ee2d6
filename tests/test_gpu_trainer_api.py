"""GPU: the trainer-level behaviours the reference's own test-suite exercises (tests/test_train.py there: PSNR / SSIM
helpers, rotation statistics, MetricLogger, train_one_epoch / evaluate on a small stand-in model, TensorBoard helpers,
atom-position accuracy), run against the drop-in.  The stand-in is an ordinary torch model -- the loops accept any
module with the 3- or 5-tuple output contract (train.py:80-96); only the metric arithmetic runs in this repo's kernels."""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


class TinyVAE(nn.Module):
    """3-tuple output contract (recon, mu, logvar) on plain ATen layers"""

    def __init__(self, rvae=False):
        super().__init__()
        self.enc = nn.Sequential(nn.Flatten(), nn.Linear(32 * 32, 8))
        self.dec = nn.Sequential(nn.Linear(4, 32 * 32), nn.Sigmoid())
        self.rvae = rvae

    def forward(self, x):
        h = self.enc(x)
        mu, logvar = h[:, :4], h[:, 4:]
        recon = self.dec(mu + torch.randn_like(mu) * torch.exp(0.5 * logvar)).view(-1, 1, 32, 32)
        if self.rvae:
            theta = mu[:, :1] * 0.1
            return recon, recon, theta, mu, logvar
        return recon, mu, logvar


def _loader(n=12, bs=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(bs, 1, 32, 32, generator=g) for _ in range(n // bs)]


def test_psnr_and_ssim_helpers():
    from livae.train import compute_psnr, compute_ssim
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 1, 32, 32, generator=g)                  # HOST tensors, as the reference's tests pass them
    other = torch.rand(1, 1, 32, 32, generator=g)
    assert compute_psnr(img, img) == float("inf")
    p = compute_psnr(img, other)
    assert 0 < p < 100 and abs(p - 10 * math.log10(1.0 / float(((img - other) ** 2).mean()))) < 1e-3
    assert compute_psnr(img, img + 1e-3) > 50
    assert compute_psnr(img * 255, other * 255, max_val=255.0) == pytest.approx(p, abs=1e-3)
    assert compute_ssim(img, img) == pytest.approx(1.0, abs=1e-5)
    assert compute_ssim(img, other) < 0.9
    assert compute_ssim(img, img + 1e-3 * other) > 0.99
    assert compute_ssim(img, other, window_size=7) != compute_ssim(img, other, window_size=11)
    assert compute_ssim(img.to(DEV), other.to(DEV)) == pytest.approx(compute_ssim(img, other), abs=1e-6)


def test_rotation_stats_and_metric_logger():
    from livae.train import MetricLogger, get_rotation_stats
    rot = torch.tensor([[1.0, 0.0]] * 5)
    m, s = get_rotation_stats(rot)
    assert m == pytest.approx(0.0, abs=1e-5) and s == pytest.approx(0.0, abs=1e-5)
    ang = torch.tensor([0.0, 90.0, 135.0]) * math.pi / 180
    m, s = get_rotation_stats(torch.stack([torch.cos(ang), torch.sin(ang)], 1))
    assert m == pytest.approx(75.0, abs=1e-3) and s > 10
    lg = MetricLogger()
    assert len(lg.metrics) == 0
    lg.update(a=1.0, b=torch.tensor(2.0))
    lg.update(a=3.0, b=4.0)
    assert lg.metrics["a"] == [1.0, 3.0] and lg.get_averages() == {"a": 2.0, "b": 3.0}
    lg.reset()
    assert len(lg.metrics) == 0


@pytest.mark.parametrize("rvae", [False, True])
def test_train_one_epoch_and_evaluate_on_a_plain_torch_model(rvae):
    from livae.loss import VAELoss
    from livae.train import MetricLogger, evaluate, train_one_epoch
    torch.manual_seed(0)
    model = TinyVAE(rvae).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    before = [p.detach().clone() for p in model.parameters()]
    tl, vl = MetricLogger(), MetricLogger()
    for _ in range(2):
        train_one_epoch(model, _loader(), opt, VAELoss(beta=1.0), tl, DEV)
        assert model.training
    tm = tl.get_averages()
    for k in ("train_loss", "train_recon_loss", "train_kld_loss", "train_psnr", "train_ssim", "train_grad_norm",
              "train_latent_mean_abs", "train_latent_std", "train_rotation_std"):
        assert k in tm and np.isfinite(tm[k]), k
    assert len(tl.metrics["train_loss"]) == 2 and tm["train_psnr"] > 0 and -1 <= tm["train_ssim"] <= 1
    assert ("train_canonical_psnr" in tm) == rvae
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    snap = [p.detach().clone() for p in model.parameters()]
    evaluate(model, _loader(seed=1), VAELoss(beta=1.0), vl, DEV)
    assert not model.training
    vm = vl.get_averages()
    assert "val_loss" in vm and "val_psnr" in vm and "val_grad_norm" not in vm and np.isfinite(vm["val_loss"])
    assert all(torch.equal(a, b) for a, b in zip(snap, model.parameters()))

    class Bad(nn.Module):
        def forward(self, x):
            return x, x
    with pytest.raises(ValueError):
        evaluate(Bad(), _loader(), VAELoss(), MetricLogger(), DEV)


def test_tensorboard_helpers_and_atom_accuracy():
    from livae.train import compute_atom_position_accuracy, log_reconstructions_tensorboard, log_scalar_metrics_tensorboard

    class W:
        def __init__(self):
            self.images, self.scalars = [], []

        def add_image(self, tag, img, step):
            self.images.append((tag, tuple(img.shape), step))

        def add_scalar(self, tag, v, step):
            self.scalars.append((tag, v, step))

    w = W()
    model = TinyVAE().to(DEV)
    log_reconstructions_tensorboard(model, torch.rand(4, 1, 32, 32), w, 3, DEV, tag="t")
    assert len(w.images) == 1 and w.images[0][0] == "t/original_recon_diff" and w.images[0][2] == 3
    log_scalar_metrics_tensorboard(w, {"x": 1.5}, 9, prefix="p/")
    assert w.scalars == [("p/x", 1.5, 9)]
    # peaks on a 12 px square grid, reconstruction shifted by one pixel: all found, 1 px mean error
    img = torch.zeros(1, 64, 64)
    for y in range(8, 64, 12):
        for x in range(8, 64, 12):
            img[0, y, x] = 1.0
    rec = torch.roll(img, shifts=1, dims=2)
    acc = compute_atom_position_accuracy(img, rec, lattice_spacing=12.0)
    assert acc["n_original_atoms"] == acc["n_reconstructed_atoms"] > 0
    assert acc["atom_position_accuracy"] == 1.0 and acc["atom_mean_position_error"] == pytest.approx(1.0)
    with pytest.raises(ValueError):
        compute_atom_position_accuracy(img, rec, lattice_spacing=0.0)
