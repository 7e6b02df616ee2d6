"""GPU: the reference's scripts' wiring run against the drop-in package (scripts/train_rvae.py:27-96 make_dataloaders
+ 99-317 run_training, train_vae.py, pretrain_stn.py:59-163), the Dataset classes' items against the reference's
own items, and the trainer's logged metrics / evaluate loops against the reference's (tests/golden/eval.npz,
rvae_step_p*.npz `metric/*`)."""
import random

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, random_split

from oracle import rvae as O
from tests.golden.make_golden import synth_image
from tests.golden.make_golden_r2 import synth_lattice
from tests.test_gpu_step import FixedEps
from tests.util import load_golden

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


class MockWriter:
    def __init__(self):
        self.images, self.scalars = [], []

    def add_image(self, tag, img, step):
        self.images.append((tag, tuple(img.shape), step))

    def add_scalar(self, tag, v, step):
        self.scalars.append((tag, float(v), step))


def _loaders(ds, batch_size, val_split=0.1, num_workers=2, prefetch_factor=1):
    """scripts/train_rvae.py:71-95, verbatim wiring"""
    val_len = max(1, int(len(ds) * val_split))
    train_len = max(1, len(ds) - val_len)
    train_ds, val_ds = random_split(ds, [train_len, val_len])
    train_loader = DataLoader(train_ds, batch_size=batch_size, shuffle=True, num_workers=num_workers, pin_memory=True,
                              persistent_workers=True, prefetch_factor=prefetch_factor if num_workers > 0 else None,
                              drop_last=True)
    val_loader = DataLoader(val_ds, batch_size=batch_size, shuffle=False, num_workers=num_workers, pin_memory=True,
                            persistent_workers=True, prefetch_factor=prefetch_factor if num_workers > 0 else None)
    return train_loader, val_loader


@pytest.fixture(scope="module")
def micrographs():
    return [synth_lattice(512, 16.0, 7.0, 1), synth_lattice(512, 19.0, 33.0, 2)]


@pytest.fixture(autouse=True)
def _engine():
    import livae
    livae.set_engine("tc")
    yield
    livae.set_engine("tc")


def test_train_rvae_script_wiring_one_epoch(micrographs, tmp_path):
    """run_training of scripts/train_rvae.py (:118-317) for one epoch: constructor from raw images (site finding),
    DataLoader workers + pin_memory, AdamW, RVAELoss, beta schedule write, train / evaluate loops, TensorBoard
    helpers, checkpoint"""
    from livae.data import PairedAdaptiveLatticeDataset
    from livae.loss import RVAELoss
    from livae.model import RVAE
    from livae.train import (MetricLogger, evaluate_rvae, log_reconstructions_tensorboard,
                             log_scalar_metrics_tensorboard, train_rvae_one_epoch)
    P = 64
    ds = PairedAdaptiveLatticeDataset(micrographs, patch_size=P, padding=16)
    assert len(ds) > 500
    train_loader, val_loader = _loaders(ds, batch_size=128)
    first = next(iter(train_loader))
    assert isinstance(first, list) and len(first) == 3
    x, xr, ang = first
    assert x.is_cuda and xr.is_cuda and ang.is_cuda and x.shape == (128, 1, P, P) and ang.shape == (128,)
    assert float(x.min()) == 0.0 and float(x.max()) == 1.0                      # per-patch min-max (data.py:716-730)
    model = RVAE(latent_dim=2, in_channels=1, patch_size=P).to(DEV)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=1)
    criterion = RVAELoss(beta=10.0, gamma=10.0, use_diversity=False)
    scaler = torch.amp.GradScaler(DEV.type)                                   # the script's default (AMP on CUDA)
    writer = MockWriter()
    w0 = model.decoder.fc.weight.detach().clone()
    train_logger, val_logger = MetricLogger(), MetricLogger()
    criterion.beta = 5.0                                                       # scripts/train_rvae.py:221
    train_rvae_one_epoch(model, train_loader, optimizer, criterion, train_logger, DEV, scaler=scaler, grad_max_norm=None)
    evaluate_rvae(model, val_loader, criterion, val_logger, DEV)
    tm, vm = train_logger.get_averages(), val_logger.get_averages()
    for k in ("train_loss", "train_recon_loss", "train_kld_loss", "train_cycle_loss", "train_canonical_loss",
              "train_psnr", "train_ssim", "train_latent_mean_abs", "train_latent_std", "train_rotation_std",
              "train_grad_norm", "train_canonical_psnr", "train_canonical_ssim"):                # train.py:431-445
        assert k in tm and np.isfinite(tm[k]), k
    for k in ("val_loss", "val_recon_loss", "val_kld_loss", "val_cycle_loss", "val_psnr", "val_ssim"):
        assert k in vm and np.isfinite(vm[k]), k
    assert not torch.equal(w0, model.decoder.fc.weight)
    log_scalar_metrics_tensorboard(writer, tm, 1, prefix="")
    sample = next(iter(val_loader))
    log_reconstructions_tensorboard(model, sample[0][:8], writer, 1, DEV, tag="val")
    assert [t for t, _, _ in writer.images] == ["val/original_recon_diff", "val/canonical_original_recon_diff"]
    scheduler.step()
    ckpt = tmp_path / "rvae.pt"
    torch.save({"model_state": model.state_dict(), "optimizer_state": optimizer.state_dict(), "epoch": 1}, ckpt)
    sd = torch.load(ckpt)["model_state"]
    RVAE(latent_dim=2, in_channels=1, patch_size=P).load_state_dict(sd, strict=True)
    del train_loader, val_loader


def test_train_vae_script_wiring_one_epoch(micrographs):
    """scripts/train_vae.py: AdaptiveLatticeDataset(transform=default_transform) -> Adam -> train_one_epoch / evaluate"""
    from livae.data import AdaptiveLatticeDataset, default_transform
    from livae.loss import VAELoss
    from livae.model import VAE
    from livae.train import MetricLogger, evaluate, train_one_epoch
    P = 64
    ds = AdaptiveLatticeDataset(micrographs, patch_size=P, padding=16, transform=default_transform)
    train_loader, val_loader = _loaders(ds, batch_size=128)
    model = VAE(latent_dim=16, in_channels=1, patch_size=P).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    tl, vl = MetricLogger(), MetricLogger()
    train_one_epoch(model, train_loader, opt, VAELoss(beta=1.0), tl, DEV, scaler=None)
    evaluate(model, val_loader, VAELoss(beta=1.0), vl, DEV)
    assert np.isfinite(tl.get_averages()["train_loss"]) and np.isfinite(vl.get_averages()["val_loss"])
    del train_loader, val_loader


def test_pretrain_stn_script_loop(micrographs):
    """the script's OWN inner loop (scripts/pretrain_stn.py:93-120) over the drop-in's loader, model and loss"""
    from livae.data import PairedAdaptiveLatticeDataset
    from livae.loss import cycle_consistency_loss
    from livae.model import RVAE
    P = 64
    ds = PairedAdaptiveLatticeDataset(micrographs, patch_size=P, padding=16)
    train_loader, _ = _loaders(ds, batch_size=128)
    model = RVAE(latent_dim=2, in_channels=1, patch_size=P).to(DEV)
    stn_params = list(model.encoder.rotation_stn.parameters())
    optimizer = torch.optim.AdamW(stn_params, lr=1e-3, weight_decay=1e-5)
    model.train()
    losses = []
    for batch in train_loader:
        x, x_rot, angle = batch
        x = x.to(DEV); x_rot = x_rot.to(DEV); angle = angle.to(DEV)
        optimizer.zero_grad(set_to_none=True)
        _, _, theta_orig = model.encoder(x)
        _, _, theta_rot = model.encoder(x_rot)
        loss = cycle_consistency_loss(theta_orig, theta_rot, angle)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(stn_params, max_norm=5.0)
        optimizer.step()
        losses.append(loss.item())
    assert len(losses) >= 3 and all(np.isfinite(losses))
    assert all(p.grad is None for n, p in model.named_parameters() if "rotation_stn" not in n)
    del train_loader


# ---- Dataset-class items against the reference's own items ---------------------------------------------------
def test_patch_dataset_items_with_transform_match_reference():
    """PatchDataset.__getitem__ with default_transform: crop P + 2*padding, transform(rotation=True), centre crop
    (data.py:240-248) -- against items of the reference itself (tests/golden/patchds.npz)"""
    from livae.data import PatchDataset, default_transform
    g = load_golden("patchds.npz")
    HW, P, pad = 256, 64, 16
    ds = PatchDataset.__new__(PatchDataset)
    ds.patch_size, ds.padding, ds.transform = P, pad, default_transform
    ds.images, ds.atom_coords = [synth_image(HW, 5150)], [g["sites"]]
    ds._register()
    random.seed(6000)
    got = np.stack([ds[i].numpy() for i in range(len(g["sites"]))])
    assert got.shape == g["items"].shape
    assert np.abs(got - g["items"]).max() < 1e-4
    ds.transform = None
    ds._src = None
    want = np.stack([ds.images[0][cy - P // 2:cy + P // 2, cx - P // 2:cx + P // 2].astype(np.float32)[None]
                     for cy, cx in g["sites"]])
    assert np.array_equal(np.stack([ds[i].numpy() for i in range(len(want))]), want)         # bit-exact crop


@pytest.mark.parametrize("tag,P,pad", [("a", 64, 8), ("b", 32, 16)])
def test_paired_dataset_items_match_reference(tag, P, pad):
    from livae.data import AdaptiveLatticeDataset, PairedAdaptiveLatticeDataset, default_transform
    g = load_golden("augment.npz")
    img, sites = synth_image(256, 410 + P), g[f"sites_{tag}"]
    ds = PairedAdaptiveLatticeDataset.from_sites([img], [sites], P, pad, transform=default_transform)
    random.seed(1000 + P)
    items = [ds[i] for i in range(len(sites))]
    assert all(not it[0].is_cuda and it[0].shape == (1, P, P) for it in items)
    assert np.abs(np.stack([it[0].numpy() for it in items]) - g[f"pair_{tag}_x"]).max() < 1e-4
    assert np.abs(np.stack([it[1].numpy() for it in items]) - g[f"pair_{tag}_r"]).max() < 1e-4
    assert np.abs(np.array([it[2] for it in items]) - g[f"pair_{tag}_angle"]).max() < 1e-6
    ad = AdaptiveLatticeDataset.from_sites([img], [sites], P, pad, transform=default_transform)
    random.seed(3000 + P)
    assert np.abs(np.stack([ad[i].numpy() for i in range(len(sites))]) - g[f"adapt_{tag}"]).max() < 1e-4
    with pytest.raises(IndexError):
        ds[len(sites)]
    # the worker route: a recipe batch with the same draws yields the same pixels
    from livae.data import PatchRecipe, _collate_recipes
    random.seed(1000 + P)
    recipes = []
    for i in range(len(sites)):
        t, a = ds._draws(1)
        recipes.append(PatchRecipe(ds._key, i, t, float(a[0])))
    x, r, a = _collate_recipes(recipes).pin_memory()
    assert x.is_cuda and np.abs(x.cpu().numpy() - g[f"pair_{tag}_x"]).max() < 1e-4
    assert np.abs(r.cpu().numpy() - g[f"pair_{tag}_r"]).max() < 1e-4


# ---- the trainer's logged metrics and the evaluate loops against the reference's --------------------------------
METRIC_TOL = {"train_psnr": 1e-4, "train_ssim": 1e-4, "train_canonical_psnr": 1e-4, "train_canonical_ssim": 1e-4,
              "train_grad_norm": 1e-3}


@pytest.mark.parametrize("tag", ["p32", "p128"])
def test_train_rvae_one_epoch_metrics_match_reference(tag):
    """all 13 values the reference's train_rvae_one_epoch logs for the golden batch (train.py:399-445), run exactly
    as make_golden.py ran the reference: SGD lr=0, grad_max_norm=1e30"""
    import livae
    from livae.train import MetricLogger, train_rvae_one_epoch
    livae.set_engine("f32")
    g = load_golden(f"rvae_step_{tag}.npz")
    P, L, B, seed = int(g["P"]), int(g["L"]), int(g["B"]), int(g["seed"])
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    model = livae.RVAE(latent_dim=L, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    model.to(DEV)
    lg = MetricLogger()
    with FixedEps(eps):
        train_rvae_one_epoch(model, [(x, xr, ang)], torch.optim.SGD(model.parameters(), lr=0.0),
                             livae.RVAELoss(beta=10.0, gamma=10.0), lg, DEV, canonical_weight=0.2, scaler=None,
                             grad_max_norm=1e30)
    got = lg.get_averages()
    keys = [k[len("metric/"):] for k in g.files if k.startswith("metric/")]
    assert len(keys) == 13 and set(keys) == set(got)
    for k in keys:
        want = float(g["metric/" + k])
        tol = METRIC_TOL.get(k, 1e-4)
        assert abs(got[k] - want) <= tol * max(abs(want), 1e-3), (k, got[k], want)


def test_evaluate_loops_match_reference():
    """evaluate_rvae (last-batch-only quirk included, train.py:521-541), evaluate on an rVAE and on a VAE, and
    train_one_epoch's rVAE branch (train.py:85-94) against the reference's own outputs for two seeded batches"""
    import livae
    from livae.train import MetricLogger, evaluate, evaluate_rvae, train_one_epoch
    livae.set_engine("f32")
    g = load_golden("eval.npz")
    P, L, B, seed = int(g["P"]), int(g["L"]), int(g["B"]), int(g["seed"])
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    batches = [O.make_lattice_batch(B, P, seed=seed + 1 + k) for k in range(2)]
    eps = torch.from_numpy(np.random.default_rng(seed + 9).standard_normal((B, L))).float()
    model = livae.RVAE(latent_dim=L, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    model.to(DEV)

    def check(prefix, got, tol=2e-4):
        keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
        assert keys and set(keys) == set(got), (sorted(keys), sorted(got))
        for k in keys:
            want = float(g[prefix + k])
            assert abs(got[k] - want) <= tol * max(abs(want), 1e-3), (prefix + k, got[k], want)

    lg = MetricLogger()
    with FixedEps(eps):
        evaluate_rvae(model, batches, livae.RVAELoss(beta=10.0, gamma=10.0), lg, DEV, canonical_weight=0.2)
    check("rvae/", lg.get_averages())
    lg = MetricLogger()
    with FixedEps(eps):
        evaluate(model, [b[0] for b in batches], livae.VAELoss(beta=1.0), lg, DEV, canonical_weight=0.2)
    check("rvae_evaluate/", lg.get_averages())
    lg = MetricLogger()
    with FixedEps(eps):
        train_one_epoch(model, [b[0] for b in batches], torch.optim.SGD(model.parameters(), lr=0.0),
                        livae.VAELoss(beta=1.0), lg, DEV)
    check("rvae_train_one_epoch/", lg.get_averages(), tol=1e-3)
    Pv, Lv, seedv = int(g["Pv"]), int(g["Lv"]), int(g["seedv"])
    vae = livae.VAE(latent_dim=Lv, in_channels=1, patch_size=Pv)
    vae.load_state_dict(O.make_params(O.vae_param_shapes(Pv, Lv), seed=seedv), strict=True)
    vae.to(DEV)
    vb = [O.make_lattice_batch(B, Pv, seed=seedv + 1 + k)[0] for k in range(2)]
    veps = torch.from_numpy(np.random.default_rng(seedv + 9).standard_normal((B, Lv))).float()
    lg = MetricLogger()
    with FixedEps(veps):
        evaluate(vae, vb, livae.VAELoss(beta=1.0), lg, DEV)
    check("vae/", lg.get_averages())


def test_rotation_invariance_and_atom_metrics_run():
    import livae
    from livae.train import compute_atom_position_accuracy, evaluate_rotation_invariance
    model = livae.RVAE(latent_dim=2, in_channels=1, patch_size=32).to(DEV)
    imgs = O.make_lattice_batch(3, 32, seed=5)[0]
    out = evaluate_rotation_invariance(model, imgs, angles=(0, 90, 180), device=DEV, max_batches=2)
    assert set(out) == {"rotation_latent_variance", "rotation_recon_rmse", "rotation_recon_psnr",
                        "rotation_recon_ssim", "rotation_angle_error"}
    assert all(np.isfinite(v) for v in out.values())
    acc = compute_atom_position_accuracy(imgs[0], imgs[0], lattice_spacing=10.0)
    assert acc["atom_position_accuracy"] == 1.0 and acc["atom_mean_position_error"] == 0.0
