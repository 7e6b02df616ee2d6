"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_loader.py) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
Inputs and parameters come from the deterministic numpy generators in
oracle/rvae.py, so only OUTPUTS are stored.  torch version is recorded in each
file (the reference pins torch 2.9.1; this image has 2.11.0 -- SURVEY.md 8c).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import rvae as O   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = 64  # sampled gradient entries stored per parameter tensor


def sample_idx(numel, name):
    h = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    rng = np.random.default_rng(h)
    return rng.integers(0, numel, size=min(SAMPLE, numel))


def pack_grads(named_grads):
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().reshape(-1)
        out["gnorm/" + k] = np.array(g.norm().item())
        out["gsum/" + k] = np.array(g.sum().item())
        out["gsamp/" + k] = g[torch.from_numpy(sample_idx(g.numel(), k))].numpy()
    return out


class FixedEps:
    """Replace torch.randn_like with a fixed tensor (model.py:438 draws eps there)."""
    def __init__(self, eps):
        self.eps = eps
    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda t, **k: self.eps.to(t.dtype).reshape(t.shape)
    def __exit__(self, *a):
        torch.randn_like = self.orig


def golden_rot_sample(livae):
    import torch.nn.functional as F
    res = {}
    thetas = torch.tensor([0.0, np.pi / 2, -np.pi / 2, np.pi, 0.3, 2.5, -1.7, np.pi / 4])
    for HW in (16, 32):
        rng = np.random.default_rng(7 + HW)
        B = len(thetas)
        x = torch.from_numpy(rng.random((B, 1, HW, HW))).float().requires_grad_(True)
        go = torch.from_numpy(rng.standard_normal((B, 1, HW, HW))).float()
        c = torch.cos(thetas).clone().requires_grad_(True)
        s = torch.sin(thetas).clone().requires_grad_(True)
        z = torch.zeros_like(c)
        rot = torch.stack([torch.stack([c, -s, z], 1), torch.stack([s, c, z], 1)], 1)
        grid = F.affine_grid(rot, x.size(), align_corners=False)
        out = F.grid_sample(x, grid, padding_mode="reflection", align_corners=False)
        (out * go).sum().backward()
        res[f"out{HW}"] = out.detach().numpy()
        res[f"gx{HW}"] = x.grad.numpy()
        res[f"gc{HW}"] = c.grad.numpy()
        res[f"gs{HW}"] = s.grad.numpy()
    res["thetas"] = thetas.numpy()
    np.savez_compressed(os.path.join(OUT, "rot_sample.npz"), torch_version=torch.__version__, **res)


def golden_rvae_step(livae, P, L, B, seed, tag, store_images):
    from livae.model import RVAE
    from livae.loss import RVAELoss
    from livae.train import MetricLogger, train_rvae_one_epoch
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    model = RVAE(latent_dim=L, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    crit = RVAELoss(beta=10.0, gamma=10.0)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    logger = MetricLogger()
    with FixedEps(eps):
        # the reference's own trainer body (train.py:315-445); lr=0 and an infinite
        # clip norm leave p.grad exactly as loss.backward() produced it
        train_rvae_one_epoch(model, [(x, xr, ang)], opt, crit, logger, torch.device("cpu"),
                             canonical_weight=0.2, scaler=None, grad_max_norm=1e30)
    m = {k: np.array(v[0]) for k, v in logger.metrics.items()}
    res = {"metric/" + k: v for k, v in m.items()}
    res.update(pack_grads({k: p.grad for k, p in model.named_parameters()}))
    with FixedEps(eps), torch.no_grad():
        rotated, recon, theta, mu, logvar = model(x)
        _, _, theta_rot = model.encoder(xr)
    res["theta"] = theta.numpy(); res["mu"] = mu.numpy(); res["logvar"] = logvar.numpy()
    res["theta_rot"] = theta_rot.numpy()
    if store_images:
        res["rotated_recon"] = rotated.numpy(); res["recon"] = recon.numpy()
    else:
        res["rotated_recon_sum"] = np.array(rotated.double().sum().item())
        res["recon_sum"] = np.array(recon.double().sum().item())
        res["rotated_recon_b0"] = rotated[0, 0, ::8, ::8].numpy()
        res["recon_b0"] = recon[0, 0, ::8, ::8].numpy()
    np.savez_compressed(os.path.join(OUT, f"rvae_step_{tag}.npz"), torch_version=torch.__version__,
                        P=P, L=L, B=B, seed=seed, **res)
    print(tag, {k: float(v) for k, v in m.items() if "loss" in k or "grad" in k})


def golden_vae_step(livae, P, L, B, seed):
    from livae.model import VAE
    from livae.loss import VAELoss
    from livae.train import MetricLogger, train_one_epoch
    params = O.make_params(O.vae_param_shapes(P, L), seed=seed)
    x, _, _ = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    model = VAE(latent_dim=L, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    # train_one_epoch (train.py:33-165) clips at 5.0 unconditionally, so take the raw
    # gradients from the same forward/criterion calls it makes (train.py:76-84,104)
    with FixedEps(eps):
        recon, mu, logvar = model(x)
    loss, rl, kl = VAELoss(beta=1.0)(recon, x, mu, logvar)
    loss.backward()
    res = pack_grads({k: p.grad for k, p in model.named_parameters()})
    res.update(loss=np.array(loss.item()), recon_loss=np.array(rl.item()), kld=np.array(kl.item()),
               mu=mu.detach().numpy(), logvar=logvar.detach().numpy(),
               recon=recon.detach().numpy())
    # and the trainer's own metrics for the same batch (clip 5.0 applied inside)
    model.zero_grad()
    logger = MetricLogger()
    with FixedEps(eps):
        train_one_epoch(model, [x], torch.optim.SGD(model.parameters(), lr=0.0), VAELoss(beta=1.0),
                        logger, torch.device("cpu"))
    res.update({"metric/" + k: np.array(v[0]) for k, v in logger.metrics.items()})
    np.savez_compressed(os.path.join(OUT, "vae_step_p64.npz"), torch_version=torch.__version__,
                        P=P, L=L, B=B, seed=seed, **res)
    print("vae", float(loss))


def golden_stn_pretrain(livae, P, B, seed):
    from livae.model import RVAE
    from livae.loss import cycle_consistency_loss
    params = O.make_params(O.rvae_param_shapes(P, 2), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    model = RVAE(latent_dim=2, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    # scripts/pretrain_stn.py:104-112
    _, _, th0 = model.encoder(x)
    _, _, th1 = model.encoder(xr)
    loss = cycle_consistency_loss(th0, th1, ang)
    loss.backward()
    stn = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    res = pack_grads(stn)
    res.update(loss=np.array(loss.item()), theta=th0.detach().numpy(), theta_rot=th1.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "stn_pretrain_p32.npz"), torch_version=torch.__version__,
                        P=P, B=B, seed=seed, **res)


def synth_image(HW, seed):
    """Bit-reproducible float64 test image (no transcendental functions)."""
    rng = np.random.default_rng(seed)
    return rng.random((HW, HW))


def golden_patch_gather(livae):
    from livae.data import PatchDataset, AdaptiveLatticeDataset
    res = {}
    for HW, P, pad, n in ((2048, 128, 32, 6), (1024, 64, 8, 6)):
        img = synth_image(HW, 100 + HW)
        rng = np.random.default_rng(5 + HW)
        lo = P // 2 + pad
        sites = rng.integers(lo, HW - lo + 1, size=(n, 2))
        ds = PatchDataset.__new__(PatchDataset)      # skip __init__ (site finding, skimage)
        ds.patch_size = P; ds.padding = pad; ds.transform = None
        ds.images = [img]; ds.atom_coords = [sites]
        got = np.stack([ds[i].numpy() for i in range(n)])
        res[f"sites{HW}"] = sites
        res[f"sha{HW}"] = np.array(hashlib.sha256(got.tobytes()).hexdigest())
        want = np.stack([img[cy - P // 2:cy + P // 2, cx - P // 2:cx + P // 2].astype(np.float32)[None]
                         for cy, cx in sites])
        assert np.array_equal(got, want), "reference crop is not bit-equal to the integer slice"
    # a2: sub-pixel adaptive gather, transform=None
    HW, P, pad, n = 512, 64, 8, 8
    img = synth_image(HW, 300)
    rng = np.random.default_rng(301)
    sites = rng.uniform(P, HW - P, size=(n, 2))
    sites[0] = (20.3, 30.7)            # border: exercises the zero padding of the ROI
    sites[1] = (HW - 10.2, HW - 5.5)
    ds = AdaptiveLatticeDataset.__new__(AdaptiveLatticeDataset)
    ds.patch_size = P; ds.padding = pad; ds.transform = None
    ds.images = [img]; ds.sample_coords = [sites]; ds.labels = [np.ones(n)]
    res["sub_sites"] = sites
    res["sub_out"] = np.stack([ds[i].numpy() for i in range(n)])
    np.savez_compressed(os.path.join(OUT, "patch_gather.npz"), torch_version=torch.__version__, **res)


def golden_augment(livae):
    """a3 + the paired rotation: the reference's default_transform and Paired/AdaptiveLatticeDataset items
    with Python `random` seeded; only the seeds, sites and OUTPUTS are stored (oracle/augment.draw_params
    replays the draws)."""
    import random
    from livae.data import AdaptiveLatticeDataset, PairedAdaptiveLatticeDataset, default_transform
    res = {}
    # default_transform on a bare patch, with and without rotation
    patch = torch.from_numpy(synth_image(48, 400).astype(np.float32))[None]
    for k, rot in enumerate((False, True, False, True)):
        random.seed(900 + k)
        res[f"dt{k}"] = default_transform(patch, rotation=rot).numpy()
    # dataset items: (P, pad) = (64, 8) reaches the ROI border under rotation, (32, 16) does not
    for tag, P, pad, n in (("a", 64, 8, 5), ("b", 32, 16, 5)):
        HW = 256
        img = synth_image(HW, 410 + P)
        rng = np.random.default_rng(420 + P)
        sites = rng.uniform(P, HW - P, size=(n, 2))
        sites[0] = (14.5, 17.5)                       # image border + round-half-even sites
        sites[1] = (HW - 9.25, HW - 30.5)
        res[f"sites_{tag}"] = sites
        ds = PairedAdaptiveLatticeDataset.__new__(PairedAdaptiveLatticeDataset)
        ds.patch_size = P; ds.padding = pad; ds.transform = default_transform
        ds.images = [img]; ds.sample_coords = [sites]; ds.labels = [np.ones(n)]
        random.seed(1000 + P)
        items = [ds[i] for i in range(n)]
        res[f"pair_{tag}_x"] = np.stack([it[0].numpy() for it in items])
        res[f"pair_{tag}_r"] = np.stack([it[1].numpy() for it in items])
        res[f"pair_{tag}_angle"] = np.array([it[2] for it in items])
        ds.transform = None
        random.seed(2000 + P)
        items = [ds[i] for i in range(n)]
        res[f"pairnt_{tag}_x"] = np.stack([it[0].numpy() for it in items])
        res[f"pairnt_{tag}_r"] = np.stack([it[1].numpy() for it in items])
        res[f"pairnt_{tag}_angle"] = np.array([it[2] for it in items])
        ad = AdaptiveLatticeDataset.__new__(AdaptiveLatticeDataset)
        ad.patch_size = P; ad.padding = pad; ad.transform = default_transform
        ad.images = [img]; ad.sample_coords = [sites]; ad.labels = [np.ones(n)]
        random.seed(3000 + P)
        res[f"adapt_{tag}"] = np.stack([ad[i].numpy() for i in range(n)])
    np.savez_compressed(os.path.join(OUT, "augment.npz"), torch_version=torch.__version__,
                        torchvision_version=__import__("torchvision").__version__, **res)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    livae = ref_loader.load()
    golden_rot_sample(livae)
    golden_rvae_step(livae, P=32, L=2, B=4, seed=1234, tag="p32", store_images=True)
    golden_rvae_step(livae, P=128, L=2, B=4, seed=4321, tag="p128", store_images=False)
    golden_vae_step(livae, P=64, L=16, B=4, seed=2468)
    golden_stn_pretrain(livae, P=32, B=4, seed=1357)
    golden_patch_gather(livae)
    golden_augment(livae)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
