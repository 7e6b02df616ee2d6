"""Round-2 golden vectors, produced by the UNMODIFIED reference (/root/reference via oracle/ref_loader.py) on CPU.
Run in the build container only:   python tests/golden/make_golden_r2.py     (make_golden.py's files stay untouched)

  eval.npz    -- evaluate_rvae / evaluate outputs of the reference's own loops (train.py:168-278, 448-556) on two
                 seeded batches (the rVAE loop's last-batch-only quirk therefore shows), and train_one_epoch's
                 metrics on an rVAE (train.py:85-94 branch).
  preproc.npz -- filter.py (band/low/high-pass, normalize, spectra) and utils.estimate_lattice_constant outputs.
  sites.npz   -- AdaptiveLatticeDataset / PatchDataset CONSTRUCTORS of the reference (data.py:176-202, 299-473) with
                 skimage's absent `peak_local_max` replaced by livae/sites.py's restatement: pins the vectorised
                 lattice-site extrapolation + clustering against the reference's loops on identical peaks.
  patchds.npz -- PatchDataset items with default_transform (rotation=True path, data.py:240-248), `random` seeded.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import rvae as O   # noqa: E402
from tests.golden.make_golden import FixedEps, synth_image  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synth_lattice(hw, a, ang_deg, seed):
    """uint16-like lattice micrograph: three cosine waves 60 degrees apart (row spacing a*sqrt(3)/2) + noise"""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(hw), np.arange(hw), indexing="ij")
    img = np.zeros((hw, hw))
    t0 = np.deg2rad(ang_deg)
    for k in range(3):
        t = t0 + k * np.pi / 3
        img += np.cos(2 * np.pi / (a * np.sqrt(3) / 2) * (np.cos(t) * xx + np.sin(t) * yy))
    img += rng.normal(0, 0.3, img.shape)
    return img * 1000 + 5000


def golden_eval(livae):
    from livae.loss import RVAELoss, VAELoss
    from livae.model import RVAE, VAE
    from livae.train import MetricLogger, evaluate, evaluate_rvae, train_one_epoch
    res = {}
    P, L, B, seed = 32, 2, 4, 777
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    batches = []
    for k in range(2):
        x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1 + k)
        batches.append((x, xr, ang))
    eps = torch.from_numpy(np.random.default_rng(seed + 9).standard_normal((B, L))).float()
    model = RVAE(latent_dim=L, in_channels=1, patch_size=P)
    model.load_state_dict(params, strict=True)
    lg = MetricLogger()
    with FixedEps(eps):
        evaluate_rvae(model, batches, RVAELoss(beta=10.0, gamma=10.0), lg, torch.device("cpu"), canonical_weight=0.2)
    res.update({"rvae/" + k: np.array(v[0]) for k, v in lg.metrics.items()})
    lg = MetricLogger()
    with FixedEps(eps):
        evaluate(model, [b[0] for b in batches], VAELoss(beta=1.0), lg, torch.device("cpu"), canonical_weight=0.2)
    res.update({"rvae_evaluate/" + k: np.array(v[0]) for k, v in lg.metrics.items()})
    lg = MetricLogger()
    with FixedEps(eps):
        train_one_epoch(model, [b[0] for b in batches], torch.optim.SGD(model.parameters(), lr=0.0),
                        VAELoss(beta=1.0), lg, torch.device("cpu"))
    res.update({"rvae_train_one_epoch/" + k: np.array(v[0]) for k, v in lg.metrics.items()})
    # plain VAE
    Pv, Lv, seedv = 64, 16, 888
    vparams = O.make_params(O.vae_param_shapes(Pv, Lv), seed=seedv)
    vb = [O.make_lattice_batch(B, Pv, seed=seedv + 1 + k)[0] for k in range(2)]
    veps = torch.from_numpy(np.random.default_rng(seedv + 9).standard_normal((B, Lv))).float()
    vae = VAE(latent_dim=Lv, in_channels=1, patch_size=Pv)
    vae.load_state_dict(vparams, strict=True)
    lg = MetricLogger()
    with FixedEps(veps):
        evaluate(vae, vb, VAELoss(beta=1.0), lg, torch.device("cpu"))
    res.update({"vae/" + k: np.array(v[0]) for k, v in lg.metrics.items()})
    np.savez_compressed(os.path.join(OUT, "eval.npz"), torch_version=torch.__version__, P=P, L=L, B=B, seed=seed,
                        Pv=Pv, Lv=Lv, seedv=seedv, **res)
    print("eval", {k: float(v) for k, v in res.items()})


def golden_preproc(livae):
    from livae import filter as F
    from livae.utils import estimate_lattice_constant
    res = {}
    img = synth_lattice(256, 14.0, 11.0, 5)
    rect = np.random.default_rng(6).random((96, 128)) * 1000.0
    samp = lambda a: a[::7, ::5].copy()
    res["band"] = samp(F.bandpass_filter(img, 10, 60)); res["low"] = samp(F.lowpass_filter(rect, 20))
    res["high"] = samp(F.highpass_filter(rect, 7)); res["norm"] = samp(F.normalize_image(rect))
    mag, ph = F.fft_spectra(rect)
    res["mag"] = samp(mag); res["phase"] = samp(ph)
    res["band_sum"] = np.array(F.bandpass_filter(img, 10, 60).sum())
    for k, (hw, a) in enumerate(((256, 14.0), (512, 16.0), (512, 19.0), (384, 24.0))):
        res[f"lattice_{k}"] = np.array(estimate_lattice_constant(synth_lattice(hw, a, 7.0 + 13 * k, 20 + k)))
    res["lattice_noise"] = np.array(estimate_lattice_constant(np.random.default_rng(1).random((128, 128))))
    np.savez_compressed(os.path.join(OUT, "preproc.npz"), **res)
    print("preproc", {k: float(v) for k, v in res.items() if v.ndim == 0})


def golden_sites(livae):
    import livae.data as RD
    spec = importlib.util.spec_from_file_location("livae_b200_sites", os.path.join(ROOT, "li-vae_b200", "livae", "sites.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    RD.peak_local_max = lambda img, min_distance=1, threshold_rel=None: S.peak_local_max(img, min_distance, threshold_rel)
    imgs = [synth_lattice(512, 16.0, 7.0, 1), synth_lattice(512, 19.0, 33.0, 2)]
    res = {}
    ad = RD.AdaptiveLatticeDataset(imgs, patch_size=64, padding=16, transform=None)
    pd = RD.PatchDataset(imgs, patch_size=64, padding=8, transform=None)
    for i in range(2):
        res[f"sample_coords_{i}"] = ad.sample_coords[i]; res[f"labels_{i}"] = ad.labels[i]
        res[f"atom_coords_{i}"] = pd.atom_coords[i]
        res[f"image_sum_{i}"] = np.array(ad.images[i].sum())
    res["grid"] = RD.generate_lattice_grid((100, 120), 11.3, (2.5, 1.0))
    np.savez_compressed(os.path.join(OUT, "sites.npz"), **res)
    print("sites", [len(c) for c in ad.sample_coords], [len(c) for c in pd.atom_coords])


def golden_patchds(livae):
    import random
    from livae.data import PatchDataset, default_transform
    res = {}
    HW, P, pad, n = 256, 64, 16, 5
    img = synth_image(HW, 5150)
    rng = np.random.default_rng(5151)
    lo = P // 2 + pad
    sites = rng.integers(lo, HW - lo + 1, size=(n, 2))
    ds = PatchDataset.__new__(PatchDataset)
    ds.patch_size = P; ds.padding = pad; ds.transform = default_transform
    ds.images = [img]; ds.atom_coords = [sites]
    random.seed(6000)
    res["items"] = np.stack([ds[i].numpy() for i in range(n)])
    res["sites"] = sites
    np.savez_compressed(os.path.join(OUT, "patchds.npz"), torch_version=torch.__version__, **res)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    livae = ref_loader.load()
    golden_eval(livae)
    golden_preproc(livae)
    golden_patchds(livae)
    golden_sites(livae)
    for f in ("eval.npz", "preproc.npz", "sites.npz", "patchds.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
