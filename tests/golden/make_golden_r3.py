"""Branch coverage of the oracle's step bodies: run the UNMODIFIED reference trainer (/root/reference through
oracle/ref_loader.py, CPU) on the loss / batch-format branches the main vectors (make_golden.py: beta = gamma = 10,
paired batch, canonical weight 0.2) do not reach, and store what tests/test_oracle_golden.py compares the oracle with:

  rVAE   gamma = 0 (no rotation term, loss.py:171-182)            P = 64, latent 10
         use_diversity=True (loss.py:173-175, 32-49)              P = 32, latent 4
         canonical_weight = 0 (train.py:386)                      P = 32, latent 2
         unpaired batch: the loader yields x alone (train.py:335-338), gamma > 0 -> rotation loss 0, P = 64, latent 16
         2-tuple batch (x, x_rotated) without angle (train.py:325-330)   P = 32, latent 3
  VAE    P = 32 latent 8 beta 4;  P = 128 latent 16 beta 0.5      (train.py:33-165, loss.py:104-122)

Build container only:   python tests/golden/make_golden_r3.py   ->  tests/golden/branches.npz
Inputs come from the deterministic generators in oracle/rvae.py (only outputs are stored)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import rvae as O   # noqa: E402
from tests.golden.make_golden import FixedEps, pack_grads  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# tag -> (P, L, B, seed, beta, gamma, use_diversity, canonical_weight, batch format)
RVAE_CASES = {
    "nogamma": (64, 10, 3, 11, 1.0, 0.0, False, 0.2, "paired"),
    "diversity": (32, 4, 4, 12, 2.0, 3.0, True, 0.2, "paired"),
    "nocanon": (32, 2, 4, 13, 10.0, 10.0, False, 0.0, "paired"),
    "unpaired": (64, 16, 2, 14, 10.0, 10.0, False, 0.5, "single"),
    "noangle": (32, 3, 4, 15, 5.0, 10.0, False, 0.2, "pair2"),
}
VAE_CASES = {"vae_p32": (32, 8, 4, 21, 4.0), "vae_p128": (128, 16, 2, 22, 0.5)}


def rvae_inputs(P, L, B, seed):
    params = O.make_params(O.rvae_param_shapes(P, L), seed=seed, stn_head_std=0.5)
    x, xr, ang = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    return params, x, xr, ang, eps


def batch_of(fmt, x, xr, ang):
    return {"paired": (x, xr, ang), "single": x, "pair2": (x, xr)}[fmt]


def vae_inputs(P, L, B, seed):
    params = O.make_params(O.vae_param_shapes(P, L), seed=seed)
    x, _, _ = O.make_lattice_batch(B, P, seed=seed + 1)
    eps = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((B, L))).float()
    return params, x, eps


def metric_inputs():
    rng = np.random.default_rng(31)
    mu = torch.from_numpy(rng.standard_normal((6, 5))).float()
    logvar = torch.from_numpy(rng.standard_normal((6, 5)) * 0.3 - 1.0).float()
    a = torch.from_numpy(rng.random((3, 1, 32, 32))).float()
    b = (a + torch.from_numpy(rng.standard_normal((3, 1, 32, 32))).float() * 0.05).clamp(0, 1)
    return mu, logvar, a, b


def main():
    ref_loader.load()
    from livae.loss import RVAELoss, VAELoss
    from livae.model import RVAE, VAE
    from livae.train import MetricLogger, train_rvae_one_epoch
    res = {}
    for tag, (P, L, B, seed, beta, gamma, div, cw, fmt) in RVAE_CASES.items():
        params, x, xr, ang, eps = rvae_inputs(P, L, B, seed)
        model = RVAE(latent_dim=L, in_channels=1, patch_size=P)
        model.load_state_dict(params, strict=True)
        logger = MetricLogger()
        with FixedEps(eps):
            # lr = 0 and an infinite clip norm leave p.grad exactly as loss.backward() produced it
            train_rvae_one_epoch(model, [batch_of(fmt, x, xr, ang)], torch.optim.SGD(model.parameters(), lr=0.0),
                                 RVAELoss(beta=beta, gamma=gamma, use_diversity=div), logger, torch.device("cpu"),
                                 canonical_weight=cw, scaler=None, grad_max_norm=1e30)
        for k, v in logger.metrics.items():
            res[f"{tag}/metric/{k}"] = np.array(v[0])
        grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
        for k, v in pack_grads(grads).items():
            res[f"{tag}/{k}"] = v
        print(tag, {k: float(v[0]) for k, v in logger.metrics.items() if "loss" in k})
    for tag, (P, L, B, seed, beta) in VAE_CASES.items():
        params, x, eps = vae_inputs(P, L, B, seed)
        model = VAE(latent_dim=L, in_channels=1, patch_size=P)
        model.load_state_dict(params, strict=True)
        with FixedEps(eps):
            recon, mu, logvar = model(x)                      # train.py:76-84, 104 (its clip at 5.0 is not wanted here)
        loss, rl, kl = VAELoss(beta=beta)(recon, x, mu, logvar)
        loss.backward()
        for k, v in pack_grads({k: p.grad for k, p in model.named_parameters()}).items():
            res[f"{tag}/{k}"] = v
        res[f"{tag}/loss"] = np.array(loss.item()); res[f"{tag}/recon_loss"] = np.array(rl.item())
        res[f"{tag}/kld"] = np.array(kl.item())
        res[f"{tag}/recon_sum"] = np.array(recon.double().sum().item())
        res[f"{tag}/mu"] = mu.detach().numpy()
        print(tag, float(loss))
    # livae.metrics.compute_latent_metrics / the pixel part of compute_reconstruction_metrics (metrics.py:116-194)
    from livae.metrics import compute_latent_metrics, compute_reconstruction_metrics
    mu, logvar, a, b = metric_inputs()
    for k, v in compute_latent_metrics(mu, logvar).items():
        res["metrics/" + k] = np.array(v)
    for k, v in compute_reconstruction_metrics(a, b).items():
        res["metrics/" + k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, "branches.npz"), torch_version=torch.__version__, **res)


if __name__ == "__main__":
    main()
