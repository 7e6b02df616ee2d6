"""livae.metrics -- evaluation metrics the reference exports from its package root (src/livae/metrics.py).

PSNR and box-filter SSIM are the device implementations of livae.train (fused MSE reduction, csrc/ssim.cu); the
dictionaries below are post-hoc analysis helpers with the reference's keys.
"""
from __future__ import annotations

import numpy as np
import torch

from .train import compute_atom_position_accuracy, compute_psnr, compute_ssim

__all__ = ["compute_psnr", "compute_ssim", "compute_reconstruction_metrics", "compute_latent_metrics",
           "compute_atom_detection_metrics", "compute_all_metrics"]


def compute_reconstruction_metrics(original: torch.Tensor, reconstruction: torch.Tensor) -> dict:
    """mse / rmse / mae / psnr / ssim of a batch (metrics.py:116-150)"""
    diff = original - reconstruction
    mse = torch.mean(diff ** 2).item()
    return {"mse": mse, "rmse": float(np.sqrt(mse)), "mae": torch.mean(diff.abs()).item(),
            "psnr": compute_psnr(original, reconstruction), "ssim": compute_ssim(original, reconstruction)}


def compute_latent_metrics(mu: torch.Tensor, logvar: torch.Tensor) -> dict:
    """spread of the posterior means and standard deviations, KL per dimension (metrics.py:153-194)"""
    std = torch.exp(0.5 * logvar)
    return {"latent_mean_abs": mu.abs().mean().item(), "latent_mean_std": torch.std(mu).item(),
            "latent_std_mean": std.mean().item(), "latent_std_std": torch.std(std).item(),
            "latent_kl_per_dim": (-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())).item()}


def compute_atom_detection_metrics(original, reconstruction, lattice_spacing: float, threshold_ratio: float = 0.35):
    """metrics.py:197-285 -- the same comparison as livae.train.compute_atom_position_accuracy"""
    return compute_atom_position_accuracy(original, reconstruction, lattice_spacing, threshold_ratio)


def compute_all_metrics(model, images: torch.Tensor, device, lattice_spacing: float | None = None) -> dict:
    """reconstruction + latent (+ atom) metrics of one batch, VAE or rVAE (metrics.py:288-348)"""
    model.eval()
    out = {}
    with torch.no_grad():
        images = images.to(device)
        res = model(images)
        if len(res) == 3:
            recon, mu, logvar = res
        elif len(res) == 5:
            recon, _, _, mu, logvar = res
        else:
            raise ValueError(f"Unexpected model output length: {len(res)}")
        out.update(compute_reconstruction_metrics(images, recon))
        out.update(compute_latent_metrics(mu, logvar))
        if lattice_spacing is not None:
            out.update(compute_atom_detection_metrics(images[0], recon[0], lattice_spacing))
    return out
