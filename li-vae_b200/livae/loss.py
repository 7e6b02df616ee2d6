"""livae.loss -- drop-in for the reference's src/livae/loss.py; reductions run in the fused ELBO
kernel (csrc/elbo.cu).  Same signatures and return tuples as the reference (loss.py:97-186)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

__all__ = ["VAELoss", "RVAELoss", "cycle_consistency_loss", "rotation_diversity_loss", "circular_distance"]


def circular_distance(theta1: torch.Tensor, theta2: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """mean shortest angular distance between two sets of angles, in [0, pi] (reference loss.py:6-29; analysis
    helper, not on the training path).  The difference is wrapped into [-pi, pi) once instead of taking
    min(|d|, 2*pi - |d|); both agree for |d| <= 2*pi, which is all atan2 outputs can produce."""
    d = theta1.reshape(theta1.shape[0], -1) - theta2.reshape(theta2.shape[0], -1)
    return (torch.remainder(d + torch.pi, 2 * torch.pi) - torch.pi).abs().mean()


def rotation_diversity_loss(theta: torch.Tensor, target_std: float = 1.0) -> torch.Tensor:
    """reference loss.py:32-49: (std(theta) - target)^2 with the unbiased batch std.
    A [B]-element reduction used only with --use-diversity-loss; kept on torch scalar ops."""
    return (torch.std(theta) - target_std) ** 2


def cycle_consistency_loss(theta_original, theta_rotated, expected_angle):
    """reference loss.py:52-94: mean(1 - cos(theta_rot - theta + angle))"""
    if expected_angle.dim() == 0:
        expected_angle = expected_angle.unsqueeze(0)
    expected_angle = expected_angle.to(torch.float32)
    if expected_angle.numel() == 1 and theta_original.numel() > 1:
        expected_angle = expected_angle.expand(theta_original.numel()).contiguous()
    return ops.cycle_loss(theta_original.reshape(-1), theta_rotated.reshape(-1), expected_angle.reshape(-1))


class VAELoss(nn.Module):
    """reference loss.py:97-122: mean MSE + beta * mean KLD"""

    def __init__(self, beta: float = 1.0) -> None:
        super().__init__()
        self.beta = beta

    def forward(self, recon_x, x, mu, logvar):
        sums = ops.elbo_sums(recon_x, x, mu, logvar)
        recon_loss = sums[0] / x.numel()
        kld_loss = sums[1] / mu.numel()
        total_loss = recon_loss + self.beta * kld_loss
        return total_loss, recon_loss, kld_loss


class RVAELoss(nn.Module):
    """reference loss.py:125-186: sum-MSE / B + beta * mean_b KLD + gamma * rotation loss"""

    def __init__(self, beta: float = 1.0, gamma: float = 0.0, use_diversity: bool = False) -> None:
        super().__init__()
        self.beta = beta
        self.gamma = gamma
        self.use_diversity = use_diversity

    def forward(self, recon_x, x, mu, logvar, theta=None, theta_rotated=None, expected_angle=None):
        n = x.size(0)
        sums = ops.elbo_sums(recon_x, x, mu, logvar)             # one fused launch: sum (r-x)^2 and the KLD sum
        recon_loss, kld_loss = sums[0] / n, sums[1] / n
        rotation_loss = self._rotation_term(theta, theta_rotated, expected_angle, recon_x.device)
        return recon_loss + self.beta * kld_loss + self.gamma * rotation_loss, recon_loss, kld_loss, rotation_loss

    def _rotation_term(self, theta, theta_rotated, expected_angle, device):
        """loss.py:171-182: nothing when gamma == 0; the batch-spread term with use_diversity; otherwise the cycle
        term, which needs the rotated partner's angle and the applied rotation"""
        if self.gamma > 0 and theta is not None:
            if self.use_diversity:
                return rotation_diversity_loss(theta, target_std=1.0)
            if theta_rotated is not None and expected_angle is not None:
                return cycle_consistency_loss(theta, theta_rotated, expected_angle)
        return torch.tensor(0.0, device=device)
