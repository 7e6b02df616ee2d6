"""livae.loss -- drop-in for the reference's src/livae/loss.py; reductions run in the fused ELBO
kernel (csrc/elbo.cu).  Same signatures and return tuples as the reference (loss.py:97-186)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

__all__ = ["VAELoss", "RVAELoss", "cycle_consistency_loss", "rotation_diversity_loss", "circular_distance"]


def circular_distance(theta1: torch.Tensor, theta2: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """reference loss.py:6-29 (analysis helper, not on the training path)"""
    if theta1.dim() == 1:
        theta1 = theta1.unsqueeze(1)
    if theta2.dim() == 1:
        theta2 = theta2.unsqueeze(1)
    diff = torch.abs(theta1 - theta2)
    diff = torch.min(diff, 2 * torch.pi - diff)
    return torch.mean(diff)


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
        batch_size = x.size(0)
        sums = ops.elbo_sums(recon_x, x, mu, logvar)
        recon_loss = sums[0] / batch_size
        kld_loss = sums[1] / batch_size
        if self.gamma > 0:
            if self.use_diversity and theta is not None:
                rotation_loss = rotation_diversity_loss(theta, target_std=1.0)
            elif theta is not None and theta_rotated is not None and expected_angle is not None:
                rotation_loss = cycle_consistency_loss(theta, theta_rotated, expected_angle)
            else:
                rotation_loss = torch.tensor(0.0, device=recon_x.device)
        else:
            rotation_loss = torch.tensor(0.0, device=recon_x.device)
        total_loss = recon_loss + self.beta * kld_loss + self.gamma * rotation_loss
        return total_loss, recon_loss, kld_loss, rotation_loss
