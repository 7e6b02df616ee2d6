"""livae on B200: drop-in for the training-step surface of jerrydzhang/LI-VAE's `livae` package
(model, loss, train, data), backed by hand-written sm_100a CUDA kernels behind a C ABI
(include/livae_b200.h).  No CPU fallback: the ops raise if the CUDA library or device is missing.
"""
from livae.loss import RVAELoss, VAELoss, cycle_consistency_loss
from livae.model import RVAE, VAE, Decoder, Encoder, RotationSTN, VAEDecoder, VAEEncoder, get_engine, set_engine
from livae.train import (
    MetricLogger,
    compute_psnr,
    compute_ssim,
    evaluate,
    evaluate_rvae,
    rotate_to_canonical,
    train_one_epoch,
    train_rvae_one_epoch,
)

__version__ = "0.1.0"

__all__ = [
    "VAELoss", "RVAELoss", "cycle_consistency_loss",
    "VAE", "RVAE", "Encoder", "Decoder", "RotationSTN", "VAEEncoder", "VAEDecoder",
    "train_one_epoch", "train_rvae_one_epoch", "evaluate", "evaluate_rvae", "rotate_to_canonical",
    "MetricLogger", "compute_psnr", "compute_ssim", "set_engine", "get_engine",
]
