"""livae on B200: drop-in for jerrydzhang/LI-VAE's `livae` package (model, loss, train, data, utils, filter,
metrics -- every name the reference's scripts and package root import), with the rVAE / VAE training step
backed by hand-written sm_100a CUDA kernels behind a C ABI (include/livae_b200.h).  No CPU fallback on the
training path: the ops raise if the CUDA library or device is missing.
"""
from livae.data import PatchDataset, default_transform
from livae.filter import bandpass_filter, fft_spectra, highpass_filter, lowpass_filter, normalize_image
from livae.loss import RVAELoss, VAELoss, cycle_consistency_loss
from livae.metrics import (
    compute_all_metrics,
    compute_atom_detection_metrics,
    compute_latent_metrics,
    compute_reconstruction_metrics,
)
from livae.model import RVAE, VAE, Decoder, Encoder, RotationSTN, VAEDecoder, VAEEncoder, get_engine, set_engine
from livae.train import (
    MetricLogger,
    compute_psnr,
    compute_ssim,
    evaluate,
    evaluate_rotation_invariance,
    evaluate_rvae,
    log_reconstructions_tensorboard,
    log_scalar_metrics_tensorboard,
    rotate_to_canonical,
    train_one_epoch,
    train_rvae_one_epoch,
)
from livae.utils import estimate_lattice_constant, load_image_from_h5

__version__ = "0.1.0"

__all__ = [
    "PatchDataset", "default_transform",
    "normalize_image", "bandpass_filter", "fft_spectra", "lowpass_filter", "highpass_filter",
    "VAELoss", "RVAELoss", "cycle_consistency_loss",
    "VAE", "RVAE", "Encoder", "Decoder", "RotationSTN", "VAEEncoder", "VAEDecoder",
    "train_one_epoch", "train_rvae_one_epoch", "evaluate", "evaluate_rvae", "evaluate_rotation_invariance",
    "log_reconstructions_tensorboard", "log_scalar_metrics_tensorboard", "rotate_to_canonical", "MetricLogger",
    "compute_psnr", "compute_ssim", "compute_reconstruction_metrics", "compute_latent_metrics",
    "compute_atom_detection_metrics", "compute_all_metrics",
    "load_image_from_h5", "estimate_lattice_constant", "set_engine", "get_engine",
]
