"""livae.model -- drop-in for the reference's src/livae/model.py on B200.

Same class names, constructor signatures, forward return tuples and state_dict keys as the
reference, so scripts/train_rvae.py, train_vae.py and pretrain_stn.py run
unchanged and checkpoints are interchangeable.  The nn.Conv2d / nn.Linear / nn.ConvTranspose2d
children exist only as PARAMETER CONTAINERS (torch's default initialisation, torch's weight
layouts); every forward/backward FLOP runs in the hand-written sm_100a kernels behind
livae.ops.  There is no CPU or ATen fallback: tensors must be CUDA float32.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops, tc
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID

# Convolution engine: "tc" = tcgen05/TMA tensor cores with bf16 storage between layers (default; north_star's "bf16 GEMM
# inputs": ELBO 1e-3, reconstructions 1e-2, gradients max(1e-2, operand-rounding floor) -- livae/tc.py),
# "f32" = exact fp32 SIMT engine (1e-4 class).
_ENGINE = "tc"


def set_engine(name: str) -> None:
    global _ENGINE
    if name not in ("tc", "f32"):
        raise ValueError(f"unknown engine {name!r} (expected 'tc' or 'f32')")
    _ENGINE = name


def get_engine() -> str:
    return _ENGINE

__all__ = ["VAEEncoder", "VAEDecoder", "VAE", "RotationSTN", "Encoder", "Decoder", "RVAE"]


def _to_nhwc(x):
    """[B,C,H,W] -> [B,H,W,C]; free for C == 1"""
    B, Cc, H, W = x.shape
    if Cc == 1:
        return x.reshape(B, H, W, 1)
    return x.permute(0, 2, 3, 1).contiguous()


def _to_nchw(h):
    B, H, W, Cc = h.shape
    if Cc == 1:
        return h.reshape(B, 1, H, W)
    return h.permute(0, 3, 1, 2).contiguous()


def _encoder_stack():
    # reference model.py:29-38 / 289-298
    def make(in_channels):
        return nn.Sequential(
            nn.Conv2d(in_channels, 32, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.Conv2d(32, 64, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.Conv2d(64, 128, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.Conv2d(128, 256, kernel_size=4, stride=2, padding=1), nn.ReLU(),
        )
    return make


def _run_encoder_convs(conv_layers, x_nchw):
    h = _to_nhwc(x_nchw)
    for i in (0, 2, 4, 6):
        m = conv_layers[i]
        h = ops.conv2d(h, m.weight, m.bias, 4, 4, 2, 1, ACT_RELU)
    return h


class VAEEncoder(nn.Module):
    """reference model.py:9-61"""

    def __init__(self, in_channels: int = 1, latent_dim: int = 10, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.patch_size = patch_size
        self.conv_layers = _encoder_stack()(in_channels)
        flat_size = 256 * (patch_size // 16) * (patch_size // 16)
        self.fc_mu = nn.Linear(flat_size, latent_dim)
        self.fc_logvar = nn.Linear(flat_size, latent_dim)

    def forward(self, x):
        if _ENGINE == "tc" and tc.supported(self.patch_size, self.latent_dim, x.shape[1]):
            cl = self.conv_layers
            return tc.VAEEncoderTc.apply(x, cl[0].weight, cl[0].bias, cl[2].weight, cl[2].bias, cl[4].weight, cl[4].bias,
                                         cl[6].weight, cl[6].bias, self.fc_mu.weight, self.fc_mu.bias,
                                         self.fc_logvar.weight, self.fc_logvar.bias)
        h = _run_encoder_convs(self.conv_layers, x)
        mu = ops.linear_nhwc(h, self.fc_mu.weight, self.fc_mu.bias)
        logvar = ops.linear_nhwc(h, self.fc_logvar.weight, self.fc_logvar.bias)
        return mu, logvar


class VAEDecoder(nn.Module):
    """reference model.py:64-113: relu(fc) -> 4x ConvTranspose2d(k4,s2,p1) + ReLU/Sigmoid"""

    def __init__(self, latent_dim: int = 10, out_channels: int = 1, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.out_channels = out_channels
        self.patch_size = patch_size
        inter_size = 256 * (patch_size // 16) * (patch_size // 16)
        self.fc = nn.Linear(latent_dim, inter_size)
        self.deconv_layers = nn.Sequential(
            nn.ConvTranspose2d(256, 128, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.ConvTranspose2d(128, 64, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.ConvTranspose2d(64, 32, kernel_size=4, stride=2, padding=1), nn.ReLU(),
            nn.ConvTranspose2d(32, out_channels, kernel_size=4, stride=2, padding=1), nn.Sigmoid(),
        )

    def forward(self, z):
        q = self.patch_size // 16
        if _ENGINE == "tc" and tc.supported(self.patch_size, self.latent_dim, self.out_channels):
            dl = self.deconv_layers
            return tc.VAEDecoderTc.apply(z, self.fc.weight, self.fc.bias, dl[0].weight, dl[0].bias, dl[2].weight,
                                         dl[2].bias, dl[4].weight, dl[4].bias, dl[6].weight, dl[6].bias)
        h = ops.decoder_fc(z, self.fc.weight, self.fc.bias, 256, q)
        for n, i in enumerate((0, 2, 4, 6)):
            m = self.deconv_layers[i]
            h = ops.conv_transpose2d(h, m.weight, m.bias, 4, 4, 2, 1, ACT_SIGMOID if n == 3 else ACT_RELU)
        return _to_nchw(h)


class VAE(nn.Module):
    """reference model.py:116-182"""

    def __init__(self, latent_dim: int = 10, in_channels: int = 1, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.in_channels = in_channels
        self.patch_size = patch_size
        self.encoder = VAEEncoder(in_channels, latent_dim, patch_size)
        self.decoder = VAEDecoder(latent_dim, in_channels, patch_size)

    def reparameterize(self, mu, logvar):
        # eps comes from torch's generator exactly where the reference draws it (model.py:158)
        eps = torch.randn_like(logvar)
        return ops.reparam(mu, logvar, eps)

    def forward(self, x):
        mu, logvar = self.encoder(x)
        z = self.reparameterize(mu, logvar)
        recon = self.decoder(z)
        return recon, mu, logvar


class RotationSTN(nn.Module):
    """reference model.py:185-262"""

    def __init__(self, input_shape=(1, 64, 64)):
        super().__init__()
        self.c, self.h, self.w = input_shape
        self.localization = nn.Sequential(
            nn.Conv2d(self.c, 16, kernel_size=5, stride=1, padding=2),
            nn.ReLU(True),
            nn.MaxPool2d(2, stride=2),
            nn.Conv2d(16, 32, kernel_size=5, stride=1, padding=2),
            nn.ReLU(True),
            nn.MaxPool2d(2, stride=2),
            nn.Flatten(),
            nn.Linear(32 * (self.h // 4) * (self.w // 4), 32),
            nn.ReLU(True),
            nn.Linear(32, 2),
        )
        nn.init.normal_(self.localization[-1].weight, mean=0.0, std=0.01)
        nn.init.zeros_(self.localization[-1].bias)

    def get_rotation_matrix(self, theta):
        """[B,1] angle -> [B,2,3] pure-rotation affine matrix (reference model.py:220-235)"""
        cs = ops.angle_to_cs(theta)
        c, s = cs[:, 0:1], cs[:, 1:2]
        zero = torch.zeros_like(c)
        return torch.stack([torch.cat([c, -s, zero], dim=1), torch.cat([s, c, zero], dim=1)], dim=1)

    def localize(self, x):
        """localisation CNN -> (cos, sin) [B,2], theta [B,1]"""
        loc = self.localization
        h = _to_nhwc(x)
        h = ops.conv2d(h, loc[0].weight, loc[0].bias, 5, 5, 1, 2, ACT_RELU, pool=True)
        h = ops.conv2d(h, loc[3].weight, loc[3].bias, 5, 5, 1, 2, ACT_RELU, pool=True)
        h = ops.linear_nhwc(h, loc[7].weight, loc[7].bias, ACT_RELU)
        vec = ops.linear_nhwc(h.view(h.shape[0], 1, 1, -1), loc[9].weight, loc[9].bias)
        return ops.stn_head(vec)

    def forward(self, x):
        cs, theta = self.localize(x)
        x_rotated = ops.rot_sample(x, cs, 1.0)
        return x_rotated, theta


class Encoder(nn.Module):
    """reference model.py:265-326"""

    def __init__(self, in_channels: int = 1, latent_dim: int = 10, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.patch_size = patch_size
        self.rotation_stn = RotationSTN((in_channels, patch_size, patch_size))
        self.conv_layers = _encoder_stack()(in_channels)
        flat_size = 256 * (patch_size // 16) * (patch_size // 16)
        self.fc_mu = nn.Linear(flat_size, latent_dim)
        self.fc_logvar = nn.Linear(flat_size, latent_dim)

    def forward(self, x):
        if _ENGINE == "tc" and tc.supported(self.patch_size, self.latent_dim, x.shape[1]):
            loc, cl = self.rotation_stn.localization, self.conv_layers
            mu, logvar, theta, x_rot = tc.EncoderTc.apply(
                x, loc[0].weight, loc[0].bias, loc[3].weight, loc[3].bias, loc[7].weight,
                loc[7].bias, loc[9].weight, loc[9].bias, cl[0].weight, cl[0].bias,
                cl[2].weight, cl[2].bias, cl[4].weight, cl[4].bias, cl[6].weight, cl[6].bias,
                self.fc_mu.weight, self.fc_mu.bias, self.fc_logvar.weight, self.fc_logvar.bias, False)
            # one-shot cache for the trainer (see _take_canonical); training forwards only, so that an idle or
            # evaluating model holds no batch and no autograd graph (deepcopy / pickling of the module stay possible)
            self._canonical = (x, theta, x_rot) if torch.is_grad_enabled() else None
            return mu, logvar, theta
        x_rotated, theta = self.rotation_stn(x)
        h = _run_encoder_convs(self.conv_layers, x_rotated)
        mu = ops.linear_nhwc(h, self.fc_mu.weight, self.fc_mu.bias)
        logvar = ops.linear_nhwc(h, self.fc_logvar.weight, self.fc_logvar.bias)
        return mu, logvar, theta


def _theta_only(self, x):
    """theta of the STN alone, for callers that discard mu / logvar (same value and gradient as self(x)[2])"""
    if _ENGINE == "tc" and tc.supported(self.patch_size, self.latent_dim, x.shape[1]):
        loc, cl = self.rotation_stn.localization, self.conv_layers
        return tc.EncoderTc.apply(
            x, loc[0].weight, loc[0].bias, loc[3].weight, loc[3].bias, loc[7].weight, loc[7].bias, loc[9].weight,
            loc[9].bias, cl[0].weight, cl[0].bias, cl[2].weight, cl[2].bias, cl[4].weight, cl[4].bias, cl[6].weight,
            cl[6].bias, self.fc_mu.weight, self.fc_mu.bias, self.fc_logvar.weight, self.fc_logvar.bias, True)[2]
    return self.rotation_stn.localize(x)[1]


def _take_canonical(self, x, theta):
    """The STN's own rotated input, if the last forward of this encoder was for exactly (x, theta): it equals
    rotate_to_canonical(x, theta) (reference train.py:670-677) in value and, through autograd, in gradient (the
    normalise backward projects onto the tangent of the unit circle exactly as atan2 -> cos/sin does).  One-shot:
    the cache is dropped so no batch is kept alive.  None when unavailable (fp32 engine, different tensors)."""
    c, self._canonical = getattr(self, "_canonical", None), None
    if c is not None and c[0] is x and c[1] is theta:
        return c[2]
    return None


Encoder.take_canonical = _take_canonical
Encoder.theta_only = _theta_only


class Decoder(nn.Module):
    """reference model.py:329-388: relu(fc) -> 4x [Upsample x2 bilinear -> ReflectionPad2d(1) ->
    Conv3x3] + ReLU/Sigmoid"""

    def __init__(self, latent_dim: int = 10, out_channels: int = 1, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.out_channels = out_channels
        self.patch_size = patch_size
        inter_size = 256 * (patch_size // 16) * (patch_size // 16)
        self.fc = nn.Linear(latent_dim, inter_size)
        layers = []
        for ci, co, last in ((256, 128, False), (128, 64, False), (64, 32, False), (32, out_channels, True)):
            layers += [nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                       nn.ReflectionPad2d(1),
                       nn.Conv2d(ci, co, kernel_size=3, stride=1, padding=0),
                       nn.Sigmoid() if last else nn.ReLU()]
        self.deconv_layers = nn.Sequential(*layers)

    def forward(self, z):
        q = self.patch_size // 16
        if _ENGINE == "tc" and tc.supported(self.patch_size, self.latent_dim, self.out_channels):
            dl = self.deconv_layers
            return tc.DecoderTc.apply(z, self.fc.weight, self.fc.bias, dl[2].weight, dl[2].bias, dl[6].weight,
                                      dl[6].bias, dl[10].weight, dl[10].bias, dl[14].weight, dl[14].bias)
        h = ops.decoder_fc(z, self.fc.weight, self.fc.bias, 256, q)
        for n, i in enumerate((2, 6, 10, 14)):
            m = self.deconv_layers[i]
            h = ops.upsample_pad(h)
            h = ops.conv2d(h, m.weight, m.bias, 3, 3, 1, 0, ACT_SIGMOID if n == 3 else ACT_RELU)
        return _to_nchw(h)


class RVAE(nn.Module):
    """reference model.py:391-472"""

    def __init__(self, latent_dim: int = 10, in_channels: int = 1, patch_size: int = 64):
        super().__init__()
        self.latent_dim = latent_dim
        self.in_channels = in_channels
        self.patch_size = patch_size
        self.encoder = Encoder(in_channels, latent_dim, patch_size)
        self.decoder = Decoder(latent_dim, in_channels, patch_size)

    def reparameterize(self, mu, logvar):
        # eps comes from torch's generator exactly where the reference draws it (model.py:438)
        eps = torch.randn_like(logvar)
        return ops.reparam(mu, logvar, eps)

    def forward(self, x):
        mu, logvar, theta = self.encoder(x)
        z = self.reparameterize(mu, logvar)
        recon = self.decoder(z)
        # inverse rotation back to the input frame: get_rotation_matrix(-theta) (model.py:465-470)
        rotated_recon = ops.rot_sample(recon, ops.angle_to_cs(theta), -1.0)
        return rotated_recon, recon, theta, mu, logvar
