"""Oracle: functional restatement of the rVAE / VAE forward, losses and train step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure functions over a ``dict[str, Tensor]`` keyed by the reference's
``state_dict`` names.  Gradients come from torch
autograd on CPU.  ``eps`` (the reparameterisation noise) is always an explicit
argument: the reference draws it with ``torch.randn_like`` from the global
generator (model.py:438), and CPU/CUDA generators differ, so parity runs inject
the same tensor on both sides.

``quant`` is an optional callable applied to every GEMM operand (activations
and weights of conv / linear layers) before the op; it is used only to
*emulate* 16-bit tensor-core operand rounding when choosing tolerances.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _id(t):
    return t


# ----------------------------------------------------------------------------
# rotate + bilinear sample, differentiable torch restatement
# (same maths as oracle/rot_sample.py; reference model.py:250-258, 465-470,
#  train.py:670-677)
# ----------------------------------------------------------------------------
def rot_sample_t(img: torch.Tensor, c: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """img [B,C,H,W], c/s [B] or [B,1] -> rotated image, reflection padding."""
    B, C, H, W = img.shape
    dt = img.dtype
    c = c.reshape(B, 1, 1).to(dt)
    s = s.reshape(B, 1, 1).to(dt)
    xs = ((2.0 * torch.arange(W, dtype=dt) + 1.0) / W - 1.0).view(1, 1, W)
    ys = ((2.0 * torch.arange(H, dtype=dt) + 1.0) / H - 1.0).view(1, H, 1)
    gx = c * xs - s * ys
    gy = s * xs + c * ys

    def src(g, n):
        u = ((g + 1.0) * n - 1.0) / 2.0
        v = torch.abs(u + 0.5)
        extra = torch.fmod(v, float(n))
        flips = torch.floor(v / n)
        odd = (flips.to(torch.int64) % 2) == 1
        r = torch.where(odd, n - extra - 0.5, extra - 0.5)
        return r.clamp(0.0, n - 1.0)

    ix = src(gx, W)
    iy = src(gy, H)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx = (ix - x0).unsqueeze(1)
    fy = (iy - y0).unsqueeze(1)
    x0 = x0.to(torch.int64)
    y0 = y0.to(torch.int64)
    flat = img.reshape(B, C, H * W)

    def tap(yy, xx):
        ok = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)).unsqueeze(1)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, H * W)
        v = torch.gather(flat, 2, idx.expand(B, C, H * W)).reshape(B, C, H, W)
        return torch.where(ok, v, torch.zeros((), dtype=dt))

    return (tap(y0, x0) * (1 - fx) * (1 - fy) + tap(y0, x0 + 1) * fx * (1 - fy)
            + tap(y0 + 1, x0) * (1 - fx) * fy + tap(y0 + 1, x0 + 1) * fx * fy)


# ----------------------------------------------------------------------------
# RotationSTN (reference model.py:203-218 localisation, 237-262 forward)
# ----------------------------------------------------------------------------
def stn_vec(p, x, quant=_id, prefix="encoder.rotation_stn.localization."):
    """Localisation CNN -> unnormalised (cos, sin) vector [B,2]."""
    w = lambda n: quant(p[prefix + n + ".weight"])
    b = lambda n: p[prefix + n + ".bias"]
    h = F.max_pool2d(F.relu(F.conv2d(quant(x), w("0"), b("0"), padding=2)), 2, 2)
    h = F.max_pool2d(F.relu(F.conv2d(quant(h), w("3"), b("3"), padding=2)), 2, 2)
    h = h.flatten(1)                                 # NCHW flatten (model.py:210)
    h = F.relu(F.linear(quant(h), w("7"), b("7")))
    return F.linear(h, p[prefix + "9.weight"], b("9"))


def stn_forward(p, x, quant=_id):
    """-> (x_canonical, theta [B,1], cos [B], sin [B])."""
    vec = stn_vec(p, x, quant)
    n = vec.norm(dim=1, keepdim=True).clamp_min(1e-6)   # F.normalize eps (model.py:245)
    u = vec / n
    c, s = u[:, 0], u[:, 1]
    x_can = rot_sample_t(x, c, s)                        # model.py:250-258
    theta = torch.atan2(s, c).unsqueeze(1)               # model.py:261
    return x_can, theta, c, s


# ----------------------------------------------------------------------------
# Encoder conv stack + heads (reference model.py:289-303, 319-324; VAE 29-43)
# ----------------------------------------------------------------------------
def enc_convs(p, x, quant=_id):
    h = x
    for i in ("0", "2", "4", "6"):
        h = F.relu(F.conv2d(quant(h), quant(p[f"encoder.conv_layers.{i}.weight"]),
                            p[f"encoder.conv_layers.{i}.bias"], stride=2, padding=1))
    h = quant(h.flatten(1))
    mu = F.linear(h, quant(p["encoder.fc_mu.weight"]), p["encoder.fc_mu.bias"])
    logvar = F.linear(h, quant(p["encoder.fc_logvar.weight"]), p["encoder.fc_logvar.bias"])
    return mu, logvar


def encoder_forward(p, x, quant=_id):
    """reference Encoder.forward model.py:305-326 -> (mu, logvar, theta, x_can, c, s)."""
    x_can, theta, c, s = stn_forward(p, x, quant)
    mu, logvar = enc_convs(p, x_can, quant)
    return mu, logvar, theta, x_can, c, s


def reparam(mu, logvar, eps):
    """reference model.py:426-440 with eps injected."""
    return mu + eps * torch.exp(0.5 * logvar)


# ----------------------------------------------------------------------------
# Decoder (reference model.py:353-373, 383-388): relu(fc) -> 4x
# [bilinear x2 (align_corners=False) -> ReflectionPad2d(1) -> conv3x3] + ReLU/Sigmoid
# ----------------------------------------------------------------------------
def decoder_forward(p, z, patch_size, quant=_id):
    q = patch_size // 16
    h = F.relu(F.linear(quant(z), quant(p["decoder.fc.weight"]), p["decoder.fc.bias"]))
    h = h.view(-1, 256, q, q)
    for n, i in enumerate(("2", "6", "10", "14")):
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)
        h = F.pad(h, (1, 1, 1, 1), mode="reflect")
        h = F.conv2d(quant(h), quant(p[f"decoder.deconv_layers.{i}.weight"]),
                     p[f"decoder.deconv_layers.{i}.bias"])
        h = torch.sigmoid(h) if n == 3 else F.relu(h)
    return h


def rvae_forward(p, x, eps, quant=_id):
    """reference RVAE.forward model.py:442-472 ->
    (rotated_recon, recon, theta, mu, logvar, x_can)."""
    P = x.shape[-1]
    mu, logvar, theta, x_can, c, s = encoder_forward(p, x, quant)
    z = reparam(mu, logvar, eps)
    recon = decoder_forward(p, z, P, quant)
    # inverse rotation: get_rotation_matrix(-theta) (model.py:220-235, 465-470)
    ci = torch.cos(-theta).squeeze(1)
    si = torch.sin(-theta).squeeze(1)
    rotated = rot_sample_t(recon, ci, si)
    return rotated, recon, theta, mu, logvar, x_can


# ----------------------------------------------------------------------------
# Losses (reference loss.py:52-94 cycle, 32-49 diversity, 138-186 RVAELoss,
# 104-122 VAELoss)
# ----------------------------------------------------------------------------
def cycle_loss(theta, theta_rot, angle):
    d = (theta_rot.reshape(-1) - theta.reshape(-1)) + angle.reshape(-1)
    return (1.0 - torch.cos(d)).mean()


def rvae_loss(rotated, x, mu, logvar, theta=None, theta_rot=None, angle=None,
              beta=1.0, gamma=0.0, use_diversity=False):
    B = x.shape[0]
    recon = ((rotated - x) ** 2).sum() / B
    kld = (-0.5 * (1 + logvar - mu ** 2 - logvar.exp()).sum(1)).mean()
    rot = torch.zeros((), dtype=x.dtype)
    if gamma > 0:
        if use_diversity and theta is not None:
            rot = (theta.std() - 1.0) ** 2
        elif theta is not None and theta_rot is not None and angle is not None:
            rot = cycle_loss(theta, theta_rot, angle)
    return recon + beta * kld + gamma * rot, recon, kld, rot


def vae_loss(recon_x, x, mu, logvar, beta=1.0):
    recon = ((recon_x - x) ** 2).mean()
    kld = -0.5 * (1 + logvar - mu ** 2 - logvar.exp()).mean()
    return recon + beta * kld, recon, kld


# ----------------------------------------------------------------------------
# Full train-step bodies
# ----------------------------------------------------------------------------
def rvae_full_step(p, x, x_rot, angle, eps, beta=10.0, gamma=10.0,
                   canonical_weight=0.2, use_diversity=False, quant=_id):
    """reference train_rvae_one_epoch batch body, train.py:373-395 (no-AMP branch),
    without clip/optimizer/metrics.  Returns (outs dict, grads dict)."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    rotated, recon, theta, mu, logvar, x_can = rvae_forward(p, x, eps, quant)
    theta_rot = None
    if x_rot is not None:                                   # train.py:376-377
        _, theta_rot, _, _ = stn_forward(p, x_rot, quant)
    total, rl, kl, cyc = rvae_loss(rotated, x, mu, logvar, theta, theta_rot, angle,
                                   beta, gamma, use_diversity)
    can = torch.zeros(())
    if canonical_weight > 0:                                # train.py:386-394
        # rotate_to_canonical: get_rotation_matrix(theta) -> cos/sin of atan2
        can_in = rot_sample_t(x, torch.cos(theta).squeeze(1), torch.sin(theta).squeeze(1))
        can = ((recon - can_in) ** 2).mean()
        total = total + canonical_weight * can
    total.backward()
    outs = dict(rotated_recon=rotated, recon=recon, theta=theta, mu=mu, logvar=logvar,
                x_can=x_can, theta_rot=theta_rot, loss=total, recon_loss=rl, kld=kl,
                cycle=cyc, canonical=can)
    outs = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in outs.items()}
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v))
             for k, v in p.items()}
    return outs, grads


def vae_encoder(p, x, quant=_id):
    """reference VAEEncoder.forward model.py:45-61."""
    return enc_convs(p, x, quant)


def vae_decoder(p, z, patch_size, quant=_id):
    """reference VAEDecoder.forward model.py:100-113: relu(fc) -> 4x ConvTranspose2d(k4,s2,p1)."""
    q = patch_size // 16
    h = F.relu(F.linear(quant(z), quant(p["decoder.fc.weight"]), p["decoder.fc.bias"]))
    h = h.view(-1, 256, q, q)
    for n, i in enumerate(("0", "2", "4", "6")):
        h = F.conv_transpose2d(quant(h), quant(p[f"decoder.deconv_layers.{i}.weight"]),
                               p[f"decoder.deconv_layers.{i}.bias"], stride=2, padding=1)
        h = torch.sigmoid(h) if n == 3 else F.relu(h)
    return h


def vae_full_step(p, x, eps, beta=1.0, quant=_id):
    """reference train_one_epoch VAE branch, train.py:76-84 + backward."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    mu, logvar = vae_encoder(p, x, quant)
    z = reparam(mu, logvar, eps)
    recon = vae_decoder(p, z, x.shape[-1], quant)
    total, rl, kl = vae_loss(recon, x, mu, logvar, beta)
    total.backward()
    outs = dict(recon=recon.detach(), mu=mu.detach(), logvar=logvar.detach(),
                loss=total.detach(), recon_loss=rl.detach(), kld=kl.detach())
    grads = {k: v.grad.detach() for k, v in p.items()}
    return outs, grads


def stn_pretrain_step(p, x, x_rot, angle, quant=_id):
    """reference scripts/pretrain_stn.py:104-112: two encoder passes, cycle loss,
    backward (reaches the STN localisation only)."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    _, th0, _, _ = stn_forward(p, x, quant)
    _, th1, _, _ = stn_forward(p, x_rot, quant)
    loss = cycle_loss(th0, th1, angle)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items() if v.grad is not None}
    return dict(theta=th0.detach(), theta_rot=th1.detach(), loss=loss.detach()), grads


# ----------------------------------------------------------------------------
# Deterministic, library-independent parameter / input generators so golden
# vectors only need to store OUTPUTS.  numpy PCG64 streams are stable.
# ----------------------------------------------------------------------------
def rvae_param_shapes(P: int, L: int) -> dict[str, tuple]:
    q = P // 16
    s = {}
    loc = "encoder.rotation_stn.localization."
    s[loc + "0.weight"] = (16, 1, 5, 5); s[loc + "0.bias"] = (16,)
    s[loc + "3.weight"] = (32, 16, 5, 5); s[loc + "3.bias"] = (32,)
    s[loc + "7.weight"] = (32, 32 * (P // 4) ** 2); s[loc + "7.bias"] = (32,)
    s[loc + "9.weight"] = (2, 32); s[loc + "9.bias"] = (2,)
    for i, (ci, co) in zip(("0", "2", "4", "6"), ((1, 32), (32, 64), (64, 128), (128, 256))):
        s[f"encoder.conv_layers.{i}.weight"] = (co, ci, 4, 4)
        s[f"encoder.conv_layers.{i}.bias"] = (co,)
    for h in ("fc_mu", "fc_logvar"):
        s[f"encoder.{h}.weight"] = (L, 256 * q * q); s[f"encoder.{h}.bias"] = (L,)
    s["decoder.fc.weight"] = (256 * q * q, L); s["decoder.fc.bias"] = (256 * q * q,)
    for i, (ci, co) in zip(("2", "6", "10", "14"), ((256, 128), (128, 64), (64, 32), (32, 1))):
        s[f"decoder.deconv_layers.{i}.weight"] = (co, ci, 3, 3)
        s[f"decoder.deconv_layers.{i}.bias"] = (co,)
    return s


def vae_param_shapes(P: int, L: int) -> dict[str, tuple]:
    q = P // 16
    s = {}
    for i, (ci, co) in zip(("0", "2", "4", "6"), ((1, 32), (32, 64), (64, 128), (128, 256))):
        s[f"encoder.conv_layers.{i}.weight"] = (co, ci, 4, 4)
        s[f"encoder.conv_layers.{i}.bias"] = (co,)
    for h in ("fc_mu", "fc_logvar"):
        s[f"encoder.{h}.weight"] = (L, 256 * q * q); s[f"encoder.{h}.bias"] = (L,)
    s["decoder.fc.weight"] = (256 * q * q, L); s["decoder.fc.bias"] = (256 * q * q,)
    for i, (ci, co) in zip(("0", "2", "4", "6"), ((256, 128), (128, 64), (64, 32), (32, 1))):
        s[f"decoder.deconv_layers.{i}.weight"] = (ci, co, 4, 4)   # ConvTranspose2d layout
        s[f"decoder.deconv_layers.{i}.bias"] = (co,)
    return s


def make_params(shapes: dict[str, tuple], seed: int, stn_head_std: float | None = None,
                dtype=torch.float32) -> dict[str, torch.Tensor]:
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for everything (torch's default scale);
    ``stn_head_std`` overrides the STN's last Linear with N(0, std) / zero bias the
    way model.py:217-218 does (std=0.01 there; tests use a larger std so theta is
    not degenerate)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = {}
    fan = 1
    for name, shp in shapes.items():
        if name.endswith(".weight"):
            if "decoder.deconv_layers" in name and len(shp) == 4 and shp[2] == 4:
                fan = shp[1] * shp[2] * shp[3]            # ConvTranspose2d fan_in as torch computes it
            else:
                fan = int(math.prod(shp[1:]))
        bound = 1.0 / math.sqrt(fan)
        a = rng.uniform(-bound, bound, size=shp)
        if stn_head_std is not None and "localization.9." in name:
            a = rng.normal(0.0, stn_head_std, size=shp) if name.endswith("weight") else np.zeros(shp)
        out[name] = torch.from_numpy(a).to(dtype)
    return out


def make_lattice_batch(B: int, P: int, seed: int, dtype=torch.float32):
    """Synthetic lattice-like patches in [0,1] plus a rotated partner and the angle.
    Cheap stand-in for the paired dataset (data.py:617-735) for parity tests."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(P) - P / 2 + 0.5, np.arange(P) - P / 2 + 0.5, indexing="ij")
    x = np.zeros((B, 1, P, P)); xr = np.zeros((B, 1, P, P))
    ang = rng.uniform(0, 2 * np.pi, size=B)
    for b in range(B):
        a = rng.uniform(8.0, 14.0); ph = rng.uniform(0, 2 * np.pi, size=3); t0 = rng.uniform(0, np.pi)
        def img(rot):
            v = np.zeros((P, P))
            for k in range(3):
                t = t0 + rot + k * np.pi / 3
                v += np.cos(2 * np.pi / a * (np.cos(t) * xx + np.sin(t) * yy) + ph[k])
            v = v + rng.normal(0, 0.15, size=(P, P))
            return (v - v.min()) / (v.max() - v.min())
        x[b, 0] = img(0.0); xr[b, 0] = img(ang[b])
    return (torch.from_numpy(x).to(dtype), torch.from_numpy(xr).to(dtype),
            torch.from_numpy(ang).to(dtype))
