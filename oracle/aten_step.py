"""Like-for-like baseline: the reference's rVAE train-step body executed by STOCK ATen / cuDNN ops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's `gpu_baseline` leg (and the CPU
reference arm) as the thing the hand-written kernels are compared WITH, never by the product.

/root/reference does not exist on the GPU box, so its modules cannot be imported there.  This file
restates them over the functional pieces of oracle/rvae.py, with one difference that matters for a
baseline: the rotation is done by the very ops the reference calls, `F.affine_grid` +
`F.grid_sample(padding_mode="reflection", align_corners=False)` (model.py:254-258, 467-470,
train.py:675-677), instead of the hand restatement in `oracle.rvae.rot_sample_t`; convolutions,
linears, pooling, upsampling and padding are the same ATen ops the reference's nn.Modules dispatch to.
Under `amp=torch.float16/bfloat16` the forward runs inside `torch.autocast` with a GradScaler exactly
as the reference's default CUDA branch does (train.py:343-371); with `amp=None` it is the `--no-amp`
fp32 branch (train.py:372-397).  Optimiser: torch.optim.AdamW(lr=1e-3, weight_decay=1e-5)
(scripts/train_rvae.py:157-159); clipping: clip_grad_norm_(20) (train.py:313, 396).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import rvae as O


def rot_sample_aten(img, c, s):
    """grid_sample(img, affine_grid([[c,-s,0],[s,c,0]])) -- the reference's own two calls"""
    B = img.shape[0]
    c = c.reshape(B, 1)
    s = s.reshape(B, 1)
    z = torch.zeros_like(c)
    mat = torch.stack([torch.cat([c, -s, z], 1), torch.cat([s, c, z], 1)], 1)
    grid = F.affine_grid(mat, list(img.shape), align_corners=False)
    return F.grid_sample(img, grid, padding_mode="reflection", align_corners=False)


def _stn(p, x):
    vec = O.stn_vec(p, x)
    u = F.normalize(vec, dim=1, eps=1e-6)                       # model.py:245
    c, s = u[:, 0:1], u[:, 1:2]
    return rot_sample_aten(x, c, s), torch.atan2(s, c)           # model.py:250-261


def _encoder(p, x):
    x_can, theta = _stn(p, x)
    mu, logvar = O.enc_convs(p, x_can)
    return mu, logvar, theta


def loss_fn(p, x, x_rot, angle, eps, beta=10.0, gamma=10.0, canonical_weight=0.2):
    """model(x) + model.encoder(x_rot) + RVAELoss + canonical MSE (train.py:373-394) -> (loss, dict)"""
    P = x.shape[-1]
    mu, logvar, theta = _encoder(p, x)
    z = O.reparam(mu, logvar, eps.to(mu.dtype))
    recon = O.decoder_forward(p, z, P)
    rotated = rot_sample_aten(recon, torch.cos(-theta), torch.sin(-theta))        # model.py:465-470
    theta_rot = None
    if x_rot is not None:
        # train.py:376-377 runs the FULL encoder on x_rot although only theta is used; so does this
        _, _, theta_rot = _encoder(p, x_rot)
    total, rl, kl, cyc = O.rvae_loss(rotated.float(), x, mu.float(), logvar.float(), theta.float(),
                                     None if theta_rot is None else theta_rot.float(), angle, beta, gamma)
    can = torch.zeros((), device=x.device)
    if canonical_weight > 0:
        can_in = rot_sample_aten(x, torch.cos(theta).to(x.dtype), torch.sin(theta).to(x.dtype))  # train.py:675
        can = F.mse_loss(recon.float(), can_in.float(), reduction="mean")
        total = total + canonical_weight * can
    return total, dict(recon=rl, kld=kl, cycle=cyc, canonical=can, theta=theta, mu=mu, logvar=logvar,
                       rotated_recon=rotated, canonical_recon=recon)


class AtenTrainer:
    """Holds parameters (leaf tensors keyed by the reference's state_dict names), AdamW and the GradScaler."""

    def __init__(self, params: dict, device, amp=None, lr=1e-3, weight_decay=1e-5, max_norm=20.0):
        self.device = torch.device(device)
        self.p = {k: v.detach().to(self.device).clone().requires_grad_(True) for k, v in params.items()}
        self.opt = torch.optim.AdamW(list(self.p.values()), lr=lr, weight_decay=weight_decay)
        self.amp = amp
        self.max_norm = max_norm
        self.scaler = torch.amp.GradScaler("cuda") if (amp is not None and self.device.type == "cuda") else None

    def grads(self):
        return {k: v.grad for k, v in self.p.items()}

    def forward_backward(self, x, x_rot, angle, eps):
        self.opt.zero_grad(set_to_none=True)
        if self.amp is not None:
            with torch.autocast(self.device.type, dtype=self.amp):
                loss, aux = loss_fn(self.p, x, x_rot, angle, eps)
            if self.scaler is not None:
                self.scaler.scale(loss).backward()
                self.scaler.unscale_(self.opt)
            else:
                loss.backward()
        else:
            loss, aux = loss_fn(self.p, x, x_rot, angle, eps)
            loss.backward()
        return loss, aux

    def step(self, x, x_rot, angle, eps):
        loss, aux = self.forward_backward(x, x_rot, angle, eps)
        torch.nn.utils.clip_grad_norm_(list(self.p.values()), max_norm=self.max_norm)
        if self.scaler is not None:
            self.scaler.step(self.opt)
            self.scaler.update()
        else:
            self.opt.step()
        return loss


def time_gpu_baseline(params, batches, eps_list, amp, steps=5, warmup=2):
    """patches/s of the stock-ATen step (fwd + bwd + clip + AdamW) on resident device batches, CUDA events.
    -> (patches_per_s, ms_per_step) or raises torch.cuda.OutOfMemoryError"""
    dev = batches[0][0].device
    t = AtenTrainer(params, dev, amp=amp)
    B = batches[0][0].shape[0]
    for i in range(warmup):
        x, xr, ang = batches[i % len(batches)]
        t.step(x, xr, ang, eps_list[i % len(eps_list)])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        x, xr, ang = batches[i % len(batches)]
        t.step(x, xr, ang, eps_list[i % len(eps_list)])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    del t
    torch.cuda.empty_cache()
    return B / (ms * 1e-3), ms
